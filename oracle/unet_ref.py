"""Oracle (test infrastructure): functional fp32 restatement of the reference UNet.

Follows ``/root/reference/v_diffusion/models/unet.py`` (UNet.forward 286-322,
ResidualBlock.forward 137-148, AttentionBlock.forward 73-81,
BaseAttentionBlock.scaled_dot_product 55-64) and ``modules.py`` (OneHot 184-201)
without sharing code with it: the model is described by a flat state_dict plus
the constructor integers, and evaluated as one straight-line function.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .diffusion_ref import timestep_embedding

GN_GROUPS = 32      # unet.py:28-30
GN_EPS = 1e-6       # unet.py:29  (not torch's default 1e-5)


def unet_config_from_json(model_cfg: dict, in_channels: int, out_channels: int,
                          num_classes: int = 0, multitags: bool = False) -> dict:
    """Normalise a ``config["model"]`` block (after defaults merge) into ctor integers
    the way ``UNet.__init__`` does (unet.py:155-200)."""
    cfg = dict(model_cfg)
    cfg.pop("use_xformers", None)
    cfg.pop("drop_rate", None)
    cfg.pop("resample_with_res", None)
    hid = cfg["hid_channels"]
    levels = len(cfg["ch_multipliers"])
    apply_attn = cfg["apply_attn"]
    if isinstance(apply_attn, bool):
        apply_attn = [apply_attn] * levels
    head_dim, num_heads = cfg.get("head_dim"), cfg.get("num_heads")
    if head_dim is None and num_heads is None:
        num_heads = 1
    return dict(
        in_channels=cfg.get("in_channels", in_channels), hid_channels=hid, out_channels=out_channels,
        ch_multipliers=list(cfg["ch_multipliers"]), num_res_blocks=cfg["num_res_blocks"],
        apply_attn=list(apply_attn), embedding_dim=cfg.get("embedding_dim") or 4 * hid,
        head_dim=head_dim, num_heads=num_heads, num_classes=num_classes, multitags=multitags)


def _attn_dims(c: int, head_dim: Optional[int], num_heads: Optional[int]):
    # unet.py:43-53
    if head_dim is None:
        head_dim = c // num_heads
    if num_heads is None:
        num_heads = c // head_dim
    return head_dim, num_heads


def block_plan(cfg: dict) -> List[dict]:
    """Execution order of the UNet as a list of ops (unet.py:250-283, 297-321).

    Each entry: ``{"kind": "res"|"attn", "name": state_dict prefix, "cin", "cout",
    "resample": "none"|"down"|"up", "concat": bool, "push": bool}``.
    """
    hid, mult, nrb = cfg["hid_channels"], cfg["ch_multipliers"], cfg["num_res_blocks"]
    attn, levels = cfg["apply_attn"], len(cfg["ch_multipliers"])
    chs = [hid * m for m in mult]
    plan: List[dict] = []

    def add(prefix, cin, cout, level, resample="none", concat=False, push=False):
        has_attn = attn[level] if level is not None else False
        rp = prefix + (".0" if has_attn else "")
        plan.append(dict(kind="res", name=rp, cin=cin, cout=cout, resample=resample,
                         concat=concat, push=push and not has_attn))
        if has_attn:
            plan.append(dict(kind="attn", name=prefix + ".1", cin=cout, cout=cout,
                             resample="none", concat=False, push=push))

    for i in range(levels):
        prev = chs[i - 1] if i else hid
        add(f"downsamples.level_{i}.0", prev, chs[i], i, push=True)
        for j in range(1, nrb):
            add(f"downsamples.level_{i}.{j}", chs[i], chs[i], i, push=True)
        if i != levels - 1:
            add(f"downsamples.level_{i}.{nrb}", chs[i], chs[i], i, resample="down", push=True)
    mid = chs[-1]
    plan.append(dict(kind="res", name="middle.0", cin=mid, cout=mid, resample="none", concat=False, push=False))
    plan.append(dict(kind="attn", name="middle.1", cin=mid, cout=mid, resample="none", concat=False, push=False))
    plan.append(dict(kind="res", name="middle.2", cin=mid, cout=mid, resample="none", concat=False, push=False))
    for i in range(levels - 1, -1, -1):
        nxt = hid if i == 0 else chs[i - 1]
        prev = chs[-1] if i == levels - 1 else chs[i + 1]
        cur = chs[i]
        add(f"upsamples.level_{i}.0", prev + cur, cur, i, concat=True)
        for j in range(1, nrb):
            add(f"upsamples.level_{i}.{j}", 2 * cur, cur, i, concat=True)
        add(f"upsamples.level_{i}.{nrb}", nxt + cur, cur, i, concat=True)
        if i != 0:
            add(f"upsamples.level_{i}.{nrb + 1}", cur, cur, i, resample="up")
    return plan


def state_dict_shapes(cfg: dict) -> Dict[str, tuple]:
    """Key -> shape of the reference state_dict (checked against the real
    module in tests/golden/make_golden.py)."""
    hid, E = cfg["hid_channels"], cfg["embedding_dim"]
    shapes: Dict[str, tuple] = {
        "time_embed.0.weight": (E, hid), "time_embed.0.bias": (E,),
        "time_embed.2.weight": (E, E), "time_embed.2.bias": (E,),
    }
    if cfg["num_classes"] > 0:
        p = "class_embed" if cfg["multitags"] else "class_embed.1"
        shapes[p + ".weight"] = (E, cfg["num_classes"])
        shapes[p + ".bias"] = (E,)
    shapes["in_conv.weight"] = (hid, cfg["in_channels"], 3, 3)
    shapes["in_conv.bias"] = (hid,)
    for op in block_plan(cfg):
        n, cin, cout = op["name"], op["cin"], op["cout"]
        if op["kind"] == "res":
            shapes[n + ".norm1.weight"] = (cin,); shapes[n + ".norm1.bias"] = (cin,)
            shapes[n + ".conv1.weight"] = (cout, cin, 3, 3); shapes[n + ".conv1.bias"] = (cout,)
            shapes[n + ".fc.weight"] = (2 * cout, E); shapes[n + ".fc.bias"] = (2 * cout,)
            shapes[n + ".norm2.weight"] = (cout,); shapes[n + ".norm2.bias"] = (cout,)
            shapes[n + ".conv2.weight"] = (cout, cout, 3, 3); shapes[n + ".conv2.bias"] = (cout,)
            if cin != cout:
                shapes[n + ".skip.weight"] = (cout, cin, 1, 1); shapes[n + ".skip.bias"] = (cout,)
        else:
            hd, nh = _attn_dims(cin, cfg["head_dim"], cfg["num_heads"])
            shapes[n + ".norm.weight"] = (cin,); shapes[n + ".norm.bias"] = (cin,)
            shapes[n + ".proj_in.weight"] = (3 * hd * nh, cin, 1, 1); shapes[n + ".proj_in.bias"] = (3 * hd * nh,)
            shapes[n + ".proj_out.weight"] = (cin, hd * nh, 1, 1); shapes[n + ".proj_out.bias"] = (cin,)
    c0 = hid * cfg["ch_multipliers"][0]
    shapes["out_conv.0.weight"] = (c0,); shapes["out_conv.0.bias"] = (c0,)
    shapes["out_conv.2.weight"] = (cfg["out_channels"], c0, 3, 3); shapes["out_conv.2.bias"] = (cfg["out_channels"],)
    return shapes


def make_state_dict(cfg: dict, seed: int = 0, residual_gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic weights with NO all-zero tensors (the
    reference's own init zeroes conv2/proj_out/out_conv so its random-init output is
    identically 0 and exercises nothing).  Recipe: for every key in sorted order, one
    CPU generator seeded with ``seed``: >=2-D tensors ~ N(0, 1/fan_in), norm weights
    ~ 1 + 0.1 N(0,1), every bias ~ 0.1 N(0,1).  ``residual_gain`` scales conv2 /
    proj_out (the tensors the reference zero-inits) to control how fast the residual
    stream grows through the 27 blocks."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    shapes = state_dict_shapes(cfg)
    for k in sorted(shapes):
        shp = shapes[k]
        if len(shp) >= 2:
            fan_in = shp[1] * (shp[2] * shp[3] if len(shp) == 4 else 1)
            w = torch.randn(shp, generator=g) / math.sqrt(fan_in)
            if k.endswith("conv2.weight") or k.endswith("proj_out.weight"):
                w = w * residual_gain
            sd[k] = w
        elif (".norm" in k and k.endswith("weight")) or k == "out_conv.0.weight":
            sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            sd[k] = 0.1 * torch.randn(shp, generator=g)
    return sd


def dezero_(sd: Dict[str, torch.Tensor], seed: int = 7) -> Dict[str, torch.Tensor]:
    """In-place: replace every all-zero >=2-D tensor of a reference-initialised
    state_dict by N(0, 1/fan_in) draws (SURVEY §8d cfg 1 recipe, generator seed 7)."""
    g = torch.Generator().manual_seed(seed)
    for k in sorted(sd):
        v = sd[k]
        if v.ndim >= 2 and not bool(v.any()):
            fan_in = v.shape[1] * (v.shape[2] * v.shape[3] if v.ndim == 4 else 1)
            sd[k] = torch.randn(v.shape, generator=g) / math.sqrt(fan_in)
    return sd


# Optional numerics model of the CUDA path (test / calibration aid, off by default): ``operand_round`` of
# unet_forward names a 16-bit format ("fp16" | "bf16"); every tensor the CUDA path holds in that format -- GEMM operands
# (normalised activations, packed weights, the raw concat feeding the 1x1 skip conv), conv1's output, q / k / v,
# the softmax probabilities and the attention output -- is rounded to it here, everything else stays fp32.
_ROUND = {"fp16": torch.float16, "bf16": torch.bfloat16}
_round_dtype = None


def _q(x):
    if _round_dtype is None:
        return x
    if _round_dtype == torch.float16:
        x = x.clamp(-65504.0, 65504.0)                                # the CUDA conversions saturate
    return x.to(_round_dtype).float()


def _conv(x, w, b, padding=0):
    return F.conv2d(_q(x), _q(w), b, padding=padding)


def _gn(x, sd, prefix):
    return F.group_norm(x, GN_GROUPS, sd[prefix + ".weight"], sd[prefix + ".bias"], GN_EPS)


def _resample(x, mode):
    if mode == "down":
        return F.avg_pool2d(x, 2)                                   # unet.py:130
    if mode == "up":
        return F.interpolate(x, scale_factor=2, mode="nearest")     # unet.py:128
    return x


def _res_block(x, emb_act, sd, op):
    n = op["name"]
    skip = _resample(x, op["resample"])                             # unet.py:138
    if op["cin"] != op["cout"]:
        skip = _conv(skip, sd[n + ".skip.weight"], sd[n + ".skip.bias"])
    h = _resample(F.silu(_gn(x, sd, n + ".norm1")), op["resample"])  # unet.py:141 norm->act->resample
    h = _q(_conv(h, sd[n + ".conv1.weight"], sd[n + ".conv1.bias"], padding=1))    # kept in 16 bits on the CUDA path
    film = F.linear(emb_act, sd[n + ".fc.weight"], sd[n + ".fc.bias"])[:, :, None, None]
    shift, scale = film.chunk(2, dim=1)                             # unet.py:145 (shift first)
    h = (1 + scale) * _gn(h, sd, n + ".norm2") + shift
    h = _conv(F.silu(h), sd[n + ".conv2.weight"], sd[n + ".conv2.bias"], padding=1)
    return h + skip


def _attn_block(x, sd, op, cfg):
    n = op["name"]
    B, C, H, W = x.shape
    hd, nh = _attn_dims(C, cfg["head_dim"], cfg["num_heads"])
    qkv = _q(_conv(_gn(x, sd, n + ".norm"), sd[n + ".proj_in.weight"], sd[n + ".proj_in.bias"]))
    q, k, v = qkv.reshape(B, 3 * nh, hd, H * W).chunk(3, dim=1)     # unet.py:76-78: all-q | all-k | all-v
    w = torch.einsum("bncq,bnck->bnqk", q, k) / math.sqrt(hd)       # unet.py:58-60
    if _round_dtype is None:
        w = torch.softmax(w, dim=-1)
    else:                                                           # un-normalised probabilities are rounded, the sum is fp32
        e = torch.exp(w - w.amax(dim=-1, keepdim=True))
        w = _q(e) / e.sum(dim=-1, keepdim=True)
    o = _q(torch.einsum("bnqk,bnck->bncq", w, v).reshape(B, nh * hd, H, W))
    o = _conv(o, sd[n + ".proj_out.weight"], sd[n + ".proj_out.bias"])
    return o + x


def embedding(sd: Dict[str, torch.Tensor], cfg: dict, t: torch.Tensor, y: Optional[torch.Tensor]):
    """time/class embedding (unet.py:287-295); returns the *pre-activation* t_emb."""
    e = timestep_embedding(t, cfg["hid_channels"])
    e = F.linear(e, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    e = F.linear(F.silu(e), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    if cfg["num_classes"] and y is not None:
        if cfg["multitags"]:
            yy = y.float()
            yy = yy / torch.count_nonzero(yy, dim=1).clamp(min=1.).sqrt().unsqueeze(1)
            e = e + F.linear(yy, sd["class_embed.weight"], sd["class_embed.bias"])
        else:
            yl = y.long()
            onehot = F.one_hot((yl - 1).clamp(min=0), cfg["num_classes"]).float()
            onehot[yl == 0] = 0                                      # modules.py:193-197
            e = e + F.linear(onehot, sd["class_embed.1.weight"], sd["class_embed.1.bias"])
    return e


@torch.no_grad()
def unet_forward(sd: Dict[str, torch.Tensor], cfg: dict, x: torch.Tensor, t: torch.Tensor,
                 y: Optional[torch.Tensor] = None, trace: Optional[dict] = None,
                 operand_round: Optional[str] = None) -> torch.Tensor:
    """fp32 NCHW in, fp32 NCHW out.  ``t`` may be fp64 (sampler) — the sinusoid is
    evaluated in t's dtype (functions.py:20-25).  ``trace`` (optional dict) receives
    the output of every block keyed by its state_dict prefix, for layer-wise checks."""
    global _round_dtype
    _round_dtype = _ROUND[operand_round] if operand_round else None
    try:
        return _unet_forward(sd, cfg, x, t, y, trace)
    finally:
        _round_dtype = None


def _unet_forward(sd, cfg, x, t, y, trace):
    x = x.float()
    emb_act = F.silu(embedding(sd, cfg, t, y))                      # unet.py:142: fc(act(t_emb))
    hs = [_conv(x, sd["in_conv.weight"], sd["in_conv.bias"], padding=1)]
    if trace is not None:
        trace["in_conv"] = hs[0]
    h = hs[0]
    for op in block_plan(cfg):
        if op["concat"]:
            h = torch.cat([h, hs.pop()], dim=1)                     # unet.py:315 (h first)
        h = _res_block(h, emb_act, sd, op) if op["kind"] == "res" else _attn_block(h, sd, op, cfg)
        if trace is not None:
            trace[op["name"]] = h
        if op["push"]:
            hs.append(h)
    assert len(hs) == 0, len(hs)
    h = F.silu(_gn(h, sd, "out_conv.0"))
    return _conv(h, sd["out_conv.2.weight"], sd["out_conv.2.bias"], padding=1)
