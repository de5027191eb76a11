"""Drop-in for ``v_diffusion.models.unet.UNet`` (unet.py:151-322) on B200.

Same constructor, same ``forward(x, t, y=None)``, same parameter names and shapes (so a reference
``state_dict`` / checkpoint loads unchanged) — but the module tree only *holds*
parameters.  ``forward`` hands device pointers to the C ABI (``vdt_unet_forward``), which runs the
hand-written sm_100a kernels.  There is no PyTorch compute path and no CPU fallback.
"""
import ctypes as C
import math
import os

import torch
import torch.nn as nn

from . import _lib


def _lecun_normal_(w, scale=1.):
    # modules.py:25-35: truncated normal (+-2 std) scaled by sqrt(scale / fan_in); scale 0 -> zeros
    nn.init.trunc_normal_(w, mean=0., std=1., a=-2., b=2.)
    fan_in = w.shape[1] * (math.prod(w.shape[2:]) if w.ndim > 2 else 1)
    with torch.no_grad():
        w.mul_(math.sqrt(scale / fan_in))
    return w


class _Affine(nn.Module):
    """weight/bias holder for Linear, Conv2d (OIHW) and GroupNorm."""

    def __init__(self, shape, init_scale=1., norm=False):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(shape, dtype=torch.float32))
        self.bias = nn.Parameter(torch.zeros(shape[0], dtype=torch.float32))
        if norm:
            nn.init.ones_(self.weight)
        else:
            _lecun_normal_(self.weight, init_scale)


class _Slot(nn.Module):
    """parameter-less placeholder keeping the reference's Sequential indices (SiLU, OneHot)."""


class _Res(nn.Module):
    def __init__(self, cin, cout, embed_dim):
        super().__init__()
        self.norm1 = _Affine((cin,), norm=True)
        self.conv1 = _Affine((cout, cin, 3, 3))
        self.fc = _Affine((2 * cout, embed_dim))
        self.norm2 = _Affine((cout,), norm=True)
        self.conv2 = _Affine((cout, cout, 3, 3), init_scale=0.)
        if cin != cout:
            self.skip = _Affine((cout, cin, 1, 1))


class _Attn(nn.Module):
    def __init__(self, cin, head_dim, num_heads):
        super().__init__()
        if head_dim is None:                                    # unet.py:43-49
            assert num_heads is not None and cin % num_heads == 0
            head_dim = cin // num_heads
        if num_heads is None:
            assert head_dim is not None and cin % head_dim == 0
            num_heads = cin // head_dim
        hid = head_dim * num_heads
        self.norm = _Affine((cin,), norm=True)
        self.proj_in = _Affine((3 * hid, cin, 1, 1))
        self.proj_out = _Affine((cin, hid, 1, 1), init_scale=0.)


class UNet(nn.Module):
    def __init__(self, in_channels, hid_channels, out_channels, ch_multipliers, num_res_blocks, apply_attn,
                 embedding_dim=None, drop_rate=0., head_dim=None, num_heads=None, num_classes=0, multitags=False,
                 resample_with_res=True, use_xformers=False):
        super().__init__()
        if not resample_with_res:
            raise NotImplementedError("resample_with_res=False is not used by any reference config")
        self.in_channels, self.hid_channels, self.out_channels = in_channels, hid_channels, out_channels
        self.embedding_dim = embedding_dim or 4 * hid_channels
        self.levels = levels = len(ch_multipliers)
        self.ch_multipliers = list(ch_multipliers)
        if isinstance(apply_attn, bool):
            apply_attn = [apply_attn] * levels
        self.apply_attn = list(apply_attn)
        self.num_res_blocks = num_res_blocks
        self.drop_rate = drop_rate                              # active in .train(); identity in .eval()
        if head_dim is None and num_heads is None:
            num_heads = 1
        self.head_dim, self.num_heads = head_dim, num_heads
        self.num_classes, self.multitags = num_classes, multitags
        E, hid = self.embedding_dim, hid_channels

        self.time_embed = nn.ModuleList([_Affine((E, hid)), _Slot(), _Affine((E, E))])
        if num_classes > 0:
            if multitags:
                self.class_embed = nn.Linear(num_classes, E)          # stock nn.Linear and init (unet.py:209-210)
            else:
                self.class_embed = nn.ModuleList([_Slot(), _Affine((E, num_classes))])
        self.in_conv = _Affine((hid, in_channels, 3, 3))
        chs = [hid * m for m in ch_multipliers]

        def block(level, cin, cout):
            if self.apply_attn[level]:
                return nn.ModuleList([_Res(cin, cout, E), _Attn(cout, head_dim, num_heads)])
            return _Res(cin, cout, E)

        self.downsamples = nn.ModuleDict()
        for i in range(levels):
            prev = chs[i - 1] if i else hid
            mods = [block(i, prev, chs[i])] + [block(i, chs[i], chs[i]) for _ in range(num_res_blocks - 1)]
            if i != levels - 1:
                mods.append(block(i, chs[i], chs[i]))
            self.downsamples[f"level_{i}"] = nn.ModuleList(mods)
        mid = chs[-1]
        self.middle = nn.ModuleList([_Res(mid, mid, E), _Attn(mid, head_dim, num_heads), _Res(mid, mid, E)])
        self.upsamples = nn.ModuleDict()
        for i in range(levels):
            nxt = hid if i == 0 else chs[i - 1]
            prev = chs[-1] if i == levels - 1 else chs[i + 1]
            cur = chs[i]
            mods = [block(i, prev + cur, cur)] + [block(i, 2 * cur, cur) for _ in range(num_res_blocks - 1)]
            mods.append(block(i, nxt + cur, cur))
            if i != 0:
                mods.append(block(i, cur, cur))
            self.upsamples[f"level_{i}"] = nn.ModuleList(mods)
        self.out_conv = nn.ModuleList([_Affine((chs[0],), norm=True), _Slot(),
                                       _Affine((out_channels, chs[0], 3, 3), init_scale=0.)])
        # runtime state (not part of the reference surface)
        self.max_rows = int(os.environ.get("VDT_MAX_ROWS", "1024"))
        # 16-bit tensor-core operand format: "fp16" (default: 10-bit mantissa like the TF32 the reference's own
        # cuDNN convs use on GPU, same tcgen05 rate as bf16) or "bf16"
        self.operand_dtype = os.environ.get("VDT_OPERAND", "fp16")
        # autograd = True: forward() under torch.enable_grad() returns a tensor attached to the autograd graph -- its backward
        # is the explicit kernel tape of training.UNetTrainGraph and fills every parameter's .grad -- so the reference's own
        # training loop (loss.backward(); clip_grad_norm_; optimizer.step(), train_utils.py:149-166) runs unchanged on this
        # module.  Off by default: forward() is then inference / forward-only, whatever the grad mode.
        self.autograd = False
        self._plans = {}
        self._weights_epoch = 0            # bumped by whoever rewrites parameters behind torch's version counters (the optimizer kernel)

    # ------------------------------------------------------------------ plan management
    def _weights_signature(self):
        return (self._weights_epoch,) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def mark_weights_changed(self):
        """Parameters were updated in place by a kernel (optim.AdamWEMA): the next forward / p_sample re-packs them."""
        self._weights_epoch += 1

    def plan_for(self, resolution, device):
        """C-side plan (block list, packed bf16 weights, workspace) for one image resolution."""
        if device.type != "cuda":
            raise RuntimeError("v_diffusion_b200.UNet runs on CUDA (sm_100a) only; there is no CPU fallback")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        key = (int(resolution), device.index, int(self.max_rows), self.operand_dtype)
        sig = self._weights_signature()
        ent = self._plans.get(key)
        L = _lib.lib()
        if ent is None:
            cfg = _lib.UNetConfig()
            cfg.in_channels, cfg.hid_channels, cfg.out_channels = self.in_channels, self.hid_channels, self.out_channels
            cfg.num_levels = self.levels
            for i, m in enumerate(self.ch_multipliers):
                cfg.ch_multipliers[i] = int(m)
                cfg.apply_attn[i] = int(bool(self.apply_attn[i]))
            cfg.num_res_blocks = self.num_res_blocks
            cfg.embedding_dim = self.embedding_dim
            cfg.head_dim = self.head_dim or 0
            cfg.num_heads = self.num_heads or 0
            cfg.num_classes, cfg.multitags = self.num_classes, int(self.multitags)
            cfg.resolution, cfg.max_rows = int(resolution), int(self.max_rows)
            cfg.operand_dtype = _lib.OPERAND_DTYPES[self.operand_dtype]
            handle = C.c_void_p()
            with torch.cuda.device(device):
                _lib.check(L.vdt_plan_create(C.byref(cfg), C.byref(handle)))
            ent = {"handle": handle, "sig": None}
            while len(self._plans) >= 4:                       # a plan owns packed weights + workspace: keep a few
                old_key = next(iter(self._plans))
                L.vdt_plan_destroy(self._plans.pop(old_key)["handle"])
            self._plans[key] = ent
        if ent["sig"] != sig:
            with torch.cuda.device(device):
                torch.cuda.current_stream().synchronize()
                for name, p in self.named_parameters():
                    if p.device != device or p.dtype != torch.float32:
                        raise RuntimeError(f"parameter {name} must be fp32 on {device} (got {p.dtype} on {p.device})")
                    t = p.detach().contiguous()
                    _lib.check(L.vdt_plan_load_weight(ent["handle"], name.encode(), _lib.ptr(t), t.numel(), 1))
                _lib.check(L.vdt_plan_finalize(ent["handle"]))
            ent["sig"] = sig
        return ent["handle"]

    def __del__(self):
        try:
            L = _lib.lib()
            for ent in self._plans.values():
                L.vdt_plan_destroy(ent["handle"])
        except Exception:
            pass

    # ------------------------------------------------------------------ forward (unet.py:286-322)
    def forward(self, x, t, y=None):
        if self.autograd and torch.is_grad_enabled():
            from .training import unet_autograd_forward
            return unet_autograd_forward(self, x, t, y)
        return self._forward_plan(x, t, y)

    @torch.no_grad()
    def _forward_plan(self, x, t, y=None):
        # .train(): forward with dropout active (unet.py:135, 146); the masks come from the library's Philox stream, seeded
        # per call from torch's global CPU generator (nn.Dropout draws from the global generator too).  Forward only: the
        # differentiable path is training.UNetTrainGraph (self.autograd = True routes forward() to it under grad mode).
        train = bool(self.training and self.drop_rate > 0.)
        if x.ndim != 4 or x.shape[1] != self.in_channels or x.shape[2] != x.shape[3]:
            raise ValueError(f"expected x of shape (B, {self.in_channels}, R, R), got {tuple(x.shape)}")
        dev = x.device
        plan = self.plan_for(x.shape[2], dev)
        B = x.shape[0]
        x = x.to(torch.float32).contiguous()
        t = t.reshape(-1).to(device=dev, dtype=torch.float64).contiguous()
        if t.numel() != B:
            raise ValueError("t must have one entry per batch row")
        if self.num_classes and y is not None:
            if self.multitags:
                assert y.ndim == 2                                     # unet.py:291
                y = y.to(device=dev, dtype=torch.float32).contiguous()
                if tuple(y.shape) != (B, self.num_classes):
                    raise ValueError(f"multitag y must have shape ({B}, {self.num_classes})")
            else:
                y = y.to(device=dev, dtype=torch.int64).contiguous()  # OneHot casts to long (modules.py:191-192)
                if y.numel() != B:
                    raise ValueError("y must have one entry per batch row")
                lo, hi = int(y.min()), int(y.max())                   # F.one_hot raises on these too (modules.py:193-196)
                if lo < 0 or hi > self.num_classes:
                    raise RuntimeError(f"class ids must lie in [0, {self.num_classes}] (0 = unconditional), got [{lo}, {hi}]")
        else:
            y = None
        out = torch.empty((B, self.out_channels, x.shape[2], x.shape[3]), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            if train:
                seed = int(torch.randint(0, 2 ** 63 - 1, (1,), dtype=torch.int64).item())
                _lib.check(_lib.lib().vdt_unet_forward_train(plan, _lib.ptr(x), _lib.ptr(t), _lib.ptr(y), _lib.ptr(out), B,
                                                             float(self.drop_rate), seed, _lib.current_stream_ptr()))
            else:
                _lib.check(_lib.lib().vdt_unet_forward(plan, _lib.ptr(x), _lib.ptr(t), _lib.ptr(y), _lib.ptr(out), B,
                                                       _lib.current_stream_ptr()))
        return out
