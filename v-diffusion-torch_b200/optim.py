"""Optimizer step of the reference trainer on the GPU without a host sync (train_utils.py:159-166):

    nn.utils.clip_grad_norm_(model.parameters(), max_norm=grad_norm)      # train_utils.py:161
    optimizer.step()            # torch.optim.AdamW(lr, betas, weight_decay), train.py:158
    ema.update()                # shadow += (1 - decay) * (param - shadow), utils.py:144-149

as two kernels per parameter tensor: a deterministic sum of squared gradients into one device scalar, then clip + decoupled
weight decay + Adam moments + bias-corrected update + EMA in one pass (vdt_grad_sq_accumulate, vdt_adamw_ema_step).
The learning-rate schedule (LambdaLR warm-up, train.py:161-162) stays with the caller: pass ``lr`` per step.
"""
import torch

from . import _lib


def torch_adamw_state(names, exp_avg, exp_avg_sq, step, lr, betas, eps, weight_decay):
    """The moments in ``torch.optim.AdamW.state_dict()`` form (what the reference's checkpoints hold under "optimizer",
    train_utils.py:328-352): parameter i of ``model.parameters()`` <-> names[i]."""
    state = {}
    if step > 0:
        for i, k in enumerate(names):
            state[i] = {"step": torch.tensor(float(step)), "exp_avg": exp_avg[k], "exp_avg_sq": exp_avg_sq[k]}
    group = {"lr": lr, "betas": tuple(betas), "eps": eps, "weight_decay": weight_decay, "amsgrad": False, "maximize": False,
             "foreach": None, "capturable": False, "differentiable": False, "fused": None, "params": list(range(len(names)))}
    return {"state": state, "param_groups": [group]}


def read_torch_adamw_state(sd, names):
    """Inverse of torch_adamw_state for a checkpoint written by torch.optim.AdamW (one parameter group, every parameter in it,
    in ``model.parameters()`` order).  Returns (exp_avg, exp_avg_sq, step, param_group); (None, None, 0, group) for a fresh one."""
    groups = sd["param_groups"]
    if len(groups) != 1 or list(groups[0]["params"]) != list(range(len(names))):
        raise RuntimeError("expected the reference's single AdamW parameter group over model.parameters() (train.py:158)")
    if not sd["state"]:
        return None, None, 0, groups[0]
    steps = {int(float(sd["state"][i]["step"])) for i in range(len(names))}
    if len(steps) != 1:
        raise RuntimeError(f"parameters disagree on the step count: {sorted(steps)}")
    exp_avg = {k: sd["state"][i]["exp_avg"] for i, k in enumerate(names)}
    exp_avg_sq = {k: sd["state"][i]["exp_avg_sq"] for i, k in enumerate(names)}
    return exp_avg, exp_avg_sq, steps.pop(), groups[0]


def strip_module_prefix(d):
    """DDP checkpoints carry a "module." prefix on every key (generate.py:40-42, train_utils.py:320-324)."""
    return {(k.split(".", 1)[1] if k.startswith("module.") else k): v for k, v in d.items()}


class AdamWEMA:
    def __init__(self, named_params, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_norm=1.0, ema_decay=0.9999,
                 use_ema=True):
        self.params = dict(named_params)
        for k, p in self.params.items():
            if p.device.type != "cuda" or p.dtype != torch.float32 or not p.is_contiguous():
                raise ValueError(f"{k}: parameters must be contiguous fp32 CUDA tensors (there is no CPU fallback)")
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.grad_norm, self.decay = grad_norm, ema_decay
        self.exp_avg = {k: torch.zeros_like(p) for k, p in self.params.items()}
        self.exp_avg_sq = {k: torch.zeros_like(p) for k, p in self.params.items()}
        # the reference keeps the EMA on the leader rank only (train_utils.py:127-130): use_ema=False skips the shadow
        self.shadow = {k: p.detach().clone() for k, p in self.params.items()} if use_ema else {}    # utils.py:131-138
        self.step_count = 0
        self.num_updates = 0
        dev = next(iter(self.params.values())).device
        self._sq = torch.zeros(1, dtype=torch.float64, device=dev)
        nmax = max(p.numel() for p in self.params.values())
        self._scratch = torch.empty(int(_lib.lib().vdt_grad_sq_scratch_bytes(nmax)), dtype=torch.uint8, device=dev)

    @torch.no_grad()
    def step(self, grads, lr=None):
        """``grads``: dict name -> fp32 gradient tensor (same shapes as the parameters).  Returns the device scalar holding the
        squared total gradient norm before clipping (what clip_grad_norm_ returns, squared)."""
        L = _lib.lib()
        lr = self.lr if lr is None else lr
        self.step_count += 1
        self.num_updates += 1
        decay = min(self.decay, (1 + self.num_updates) / (10 + self.num_updates))          # utils.py:146
        st = _lib.current_stream_ptr()
        clip = self.grad_norm is not None and self.grad_norm > 0
        if clip:
            self._sq.zero_()
            for k, p in self.params.items():
                g = grads[k].contiguous()
                _lib.check(L.vdt_grad_sq_accumulate(_lib.ptr(g), g.numel(), _lib.ptr(self._scratch), self._scratch.numel(),
                                                    _lib.ptr(self._sq), st))
        for k, p in self.params.items():
            g = grads[k].contiguous()
            if g.shape != p.shape or g.dtype != torch.float32:
                raise ValueError(f"{k}: gradient must be fp32 with shape {tuple(p.shape)}")
            _lib.check(L.vdt_adamw_ema_step(_lib.ptr(p), _lib.ptr(g), _lib.ptr(self.exp_avg[k]), _lib.ptr(self.exp_avg_sq[k]),
                                            _lib.ptr(self.shadow.get(k)), p.numel(), lr, self.betas[0], self.betas[1], self.eps,
                                            self.weight_decay, self.step_count, _lib.ptr(self._sq) if clip else None,
                                            float(self.grad_norm or 0.0), decay, st))
        return self._sq

    # ---- checkpoint interop with the reference trainer (train_utils.py:309-352): torch.optim.AdamW's and EMA's own formats
    def state_dict(self):
        return torch_adamw_state(list(self.params), self.exp_avg, self.exp_avg_sq, self.step_count, self.lr, self.betas, self.eps,
                                 self.weight_decay)

    @torch.no_grad()
    def load_state_dict(self, sd):
        exp_avg, exp_avg_sq, step, group = read_torch_adamw_state(sd, list(self.params))
        self.lr, self.betas, self.eps, self.weight_decay = group["lr"], tuple(group["betas"]), group["eps"], group["weight_decay"]
        self.step_count = step
        for k in self.params:
            if exp_avg is None:
                self.exp_avg[k].zero_(); self.exp_avg_sq[k].zero_()
            else:
                self.exp_avg[k].copy_(exp_avg[k]); self.exp_avg_sq[k].copy_(exp_avg_sq[k])

    @torch.no_grad()
    def load_ema_state_dict(self, d):
        """``ckpt["ema"]`` of a reference checkpoint (utils.py:168-190); strict on the key set like EMA.load_state_dict."""
        shadow = strip_module_prefix(d["shadow"])
        if set(shadow) != set(self.params):
            raise RuntimeError(f"EMA key mismatch: {sorted(set(shadow) ^ set(self.params))[:4]} ...")
        if not self.shadow:
            self.shadow = {k: p.detach().clone() for k, p in self.params.items()}
        for k in self.params:
            self.shadow[k].copy_(shadow[k])
        self.decay, self.num_updates = d["decay"], int(d["num_updates"])

    def ema_state_dict(self):
        """The reference checkpoint's ``ckpt["ema"]`` entry (utils.py:168-173)."""
        return {"decay": self.decay, "shadow": self.shadow, "num_updates": self.num_updates}
