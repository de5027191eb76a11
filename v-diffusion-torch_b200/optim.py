"""Optimizer step of the reference trainer on the GPU without a host sync (train_utils.py:159-166):

    nn.utils.clip_grad_norm_(model.parameters(), max_norm=grad_norm)      # train_utils.py:161
    optimizer.step()            # torch.optim.AdamW(lr, betas, weight_decay), train.py:158
    ema.update()                # shadow += (1 - decay) * (param - shadow), utils.py:144-149

as two kernels per parameter tensor: a deterministic sum of squared gradients into one device scalar, then clip + decoupled
weight decay + Adam moments + bias-corrected update + EMA in one pass (vdt_grad_sq_accumulate, vdt_adamw_ema_step).
The learning-rate schedule (LambdaLR warm-up, train.py:161-162) stays with the caller: pass ``lr`` per step.
"""
import ctypes as C

import torch

from . import _lib


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

    def ema_state_dict(self):
        """The reference checkpoint's ``ckpt["ema"]`` entry (utils.py:168-173)."""
        return {"decay": self.decay, "shadow": self.shadow, "num_updates": self.num_updates}
