"""Oracle (test infrastructure): schedule, posterior coefficients and the sampling loop.

Restates ``/root/reference/v_diffusion/diffusion.py`` (get_logsnr_schedule 42-112,
stable_log1mexp 115-123, logsnr_to_posterior 126-163, logsnr_to_posterior_ddim
169-203, pred_x0_from_* 206-234, p_mean_var 317-356, p_sample_step 360-392,
p_sample 394-414) and ``functions.py`` get_timestep_embedding (11-29).

Scalars follow the reference's rounding chain exactly: log-SNR in
fp64 -> rounded to fp32 by ``broadcast_to`` -> re-upcast to fp64 inside the posterior
-> coefficients rounded to fp32; alpha/sigma evaluated in fp32 on the fp32 log-SNR.
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import numpy as np
import torch


# ----------------------------------------------------------------------------- schedule
def logsnr_schedule(t, schedule: str = "cosine", logsnr_min: float = -20., logsnr_max: float = 20.):
    """fp64 log-SNR at time t in [0,1] (diffusion.py:42-112; rescale is off in every config)."""
    t = np.asarray(t, dtype=np.float64)
    if schedule == "cosine":
        t_from = np.arctan(np.exp(-0.5 * logsnr_max)) / (0.5 * np.pi)
        t_to = np.arctan(np.exp(-0.5 * logsnr_min)) / (0.5 * np.pi)
        tt = t_from + t * (t_to - t_from)            # torch.lerp(start, end, w) = start + w (end-start)
        return -2.0 * np.log(np.tan(tt * np.pi * 0.5))
    if schedule == "linear":
        sig = lambda z: 1.0 / (1.0 + np.exp(-z))
        t_from, t_to = sig(logsnr_max), sig(logsnr_min)
        tt = t_from + t * (t_to - t_from)
        return np.log(tt) - np.log1p(-tt)
    if schedule == "sigmoid":
        rng = logsnr_max - logsnr_min
        t_from, t_to = 0.0, 1.0
        tt = t_from + t * (t_to - t_from)
        return logsnr_max - tt * rng
    if schedule == "legacy":
        x_from = x_max = 0.9999
        x_min, slope = 0.98, -0.0199
        x_to = x_max + t * (x_min - x_max)
        log_alpha = 1000 / slope * (x_to * np.log(x_to) - x_to - x_from * np.log(x_from) + x_from)
        return log_alpha - _log1mexp(log_alpha - 1e-9)
    raise NotImplementedError(schedule)


def _logsigmoid(x):
    x = np.asarray(x, dtype=np.float64)
    return -np.logaddexp(0.0, -x)


def _log1mexp(x):
    x = np.asarray(x, dtype=np.float64)
    assert np.all(x < 0)                             # diffusion.py:119
    with np.errstate(divide="ignore"):
        return np.where(x < -9, np.log1p(-np.exp(x)), np.log(-np.expm1(x)))


def step_coefficients(T: int, use_ddim: bool, var_type: str = "fixed_large", intp_frac=None,
                      schedule: str = "cosine", logsnr_min: float = -20., logsnr_max: float = 20.,
                      x0eps_coef: bool = False, t_fp32: bool = False):
    """Per-step fp32 scalars for step index i = 0..T-1 (the loop visits T-1 .. 0).

    ``t_fp32``: the step tensor is fp32 as in p_sample_progressive (diffusion.py:421, ``t = torch.empty(B)``): s and t
    are fp32 quotients (the schedule upcasts them, diffusion.py:101) and ``t_model`` is an fp32 array, so the network's
    sinusoidal embedding is evaluated in fp32 (functions.py:20-25).  p_sample itself carries fp64 (diffusion.py:399).


    Returns dict of float32 arrays of length T: ``logsnr_s, logsnr_t, alpha_t, sigma_t
    (fp32 math on the fp32 log-SNR: diffusion.py:233-234), c1, c2, logvar, std``
    (std = exp(0.5 logvar) in fp32, 0 for DDIM) and fp64 ``t_model`` = (i+1)/T, the time
    fed to the network (diffusion.py:364, 374).

    ``x0eps_coef`` (diffusion.py:137-140, 180-182): the posterior mean is written as c1 * eps + c2 * x0.  The
    reference's DDIM branch returns the two coefficients as LOGARITHMS (the ``.exp_()`` at diffusion.py:199 is
    only reached for eta != 0); that is restated as is."""
    if t_fp32:
        i32 = np.arange(T, dtype=np.float32)
        s, t = i32 / np.float32(T), (i32 + np.float32(1)) / np.float32(T)      # fp32 divisions (diffusion.py:363-364)
        t_model = t.copy()
        s, t = s.astype(np.float64), t.astype(np.float64)
    else:
        i = np.arange(T, dtype=np.float64)
        s, t = i / T, (i + 1) / T
        t_model = t
    ls32 = logsnr_schedule(s, schedule, logsnr_min, logsnr_max).astype(np.float32)
    lt32 = logsnr_schedule(t, schedule, logsnr_min, logsnr_max).astype(np.float32)
    ls, lt = ls32.astype(np.float64), lt32.astype(np.float64)
    logr = lt - ls
    if use_ddim:                                     # eta = 0 branch, diffusion.py:178-187
        if x0eps_coef:                               # diffusion.py:180-182 (not exponentiated, see docstring)
            c1 = 0.5 * _logsigmoid(-ls)
            c2 = 0.5 * _logsigmoid(ls)
        else:
            c1 = np.exp(0.5 * (_logsigmoid(-ls) - _logsigmoid(-lt)))
            c2 = np.exp(_log1mexp(0.5 * logr) + 0.5 * _logsigmoid(ls))
        logvar = np.full(T, -np.inf)
    else:                                            # diffusion.py:133-161
        log_alpha_st = 0.5 * (_logsigmoid(ls) - _logsigmoid(lt))
        l1mr = _log1mexp(logr)
        if x0eps_coef:                               # diffusion.py:137-140
            c1 = np.exp(0.5 * (_logsigmoid(ls) - lt) + logr)
            c2 = np.sqrt(1.0 / (1.0 + np.exp(-ls)))
        else:
            c1 = np.exp(logr + log_alpha_st)
            c2 = np.exp(l1mr + 0.5 * _logsigmoid(ls))
        if var_type == "fixed_large":
            logvar = l1mr + _logsigmoid(-lt)
        elif var_type == "fixed_small":
            logvar = l1mr + _logsigmoid(-ls)
        elif var_type == "fixed_medium":
            lo, hi = l1mr + _logsigmoid(-ls), l1mr + _logsigmoid(-lt)
            logvar = lo + intp_frac * (hi - lo)
        else:
            raise NotImplementedError(var_type)
    lt_t = torch.from_numpy(lt32)
    alpha = torch.sigmoid(lt_t).sqrt().numpy()
    sigma = torch.sigmoid(-lt_t).sqrt().numpy()
    logvar32 = logvar.astype(np.float32)
    std = torch.exp(0.5 * torch.from_numpy(logvar32)).numpy()
    return dict(logsnr_s=ls32, logsnr_t=lt32, alpha_t=alpha, sigma_t=sigma,
                c1=c1.astype(np.float32), c2=c2.astype(np.float32), logvar=logvar32, std=std,
                t_model=t_model)


# ----------------------------------------------------------------------------- embedding
def timestep_embedding(t: torch.Tensor, dim: int, scale: float = 1000.) -> torch.Tensor:
    """[sin | cos] of scale*t*exp(-k ln(1e4)/(half-1)), evaluated in t's dtype, cast to
    fp32 (functions.py:11-29)."""
    t = t.reshape(-1)
    half = dim // 2
    f = torch.exp(-torch.arange(half, dtype=t.dtype) * (math.log(10000) / (half - 1)))
    ang = torch.outer(scale * t, f)
    e = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1).to(torch.float32)
    if dim % 2 == 1:
        e = torch.nn.functional.pad(e, [0, 1])
    return e


# ----------------------------------------------------------------------------- sampler
def _pred_x0(x_t, out, lt, model_out_type):
    if model_out_type == "x0":
        return out
    if model_out_type == "v":                        # diffusion.py:233-234
        return x_t * torch.sigmoid(lt).sqrt() - out * torch.sigmoid(-lt).sqrt()
    if model_out_type == "eps":                      # diffusion.py:207-208
        return x_t * torch.sigmoid(lt).rsqrt() - out * (-lt * 0.5).exp()
    if model_out_type == "both":                     # diffusion.py:211-214
        x0, eps = out.chunk(2, dim=1)
        x0e = x_t * torch.sigmoid(lt).rsqrt() - eps * (-lt * 0.5).exp()
        return x0 * torch.sigmoid(-lt) + x0e * torch.sigmoid(lt)
    raise NotImplementedError(model_out_type)


@torch.no_grad()
def p_sample(denoise_fn: Callable, shape, noise: torch.Tensor, label: Optional[torch.Tensor], *,
             T: int, model_out_type: str, w_guide: float = 0., use_ddim: bool = True,
             var_type: str = "fixed_large", intp_frac=None, step_noise: Optional[torch.Tensor] = None,
             schedule: str = "cosine", logsnr_min: float = -20., logsnr_max: float = 20.,
             record: Optional[list] = None, pred_record: Optional[list] = None,
             x0eps_coef: bool = False, t_fp32: bool = False) -> torch.Tensor:
    """Reverse loop (diffusion.py:394-414 + 360-392; with ``t_fp32`` and ``pred_record`` the loop of
    p_sample_progressive, 416-441).  ``noise``: initial x_T.
    ``step_noise``: (T, B, C, H, W) pre-drawn per-step normal draws, indexed by the step
    index ti (required for ancestral sampling so both sides inject identical noise);
    DDIM multiplies them by exp(-inf)=0.  ``record`` collects (ti, model_out)."""
    B = shape[0]
    co = step_coefficients(T, use_ddim, var_type, intp_frac, schedule, logsnr_min, logsnr_max, x0eps_coef, t_fp32)
    x_t = noise.clone().float()
    use_cfg = (w_guide > 0) and (label is not None)
    for ti in reversed(range(T)):
        lt = torch.tensor(co["logsnr_t"][ti])
        c1, c2 = float(co["c1"][ti]), float(co["c2"][ti])
        t = torch.full((B,), float(co["t_model"][ti]), dtype=torch.float32 if t_fp32 else torch.float64)
        if use_cfg:                                   # rows 2i cond, 2i+1 uncond (diffusion.py:368-372)
            xin = x_t.repeat_interleave(2, dim=0)
            tin = t.repeat_interleave(2)
            yin = label.repeat_interleave(2, dim=0).clone()
            yin[1::2] = 0
        else:
            xin, tin, yin = x_t, t, label
        out = denoise_fn(xin, tin, yin)
        if record is not None:
            record.append((ti, out.clone()))
        x0 = _pred_x0(xin, out, lt, model_out_type).clamp(-1., 1.)
        xt_or_eps = xin
        if x0eps_coef:                                # eps re-derived from the clipped x0 (diffusion.py:335-343, 222-223)
            xt_or_eps = xin * torch.sigmoid(-lt).rsqrt() - x0 * (lt * 0.5).exp()
        mean = torch.tensor(c1) * xt_or_eps + torch.tensor(c2) * x0
        if ti == 0:                                   # where(cond, mean, pred_x_0)  diffusion.py:378
            mean = x0
        pred = x0
        if use_cfg:
            mc, mu = mean[0::2], mean[1::2]
            mean = mc + w_guide * (mc - mu)           # not re-clipped
            pred = x0[0::2] + w_guide * (x0[0::2] - x0[1::2])       # diffusion.py:385
        if pred_record is not None:
            pred_record.append((ti, pred.clone()))
        if ti > 0 and not use_ddim:
            assert step_noise is not None, "ancestral sampling needs injected step noise"
            mean = mean + torch.tensor(co["std"][ti]) * step_noise[ti]
        x_t = mean
    return x_t


# ----------------------------------------------------------------------------- training loss
def train_loss(denoise_fn: Callable, x_0: torch.Tensor, t: torch.Tensor, y, noise: torch.Tensor, *, model_out_type: str,
               reweight_type: str = "snr_trunc", schedule: str = "cosine", logsnr_min: float = -20., logsnr_max: float = 20.):
    """GaussianDiffusion.train_loss for loss_type "mse" (diffusion.py:492-545) and the converters it uses
    (from_model_out_to_pred 466-490, q_sample 242-245, pred_*_from_* 206-239).  Returns (per-sample loss, x_t).
    The p_uncond label mask acts on ``y`` after the model call (diffusion.py:527-529) and is not restated: it cannot
    change the loss."""
    lt = torch.from_numpy(logsnr_schedule(t.double().numpy(), schedule, logsnr_min, logsnr_max).astype(np.float32))
    lt = lt.reshape(-1, 1, 1, 1)                                      # broadcast_to: cast to x's dtype (diffusion.py:23-26)
    x_t = x_0 * torch.sigmoid(lt).sqrt() + noise * torch.sigmoid(-lt).sqrt()
    out = denoise_fn(x_t, t, y)
    if model_out_type == "v":
        x0p = x_t * torch.sigmoid(lt).sqrt() - out * torch.sigmoid(-lt).sqrt()
        epsp = x_t * torch.sigmoid(-lt).sqrt() + out * torch.sigmoid(lt).sqrt()
        vp = out
    else:
        if model_out_type == "x0":
            x0p = out
            epsp = x_t * torch.sigmoid(-lt).rsqrt() - x0p * (lt * 0.5).exp()
        elif model_out_type == "eps":
            epsp = out
            x0p = x_t * torch.sigmoid(lt).rsqrt() - epsp * (-lt * 0.5).exp()
        elif model_out_type == "both":
            x0p = _pred_x0(x_t, out, lt, "both")
            epsp = x_t * torch.sigmoid(-lt).rsqrt() - x0p * (lt * 0.5).exp()
        else:
            raise NotImplementedError(model_out_type)
        vp = -x0p * torch.sigmoid(-lt).sqrt() + epsp * torch.sigmoid(lt).sqrt()
    fm = lambda z: z.flatten(1).mean(dim=1)
    if reweight_type == "snr_trunc":
        return torch.maximum(fm((x_0 - x0p) ** 2), fm((noise - epsp) ** 2)), x_t
    # single-target reweightings compare with the RAW model output (diffusion.py:541), reproduced as is
    target = {"constant": x_0, "snr": noise,
              "snr_1plus": -x_0 * torch.sigmoid(-lt).sqrt() + noise * torch.sigmoid(lt).sqrt()}[reweight_type]
    return fm((target - out) ** 2), x_t
