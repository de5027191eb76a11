#!/usr/bin/env python
"""Launches the dominant kernels at BASELINE configs[1] shapes (one 1024-row chunk) through the kernel-level
C-ABI hooks, for `ncu --set full` captures and isolated CUDA-event timings.

    python scripts/prof_kernels.py            # prints event timings (TFLOP/s, GB/s)
    ncu --set full --clock-control none --import-source on -k regex:'conv_gemm|groupnorm_kernel|attention' \
        -o gpurun_out/r1_kernels python scripts/prof_kernels.py --once
"""
import ctypes as C
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from v_diffusion_b200 import _lib  # noqa: E402

L = _lib.lib()
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
once = "--once" in sys.argv
R = 1024          # UNet rows in one chunk
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    if once:
        return 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def conv_case(H, cin, cout, k, resid, stats, out16=False):
    x = torch.randn(R, H, H, cin, device=dev, generator=g).half()
    w = torch.randn(cout, cin, k, k, device=dev, generator=g) / math.sqrt(cin * k * k)
    b = torch.randn(cout, device=dev, generator=g)
    res = torch.randn(R, H, H, cout, device=dev, generator=g) if resid else None
    out = torch.empty(R, H, H, cout, device=dev)
    o16 = torch.empty(R, H, H, cout, device=dev, dtype=torch.float16) if out16 else None
    st = torch.empty(R * H * H // 32, cout // 4, 2, device=dev) if stats else None
    ms = timed(lambda: L.vdt_op_conv(p(x), R, H, H, cin, p(w), cout, k, p(b), p(res), p(out), 1, p(o16), p(st), 4, None), 5)
    fl = 2.0 * R * H * H * cout * cin * k * k
    print(f"conv {k}x{k} {cin}->{cout} @{H}x{H} rows={R} resid={resid} stats={stats} out16={out16}: {ms:.3f} ms "
          f"{fl / ms / 1e9 if ms else 0:.0f} TFLOP/s (includes weight pack + sync of the hook)")
    return st, out, o16


def gn_case(H, c, st, src, in16, film):
    gamma = torch.ones(c, device=dev); beta = torch.zeros(c, device=dev)
    ftab = torch.randn(R, 2 * c, device=dev, generator=g) * 0.1 if film else None
    oa = torch.empty(R, H, H, c, device=dev, dtype=torch.float16)
    ms = timed(lambda: L.vdt_op_groupnorm(p(src), c, None, 0, R, H, H, p(gamma), p(beta), p(ftab), 2 * c, 0, 1, 0, p(oa),
                                          None, None, 1, p(st), None, 4, int(in16), None), 5)
    byts = R * H * H * c * ((2 if in16 else 4) + 2)
    print(f"groupnorm fused C={c} @{H}x{H} in16={in16}: {ms:.3f} ms {byts / ms / 1e6 if ms else 0:.0f} GB/s (hook allocs scratch)")


def attn_case(N, d):
    qkv = torch.randn(R * N, 3 * d, device=dev, generator=g).half()
    o = torch.empty(R * N, d, device=dev, dtype=torch.float16)
    ms = timed(lambda: L.vdt_op_attention(p(qkv), p(o), R, N, 1, d, 1, None), 5)
    fl = 4.0 * R * N * N * d
    print(f"attention N={N} d={d}: {ms:.3f} ms {fl / ms / 1e9 if ms else 0:.0f} TFLOP/s")


def sampler_case(B=4096, C=3, H=32):
    """the fused per-step sampler update at BASELINE configs[1] size: v (2B rows), x_t in, x_s out"""
    from v_diffusion_b200 import GaussianDiffusion, get_logsnr_schedule
    diff = GaussianDiffusion(get_logsnr_schedule("cosine"), 100, "v", "fixed_medium", "snr_trunc", "mse", intp_frac=0.3, w_guide=1.0)
    coefs = diff.step_coefficients(use_ddim=True)
    mo = torch.randn(2 * B, C, H, H, device=dev, generator=g)
    x = torch.randn(B, C, H, H, device=dev, generator=g)
    out = torch.empty_like(x)
    row = coefs[50].contiguous()
    ms = timed(lambda: L.vdt_op_sampler_step(p(mo), p(x), None, p(out), B, C, H * H, 1, 3, 50, p(row), 1.0, None), 5)
    byts = 4.0 * B * C * H * H * 4
    print(f"sampler_step B={B} CFG: {ms:.4f} ms {byts / ms / 1e6 if ms else 0:.0f} GB/s (the hook allocates / frees its state and syncs)")


def backward_case(B=128, H=32, cin=256, cout=256, k=3):
    """conv backward at the BASELINE configs[4] batch (128 per GPU): dgrad on the forward kernel, wgrad kernel"""
    x = torch.randn(B, H, H, cin, device=dev, generator=g).half()
    dy = torch.randn(B, H, H, cout, device=dev, generator=g).half()
    w = torch.randn(cout, cin, k, k, device=dev, generator=g) / math.sqrt(cin * k * k)
    dx = torch.empty(B, H, H, cin, device=dev)
    dw = torch.empty(cout, cin, k, k, device=dev)
    for _ in range(1 if once else 3):
        L.vdt_op_conv_dgrad(p(dy), B, H, H, cin, p(w), cout, k, p(dx), 1, None)
        L.vdt_op_conv_wgrad(p(x), p(dy), B, H, H, cin, cout, k, p(dw), None, 1, None)
    torch.cuda.synchronize()
    print(f"conv backward {cin}->{cout} k{k} @{H} B={B}: {2.0 * B * H * H * cout * cin * k * k / 1e12:.3f} TFLOP each for dgrad and wgrad")


def gn_backward_case(B=128, H=32, Cc=256):
    """GroupNorm -> FiLM -> SiLU -> dropout backward (gn_backward.cu) at the configs[4] batch"""
    import ctypes as C
    x = torch.randn(B, H, H, Cc, device=dev, generator=g)
    go = torch.randn(B, H, H, Cc, device=dev, generator=g)
    gamma, beta = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev)
    film = torch.randn(B, 2 * Cc, device=dev, generator=g) * 0.1
    gx, gg, gb, gf = torch.empty_like(x), torch.empty(Cc, device=dev), torch.empty(Cc, device=dev), torch.empty_like(film)
    for _ in range(1 if once else 3):
        L.vdt_op_groupnorm_backward(p(x), p(go), Cc, B, H, H, p(gamma), p(beta), p(film), 1, C.c_float(0.2), 7, 3, p(gx), p(gg), p(gb), p(gf), None)
    torch.cuda.synchronize()
    # (the op allocates its scratch per call: kernel times come from the ncu launch list, scripts/gpu_profile_round.sh step 4)
    print(f"groupnorm backward C{Cc} @{H} B={B}: two streaming passes, {x.numel() * 4 * 5 / 1e6:.0f} MB algorithmic (x and dA read twice, dx written)")


if "--backward" in sys.argv:
    gn_backward_case()
    backward_case()
    backward_case(128, 16, 256, 256, 3)
    backward_case(128, 32, 512, 256, 3)
    sys.exit(0)
if "--sampler" in sys.argv:
    sampler_case()
    torch.cuda.synchronize()
    sys.exit(0)
if "--small" in sys.argv:
    conv_case(32, 64, 256, 1, False, True)                 # in_conv-like: K = 64, fp32 out + stats
    conv_case(16, 256, 512, 1, False, False, out16=True)   # q|k projection: K = 256, 16-bit out
    conv_case(16, 256, 256, 1, True, True)                 # proj_out: K = 256, fp32 out + residual
    torch.cuda.synchronize()
    sys.exit(0)
st, out, _ = conv_case(32, 256, 256, 3, True, True)
gn_case(32, 256, st, out, False, False)
st16, _, o16 = conv_case(32, 256, 256, 3, False, True, out16=True)
gn_case(32, 256, st16, o16, True, True)
attn_case(1024, 256)
if not once:
    conv_case(32, 512, 256, 3, False, True, out16=True)
    conv_case(16, 256, 256, 3, True, True)
    conv_case(8, 256, 256, 3, True, True)
    conv_case(32, 256, 768, 1, False, False, out16=True)
    conv_case(16, 256, 256, 1, True, True)
    attn_case(256, 256)
    attn_case(64, 256)
torch.cuda.synchronize()
