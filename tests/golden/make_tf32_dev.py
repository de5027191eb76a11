"""How far does the UNMODIFIED REFERENCE move under its own default GPU numerics?

PyTorch's defaults on Ampere-or-newer GPUs run every ``F.conv2d`` in TF32 (``torch.backends.cudnn.allow_tf32``
is True by default: both conv operands are rounded to a 10-bit mantissa, fp32 accumulation) and keep matmuls in
fp32.  This script replays the sampler fixtures through the reference on CPU twice -- plain fp32 (that is the
committed golden) and with ``F.conv2d`` wrapped so that its input and weight are rounded to TF32 first -- and
stores the max-abs distance of the two final samples per fixture in ``ref_tf32_deviation.json``.  The numbers
calibrate the sample tolerance: a 10-bit-mantissa operand format (fp16 here, TF32 there) cannot track the fp32
trajectory of a guidance-amplified sampler any closer than the reference's own GPU path does.

Run in the build container only (needs /root/reference):   python tests/golden/make_tf32_dev.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (imports the reference, stubs matplotlib)
from tests.cases import UNET_CASES, SAMPLE_CASES, build_sample_inputs  # noqa: E402


def round_tf32(x):
    """fp32 -> nearest TF32 (10 explicit mantissa bits, ties away from zero like the tensor-core conversion)."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


_conv2d = F.conv2d


def conv2d_tf32(input, weight, bias=None, *a, **k):
    return _conv2d(round_tf32(input), round_tf32(weight), bias, *a, **k)


def run(case, tf32):
    cfg = UNET_CASES[case["unet"]]["cfg"]
    net = mg.ref_unet(cfg, UNET_CASES[case["unet"]]["seed"])
    noise, label, _ = build_sample_inputs(case, cfg)
    diff = mg.GaussianDiffusion(
        logsnr_fn=mg.get_logsnr_schedule("cosine", -20., 20., rescale=False), sample_timesteps=case["T"],
        model_out_type=case["model_out_type"], model_var_type=case["var_type"], reweight_type="snr_trunc",
        loss_type="mse", intp_frac=case.get("intp_frac"), w_guide=case["w_guide"], x0eps_coef=case.get("x0eps_coef", False))
    F.conv2d = conv2d_tf32 if tf32 else _conv2d
    torch.conv2d, saved = (conv2d_tf32 if tf32 else torch.conv2d), torch.conv2d
    try:
        with torch.no_grad():
            return diff.p_sample(net, shape=tuple(noise.shape), noise=noise, label=label, device="cpu",
                                 seed=case["seed"], use_ddim=case["use_ddim"])
    finally:
        F.conv2d = _conv2d
        torch.conv2d = saved


def main():
    out = {}
    for name, case in SAMPLE_CASES.items():
        ref = run(case, False)
        gold = np.load(os.path.join(HERE, f"sample_{name}.npz"))["out"]
        assert np.array_equal(ref.numpy(), gold), f"{name}: fp32 replay does not reproduce the committed golden"
        tf = run(case, True)
        out[name] = {"max_abs": float((tf - ref).abs().max()), "rel_l2": float((tf - ref).norm() / ref.norm()),
                     "w_guide": case["w_guide"], "T": case["T"]}
        print(name, out[name])
    json.dump(out, open(os.path.join(HERE, "ref_tf32_deviation.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
