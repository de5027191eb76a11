"""Gradient goldens of the training step FROM THE UNMODIFIED REFERENCE (run in the build container: needs /root/reference):

    python tests/golden/make_train_grad_golden.py

The reference's own UNet (the oracle's deterministic weights strict-loaded, drop_rate 0) under the reference's own
GaussianDiffusion.train_loss and torch autograd, `loss.mean().backward()` as in Trainer.step (train_utils.py:149-151), on the
seeded inputs of tests/cases.py:TRAIN_GRAD_CASES (a small network with every block type, and cifar10_cond.json's own).  Stored: the per-sample loss and, for every parameter, the gradient's norm and
its projection on a seeded probe direction (cases.grad_probe); tensors of at most 4096 (cifar: 1024) entries are stored in full.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
_m = types.ModuleType("matplotlib"); _m.rcParams = {}
sys.modules.setdefault("matplotlib", _m)
sys.modules.setdefault("matplotlib.pyplot", types.ModuleType("matplotlib.pyplot"))
sys.path.insert(0, "/root/reference")
from v_diffusion import UNet, GaussianDiffusion, get_logsnr_schedule      # noqa: E402

from oracle.unet_ref import make_state_dict                              # noqa: E402
from tests.cases import TRAIN_GRAD_CASES, build_train_grad_inputs, grad_probe   # noqa: E402


def main():
    for name, case in TRAIN_GRAD_CASES.items():
        one(name, case)


def one(name, case):
    cfg = case["cfg"]
    net = UNet(cfg["in_channels"], cfg["hid_channels"], cfg["out_channels"], cfg["ch_multipliers"], cfg["num_res_blocks"],
               cfg["apply_attn"], embedding_dim=cfg["embedding_dim"], drop_rate=0., head_dim=cfg["head_dim"],
               num_heads=cfg["num_heads"], num_classes=cfg["num_classes"], multitags=cfg["multitags"])
    net.load_state_dict(make_state_dict(cfg, case["wseed"]), strict=True)
    net.train()
    diff = GaussianDiffusion(logsnr_fn=get_logsnr_schedule("cosine", -20., 20., rescale=False), sample_timesteps=1000,
                             model_out_type=case["model_out_type"], model_var_type="fixed_medium", reweight_type=case["reweight_type"],
                             loss_type="mse", intp_frac=0.3, p_uncond=0.1)
    x0, t, noise, y = build_train_grad_inputs(case)
    torch.manual_seed(case["seed"])
    loss = diff.train_loss(net, x_0=x0, t=t, y=y.clone(), noise=noise)
    loss.mean().backward()
    out = {"loss": loss.detach().numpy()}
    names = []
    for k, p in net.named_parameters():
        g = p.grad.double()
        names.append(k)
        out["norm/" + k] = np.float64(g.norm().item())
        out["proj/" + k] = np.float64((g * grad_probe(k, g.shape)).sum().item())
        if g.numel() <= (4096 if name == "small" else 1024):
            out["full/" + k] = p.grad.numpy()
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, f"train_grads_{name}.npz"), **out)
    print(f"train_grads_{name}:", len(names), "parameters, loss", loss.detach().numpy(),
          "total grad norm", float(torch.sqrt(sum(p.grad.double().pow(2).sum() for p in net.parameters()))))


if __name__ == "__main__":
    main()
