"""GPU: the composed training step (v-diffusion-torch_b200/training.py; SURVEY §8 f2, BASELINE configs[4]) on the real kernels.

The kernels it composes are each parity-tested on their own (tests/test_gpu_kernels.py) and the composition is checked against
autograd on the CPU with a contract stand-in (tests/test_training_graph.py); here the whole thing runs on the B200.  Each case
runs in a child process (tests/train_step_worker.py) so a device fault in this newest path cannot poison the tests after it.

Tolerances (the reference is fp32 autograd through the oracle UNet on the CPU; measured values in
profiles/r2k_train_step_parity.txt): fp16 operands -- network output rel-L2 <= 3e-3 (measured 7.8e-4 on cifar10_cond, 9.3e-4 on
the small net), every parameter gradient rel-L2 <= 8e-3 (measured worst 2.3e-3, median 1.4e-3 over 414 tensors); bf16 operands --
2.5e-2 / 5e-2 (measured 7.4e-3 / 1.6e-2).  Two cases have not run on hardware and are the only ones marked
xfail(strict=False): the two-rank NCCL step (the round's budget ended at one GPU) and the autograd-mode wrapper (added after it).
"""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
BARS = {"fp16": (3e-3, 8e-3), "bf16": (2.5e-2, 5e-2)}          # (output rel-L2, worst parameter-gradient rel-L2)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=900):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-m", "tests.train_step_worker", *map(str, args)], cwd=ROOT, env=env, capture_output=True,
                       text=True, timeout=timeout)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
    assert r.returncode == 0 and lines, (r.stdout[-1500:] + "\n" + r.stderr[-2500:])
    res = json.loads(lines[-1][len("RESULT "):])
    print(" ".join(map(str, args)), "->", json.dumps(res)[:1200])
    return res


@pytest.mark.parametrize("case,B,operand", [("small", 4, "fp16"), ("small", 3, "bf16"), ("cifar_cond", 4, "fp16")])
def test_unet_forward_backward_on_kernels_vs_autograd(case, B, operand):
    """Output and every parameter gradient of UNet forward + backward composed from the library's kernels against fp32
    autograd through the oracle UNet; `cifar_cond` is BASELINE configs[4]'s network (cifar10_cond.json) at batch 4."""
    res = _run("graph_parity", case, B, operand)
    assert res["finite"] and res["launches"] > 100
    assert res["out_rel"] <= BARS[operand][0], res
    assert res["grad_rel_worst"] <= BARS[operand][1], res["worst"]


def test_training_steps_follow_the_reference_recipe():
    """Three optimizer steps of TrainingStep.step (Trainer.loss + Trainer.step, train_utils.py:137-166) against the same steps
    with the oracle UNet under autograd + clip_grad_norm_ + torch.optim.AdamW + the reference EMA, on the same (t, noise) draws:
    loss within 1e-3 (measured 1.8e-4), total gradient norm within 2e-3 (2.3e-4), parameter updates aligned (cosine >= 0.999,
    measured 0.9999; AdamW's first steps are sign-like -- lr * g / |g| -- so entries whose gradient is below the fp16 rounding
    error may flip by 2 lr), EMA shadow within 3e-3 relative (7.3e-4: it inherits those flips; the optimizer / EMA kernel
    itself is pinned to 1.6e-7 in test_gpu_unet.py)."""
    res = _run("train_steps", "small", 8, "fp16")
    assert res["loss_rel_worst"] <= 1e-3, res
    assert res["gnorm_rel_worst"] <= 2e-3, res
    assert res["cosine_of_updates"] >= 0.999 and 0.999 <= res["update_norm_ratio"] <= 1.001, res
    assert res["ema_rel_worst"] <= 3e-3, res


def test_training_dropout_is_a_reproducible_stream():
    res = _run("dropout", "small", 4, "fp16")
    assert res["reproducible"] and res["finite"], res
    assert res["seed_changes_output"] > 1e-2 and res["eval_differs"] > 1e-2, res
    assert res["out_bias_grad_rel"] <= 1e-3, res


@pytest.mark.xfail(strict=False, reason="not yet run on hardware: added after the round's GPU budget was spent (the same tape ran "
                                        "green above; the autograd wrapper around it is checked on the CPU in test_training_graph.py)")
def test_autograd_mode_under_the_reference_train_loss():
    """UNet.autograd = True: the unmodified reference's train_loss + loss.mean().backward() drive this package's UNet on the
    GPU; loss within 1e-3 and every .grad within the fp16 gradient bar of the same lines run with the oracle UNet on the CPU."""
    res = _run("autograd_step", "small", 4, "fp16")
    if "skipped" in res:
        pytest.skip(res["skipped"])
    assert res["plan_path_finite"]
    assert res["loss_rel"] <= 1e-3, res
    assert res["grad_rel_worst"] <= BARS["fp16"][1], res["worst"]


@pytest.mark.xfail(strict=False, reason="not yet run on hardware: the round's GPU budget ended before a 2-GPU box could be used")
def test_two_rank_training_step_keeps_replicas_identical():
    """DDP semantics over NCCL: two ranks with different data end every step with bit-identical parameters."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29671", "-m", "tests.train_step_ddp_worker"], cwd=ROOT, env=env, capture_output=True,
                       text=True, timeout=900)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
    assert r.returncode == 0 and lines, (r.stdout[-1500:] + "\n" + r.stderr[-2500:])
    res = json.loads(lines[-1][len("RESULT "):])
    assert res["identical"] and res["moved"] > 0 and res["finite"], res
