"""torchrun child of tests/test_train_step_gpu.py::test_two_rank_training_step_keeps_replicas_identical: one rank per GPU,
NCCL; each rank trains on its own data for three steps; the parameters must stay bit-identical across ranks (the gradient
average is the step's only collective) and must have moved."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests.train_step_worker import CASES, build                          # noqa: E402
from v_diffusion_b200 import GaussianDiffusion, get_logsnr_schedule      # noqa: E402
from v_diffusion_b200.training import TrainingStep                       # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    cfg, R = CASES["small"]["cfg"], CASES["small"]["res"]
    sd, net = build(cfg, 31, "fp16")
    net.train()
    diff = GaussianDiffusion(get_logsnr_schedule("cosine", -20., 20.), 100, "v", "fixed_medium", "snr_trunc", "mse", intp_frac=0.3)
    ts = TrainingStep(net, diff, timesteps=0, lr=2e-4, weight_decay=0.001, distributed=True, rank=rank, world_size=world)
    g = torch.Generator().manual_seed(500 + rank)
    finite = True
    for _ in range(3):
        x = torch.randn(4, cfg["in_channels"], R, R, generator=g).clamp(-1, 1).cuda()
        y = torch.randint(1, cfg["num_classes"] + 1, (4,), generator=g).cuda()
        loss = ts.step(x, y)
        finite = finite and bool(torch.isfinite(loss))
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    init = torch.cat([sd[k].reshape(-1) for k, _ in net.named_parameters()]).cuda()
    both = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    if rank == 0:
        print("RESULT " + json.dumps(dict(identical=all(bool(torch.equal(both[0], b)) for b in both[1:]),
                                          moved=float((flat - init).abs().max()), finite=finite)))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
