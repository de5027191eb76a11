"""One warm-up + one timed training step at batch 128 on cuda:0 (cifar10_cond network, drop_rate 0.2), printing as it goes."""
import sys, os, time
t00 = time.perf_counter()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from v_diffusion_b200.training import TrainingStep
print(f"imports {time.perf_counter() - t00:.2f}s", flush=True)
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
net, diff = bench.build_model(dev, seed=0); net.train()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g = torch.Generator().manual_seed(4321)
x = torch.randn(B, 3, 32, 32, generator=g).clamp(-1, 1).to(dev); y = (torch.randint(10, (B,), generator=g) + 1).to(dev)
ts = TrainingStep(net, diff, timesteps=0, lr=2e-4, weight_decay=0.001, grad_norm=1.0, use_ema=True)
torch.cuda.synchronize(); print(f"model ready {time.perf_counter() - t00:.2f}s", flush=True)
for i in range(3):
    t0 = time.perf_counter(); loss = ts.step(x, y.clone()); torch.cuda.synchronize()
    print(f"step {i}: {time.perf_counter() - t0:.3f}s loss {float(loss):.4f} gnorm {float(ts.last_grad_sq) ** 0.5:.3f} "
          f"mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB  t={time.perf_counter() - t00:.1f}s", flush=True)
