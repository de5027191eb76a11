"""Debug helper: position independence of a sample's trajectory (see tests/test_gpu_unet.py::test_full_size_chunk_properties)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.cases import CIFAR_COND
from tests.test_gpu_unet import _model
from v_diffusion_b200 import GaussianDiffusion, get_logsnr_schedule, _lib
net = _model(CIFAR_COND, 13)
diff = GaussianDiffusion(get_logsnr_schedule("cosine", -20., 20.), 100, "v", "fixed_medium", "snr_trunc", "mse", intp_frac=0.3, w_guide=1.0)
g = torch.Generator().manual_seed(77)
B = int(os.environ.get('DBG_B', '520'))
MR = int(os.environ.get('DBG_MR', '1024'))
noise = torch.randn(B, 3, 32, 32, generator=g)
label = torch.randint(0, 11, (B,), generator=g)
sc = diff.sampler_config(use_ddim=True)
def run(n_idx, max_rows, steps=2):
    net.max_rows = max_rows
    plan = net.plan_for(32, torch.device("cuda", 0))
    x = noise[n_idx].cuda().contiguous().clone(); y = label[n_idx].cuda().contiguous()
    _lib.check(_lib.lib().vdt_p_sample_range(plan, C.byref(sc), _lib.ptr(x), _lib.ptr(y), None, x.shape[0], 99, steps, None, None))
    torch.cuda.synchronize(); return x.cpu()
for steps in ((1,) if B < 520 else (1, 2)):
    full = run(torch.arange(B), MR, steps); again = run(torch.arange(B), MR, steps)
    print("steps", steps, "deterministic", torch.equal(full, again))
    pick = torch.tensor([0, 255, 511, 512, 519]) if B >= 520 else torch.tensor([0, 1, B // 2, B - 2, B - 1])
    alone = run(pick, MR, steps)
    print("  alone per-image diff", [(alone[i] - full[pick[i]]).abs().max().item() for i in range(5)])
    sm = run(torch.arange(B), max(2, MR // 8), steps)
    d = (sm - full).abs().flatten(1).max(1).values
    print("  small chunks: max", d.max().item(), "images differing", int((d > 0).sum()), "first", d.nonzero().flatten()[:10].tolist())
