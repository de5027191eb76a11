"""CPU study for the next round: would a Winograd F(2x2, 3x3) formulation of the 3x3 convolutions (2.25x fewer tensor-core
MACs) stay inside the parity bars with fp16 operands?  Compares, against an fp64 convolution, (a) the direct product
with both operands rounded to fp16 (what conv_gemm.cu does) and (b) Winograd with the transformed input tiles and the
transformed weights rounded to fp16 and fp32 accumulation.  Activations are SiLU(N(0,1)) like a GroupNorm+SiLU output.

    python scripts/winograd_error_study.py
"""
import torch
import torch.nn.functional as F

torch.manual_seed(0)
B, C, K, H = 2, 256, 256, 32
x = F.silu(torch.randn(B, C, H, H, dtype=torch.float64))
w = torch.randn(K, C, 3, 3, dtype=torch.float64) / (C * 9) ** 0.5
ref = F.conv2d(x, w, padding=1)

r16 = lambda t: t.to(torch.float16).to(torch.float64)
direct = F.conv2d(r16(x), r16(w), padding=1)

Bt = torch.tensor([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], dtype=torch.float64)
G = torch.tensor([[1, 0, 0], [.5, .5, .5], [.5, -.5, .5], [0, 0, 1]], dtype=torch.float64)
At = torch.tensor([[1, 1, 1, 0], [0, 1, -1, -1]], dtype=torch.float64)


def winograd(x, w, round_ops):
    xp = F.pad(x, (1, 1, 1, 1))
    tiles = xp.unfold(2, 4, 2).unfold(3, 4, 2)                  # B, C, H/2, W/2, 4, 4
    V = Bt @ tiles @ Bt.T
    U = G @ w @ G.T                                              # K, C, 4, 4
    if round_ops:
        V, U = r16(V), r16(U)
    M = torch.einsum("bcijuv,kcuv->bkijuv", V, U)                # 16 independent GEMMs over C
    Y = At @ M @ At.T                                            # B, K, H/2, W/2, 2, 2
    return Y.permute(0, 1, 2, 4, 3, 5).reshape(x.shape[0], w.shape[0], x.shape[2], x.shape[3])


wino_exact = winograd(x, w, False)
wino16 = winograd(x, w, True)
rel = lambda a: ((a - ref).norm() / ref.norm()).item()
print(f"winograd in fp64 (sanity)           rel-L2 {rel(wino_exact):.2e}")
print(f"direct, fp16 operands               rel-L2 {rel(direct):.2e}   max-abs {(direct - ref).abs().max().item():.2e}")
print(f"winograd F(2x2,3x3), fp16 operands  rel-L2 {rel(wino16):.2e}   max-abs {(wino16 - ref).abs().max().item():.2e}")
print(f"error ratio winograd / direct       {rel(wino16) / rel(direct):.2f}")
