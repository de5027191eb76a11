// Training-step kernels around the model call: GaussianDiffusion.train_loss (diffusion.py:492-545, loss_type "mse").
//
//   q_sample_kernel      x_t = x_0 * sqrt(sigmoid(l_t)) + eps * sqrt(sigmoid(-l_t))            (diffusion.py:242-245)
//   train_loss_kernel    from_model_out_to_pred (466-490) + the re-weighted MSE per sample (518-541), and optionally the
//                        gradient of loss.mean() with respect to the model output (what autograd hands the UNet's backward)
//
// Every per-sample scalar comes from a [B][16] coefficient table computed on the host in the reference's rounding chain
// (vdt_train_coefficients: log-SNR in fp64 -> fp32, sigmoid / sqrt / exp in fp32).  Products and sums are rounded
// separately (no FMA contraction), as the reference's chain of element-wise tensor ops rounds them.
#include "kernels.cuh"
#include "ptx.cuh"

namespace vdt {
namespace {

__global__ void __launch_bounds__(256) q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ eps,
                                                       const float* __restrict__ coef, float* __restrict__ x_t, long long n, int chw) {
    const long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 4;
    if (i >= n) return;
    const int b = static_cast<int>(i / chw);
    const float a = coef[b * kCoefStride + 0], s = coef[b * kCoefStride + 1];
    const float4 x = *reinterpret_cast<const float4*>(x0 + i), e = *reinterpret_cast<const float4*>(eps + i);
    float4 o;
    o.x = __fadd_rn(__fmul_rn(x.x, a), __fmul_rn(e.x, s)); o.y = __fadd_rn(__fmul_rn(x.y, a), __fmul_rn(e.y, s));
    o.z = __fadd_rn(__fmul_rn(x.z, a), __fmul_rn(e.z, s)); o.w = __fadd_rn(__fmul_rn(x.w, a), __fmul_rn(e.w, s));
    *reinterpret_cast<float4*>(x_t + i) = o;
}

// predictions of (x_0, eps) from the model output (diffusion.py:466-490).  o = output channel, o2 = second half of a
// "both" output; cf = the sample's coefficient row.
struct Pred { float x0, eps; };
__device__ __forceinline__ Pred predict(int type, float xt, float o, float o2, const float* cf) {
    Pred p;
    if (type == 3) {                                     // v: pred_x0_from_v, pred_eps_from_v (233-239)
        p.x0 = __fsub_rn(__fmul_rn(xt, cf[0]), __fmul_rn(o, cf[1]));
        p.eps = __fadd_rn(__fmul_rn(xt, cf[1]), __fmul_rn(o, cf[0]));
    } else if (type == 0) {                              // x0: eps = pred_eps_from_x0 (222-223)
        p.x0 = o;
        p.eps = __fsub_rn(__fmul_rn(xt, cf[12]), __fmul_rn(o, cf[13]));
    } else if (type == 1) {                              // eps: x0 = pred_x0_from_eps (207-208)
        p.eps = o;
        p.x0 = __fsub_rn(__fmul_rn(xt, cf[2]), __fmul_rn(o, cf[3]));
    } else {                                             // both: pred_x0_from_x0eps (211-214), then pred_eps_from_x0
        const float xe = __fsub_rn(__fmul_rn(xt, cf[2]), __fmul_rn(o2, cf[3]));
        p.x0 = __fadd_rn(__fmul_rn(o, cf[5]), __fmul_rn(xe, cf[4]));
        p.eps = __fsub_rn(__fmul_rn(xt, cf[12]), __fmul_rn(p.x0, cf[13]));
    }
    return p;
}

__device__ __forceinline__ float block_sum(float v, float* red) {       // 256 threads, fixed order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k];
    return t;
}

// One CTA per sample.  reweight: 0 constant, 1 snr, 2 snr_trunc, 3 snr_1plus.  The single-target reweightings compare the
// target with the RAW model output (diffusion.py:541 uses model_out, not `predict`), reproduced as is.
__global__ void __launch_bounds__(256) train_loss_kernel(const float* __restrict__ model_out, const float* __restrict__ x0,
                                                         const float* __restrict__ noise, const float* __restrict__ x_t,
                                                         const float* __restrict__ coef, float* __restrict__ loss,
                                                         float* __restrict__ grad_out, int B, int C, int HW, int type, int reweight) {
    __shared__ float red[8];
    __shared__ float cf[kCoefStride];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid < kCoefStride) cf[tid] = coef[b * kCoefStride + tid];
    __syncthreads();
    const int N = C * HW, Cm = (type == 2) ? 2 * C : C;
    const float* mo = model_out + static_cast<size_t>(b) * Cm * HW;
    const float* x0b = x0 + static_cast<size_t>(b) * N;
    const float* nb = noise + static_cast<size_t>(b) * N;
    const float* xtb = x_t + static_cast<size_t>(b) * N;
    float s0 = 0.f, s1 = 0.f;
    for (int i = tid; i < N; i += 256) {
        const float o = mo[i], o2 = (type == 2) ? mo[N + i] : 0.f;
        if (reweight == 2) {
            const Pred p = predict(type, xtb[i], o, o2, cf);
            const float d0 = x0b[i] - p.x0, d1 = nb[i] - p.eps;
            s0 += d0 * d0; s1 += d1 * d1;
        } else {
            const float tgt = reweight == 0 ? x0b[i] : reweight == 1 ? nb[i]
                              : __fadd_rn(-__fmul_rn(x0b[i], cf[1]), __fmul_rn(nb[i], cf[0]));   // pred_v_from_x0eps (226-227)
            const float d = tgt - o;
            s0 += d * d;
        }
    }
    const float m0 = block_sum(s0, red) / static_cast<float>(N);
    const float m1 = (reweight == 2) ? block_sum(s1, red) / static_cast<float>(N) : 0.f;
    const bool eps_branch = (reweight == 2) && (m1 > m0);              // torch.maximum(mse_x0, mse_eps)
    if (tid == 0) loss[b] = (reweight == 2) ? fmaxf(m0, m1) : m0;
    if (grad_out == nullptr) return;
    // d mean_b(loss_b) / d model_out: 2 (pred - target) / (N B) through the prediction's dependence on the output
    const float g = 2.0f / (static_cast<float>(N) * static_cast<float>(B));
    float* go = grad_out + static_cast<size_t>(b) * Cm * HW;
    for (int i = tid; i < N; i += 256) {
        const float o = mo[i], o2 = (type == 2) ? mo[N + i] : 0.f;
        if (reweight != 2) {
            const float tgt = reweight == 0 ? x0b[i] : reweight == 1 ? nb[i]
                              : __fadd_rn(-__fmul_rn(x0b[i], cf[1]), __fmul_rn(nb[i], cf[0]));
            go[i] = g * (o - tgt);
            continue;
        }
        const Pred p = predict(type, xtb[i], o, o2, cf);
        if (!eps_branch) {                                             // d pred_x0 / d out
            const float d = g * (p.x0 - x0b[i]);
            if (type == 3) go[i] = -d * cf[1];
            else if (type == 0) go[i] = d;
            else if (type == 1) go[i] = -d * cf[3];
            else { go[i] = d * cf[5]; go[N + i] = -d * cf[3] * cf[4]; }
        } else {                                                       // d pred_eps / d out
            const float d = g * (p.eps - nb[i]);
            if (type == 3) go[i] = d * cf[0];
            else if (type == 0) go[i] = -d * cf[13];
            else if (type == 1) go[i] = d;
            else { const float dx = -d * cf[13]; go[i] = dx * cf[5]; go[N + i] = -dx * cf[3] * cf[4]; }
        }
    }
}

}  // namespace

cudaError_t launch_q_sample(const float* x0, const float* eps, const float* coef, float* x_t, int B, int chw, cudaStream_t stream) {
    const long long n = static_cast<long long>(B) * chw;
    if (n == 0) return cudaSuccess;
    if (chw % 4 != 0) return cudaErrorInvalidValue;
    q_sample_kernel<<<static_cast<unsigned>((n / 4 + 255) / 256), 256, 0, stream>>>(x0, eps, coef, x_t, n, chw);
    return cudaGetLastError();
}

cudaError_t launch_train_loss(const float* model_out, const float* x0, const float* noise, const float* x_t, const float* coef,
                              float* loss, float* grad_out, int B, int C, int HW, int type, int reweight, cudaStream_t stream) {
    if (B == 0) return cudaSuccess;
    train_loss_kernel<<<B, 256, 0, stream>>>(model_out, x0, noise, x_t, coef, loss, grad_out, B, C, HW, type, reweight);
    return cudaGetLastError();
}

}  // namespace vdt
