// Training-step kernels around the model call: GaussianDiffusion.train_loss (diffusion.py:492-545, loss_type "mse").
//
//   q_sample_kernel      x_t = x_0 * sqrt(sigmoid(l_t)) + eps * sqrt(sigmoid(-l_t))            (diffusion.py:242-245)
//   train_loss_kernel    from_model_out_to_pred (466-490) + the re-weighted MSE per sample (518-541), and optionally the
//                        gradient of loss.mean() with respect to the model output (what autograd hands the UNet's backward)
//
// Every per-sample scalar comes from a [B][16] coefficient table computed on the host in the reference's rounding chain
// (vdt_train_coefficients: log-SNR in fp64 -> fp32, sigmoid / sqrt / exp in fp32).  Products and sums are rounded
// separately (no FMA contraction), as the reference's chain of element-wise tensor ops rounds them.
#include <cmath>

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

// ------------------------------------------------------------------------------------------------
// Optimizer step of the reference trainer (train_utils.py:159-166): clip_grad_norm_ -> AdamW.step -> EMA.update
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int kSqBlock = 256;
constexpr int kSqItemsPerBlock = kSqBlock * 16;             // fixed tiling: the result does not depend on the grid

// sum of g^2 over one tensor: fp32 products, per-thread fp32 partials over a fixed stride, block tree in fp64, one
// partial per block; the last block to finish (ticket) adds the partials in index order -> deterministic
__global__ void __launch_bounds__(kSqBlock) grad_sq_kernel(const float* __restrict__ g, long long n, double* __restrict__ partial,
                                                           unsigned int* __restrict__ ticket, double* __restrict__ accum) {
    __shared__ double red[kSqBlock];
    __shared__ bool last;
    const long long base = static_cast<long long>(blockIdx.x) * kSqItemsPerBlock;
    float s = 0.f;
    for (int k = 0; k < 16; ++k) {
        const long long i = base + k * kSqBlock + threadIdx.x;
        if (i < n) { const float v = g[i]; s = fmaf(v, v, s); }
    }
    red[threadIdx.x] = static_cast<double>(s);
    __syncthreads();
    for (int o = kSqBlock / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = red[0];
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double t = 0.0;
        for (unsigned int b = 0; b < gridDim.x; ++b) t += partial[b];
        *accum += t;                                           // tensors are accumulated one launch after the other (stream order)
        *ticket = 0u;
    }
}

// One parameter tensor: gradient clipping coefficient from the total squared norm (device scalar, no host sync),
// decoupled weight decay, Adam moments, bias-corrected update (torch/optim/adamw.py, single-tensor form) and the EMA
// shadow update shadow += (1 - decay) * (param - shadow) (utils.py:144-149).
__global__ void __launch_bounds__(256) adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, float* __restrict__ shadow, long long n,
                                                        float decay_mul, float w1, float beta2, float w2, float eps, float step_size,
                                                        float inv_bc2_sqrt, const double* __restrict__ grad_sq_total, float max_norm,
                                                        float ema_w) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    float coef = 1.f;
    if (grad_sq_total != nullptr && max_norm > 0.f) {          // clip_grad_norm_: max_norm / (total_norm + 1e-6), clamped to 1
        const float total = static_cast<float>(sqrt(*grad_sq_total));
        coef = fminf(__fdiv_rn(max_norm, __fadd_rn(total, 1e-6f)), 1.f);
    }
    const float gi = __fmul_rn(g[i], coef);
    float pi = __fmul_rn(p[i], decay_mul);                                                           // param.mul_(1 - lr * wd)
    const float mi = __fadd_rn(m[i], __fmul_rn(w1, __fsub_rn(gi, m[i])));                            // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = __fadd_rn(__fmul_rn(v[i], beta2), __fmul_rn(__fmul_rn(w2, gi), gi));            // mul_(beta2).addcmul_(g, g, 1 - beta2)
    const float denom = __fadd_rn(__fmul_rn(__fsqrt_rn(vi), inv_bc2_sqrt), eps);                      // (sqrt / scalar) is sqrt * (1 / scalar) in ATen
    pi = __fsub_rn(pi, __fmul_rn(step_size, __fdiv_rn(mi, denom)));
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (shadow != nullptr) shadow[i] = __fadd_rn(shadow[i], __fmul_rn(ema_w, __fsub_rn(pi, shadow[i])));
}

}  // namespace

size_t grad_sq_scratch_bytes(long long n) {
    const long long blocks = (n + kSqItemsPerBlock - 1) / kSqItemsPerBlock;
    return static_cast<size_t>(blocks) * sizeof(double) + 16;
}
cudaError_t launch_grad_sq(const float* g, long long n, void* scratch, double* accum, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const long long blocks = (n + kSqItemsPerBlock - 1) / kSqItemsPerBlock;
    double* partial = static_cast<double*>(scratch);
    unsigned int* ticket = reinterpret_cast<unsigned int*>(partial + blocks);
    cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(unsigned int), stream);
    if (e != cudaSuccess) return e;
    grad_sq_kernel<<<static_cast<unsigned int>(blocks), kSqBlock, 0, stream>>>(g, n, partial, ticket, accum);
    return cudaGetLastError();
}
cudaError_t launch_adamw_ema(float* p, const float* g, float* m, float* v, float* shadow, long long n, double lr, double beta1,
                             double beta2, double eps, double wd, int step, const double* grad_sq_total, float max_norm,
                             double ema_decay, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    // scalar prologue of torch.optim's single-tensor AdamW: python floats (doubles), each cast to fp32 where it meets a tensor
    const double bc1 = 1.0 - std::pow(beta1, step), bc2 = 1.0 - std::pow(beta2, step);
    const float step_size = static_cast<float>(lr / bc1);
    const float inv_bc2_sqrt = 1.0f / static_cast<float>(std::sqrt(bc2));
    adamw_ema_kernel<<<static_cast<unsigned int>((n + 255) / 256), 256, 0, stream>>>(
        p, g, m, v, shadow, n, static_cast<float>(1.0 - lr * wd), static_cast<float>(1.0 - beta1), static_cast<float>(beta2),
        static_cast<float>(1.0 - beta2), static_cast<float>(eps), step_size, inv_bc2_sqrt, grad_sq_total, max_norm,
        static_cast<float>(1.0 - ema_decay));
    return cudaGetLastError();
}

}  // namespace vdt
