// Backward of the fused GroupNorm(32, eps 1e-6) -> FiLM -> SiLU -> dropout activation of a ResidualBlock
// (unet.py:106-148; forward: groupnorm.cu).  HBM-bound: two streaming passes over (x, dA), everything else is tiny.
//
//   xhat = (x - mean_g) * rstd_g,  y = gamma * xhat + beta,  z = (1 + scale_b) * y + shift_b,  a = drop(silu(z))
//
//   dz        = dA * mask / (1 - p) * silu'(z)
//   A[b, c]   = sum_hw dz              Bc[b, c] = sum_hw dz * xhat               (pass 1, per-sample slabs, fixed order)
//   dshift    = A                      dscale   = gamma * Bc + beta * A
//   dgamma[c] = sum_b (1 + scale) Bc   dbeta[c] = sum_b (1 + scale) A
//   m1[b, g]  = mean over the group of dxhat = sum_c gamma (1 + scale) A / n,    m2 = the same with Bc
//   dx        = rstd * (gamma (1 + scale) dz - m1 - xhat * m2)                   (pass 2)
//
// The 16-bit rounding of the forward output is treated as the identity (straight-through).  The dropout mask is
// regenerated from the forward's Philox stream (same counter / key / word per element), never stored.
// All reductions run in a fixed order: results are bit-reproducible run to run.
#include "kernels.cuh"
#include "ptx.cuh"

namespace vdt {
namespace {

constexpr int kGroups = 32;
constexpr float kEps = 1e-6f;
constexpr int kThreads = 256;

// (mean, rstd) of every group of one sample: fp32 partial sums per thread, combined in fp64 (as the forward's fallback pass)
__global__ void __launch_bounds__(kThreads) gn_stats_kernel(const float* __restrict__ x, float2* __restrict__ meanrstd, int HW, int C) {
    extern __shared__ float2 part[];                       // [kThreads]
    const int b = blockIdx.x, tid = threadIdx.x;
    const int CV = C / 4, PPH = kThreads / CV;             // channel vectors, pixel phases (host: CV divides kThreads)
    const int cv = tid % CV, pp = tid / CV;
    const float* src = x + static_cast<size_t>(b) * HW * C + cv * 4;
    float s = 0.f, ss = 0.f;
    if (pp < PPH)
        for (int pix = pp; pix < HW; pix += PPH) {
            const float4 v = *reinterpret_cast<const float4*>(src + static_cast<size_t>(pix) * C);
            s += (v.x + v.y) + (v.z + v.w);
            ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ss))));
        }
    part[tid] = make_float2(s, ss);
    __syncthreads();
    if (tid < kGroups) {
        const int cpg = C / kGroups, vpg = cpg / 4;
        double ds = 0.0, dss = 0.0;
        for (int ph = 0; ph < PPH; ++ph)
            for (int j = 0; j < vpg; ++j) {
                const float2 t = part[ph * CV + tid * vpg + j];
                ds += t.x; dss += t.y;
            }
        const double n = static_cast<double>(cpg) * HW;
        const double mean = ds / n;
        double var = dss / n - mean * mean;
        if (var < 0.0) var = 0.0;
        meanrstd[b * kGroups + tid] = make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + static_cast<double>(kEps))));
    }
}

struct Elem { float xhat[4], dz[4]; };

// The per-element chain shared by both passes: recompute xhat, z and the activation's derivative, apply the mask.
struct Chain {
    float rstd, mean;
    float ga[4], be[4], fs[4], fb[4];
    int silu;
    bool drop; uint32_t thr; float dscale; unsigned long long seed; int layer;
    __device__ __forceinline__ Elem eval(const float4 xv, const float4 gv, unsigned long long e) const {
        const float x[4] = {xv.x, xv.y, xv.z, xv.w}, g[4] = {gv.x, gv.y, gv.z, gv.w};
        uint32_t w[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
        if (drop) {
            const uint4 r = philox4x32(make_uint4(static_cast<uint32_t>(e >> 2), static_cast<uint32_t>(e >> 34), static_cast<uint32_t>(layer), 0x44524f50u),
                                       make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
            w[0] = r.x; w[1] = r.y; w[2] = r.z; w[3] = r.w;
        }
        Elem o;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            o.xhat[i] = (x[i] - mean) * rstd;
            const float z = fs[i] * (ga[i] * o.xhat[i] + be[i]) + fb[i];
            float d = g[i];
            if (drop) d = (w[i] < thr) ? 0.f : d * dscale;
            if (silu) {
                const float sg = 1.f / (1.f + expf(-z));
                d *= sg * (1.f + z * (1.f - sg));
            }
            o.dz[i] = d;
        }
        return o;
    }
};

struct BwdArgs {
    const float* x; const float* grad_out; const float2* meanrstd;
    const float* gamma; const float* beta; const float* film;      // film: [B][2C] (shift | scale) or null
    int B, HW, C, silu, slabs;
    float drop_p; unsigned long long seed; int layer;
    float2* partial;            // [B][slabs][C] (sum dz, sum dz * xhat)
    float2* ab;                 // [B][C] the same, summed over the slabs
    float2* m12;                // [B][32]
    float* grad_x; float* grad_gamma; float* grad_beta; float* grad_film;
};

__device__ __forceinline__ Chain make_chain(const BwdArgs& p, int b, int c) {
    Chain ch;
    const float2 st = p.meanrstd[b * kGroups + c / (p.C / kGroups)];
    ch.mean = st.x; ch.rstd = st.y;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        ch.ga[i] = p.gamma[c + i]; ch.be[i] = p.beta[c + i];
        ch.fb[i] = p.film ? p.film[static_cast<size_t>(b) * 2 * p.C + c + i] : 0.f;
        ch.fs[i] = p.film ? 1.f + p.film[static_cast<size_t>(b) * 2 * p.C + p.C + c + i] : 1.f;
    }
    ch.silu = p.silu;
    ch.drop = p.drop_p > 0.f;
    ch.thr = ch.drop ? static_cast<uint32_t>(fminf(p.drop_p, 0.99999994f) * 4294967296.0f) : 0u;
    ch.dscale = ch.drop ? 1.f / (1.f - p.drop_p) : 1.f;
    ch.seed = p.seed; ch.layer = p.layer;
    return ch;
}

// pass 1: grid (slabs, B); a CTA owns the pixels [slab * HW / slabs, ...) of one sample
__global__ void __launch_bounds__(kThreads) gn_bwd_reduce_kernel(const BwdArgs p) {
    extern __shared__ float2 red[];                        // [kThreads][4]
    const int b = blockIdx.y, slab = blockIdx.x, tid = threadIdx.x;
    const int CV = p.C / 4, PPH = kThreads / CV;
    const int cv = tid % CV, pp = tid / CV, c = cv * 4;
    const int p0 = static_cast<int>(static_cast<long long>(slab) * p.HW / p.slabs), p1 = static_cast<int>(static_cast<long long>(slab + 1) * p.HW / p.slabs);
    float a[4] = {0.f, 0.f, 0.f, 0.f}, bc[4] = {0.f, 0.f, 0.f, 0.f};
    if (pp < PPH) {
        const Chain ch = make_chain(p, b, c);
        for (int pix = p0 + pp; pix < p1; pix += PPH) {
            const size_t off = (static_cast<size_t>(b) * p.HW + pix) * p.C + c;
            const Elem e = ch.eval(*reinterpret_cast<const float4*>(p.x + off), *reinterpret_cast<const float4*>(p.grad_out + off), off);
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] += e.dz[i]; bc[i] = fmaf(e.dz[i], e.xhat[i], bc[i]); }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) red[tid * 4 + i] = make_float2(a[i], bc[i]);
    __syncthreads();
    if (tid < CV) {                                        // fixed order over the pixel phases
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float sa = 0.f, sb = 0.f;
            for (int ph = 0; ph < PPH; ++ph) { const float2 t = red[(ph * CV + tid) * 4 + i]; sa += t.x; sb += t.y; }
            p.partial[(static_cast<size_t>(b) * p.slabs + slab) * p.C + tid * 4 + i] = make_float2(sa, sb);
        }
    }
}

// per sample: slab sums -> ab[b][c], FiLM gradients, group means m1 / m2
__global__ void __launch_bounds__(kThreads) gn_bwd_finalize_kernel(const BwdArgs p) {
    extern __shared__ double2 gsum[];                      // [C] gamma (1 + scale) (A, Bc)
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
        double sa = 0.0, sb = 0.0;
        for (int s = 0; s < p.slabs; ++s) {
            const float2 t = p.partial[(static_cast<size_t>(b) * p.slabs + s) * p.C + c];
            sa += t.x; sb += t.y;
        }
        p.ab[static_cast<size_t>(b) * p.C + c] = make_float2(static_cast<float>(sa), static_cast<float>(sb));
        const double g = p.gamma[c], be = p.beta[c];
        const double fs = p.film ? 1.0 + p.film[static_cast<size_t>(b) * 2 * p.C + p.C + c] : 1.0;
        if (p.grad_film) {
            p.grad_film[static_cast<size_t>(b) * 2 * p.C + c] = static_cast<float>(sa);                  // d shift
            p.grad_film[static_cast<size_t>(b) * 2 * p.C + p.C + c] = static_cast<float>(g * sb + be * sa);   // d scale
        }
        gsum[c] = make_double2(g * fs * sa, g * fs * sb);
    }
    __syncthreads();
    if (threadIdx.x < kGroups) {
        const int cpg = p.C / kGroups;
        double m1 = 0.0, m2 = 0.0;
        for (int j = 0; j < cpg; ++j) { m1 += gsum[threadIdx.x * cpg + j].x; m2 += gsum[threadIdx.x * cpg + j].y; }
        const double n = static_cast<double>(cpg) * p.HW;
        p.m12[b * kGroups + threadIdx.x] = make_float2(static_cast<float>(m1 / n), static_cast<float>(m2 / n));
    }
}

// d gamma / d beta: one warp per channel, lanes stride over the samples, fixed shuffle tree (deterministic)
__global__ void __launch_bounds__(256) gn_bwd_param_kernel(const BwdArgs p) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= p.C) return;
    double dg = 0.0, db = 0.0;
    for (int b = lane; b < p.B; b += 32) {
        const float2 t = p.ab[static_cast<size_t>(b) * p.C + c];
        const double fs = p.film ? 1.0 + p.film[static_cast<size_t>(b) * 2 * p.C + p.C + c] : 1.0;
        db += fs * t.x; dg += fs * t.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { dg += __shfl_xor_sync(0xffffffffu, dg, o); db += __shfl_xor_sync(0xffffffffu, db, o); }
    if (lane == 0) { p.grad_gamma[c] = static_cast<float>(dg); p.grad_beta[c] = static_cast<float>(db); }
}

// pass 2: dx
__global__ void __launch_bounds__(kThreads) gn_bwd_apply_kernel(const BwdArgs p) {
    const int b = blockIdx.y, slab = blockIdx.x, tid = threadIdx.x;
    const int CV = p.C / 4, PPH = kThreads / CV;
    const int cv = tid % CV, pp = tid / CV, c = cv * 4;
    if (pp >= PPH) return;
    const int p0 = static_cast<int>(static_cast<long long>(slab) * p.HW / p.slabs), p1 = static_cast<int>(static_cast<long long>(slab + 1) * p.HW / p.slabs);
    const Chain ch = make_chain(p, b, c);
    const float2 m = p.m12[b * kGroups + c / (p.C / kGroups)];
    float k[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) k[i] = ch.ga[i] * ch.fs[i];
    for (int pix = p0 + pp; pix < p1; pix += PPH) {
        const size_t off = (static_cast<size_t>(b) * p.HW + pix) * p.C + c;
        const Elem e = ch.eval(*reinterpret_cast<const float4*>(p.x + off), *reinterpret_cast<const float4*>(p.grad_out + off), off);
        float4 o;
        o.x = ch.rstd * (k[0] * e.dz[0] - m.x - e.xhat[0] * m.y);
        o.y = ch.rstd * (k[1] * e.dz[1] - m.x - e.xhat[1] * m.y);
        o.z = ch.rstd * (k[2] * e.dz[2] - m.x - e.xhat[2] * m.y);
        o.w = ch.rstd * (k[3] * e.dz[3] - m.x - e.xhat[3] * m.y);
        *reinterpret_cast<float4*>(p.grad_x + off) = o;
    }
}

}  // namespace

size_t groupnorm_backward_scratch_bytes(int B, int HW, int C, int* slabs_out) {
    int slabs = 1;
    while (slabs < 16 && HW / (slabs * 2) >= 64) slabs *= 2;       // at least 64 pixels per CTA
    if (slabs_out) *slabs_out = slabs;
    return (static_cast<size_t>(B) * slabs * C + static_cast<size_t>(B) * C + 2 * static_cast<size_t>(B) * kGroups) * sizeof(float2);
}

cudaError_t launch_groupnorm_backward(const GroupNormBwdParams& q, cudaStream_t stream) {
    // a thread owns 4 channels of one group, a CTA row of threads covers all channels: C in {128, 256, 512, 1024}
    if (q.C % (4 * kGroups) != 0 || q.C / 4 > kThreads || kThreads % (q.C / 4) != 0) return cudaErrorInvalidValue;
    int slabs = 1;
    groupnorm_backward_scratch_bytes(q.B, q.HW, q.C, &slabs);
    BwdArgs p{};
    p.x = q.x; p.grad_out = q.grad_out; p.gamma = q.gamma; p.beta = q.beta; p.film = q.film;
    p.B = q.B; p.HW = q.HW; p.C = q.C; p.silu = q.silu; p.slabs = slabs;
    p.drop_p = q.drop_p; p.seed = q.drop_seed; p.layer = q.drop_layer;
    float2* s = static_cast<float2*>(q.scratch);
    p.partial = s; s += static_cast<size_t>(q.B) * slabs * q.C;
    p.ab = s; s += static_cast<size_t>(q.B) * q.C;
    p.m12 = s; s += static_cast<size_t>(q.B) * kGroups;
    float2* mr = s;
    p.meanrstd = mr;
    p.grad_x = q.grad_x; p.grad_gamma = q.grad_gamma; p.grad_beta = q.grad_beta; p.grad_film = q.grad_film;
    gn_stats_kernel<<<q.B, kThreads, kThreads * sizeof(float2), stream>>>(q.x, mr, q.HW, q.C);
    gn_bwd_reduce_kernel<<<dim3(slabs, q.B), kThreads, kThreads * 4 * sizeof(float2), stream>>>(p);
    gn_bwd_finalize_kernel<<<q.B, kThreads, q.C * sizeof(double2), stream>>>(p);
    gn_bwd_param_kernel<<<(q.C + 7) / 8, 256, 0, stream>>>(p);
    gn_bwd_apply_kernel<<<dim3(slabs, q.B), kThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace vdt
