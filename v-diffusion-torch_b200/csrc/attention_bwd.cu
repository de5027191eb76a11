// Backward of the self-attention core softmax(q k^T / sqrt(d)) v (unet.py:55-64) in fp32 on CUDA cores -- the
// reference-grade counterpart of the validation-mode forward (pointwise.cu: attention_f32_kernel), here so that every op of
// an AttentionBlock has a backward with parity; it is NOT the tensor-core kernel a training step would ship
// (flash-style dQ / dK / dV on tcgen05 is the next step).
//
//   P = softmax(S), S = scale * q k^T;   O = P v;   D_i = dO_i . O_i
//   dV_j = sum_i P_ij dO_i      dP_ij = dO_i . v_j      dS_ij = P_ij (dP_ij - D_i)
//   dQ_i = scale * sum_j dS_ij k_j                      dK_j = scale * sum_i dS_ij q_i
//
// One warp per (query row, head) [pass A: softmax statistics, D_i, dQ_i] and one warp per (key row, head) [pass B: dK_j,
// dV_j]; lane l owns features l, l + 32, ...; every sum runs in a fixed order (bit-reproducible).  Nothing N x N is stored.
#include "kernels.cuh"
#include "ptx.cuh"

namespace vdt {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// pass A: per query row -- (log-sum-exp, D) to `stat`, dQ to grad_qkv's q third
template <int DPL>
__global__ void __launch_bounds__(256) attention_bwd_q_kernel(const float* __restrict__ qkv, const float* __restrict__ dout,
                                                              float* __restrict__ dqkv, float2* __restrict__ stat, int B, int N, int heads) {
    const int d = DPL * 32, hid = heads * d;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = static_cast<long long>(blockIdx.x) * 8 + warp;
    const long long qrow = item / heads;
    const int h = static_cast<int>(item % heads);
    if (qrow >= static_cast<long long>(B) * N) return;
    const long long img = qrow / N;
    const float scale = rsqrtf(static_cast<float>(d));
    float q[DPL], go[DPL], o[DPL], dq[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) {
        q[i] = qkv[qrow * 3 * hid + h * d + lane + 32 * i] * scale;
        go[i] = dout[qrow * hid + h * d + lane + 32 * i];
        o[i] = 0.f; dq[i] = 0.f;
    }
    // sweep 1: online softmax statistics and O_i (as the forward), then D_i = dO_i . O_i
    float m = -INFINITY, l = 0.f;
    for (int k = 0; k < N; ++k) {
        const float* kp = qkv + (img * N + k) * 3 * hid + hid + h * d;
        const float* vp = kp + hid;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < DPL; ++i) s = fmaf(q[i], __ldg(kp + lane + 32 * i), s);
        s = warp_sum(s);
        const float mn = fmaxf(m, s);
        const float corr = expf(m - mn), pr = expf(s - mn);
        l = l * corr + pr;
#pragma unroll
        for (int i = 0; i < DPL; ++i) o[i] = fmaf(o[i], corr, pr * __ldg(vp + lane + 32 * i));
        m = mn;
    }
    const float inv = 1.f / l;
    float dsum = 0.f;
#pragma unroll
    for (int i = 0; i < DPL; ++i) dsum = fmaf(go[i], o[i] * inv, dsum);
    const float D = warp_sum(dsum);
    const float lse = m + logf(l);
    if (lane == 0) stat[qrow * heads + h] = make_float2(lse, D);
    // sweep 2: dQ_i
    for (int k = 0; k < N; ++k) {
        const float* kp = qkv + (img * N + k) * 3 * hid + hid + h * d;
        const float* vp = kp + hid;
        float kk[DPL];
        float s = 0.f, dp = 0.f;
#pragma unroll
        for (int i = 0; i < DPL; ++i) {
            kk[i] = __ldg(kp + lane + 32 * i);
            s = fmaf(q[i], kk[i], s);
            dp = fmaf(go[i], __ldg(vp + lane + 32 * i), dp);
        }
        s = warp_sum(s); dp = warp_sum(dp);
        const float ds = expf(s - lse) * (dp - D);
#pragma unroll
        for (int i = 0; i < DPL; ++i) dq[i] = fmaf(ds, kk[i], dq[i]);
    }
#pragma unroll
    for (int i = 0; i < DPL; ++i) dqkv[qrow * 3 * hid + h * d + lane + 32 * i] = dq[i] * scale;
}

// pass B: per key row -- dK, dV to grad_qkv's k / v thirds
template <int DPL>
__global__ void __launch_bounds__(256) attention_bwd_kv_kernel(const float* __restrict__ qkv, const float* __restrict__ dout,
                                                               float* __restrict__ dqkv, const float2* __restrict__ stat, int B, int N, int heads) {
    const int d = DPL * 32, hid = heads * d;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = static_cast<long long>(blockIdx.x) * 8 + warp;
    const long long krow = item / heads;
    const int h = static_cast<int>(item % heads);
    if (krow >= static_cast<long long>(B) * N) return;
    const long long img = krow / N;
    const float scale = rsqrtf(static_cast<float>(d));
    float kk[DPL], vv[DPL], dk[DPL], dv[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) {
        kk[i] = qkv[krow * 3 * hid + hid + h * d + lane + 32 * i] * scale;
        vv[i] = qkv[krow * 3 * hid + 2 * hid + h * d + lane + 32 * i];
        dk[i] = 0.f; dv[i] = 0.f;
    }
    for (int r = 0; r < N; ++r) {
        const long long qrow = img * N + r;
        const float* qp = qkv + qrow * 3 * hid + h * d;
        const float* gp = dout + qrow * hid + h * d;
        float qq[DPL], go[DPL];
        float s = 0.f, dp = 0.f;
#pragma unroll
        for (int i = 0; i < DPL; ++i) {
            qq[i] = __ldg(qp + lane + 32 * i);
            go[i] = __ldg(gp + lane + 32 * i);
            s = fmaf(qq[i], kk[i], s);
            dp = fmaf(go[i], vv[i], dp);
        }
        s = warp_sum(s); dp = warp_sum(dp);
        const float2 st = __ldg(stat + qrow * heads + h);
        const float p = expf(s - st.x);
        const float ds = p * (dp - st.y);
#pragma unroll
        for (int i = 0; i < DPL; ++i) { dv[i] = fmaf(p, go[i], dv[i]); dk[i] = fmaf(ds, qq[i], dk[i]); }
    }
#pragma unroll
    for (int i = 0; i < DPL; ++i) {
        dqkv[krow * 3 * hid + hid + h * d + lane + 32 * i] = dk[i] * scale;
        dqkv[krow * 3 * hid + 2 * hid + h * d + lane + 32 * i] = dv[i];
    }
}

template <int DPL>
cudaError_t run(const float* qkv, const float* dout, float* dqkv, float2* stat, int B, int N, int heads, cudaStream_t stream) {
    const long long warps = static_cast<long long>(B) * N * heads;
    const unsigned grid = static_cast<unsigned>((warps + 7) / 8);
    attention_bwd_q_kernel<DPL><<<grid, 256, 0, stream>>>(qkv, dout, dqkv, stat, B, N, heads);
    attention_bwd_kv_kernel<DPL><<<grid, 256, 0, stream>>>(qkv, dout, dqkv, stat, B, N, heads);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_attention_backward_f32(const float* qkv, const float* dout, float* dqkv, void* stat_scratch, int B, int N, int heads,
                                          int d, cudaStream_t stream) {
    if (static_cast<long long>(B) * N * heads == 0) return cudaSuccess;
    float2* stat = static_cast<float2*>(stat_scratch);
    switch (d) {
        case 64: return run<2>(qkv, dout, dqkv, stat, B, N, heads, stream);
        case 128: return run<4>(qkv, dout, dqkv, stat, B, N, heads, stream);
        case 192: return run<6>(qkv, dout, dqkv, stat, B, N, heads, stream);
        case 256: return run<8>(qkv, dout, dqkv, stat, B, N, heads, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace vdt
