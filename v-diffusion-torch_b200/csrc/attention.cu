// Fused self-attention for sm_100a: softmax(q^T k / sqrt(d)) v per image and head, flash-style.
//
// Replaces BaseAttentionBlock.scaled_dot_product (unet.py:55-64: two einsums + a softmax that
// materialise the N x N score matrix in HBM).  Here one CTA owns 128 query rows; S = Q K^T and
// O += P V run on tcgen05 with S and O resident in TMEM, the softmax runs in registers (one thread
// per query row, fp32, exp2 with the 1/sqrt(d) scale folded in), P is re-quantised to bf16 into
// swizzled shared memory as the A operand of the second MMA.  O is rescaled lazily (only when a
// row maximum grows by more than 2^8), so the TMEM round trip is rare.
//
// Operands (written by the proj_in GEMM epilogue): QK bf16 [B*N, 2*hid] (q | k, heads contiguous,
// unet.py:76-78) and V^T bf16 [B*hid, N] (the v third of proj_in, stored transposed), so every MMA operand is K-major.
//
//   warp 0     TMA producer for Q (once) and the 64-key K tiles (2-stage ring, slot freed when S = Q K^T retires)
//   warp 10    TMA producer for the V^T tiles (2-stage ring, slot freed when O += P V retires)
//   warp 1     MMA issuer
//   warps 2-9  softmax + final normalise/store: two threads per query row (= TMEM lane), each owning 32
//              of the tile's 64 key columns (and half of O's columns); row maxima are exchanged through
//              shared memory once per tile
#include <atomic>

#include "kernels.cuh"
#include "ptx.cuh"

namespace vdt {
namespace {

constexpr int kQRows = 128;
constexpr int kKeys = 64;                      // keys per tile = one 128-byte swizzle row of P / V^T
constexpr int kThreads = 64 + 256 + 32;
constexpr float kRescaleThreshold = 8.0f;      // log2 units

struct Smem {          // stage pointers are computed, not indexed (no local-memory arrays)
    uint8_t* q; uint8_t* k0; uint8_t* v0; uint8_t* p0;
    int k_stride, v_stride;
    uint64_t* q_full; uint64_t* k_full; uint64_t* v_full; uint64_t* k_empty; uint64_t* v_empty; uint64_t* s_full; uint64_t* p_full;
    uint32_t* tmem_slot;
    __device__ __forceinline__ uint8_t* k(int s) const { return k0 + s * k_stride; }
    __device__ __forceinline__ uint8_t* v(int s) const { return v0 + s * v_stride; }
    __device__ __forceinline__ uint8_t* p(int s) const { return p0 + s * 16384; }
};

__host__ __device__ inline int attn_smem_bytes(int d) {
    // Q d/64 x 16K | K 2 x d/64 x 8K | V 2 x d*128 | P 2 x 16K | barriers | align slack
    return (d / 64) * 16384 + 2 * (d / 64) * 8192 + 2 * d * 128 + 2 * 16384 + 256 + 2048 /*row max / sum exchange*/;
}

template <bool F16>
__global__ void __launch_bounds__(kThreads, 1) attention_kernel(const __grid_constant__ AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];   // 128B-swizzled tiles need 1024-byte alignment
    uint8_t* base = smem_raw;
    if ((smem_u32(base) & 1023u) != 0) __trap();
    const int d = p.d, dch = d / 64;
    Smem sm;
    sm.q = base; base += dch * 16384;
    sm.k0 = base; sm.k_stride = dch * 8192; base += 2 * dch * 8192;
    sm.v0 = base; sm.v_stride = d * 128; base += 2 * d * 128;
    sm.p0 = base; base += 2 * 16384;
    uint64_t* bars = reinterpret_cast<uint64_t*>(base);
    sm.q_full = bars; sm.k_full = bars + 1; sm.v_full = bars + 3; sm.k_empty = bars + 5; sm.v_empty = bars + 7;
    sm.s_full = bars + 9; sm.p_full = bars + 11;              // s_full[2], p_full[2]
    sm.tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
    float* xch = reinterpret_cast<float*>(base + 256);      // [2][2][128]: double-buffered row-max exchange; reused for l

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = p.N;
    // packed: N == 64 puts two images into one CTA (each 64-key tile belongs to one of them).  Otherwise one image
    // per CTA with ceil(N / 128) query tiles and ceil(N / 64) key tiles; rows / keys past the image's N are loaded
    // (they belong to the next image or are zero-filled by TMA) and masked: keys to -inf, query rows at the store.
    const bool packed = (2 * N == kQRows);
    const int ipc = packed ? 2 : 1;                         // images per CTA
    const int qtiles = packed ? 1 : (N + kQRows - 1) / kQRows;
    const int nt = packed ? 2 : (N + kKeys - 1) / kKeys;
    // blockIdx.x -> (image group, head, q tile)
    const int qt = blockIdx.x % qtiles;
    const int h = (blockIdx.x / qtiles) % p.heads;
    const int b0 = (blockIdx.x / (qtiles * p.heads)) * ipc;
    const long long row0 = static_cast<long long>(b0) * N + static_cast<long long>(qt) * kQRows;   // first query row
    const long long krow0 = static_cast<long long>(b0) * N;                                          // first key row
    uint32_t tmem_cols = 128; while (tmem_cols < static_cast<uint32_t>(d + 2 * kKeys)) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        mbar_init(sm.q_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&sm.k_full[s], 1); mbar_init(&sm.v_full[s], 1); mbar_init(&sm.k_empty[s], 1);
                                      mbar_init(&sm.v_empty[s], 1); mbar_init(&sm.p_full[s], 8); }   // one arrival per softmax warp
        mbar_init(&sm.s_full[0], 1); mbar_init(&sm.s_full[1], 1);
        fence_mbar_init();
        // Q and the first two K / V^T tiles are requested right away, before the TMEM allocation and the CTA-wide
        // sync, so their L2 latency overlaps the rest of the prologue (the ring slots are trivially free)
        mbar_expect_tx(sm.q_full, static_cast<uint32_t>(dch * 16384));
        for (int c = 0; c < dch; ++c)
            tma_load_2d(sm.q + c * 16384, &p.qk_map, sm.q_full, h * d + c * 64, static_cast<int>(row0));
        for (int j = 0; j < 2 && j < nt; ++j) {
            mbar_expect_tx(&sm.k_full[j], static_cast<uint32_t>(dch * 8192));
            for (int c = 0; c < dch; ++c)
                tma_load_2d(sm.k(j) + c * 8192, &p.k_map, &sm.k_full[j], p.hid + h * d + c * 64, static_cast<int>(krow0) + j * kKeys);
            const int img = packed ? b0 + j : b0;
            const int koff = packed ? 0 : j * kKeys;
            mbar_expect_tx(&sm.v_full[j], static_cast<uint32_t>(d * 128));
            tma_load_2d(sm.v(j), &p.vt_map, &sm.v_full[j], koff, (img * p.heads + h) * d);
        }
    }
    if (warp == 1) tmem_alloc(sm.tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_slot;
    const uint32_t tmem_o = tmem_base;                      // columns [0, d)
    const uint32_t tmem_s = tmem_base + static_cast<uint32_t>(d);   // two S buffers: columns [d, d+64) and [d+64, d+128)

    if (warp == 0) {
        if (lane == 0) {
            for (int j = 2; j < nt; ++j) {
                const int s = j & 1;
                mbar_wait(&sm.k_empty[s], ((j >> 1) & 1) ^ 1);
                mbar_expect_tx(&sm.k_full[s], static_cast<uint32_t>(dch * 8192));
                for (int c = 0; c < dch; ++c)
                    tma_load_2d(sm.k(s) + c * 8192, &p.k_map, &sm.k_full[s], p.hid + h * d + c * 64,
                                static_cast<int>(krow0) + j * kKeys);
            }
        }
    } else if (warp == 10) {
        if (lane == 0) {
            for (int j = 2; j < nt; ++j) {
                const int s = j & 1;
                mbar_wait(&sm.v_empty[s], ((j >> 1) & 1) ^ 1);
                const int img = packed ? b0 + j : b0;
                const int koff = packed ? 0 : j * kKeys;
                mbar_expect_tx(&sm.v_full[s], static_cast<uint32_t>(d * 128));
                tma_load_2d(sm.v(s), &p.vt_map, &sm.v_full[s], koff, (img * p.heads + h) * d);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s = umma_idesc_16(kQRows, kKeys, (F16 ? 1 : 0));
            const uint32_t idesc_o = umma_idesc_16(kQRows, d, (F16 ? 1 : 0));
            auto issue_s = [&](int j) {
                const int s = j & 1;
                mbar_wait(&sm.k_full[s], (j >> 1) & 1);
                tc_fence_after();
                for (int kk = 0; kk < d / 16; ++kk) {
                    const uint64_t ad = umma_desc_sw128(smem_u32(sm.q + (kk >> 2) * 16384)) + 2 * (kk & 3);
                    const uint64_t bd = umma_desc_sw128(smem_u32(sm.k(s) + (kk >> 2) * 8192)) + 2 * (kk & 3);
                    umma_16(tmem_s + static_cast<uint32_t>(s * kKeys), ad, bd, idesc_s, kk != 0);
                }
                umma_commit(&sm.k_empty[s]);              // K_j can be overwritten as soon as S_j has retired
                umma_commit(&sm.s_full[s]);
            };
            mbar_wait(sm.q_full, 0);
            issue_s(0);
            if (nt > 1) issue_s(1);                       // S is double-buffered: S_{j+1} is ready before softmax j ends
            for (int j = 0; j < nt; ++j) {
                const int s = j & 1;
                mbar_wait(&sm.p_full[s], (j >> 1) & 1);   // P_j in smem, S_j consumed (its buffer is free), O rescaled if needed
                tc_fence_after();
                if (j + 2 < nt) issue_s(j + 2);
                mbar_wait(&sm.v_full[s], (j >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < kKeys / 16; ++kk) {
                    const uint64_t ad = umma_desc_sw128(smem_u32(sm.p(s))) + 2 * kk;
                    const uint64_t bd = umma_desc_sw128(smem_u32(sm.v(s))) + 2 * kk;
                    umma_16(tmem_o, ad, bd, idesc_o, (j | kk) != 0);
                }
                umma_commit(&sm.v_empty[s]);              // V_j / P_j free, O updated
            }
        }
    } else if (warp < 10) {
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;                   // which 32 key columns of the tile (and which half of O's columns)
        const int row = q * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
        const long long grow = row0 + row;
        const bool row_ok = (grow < static_cast<long long>(p.B) * N) && (packed || qt * kQRows + row < N);
        const int row_img = row / N;                        // only meaningful when packed
        const float c = p.scale_log2e;
        const int dh = d >> 1;                              // O columns owned by this thread: [half*dh, half*dh + dh)
        float m_used = -INFINITY, l = 0.f;
        for (int j = 0; j < nt; ++j) {
            mbar_wait(&sm.s_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            uint32_t r0[32];
            tmem_ld32(tmem_s + static_cast<uint32_t>((j & 1) * kKeys) + lane_off + half * 32, r0);
            tmem_ld_wait();
            float sv[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) sv[i] = __uint_as_float(r0[i]);
            // packed: a 64-key tile belongs to one image; rows of the other image ignore it entirely
            const bool tile_valid = !packed || (j == row_img);
            if (!packed && (j + 1) * kKeys > N) {           // ragged last key tile (warp-uniform): keys >= N drop out
                const int first_bad = N - j * kKeys - half * 32;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i >= first_bad) sv[i] = -INFINITY;
            }
            float m_half = -INFINITY;
            if (tile_valid) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) m_half = fmaxf(m_half, fmaxf(sv[i], sv[i + 1]));
            }
            float* xm = xch + (j & 1) * 256;
            xm[half * 128 + row] = m_half;
            asm volatile("bar.sync 1, 256;" ::: "memory");   // the 8 softmax warps only
            const float m_tile = fmaxf(m_half, xm[(half ^ 1) * 128 + row]);
            float factor = 1.f;
            bool need = false;
            if (m_tile > m_used) {
                if (m_used == -INFINITY) {
                    m_used = m_tile;                        // nothing accumulated yet for this row (O row == 0, l == 0)
                } else if ((m_tile - m_used) * c > kRescaleThreshold) {
                    factor = ex2_approx((m_used - m_tile) * c);
                    m_used = m_tile;
                    need = true;
                }
            }
            if (__any_sync(0xffffffffu, need)) {
                // O must be complete through PV_{j-1} before it is rescaled
                mbar_wait(&sm.v_empty[(j - 1) & 1], ((j - 1) >> 1) & 1);
                tc_fence_after();
                for (int c0 = half * dh; c0 < half * dh + dh; c0 += 32) {
                    uint32_t o[32];
                    tmem_ld32(tmem_o + lane_off + c0, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
                    tmem_st32(tmem_o + lane_off + c0, o);
                }
                tmem_st_wait();
                l *= factor;
            }
            // P_j goes into the buffer PV_{j-2} read: S_j was issued before PV_{j-2}, so s_full alone does not
            // order this tile's stores after that MMA's operand reads (a late V_{j-2} tile delays it arbitrarily)
            if (j >= 2) mbar_wait(&sm.v_empty[j & 1], ((j - 2) >> 1) & 1);
            // p = 2^(s*c - m*c); a row that ignores this tile writes zeros (its m may still be -inf)
            const float mc = (m_used == -INFINITY) ? 0.f : m_used * c;
            const float cc = tile_valid ? c : 0.f;
            const float off = tile_valid ? mc : 200.f;      // 2^-200 flushes to exactly 0
            uint8_t* prow = sm.p(j & 1) + row * 128;
            float lsum = 0.f;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                float e[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    e[i] = ex2_approx(fmaf(sv[ch * 8 + i], cc, -off));
                    lsum += e[i];
                }
                *reinterpret_cast<uint4*>(prow + (((half * 4 + ch) ^ (row & 7)) << 4)) =
                    make_uint4(pack_16_inrange(e[0], e[1], (F16 ? 1 : 0)), pack_16_inrange(e[2], e[3], (F16 ? 1 : 0)), pack_16_inrange(e[4], e[5], (F16 ? 1 : 0)),
                               pack_16_inrange(e[6], e[7], (F16 ? 1 : 0)));   // p <= 2^8
            }
            l += lsum;
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.p_full[j & 1]);
        }
        // ---- final: O / l -> 16-bit.  The two threads of a row add their partial sums.  (The barrier keeps a fast
        // thread from overwriting a row maximum of the last tile that its partner has not read yet: with an odd
        // number of key tiles that exchange used this same buffer.)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        xch[half * 128 + row] = l;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        l += xch[(half ^ 1) * 128 + row];
        mbar_wait(&sm.v_empty[(nt - 1) & 1], ((nt - 1) >> 1) & 1);
        tc_fence_after();
        const float inv_l = (l > 0.f) ? 1.f / l : 0.f;
        for (int c0 = half * dh; c0 < half * dh + dh; c0 += 32) {
            uint32_t o[32];
            tmem_ld32(tmem_o + lane_off + c0, o);
            tmem_ld_wait();
            if (row_ok) {
                uint4* dst = reinterpret_cast<uint4*>(p.out + grow * p.hid + h * d + c0);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    dst[i] = make_uint4(pack_16(__uint_as_float(o[8 * i]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l, (F16 ? 1 : 0)),
                                        pack_16(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l, (F16 ? 1 : 0)),
                                        pack_16(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l, (F16 ? 1 : 0)),
                                        pack_16(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l, (F16 ? 1 : 0)));
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

}  // namespace

cudaError_t launch_attention(const AttnParams& p, cudaStream_t stream) {
    if (p.d % 64 != 0 || p.d > 256 || p.N < 1) return cudaErrorInvalidValue;
    const int smem = attn_smem_bytes(p.d);
    // cudaFuncSetAttribute is per device: the configured size is tracked per device ordinal
    static std::atomic<int> smem_set[kMaxDevices];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
    if (smem > smem_set[dev].load(std::memory_order_acquire)) {
        e = cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        smem_set[dev].store(smem, std::memory_order_release);
    }
    const bool packed = (2 * p.N == kQRows);
    const int ipc = packed ? 2 : 1;
    const int qtiles = packed ? 1 : (p.N + kQRows - 1) / kQRows;
    const int groups = (p.B + ipc - 1) / ipc;
    const long long grid = static_cast<long long>(groups) * p.heads * qtiles;
    if (grid <= 0) return cudaSuccess;
    if (p.f16) attention_kernel<true><<<static_cast<unsigned>(grid), kThreads, smem, stream>>>(p);
    else attention_kernel<false><<<static_cast<unsigned>(grid), kThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace vdt
