// Implicit-GEMM 3x3 / 1x1 convolution and linear layers for sm_100a.
//
// Replaces F.conv2d / F.linear on the reference's hot path (modules.py:79-80, 141-144 as called
// from unet.py:121,125,134,70-71,203-215,217,232).  NHWC bf16 operands, fp32 accumulation in
// TMEM, fused epilogue (bias, residual add, SiLU, bf16 / transposed / NCHW stores).
//
// Persistent warp-specialised kernel, one CTA per SM, 192 threads:
//   warp 0     TMA producer   : per K-block one 4-D box of the activation (shifted by the filter
//                               tap; out-of-bounds rows/cols are zero-filled by TMA = padding) and
//                               one 2-D box of the packed weight, into a 4-stage smem ring
//   warp 1     MMA issuer     : tcgen05.mma.cta_group::1.kind::f16, M=128 x N=block_n x K=16,
//                               accumulators double-buffered in TMEM (2 x 256 columns)
//   warps 2-5  epilogue       : tcgen05.ld -> registers -> global, overlapping the next tile's mainloop
#include "kernels.cuh"
#include "ptx.cuh"

namespace vdt {

namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;                       // bf16 elements per K-block = one 128-byte swizzle row
constexpr int kStages = 4;
constexpr int kMaxBN = 256;
constexpr int kABytes = kBM * kBK * 2;        // 16 KiB
constexpr int kBBytes = kMaxBN * kBK * 2;     // 32 KiB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int kThreads = 192;
constexpr int kTmemCols = 512;

struct TileCoord {
    int c1, c2, c3;   // TMA coordinates of the un-shifted tile origin (W, H, N)
};

__device__ __forceinline__ TileCoord tile_origin(const ConvParams& p, int m_tile) {
    TileCoord t;
    if (p.pointwise) {
        t.c1 = m_tile * kBM; t.c2 = 0; t.c3 = 0;
    } else {
        t.c1 = 0;
        t.c2 = (m_tile % p.tiles_per_image) * p.box_h;
        t.c3 = (m_tile / p.tiles_per_image) * p.box_n;
    }
    return t;
}

__global__ void __launch_bounds__(kThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* full_bar = bars;                    // [kStages] TMA -> MMA
    uint64_t* empty_bar = bars + kStages;         // [kStages] MMA -> TMA
    uint64_t* acc_full = bars + 2 * kStages;      // [2] MMA -> epilogue
    uint64_t* acc_empty = bars + 2 * kStages + 2; // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 128); }
        fence_mbar_init();
        for (int s = 0; s < p.num_segs; ++s) tma_prefetch_desc(&p.a_map[s]);
        tma_prefetch_desc(&p.b_map);
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int total_tiles = p.num_m_tiles * p.num_n_tiles;
    int kblocks_total = 0;
    for (int s = 0; s < p.num_segs; ++s) kblocks_total += p.seg_taps[s] * p.seg_kblocks[s];

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t tx_bytes = static_cast<uint32_t>(p.rows_per_tile * kBK * 2 + p.block_n * kBK * 2);
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m_tile = tile / p.num_n_tiles, n_tile = tile % p.num_n_tiles;
                const TileCoord o = tile_origin(p, m_tile);
                int kcol = 0;
                for (int s = 0; s < p.num_segs; ++s) {
                    const int taps = p.seg_taps[s];
                    for (int tap = 0; tap < taps; ++tap) {
                        const int dy = (taps == 9) ? tap / 3 - 1 : 0;
                        const int dx = (taps == 9) ? tap % 3 - 1 : 0;
                        for (int kb = 0; kb < p.seg_kblocks[s]; ++kb) {
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            uint8_t* a_dst = smem + stage * kStageBytes;
                            uint8_t* b_dst = a_dst + kABytes;
                            mbar_expect_tx(&full_bar[stage], tx_bytes);
                            tma_load_4d(a_dst, &p.a_map[s], &full_bar[stage], kb * kBK, o.c1 + dx, o.c2 + dy, o.c3);
                            tma_load_2d(b_dst, &p.b_map, &full_bar[stage], kcol, n_tile * p.block_n);
                            kcol += kBK;
                            if (++stage == kStages) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_16(kBM, p.block_n, p.f16);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(&acc_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kMaxBN);
                for (int kb = 0; kb < kblocks_total; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
                    const uint64_t adesc = umma_desc_sw128(a_addr);
                    const uint64_t bdesc = umma_desc_sw128(a_addr + kABytes);
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k)
                        umma_16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    umma_commit(&empty_bar[stage]);          // smem slot reusable once these MMAs retire
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&acc_full[acc]);                 // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        const int q = warp & 3;                              // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int m_tile = tile / p.num_n_tiles, n_tile = tile % p.num_n_tiles;
            mbar_wait(&acc_full[acc], acc_phase);
            tc_fence_after();
            const long long grow = static_cast<long long>(m_tile) * p.rows_per_tile + row;
            const bool row_ok = (row < p.rows_per_tile) && (grow < p.M);
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * kMaxBN);
            for (int c0 = 0; c0 < p.block_n; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(t_row + static_cast<uint32_t>(c0), r);
                tmem_ld_wait();
                const int col0 = n_tile * p.block_n + c0;
                if (col0 >= p.Cout) continue;                // warp-uniform
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int col = col0 + j;
                    v[j] = __uint_as_float(r[j]) + ((col < p.Cout) ? __ldg(p.bias + col) : 0.f);
                }
                if (p.out_mode == kOutF32) {
                    if (row_ok) {
                        float* dst = p.out_f32 + grow * p.ld + col0;
                        if (p.residual) {
                            const float4* res = reinterpret_cast<const float4*>(p.residual + grow * p.ld + col0);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 rr = __ldg(res + j);
                                v[4 * j] += rr.x; v[4 * j + 1] += rr.y; v[4 * j + 2] += rr.z; v[4 * j + 3] += rr.w;
                            }
                        }
                        if (p.act_silu) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            reinterpret_cast<float4*>(dst)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                } else if (p.out_mode == kOutBF16) {
                    if (p.act_silu) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
                    }
                    if (row_ok) {
                        if (col0 < p.split_col) {
                            uint4* dst = reinterpret_cast<uint4*>(p.out_bf16 + grow * p.ld + col0);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                dst[j] = make_uint4(pack_16(v[8 * j], v[8 * j + 1], p.f16), pack_16(v[8 * j + 2], v[8 * j + 3], p.f16),
                                                    pack_16(v[8 * j + 4], v[8 * j + 5], p.f16), pack_16(v[8 * j + 6], v[8 * j + 7], p.f16));
                        } else {
                            const long long img = grow / p.HW;
                            const int pix = static_cast<int>(grow - img * p.HW);
                            h16* dst = p.out_t + (img * (p.Cout - p.split_col) + (col0 - p.split_col)) * p.HW + pix;
#pragma unroll
                            for (int j = 0; j < 32; ++j) dst[static_cast<long long>(j) * p.HW] = cvt_16(v[j], p.f16);
                        }
                    }
                } else {   // kOutNCHW
                    if (row_ok) {
                        const long long img = grow / p.HW;
                        const int pix = static_cast<int>(grow - img * p.HW);
                        float* dst = p.out_f32 + (img * p.Cout + col0) * p.HW + pix;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < p.Cout) dst[static_cast<long long>(j) * p.HW] = v[j];
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace

cudaError_t launch_conv_gemm(const ConvParams& p, int num_sms, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int total = p.num_m_tiles * p.num_n_tiles;
    if (total <= 0) return cudaSuccess;
    const int grid = total < num_sms ? total : num_sms;
    conv_gemm_kernel<<<grid, kThreads, kSmemBytes, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace vdt
