// bf16 implicit-GEMM convolutions on the 5th-generation tensor cores (tcgen05 + TMEM), fed by TMA.
//
//   tag_conv_tc_fwd   : y[p, co] = sum_{tap,ci} x[p + d(tap), ci] * w[co][tap][ci]      (fwd and dgrad)
//   tag_conv_tc_wgrad : dw[co][tap][ci] += sum_p dy[p, co] * x[p + d(tap), ci]
//
// Replaces the cuDNN conv fwd / dgrad / wgrad behind F.conv2d (reference models/panns.py:49-50)
// and, with taps = 1, the cuBLAS GEMMs of fc1 and the GRU input projection
// (models/audio_encoder.py:216-217).
//
// Forward tile: 128 output pixels (a TH x W patch of one clip, TH = 128 / W) x BLOCK_N output
// channels.  For every (tap, 64-channel slab) the A operand is ONE 4-D TMA box of the NHWC
// activation tensor whose start coordinate is shifted by the tap offset — the halo is produced
// by TMA out-of-bounds zero fill, so there is no im2col buffer and no padding pass.  The B
// operand is a 2-D box of the [Cout][tap*Cin] weight matrix.  Both land in 128B-swizzled
// K-major shared memory and are consumed by tcgen05.mma (M=128, N=BLOCK_N, K=16) with the fp32
// accumulator in TMEM (double-buffered: the epilogue of tile i overlaps the MMAs of tile i+1).
// Warp roles: warps 0-3 epilogue (TMEM -> registers -> bias/ReLU/round -> per-channel sum and
// sum-of-squares for the following BatchNorm -> global), warp 4 TMA producer, warp 5 MMA issuer.
//
// wgrad tile: M = 128 rows of (tap, ci) — two 64-channel boxes of x, shifted by their tap —
// x N = BLOCK_N output channels of dy, contracted over pixels.  Both operands are MN-major
// (channels contiguous, pixels strided), which is exactly how NHWC boxes land in shared memory.
#include "tc_common.cuh"
#include <cstdlib>

namespace {

constexpr int A_TILE_BYTES = 128 * 128;          // 128 pixels x 64 bf16 channels

template <int BLOCK_N, int STAGES>
struct FwdSmem {
    static constexpr int B_TILE_BYTES = BLOCK_N * 128;
    static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int STATS_OFFSET = BAR_OFFSET + 256;
    static constexpr int TOTAL = STATS_OFFSET + 2 * BLOCK_N * 4 + 1024;   // +1024: manual alignment slack
};

template <int BLOCK_N, typename TO, int STAGES>
__global__ void __launch_bounds__(192, 1)
conv_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                   TO* __restrict__ y, const float* __restrict__ bias, int relu, double* __restrict__ stats,
                   int B, int H, int W, int Cin, int Cout, int taps) {
    using L = FwdSmem<BLOCK_N, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - raw);
    const uint32_t full_bar = base + L::BAR_OFFSET;             // STAGES x 8 B
    const uint32_t empty_bar = full_bar + 8 * STAGES;           // STAGES x 8 B
    const uint32_t tmem_full = empty_bar + 8 * STAGES;          // 2 x 8 B
    const uint32_t tmem_empty = tmem_full + 16;                 // 2 x 8 B
    const uint32_t tmem_slot = tmem_empty + 16;                 // 4 B
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + L::BAR_OFFSET + 16 * STAGES + 32);
    float* s_stats = reinterpret_cast<float*>(base_ptr + L::STATS_OFFSET);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TH = 128 / W;
    const int tiles_h = (H + TH - 1) / TH;
    const int m_tiles = B * tiles_h;
    const int n_tiles = Cout / BLOCK_N;
    const int total_tiles = m_tiles * n_tiles;
    const int KC = Cin / 64;
    const int k_blocks = taps * KC;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tmem_full + 8 * a, 1); mbar_init(tmem_empty + 8 * a, 4); }
        fence_barrier_init();
    }
    if (warp == 4 && lane == 0) { prefetch_tmap(&tmap_x); prefetch_tmap(&tmap_w); }
    if (warp == 5) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x < 2 * BLOCK_N && threadIdx.x < 128) {
        for (int i = threadIdx.x; i < 2 * BLOCK_N; i += 128) s_stats[i] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (elect_one_sync()) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n_tile = tile / m_tiles, m_tile = tile - n_tile * m_tiles;
                const int b = m_tile / tiles_h, h0 = (m_tile - b * tiles_h) * TH;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    const int tap = kb / KC, kc = kb - tap * KC;
                    const int dh = taps == 9 ? tap / 3 - 1 : 0;
                    const int dw = taps == 9 ? tap % 3 - 1 : 0;
                    mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    mbar_arrive_expect_tx(full_bar + 8 * stage, L::STAGE_BYTES);
                    const uint32_t sa = base + stage * L::STAGE_BYTES;
                    tma_load_4d(sa, &tmap_x, full_bar + 8 * stage, kc * 64, dw, h0 + dh, b);
                    tma_load_2d(sa + A_TILE_BYTES, &tmap_w, full_bar + 8 * stage, kb * 64, n_tile * BLOCK_N);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc(128, BLOCK_N, 0, 0);
        int stage = 0; uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(tmem_empty + 8 * as, aphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * BLOCK_N;
            for (int kb = 0; kb < k_blocks; ++kb) {
                mbar_wait(full_bar + 8 * stage, phase);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t sa = base + stage * L::STAGE_BYTES;
                    const uint64_t adesc = make_smem_desc(sa, 16, 1024);
                    const uint64_t bdesc = make_smem_desc(sa + A_TILE_BYTES, 16, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(empty_bar + 8 * stage);          // frees the smem slot when the MMAs retire
                    if (kb == k_blocks - 1) umma_commit(tmem_full + 8 * as);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 0-3; warp w owns TMEM lanes 32w..32w+31) =====================
        int it = 0;
        int cur_n_tile = -1;
        auto flush_stats = [&](int n_tile) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i = threadIdx.x; i < BLOCK_N; i += 128) {
                atomicAdd(stats + n_tile * BLOCK_N + i, (double)s_stats[i]);
                atomicAdd(stats + Cout + n_tile * BLOCK_N + i, (double)s_stats[BLOCK_N + i]);
                s_stats[i] = 0.f;
                s_stats[BLOCK_N + i] = 0.f;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        };
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int n_tile = tile / m_tiles, m_tile = tile - n_tile * m_tiles;
            const int b = m_tile / tiles_h, h0 = (m_tile - b * tiles_h) * TH;
            if (stats != nullptr && cur_n_tile >= 0 && n_tile != cur_n_tile) flush_stats(cur_n_tile);
            cur_n_tile = n_tile;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(tmem_full + 8 * as, aphase);
            tc_fence_after();
            const int row = warp * 32 + lane;
            const int h = h0 + row / W, w = row - (row / W) * W;
            const bool valid = h < H;
            TO* yrow = y + (((long)b * H + h) * W + w) * Cout + n_tile * BLOCK_N;
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + as * BLOCK_N + c * 32, r);
                tmem_ld_wait();
                float v[32], q[32];
                float bv[32];
                if (bias != nullptr) {
                    // 8 x 16-byte loads (the same addresses in every lane: one L1 transaction each) instead of 32 scalar ones
                    const float4* bp = reinterpret_cast<const float4*>(bias + n_tile * BLOCK_N + c * 32);
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 t4 = __ldg(bp + j4);
                        bv[4 * j4] = t4.x; bv[4 * j4 + 1] = t4.y; bv[4 * j4 + 2] = t4.z; bv[4 * j4 + 3] = t4.w;
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float f = __uint_as_float(r[j]);
                    if (bias != nullptr) f += bv[j];
                    if (relu) f = fmaxf(f, 0.f);
                    f = round_to<TO>(f);
                    v[j] = valid ? f : 0.f;
                }
                if (valid) {
                    // 256-bit stores: whole 32-byte sectors per lane and instruction
                    if (sizeof(TO) == 2) {
                        uint4 pk[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) { pack4<TO>(pk[j], 0, &v[8 * j]); pack4<TO>(pk[j], 1, &v[8 * j + 4]); }
                        st32(yrow + c * 32, pk[0], pk[1]);
                        st32(yrow + c * 32 + 16, pk[2], pk[3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint4 a4, b4;
                            pack4<TO>(a4, 0, &v[j]); pack4<TO>(b4, 0, &v[j + 4]);
                            st32(yrow + c * 32 + j, a4, b4);
                        }
                    }
                }
                if (stats != nullptr) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) q[j] = v[j] * v[j];
                    const float cs = warp_transpose_sum(v, lane);
                    const float cq = warp_transpose_sum(q, lane);
                    atomicAdd(&s_stats[c * 32 + lane], cs);
                    atomicAdd(&s_stats[BLOCK_N + c * 32 + lane], cq);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + 8 * as);
        }
        if (stats != nullptr && cur_n_tile >= 0) flush_stats(cur_n_tile);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 5) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
// wgrad: D[(tap,ci) 128, co BLOCK_N] = sum over 64-pixel K tiles; both operands MN-major.
// ---------------------------------------------------------------------------------------------
// PIX = pixels (K extent) per pipeline stage: narrow N tiles get more pixels per stage so that every
// barrier round trip of the single MMA-issuing thread covers ~512 cycles of tensor work.
template <int BLOCK_N, int STAGES, int PIX>
struct WgSmem {
    static constexpr int BOX_BYTES = PIX * 128;      // PIX pixels x 64 bf16 channels
    static constexpr int A_BYTES = 2 * BOX_BYTES;
    static constexpr int B_BYTES = (BLOCK_N / 64) * BOX_BYTES;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
};

template <int BLOCK_N, int STAGES, int PIX>
__global__ void __launch_bounds__(192, 1)
conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                     float* __restrict__ dw, int B, int H, int W, int Cin, int Cout, int taps,
                     int m_blocks, int n_blocks, int k_tiles_per_split, long slab_stride) {
    using L = WgSmem<BLOCK_N, STAGES, PIX>;
    constexpr int WG_BOX_BYTES = L::BOX_BYTES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - raw);
    const uint32_t full_bar = base + L::BAR_OFFSET;
    const uint32_t empty_bar = full_bar + 8 * STAGES;
    const uint32_t tmem_full = empty_bar + 8 * STAGES;
    const uint32_t tmem_slot = tmem_full + 8;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + L::BAR_OFFSET + 16 * STAGES + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int THK = PIX / W > 0 ? PIX / W : 1;      // rows of one clip per PIX-pixel K tile
    const int tiles_h = (H + THK - 1) / THK;
    const int k_tiles = B * tiles_h;

    int wi = blockIdx.x;
    const int m_blk = wi % m_blocks; wi /= m_blocks;
    const int n_blk = wi % n_blocks; wi /= n_blocks;
    const int split = wi;
    const int kt_begin = split * k_tiles_per_split;
    const int kt_end = min(k_tiles, kt_begin + k_tiles_per_split);
    if (kt_begin >= kt_end) return;

    // the two 64-row halves of the M tile: (tap, ci0) each
    int tap_h[2], ci_h[2];
    if (Cin >= 128) {
        const int per_tap = Cin / 128;
        tap_h[0] = tap_h[1] = m_blk / per_tap;
        ci_h[0] = (m_blk % per_tap) * 128;
        ci_h[1] = ci_h[0] + 64;
    } else {                                         // Cin == 64: two taps share one M tile
        tap_h[0] = 2 * m_blk;
        tap_h[1] = min(2 * m_blk + 1, taps - 1);
        ci_h[0] = ci_h[1] = 0;
    }
    const bool half1_live = (Cin >= 128) || (2 * m_blk + 1 < taps);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 4 && lane == 0) { prefetch_tmap(&tmap_x); prefetch_tmap(&tmap_dy); }
    if (warp == 5) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 4) {
        if (elect_one_sync()) {
            int stage = 0; uint32_t phase = 0;
            for (int kt = kt_begin; kt < kt_end; ++kt) {
                const int b = kt / tiles_h, h0 = (kt - b * tiles_h) * THK;
                mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                mbar_arrive_expect_tx(full_bar + 8 * stage, L::STAGE_BYTES);
                const uint32_t sa = base + stage * L::STAGE_BYTES;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int dh = taps == 9 ? tap_h[hh] / 3 - 1 : 0;
                    const int dw = taps == 9 ? tap_h[hh] % 3 - 1 : 0;
                    tma_load_4d(sa + hh * WG_BOX_BYTES, &tmap_x, full_bar + 8 * stage, ci_h[hh], dw, h0 + dh, b);
                }
#pragma unroll
                for (int j = 0; j < BLOCK_N / 64; ++j)
                    tma_load_4d(sa + L::A_BYTES + j * WG_BOX_BYTES, &tmap_dy, full_bar + 8 * stage,
                                n_blk * BLOCK_N + j * 64, 0, h0, b);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 5) {
        constexpr uint32_t idesc = make_idesc(128, BLOCK_N, 1, 1);
        int stage = 0; uint32_t phase = 0;
        for (int kt = kt_begin; kt < kt_end; ++kt) {
            mbar_wait(full_bar + 8 * stage, phase);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint32_t sa = base + stage * L::STAGE_BYTES;
                // MN-major SWIZZLE_128B: LBO = distance between 64-channel boxes, SBO = 8 pixel rows
                const uint64_t adesc = make_smem_desc(sa, WG_BOX_BYTES, 1024);
                const uint64_t bdesc = make_smem_desc(sa + L::A_BYTES, WG_BOX_BYTES, 1024);
#pragma unroll
                for (int k = 0; k < PIX / 16; ++k)   // 16 pixels = 2 swizzle atoms = 2048 B per K step
                    umma_bf16(tmem_base, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc,
                              (kt > kt_begin || k > 0) ? 1u : 0u);
                umma_commit(empty_bar + 8 * stage);
                if (kt == kt_end - 1) umma_commit(tmem_full);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else {
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int row = warp * 32 + lane;
        const int half = row >> 6;
        const int tap = tap_h[half];
        const int ci = ci_h[half] + (row & 63);
        const bool live = half == 0 || half1_live;
        const bool slab = slab_stride != 0;
        float* dwp = dw + (long)split * slab_stride;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32, r);
            tmem_ld_wait();
            if (live) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int co = n_blk * BLOCK_N + c * 32 + j;
                    wg_put(dwp + ((long)co * taps + tap) * Cin + ci, __uint_as_float(r[j]), slab);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 5) tmem_dealloc(tmem_base, 256);
}

// ---------------------------------------------------------------------------------------------
// weight preparation: fp32 master [Co][T][Ci] -> bf16 flipped + transposed [Ci][T][Co] (dgrad operand)
// ---------------------------------------------------------------------------------------------
__global__ void weight_flip_transpose_bf16_kernel(const float* __restrict__ w, bf16* __restrict__ wt, int Co,
                                                  int Ci, int T) {
    const long n = (long)Co * Ci * T;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        int co = (int)(i % Co);
        long r = i / Co;
        int t = (int)(r % T);
        int ci = (int)(r / T);
        wt[i] = __float2bfloat16_rn(w[((long)co * T + (T - 1 - t)) * Ci + ci]);
    }
}

// ---------------------------------------------------------------------------------------------

template <int BLOCK_N, typename TO>
int launch_tc_fwd(const CUtensorMap& tx, const CUtensorMap& tw, void* y, const float* bias, int relu,
                  double* stats, int B, int H, int W, int Cin, int Cout, int taps, cudaStream_t stream) {
    constexpr int STAGES = BLOCK_N == 256 ? 4 : (BLOCK_N == 128 ? 6 : 8);
    using L = FwdSmem<BLOCK_N, STAGES>;
    auto kern = conv_tc_fwd_kernel<BLOCK_N, TO, STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int TH = 128 / W;
    const int total_tiles = B * ((H + TH - 1) / TH) * (Cout / BLOCK_N);
    const int grid = total_tiles < sm_count() ? total_tiles : sm_count();
    kern<<<grid, 192, L::TOTAL, stream>>>(tx, tw, (TO*)y, bias, relu, stats, B, H, W, Cin, Cout, taps);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

// CTA-pair variant (cta_group::2, M = 256): the two CTAs of a cluster take two adjacent 128-row M blocks — (tap, ci 0..127)
// and (tap, ci 128..255) for Cin >= 256, two taps for Cin = 128 (an odd count pads the last pair) — against the SAME dy tile, of
// which each CTA loads half the channels: the dy stream L2 -> shared memory, its shared-memory reads per MMA and the MMA
// instruction count halve.  Barriers as in conv_tc_halo2.cu (loads of both CTAs complete on the leader, multicast commits).
template <int BLOCK_N, int STAGES, int PIX>
struct Wg2Smem {
    static constexpr int BOX_BYTES = PIX * 128;
    static constexpr int A_BYTES = 2 * BOX_BYTES;
    static constexpr int B_BYTES = (BLOCK_N / 128) * BOX_BYTES;      // this CTA's half of the dy tile
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
    static_assert(BLOCK_N % 128 == 0, "the pair splits the dy tile in whole 64-channel boxes");
};

template <int BLOCK_N, int STAGES, int PIX>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
conv_tc_wgrad2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                      float* __restrict__ dw, int B, int H, int W, int Cin, int Cout, int taps,
                      int m_blocks, int n_blocks, int k_tiles_per_split, long slab_stride) {
    using L = Wg2Smem<BLOCK_N, STAGES, PIX>;
    constexpr int WG_BOX_BYTES = L::BOX_BYTES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - raw);
    const uint32_t full_bar = base + L::BAR_OFFSET;
    const uint32_t empty_bar = full_bar + 8 * STAGES;
    const uint32_t tmem_full = empty_bar + 8 * STAGES;
    const uint32_t tmem_slot = tmem_full + 8;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + L::BAR_OFFSET + 16 * STAGES + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int THK = PIX / W > 0 ? PIX / W : 1;
    const int tiles_h = (H + THK - 1) / THK;
    const int k_tiles = B * tiles_h;
    const int m_pairs = (m_blocks + 1) / 2;

    int wi = blockIdx.x >> 1;
    const int m_pair = wi % m_pairs; wi /= m_pairs;
    const int n_blk = wi % n_blocks; wi /= n_blocks;
    const int split = wi;
    const int kt_begin = split * k_tiles_per_split;
    const int kt_end = min(k_tiles, kt_begin + k_tiles_per_split);      // >= kt_begin + 1 (host: splits re-derived from kps)

    const int m_blk_raw = 2 * m_pair + (int)rank;
    const bool blk_live = m_blk_raw < m_blocks;                         // the padding member of an odd count
    const int m_blk = blk_live ? m_blk_raw : m_blocks - 1;
    const int per_tap = Cin / 128;
    const int tap = m_blk / per_tap;
    const int ci0 = (m_blk % per_tap) * 128;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 4 && lane == 0) { prefetch_tmap(&tmap_x); prefetch_tmap(&tmap_dy); }
    if (warp == 5) tmem_alloc_2sm(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 4) {
        if (elect_one_sync()) {
            const uint32_t lead_full = mapa_u32(full_bar, 0);
            int stage = 0; uint32_t phase = 0;
            const int dh = taps == 9 ? tap / 3 - 1 : 0;
            const int dwv = taps == 9 ? tap % 3 - 1 : 0;
            for (int kt = kt_begin; kt < kt_end; ++kt) {
                const int b = kt / tiles_h, h0 = (kt - b * tiles_h) * THK;
                mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                if (rank == 0) mbar_arrive_expect_tx(full_bar + 8 * stage, 2 * L::STAGE_BYTES);
                const uint32_t sa = base + stage * L::STAGE_BYTES;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh)
                    tma_load_4d_2sm(sa + hh * WG_BOX_BYTES, &tmap_x, lead_full + 8 * stage, ci0 + 64 * hh, dwv, h0 + dh, b);
#pragma unroll
                for (int j = 0; j < BLOCK_N / 128; ++j)
                    tma_load_4d_2sm(sa + L::A_BYTES + j * WG_BOX_BYTES, &tmap_dy, lead_full + 8 * stage,
                                    n_blk * BLOCK_N + (int)rank * (BLOCK_N / 2) + j * 64, 0, h0, b);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 5 && rank == 0) {
        constexpr uint32_t idesc = make_idesc(256, BLOCK_N, 1, 1);
        int stage = 0; uint32_t phase = 0;
        for (int kt = kt_begin; kt < kt_end; ++kt) {
            mbar_wait(full_bar + 8 * stage, phase);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint32_t sa = base + stage * L::STAGE_BYTES;
                const uint64_t adesc = make_smem_desc(sa, WG_BOX_BYTES, 1024);
                const uint64_t bdesc = make_smem_desc(sa + L::A_BYTES, WG_BOX_BYTES, 1024);
#pragma unroll
                for (int k = 0; k < PIX / 16; ++k)
                    umma_bf16_2sm(tmem_base, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc,
                                  (kt > kt_begin || k > 0) ? 1u : 0u);
                umma_commit_2sm(empty_bar + 8 * stage, 3);
                if (kt == kt_end - 1) umma_commit_2sm(tmem_full, 3);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp < 4) {
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int row = warp * 32 + lane;
        const int ci = ci0 + row;
        const bool slab = slab_stride != 0;
        float* dwp = dw + (long)split * slab_stride;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32, r);
            tmem_ld_wait();
            if (blk_live) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int co = n_blk * BLOCK_N + c * 32 + j;
                    wg_put(dwp + ((long)co * taps + tap) * Cin + ci, __uint_as_float(r[j]), slab);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    if (warp == 5) tmem_dealloc_2sm(tmem_base, 256);
}

template <int BLOCK_N, int PIX>
int launch_tc_wgrad2(const CUtensorMap& tx, const CUtensorMap& tdy, float* dw, int B, int H, int W, int Cin,
                     int Cout, int taps, int splits, cudaStream_t stream) {
    constexpr int STAGE_BYTES = (2 + BLOCK_N / 128) * PIX * 128;
    constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
    using L = Wg2Smem<BLOCK_N, STAGES, PIX>;
    auto kern = conv_tc_wgrad2_kernel<BLOCK_N, STAGES, PIX>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int THK = PIX / W > 0 ? PIX / W : 1;
    const int k_tiles = B * ((H + THK - 1) / THK);
    const int m_blocks = taps * (Cin / 128);
    const int m_pairs = (m_blocks + 1) / 2;
    const int n_blocks = Cout / BLOCK_N;
    if (splits > k_tiles) splits = k_tiles;
    const int kps = (k_tiles + splits - 1) / splits;
    splits = (k_tiles + kps - 1) / kps;
    const long grid = 2l * m_pairs * n_blocks * splits;
    const long n = (long)Cout * taps * Cin;
    float* target; long slab;
    int rc = splitk_target(dw, splits, n, stream, &target, &slab);
    if (rc != TAG_OK) return rc;
    kern<<<(unsigned)grid, 192, L::TOTAL, stream>>>(tx, tdy, target, B, H, W, Cin, Cout, taps, m_blocks, n_blocks, kps, slab);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return slab ? tag_splitk_reduce(target, splits, n, dw, stream) : TAG_OK;
}

template <int BLOCK_N, int PIX>
int launch_tc_wgrad(const CUtensorMap& tx, const CUtensorMap& tdy, float* dw, int B, int H, int W, int Cin,
                    int Cout, int taps, int splits, cudaStream_t stream) {
    constexpr int STAGE_BYTES = (2 + BLOCK_N / 64) * PIX * 128;
    constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
    using L = WgSmem<BLOCK_N, STAGES, PIX>;
    auto kern = conv_tc_wgrad_kernel<BLOCK_N, STAGES, PIX>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int THK = PIX / W > 0 ? PIX / W : 1;
    const int k_tiles = B * ((H + THK - 1) / THK);
    const int m_blocks = Cin >= 128 ? taps * (Cin / 128) : (taps + 1) / 2;
    const int n_blocks = Cout / BLOCK_N;
    if (splits > k_tiles) splits = k_tiles;
    const int kps = (k_tiles + splits - 1) / splits;
    splits = (k_tiles + kps - 1) / kps;
    const long grid = (long)m_blocks * n_blocks * splits;
    const long n = (long)Cout * taps * Cin;
    float* target; long slab;
    int rc = splitk_target(dw, splits, n, stream, &target, &slab);
    if (rc != TAG_OK) return rc;
    kern<<<(unsigned)grid, 192, L::TOTAL, stream>>>(tx, tdy, target, B, H, W, Cin, Cout, taps, m_blocks, n_blocks, kps, slab);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return slab ? tag_splitk_reduce(target, splits, n, dw, stream) : TAG_OK;
}

bool width_ok(int W) { return W == 1 || W == 2 || W == 4 || W == 8 || W == 16 || W == 32 || W == 64; }

}  // namespace

// x: bf16 NHWC [B,H,W,Cin]; w: bf16 [Cout][taps*Cin]; y: NHWC [B,H,W,Cout] (bf16 or fp32)
extern "C" int tag_conv_tc_fwd(const void* x, const void* w, void* y, int y_dtype, const float* bias, int relu,
                               double* stats, int B, int H, int W, int Cin, int Cout, int taps,
                               cudaStream_t stream) {
    if ((taps != 1 && taps != 9) || Cin % 64 != 0 || Cout % 64 != 0 || !width_ok(W) || B <= 0 || H <= 0)
        return TAG_ERR_BAD_ARG;
    const int TH = 128 / W;
    if (TH > 256) return TAG_ERR_BAD_ARG;
    // widest N tile that still gives about one tile per SM: a 320-row GEMM (the CLAP text tower) on 256-wide tiles
    // would run on 9 CTAs, each walking the whole K loop at single-SM TMA speed
    int block_n = Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64);
    {
        const long m_tiles = (long)B * ((H + TH - 1) / TH);
        while (block_n > 64 && m_tiles * (Cout / block_n) < 111) block_n /= 2;
    }
    CUtensorMap tx, tw;
    int rc = make_act_tmap(&tx, x, B, H, W, Cin, W, TH);
    if (rc != TAG_OK) return rc;
    rc = make_w_tmap(&tw, w, Cout, taps * Cin, block_n);
    if (rc != TAG_OK) return rc;
#define TAG_TC_FWD(BN_)                                                                                       \
    (y_dtype == TAG_DTYPE_BF16                                                                               \
         ? launch_tc_fwd<BN_, bf16>(tx, tw, y, bias, relu, stats, B, H, W, Cin, Cout, taps, stream)          \
         : launch_tc_fwd<BN_, float>(tx, tw, y, bias, relu, stats, B, H, W, Cin, Cout, taps, stream))
    if (block_n == 256) return TAG_TC_FWD(256);
    if (block_n == 128) return TAG_TC_FWD(128);
    return TAG_TC_FWD(64);
#undef TAG_TC_FWD
}

// dy: bf16 NHWC [B,H,W,Cout]; x: bf16 NHWC [B,H,W,Cin]; dw: fp32 [Cout][taps][Cin], accumulated (+=)
extern "C" int tag_conv_tc_wgrad(const void* dy, const void* x, float* dw, int B, int H, int W, int Cin,
                                 int Cout, int taps, int splits, cudaStream_t stream) {
    if ((taps != 1 && taps != 9) || Cin % 64 != 0 || Cout % 64 != 0 || !width_ok(W) || splits <= 0)
        return TAG_ERR_BAD_ARG;
    if (taps == 1 && Cin < 128) return TAG_ERR_BAD_ARG;
    if (W > 64) return TAG_ERR_BAD_ARG;
    const int block_n = Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64);
    // pixels per stage: 64 / 128 / 256 for N = 256 / 128 / 64 (3x3 convs; the GEMM form keeps 64)
    int pix = block_n == 256 ? 64 : (block_n == 128 ? 128 : 256);
    if (taps == 1 || W * 256 < pix) pix = 64;
    const int box_w = W < pix ? W : pix;
    const int THK = pix / box_w;
    if (THK > 256) return TAG_ERR_BAD_ARG;
    CUtensorMap tx, tdy;
    int rc = make_act_tmap(&tx, x, B, H, W, Cin, box_w, THK);
    if (rc != TAG_OK) return rc;
    rc = make_act_tmap(&tdy, dy, B, H, W, Cout, box_w, THK);
    if (rc != TAG_OK) return rc;
    static const bool pair_ok = getenv("TAG_B200_NO_WGRAD_PAIR") == nullptr;
    // measured (scripts/bench_conv.py wgrad): pairs win where two M blocks share a tap (Cin >= 256: 1 446 -> 1 620, 1 432 ->
    // 1 637, 1 544 -> 1 685 TFLOP/s); with Cin = 128 a pair spans two taps and the odd tap count pads a tenth of the grid
    // (1 134 -> 970), and the one-tap GEMMs of the GRU are launch-bound either way
    if (pair_ok && taps == 9 && Cin >= 256 && block_n >= 128) {
#define TAG_WG2(BN_, PIX_) launch_tc_wgrad2<BN_, PIX_>(tx, tdy, dw, B, H, W, Cin, Cout, taps, splits, stream)
        if (block_n == 256) return TAG_WG2(256, 64);
        return pix == 128 ? TAG_WG2(128, 128) : TAG_WG2(128, 64);
#undef TAG_WG2
    }
#define TAG_WG(BN_, PIX_) launch_tc_wgrad<BN_, PIX_>(tx, tdy, dw, B, H, W, Cin, Cout, taps, splits, stream)
    if (block_n == 256) return TAG_WG(256, 64);
    if (block_n == 128) return pix == 128 ? TAG_WG(128, 128) : TAG_WG(128, 64);
    return pix == 256 ? TAG_WG(64, 256) : TAG_WG(64, 64);
#undef TAG_WG
}

extern "C" int tag_weight_flip_transpose_bf16(const float* w, void* wt, int Co, int Ci, int taps,
                                              cudaStream_t stream) {
    const long n = (long)Co * Ci * taps;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 4096) blocks = 4096;
    weight_flip_transpose_bf16_kernel<<<blocks, 256, 0, stream>>>(w, (bf16*)wt, Co, Ci, taps);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}
