// 3x3 convolution forward / dgrad on tcgen05, CTA-PAIR variant of conv_tc_halo.cu (cta_group::2).
//
// Two CTAs of a cluster (the two SMs of one TPC) compute one 256-pixel x BLOCK_N output tile with M = 256 MMAs issued
// by the leader CTA alone: each CTA streams the halo boxes of ITS 16 x 8-pixel half (same box re-use across the three
// vertical taps as in conv_tc_halo.cu) but only HALF of every weight tile (BLOCK_N / 2 output channels) — the tensor
// cores of the pair read the two halves from both shared memories.  Against the one-CTA kernel this halves the weight
// traffic L2 -> shared memory (the dominant stream of the layers with Cin >= 128: 9 * Cin * BLOCK_N * 2 bytes per
// tile against 3.4 * Cin * 256 bytes of input), halves the shared-memory reads of B per MMA and halves the number of
// MMA instructions.  Synchronisation: TMA loads of both CTAs complete on the LEADER's full barriers (expect_tx of
// both halves armed by the leader's producer), tcgen05.commit is multicast to the empty / accumulator-full barriers of
// both CTAs, and the epilogue warps of both CTAs arrive on the leader's accumulator-empty barrier (remote arrive).
#include "halo_epilogue.cuh"
#include <cstdlib>

namespace {

constexpr int A2_SUB_BYTES = (TILE_H + 2) * TILE_W * 128;      // 18432

template <int BLOCK_N, int A_STAGES, int B_STAGES, int TPS>
struct Halo2Smem {
    static constexpr int B_HALF_BYTES = TPS * (BLOCK_N / 2) * 128;
    static constexpr int A_OFFSET = 0;
    static constexpr int B_OFFSET = A_STAGES * A2_SUB_BYTES;
    static constexpr int BAR_OFFSET = B_OFFSET + B_STAGES * B_HALF_BYTES;
    static constexpr int STATS_OFFSET = BAR_OFFSET + 512;
    static constexpr int TBUF_OFFSET = STATS_OFFSET + 2 * BLOCK_N * 4;
    static constexpr int TOTAL = TBUF_OFFSET + 8 * BLOCK_N * 4 + 1024;
    static_assert(8 * (2 * A_STAGES + 2 * B_STAGES + 4) + 8 <= 512, "barrier block");
    static_assert(B_HALF_BYTES % 1024 == 0, "weight half tiles must keep the 1024 B swizzle alignment");
    static_assert(TOTAL <= 227 * 1024, "shared memory budget");
};

template <int BLOCK_N, typename TO, int A_STAGES, int B_STAGES, int TPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
conv_tc_fwd_halo2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                         TO* __restrict__ y, double* __restrict__ stats, int B, int H, int W, int Cin, int Cout,
                         const bf16* __restrict__ bn_y, const uint8_t* __restrict__ pool_cnt, int dbg) {
    using L = Halo2Smem<BLOCK_N, A_STAGES, B_STAGES, TPS>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - raw);
    const uint32_t a_full = base + L::BAR_OFFSET;
    const uint32_t a_empty = a_full + 8 * A_STAGES;
    const uint32_t b_full = a_empty + 8 * A_STAGES;
    const uint32_t b_empty = b_full + 8 * B_STAGES;
    const uint32_t tmem_full = b_empty + 8 * B_STAGES;
    const uint32_t tmem_empty = tmem_full + 16;
    const uint32_t tmem_slot = tmem_empty + 16;
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(base_ptr + L::BAR_OFFSET + 16 * A_STAGES + 16 * B_STAGES + 32);
    float* s_stats = reinterpret_cast<float*>(base_ptr + L::STATS_OFFSET);
    float* t_buf = reinterpret_cast<float*>(base_ptr + L::TBUF_OFFSET);

    const int warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_ctarank();                  // 0 = leader (issues the MMAs)
    const int tiles_w = W / TILE_W;
    const int tiles_h = (H + TILE_H - 1) / TILE_H;
    const int tiles_img = tiles_w * tiles_h;
    const int m_tiles = B * tiles_img;
    const int m_pairs = (m_tiles + 1) / 2;                    // an odd tile count pads the last pair (b == B)
    const int n_tiles = Cout / BLOCK_N;
    const int total_tiles = m_pairs * n_tiles;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int KC = Cin / 64;

    if (threadIdx.x == 0) {
        for (int s = 0; s < A_STAGES; ++s) { mbar_init(a_full + 8 * s, 1); mbar_init(a_empty + 8 * s, 1); }
        for (int s = 0; s < B_STAGES; ++s) { mbar_init(b_full + 8 * s, 1); mbar_init(b_empty + 8 * s, 1); }
        // accumulator-empty: the 8 epilogue warps of BOTH CTAs arrive on the leader's barrier
        for (int a = 0; a < 2; ++a) { mbar_init(tmem_full + 8 * a, 1); mbar_init(tmem_empty + 8 * a, 16); }
        fence_barrier_init();
    }
    if (warp == 8 && (threadIdx.x & 31) == 0) { prefetch_tmap(&tmap_x); prefetch_tmap(&tmap_w); }
    if (warp == 9) tmem_alloc_2sm(tmem_slot, 512);
    if (threadIdx.x < 256)
        for (int i = threadIdx.x; i < 2 * BLOCK_N; i += 256) s_stats[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();               // barriers initialised and TMEM allocated in BOTH CTAs before any remote signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    auto decode = [&](int tile, int& n_tile, int& b, int& h0, int& w0) {
        n_tile = tile / m_pairs;
        int m_tile = 2 * (tile - n_tile * m_pairs) + (int)rank;
        b = m_tile / tiles_img;                               // == B for the padding tile of an odd count
        m_tile -= b * tiles_img;
        const int th = m_tile / tiles_w;
        h0 = th * TILE_H;
        w0 = (m_tile - th * tiles_w) * TILE_W;
    };

    if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        if (warp == 8) {
            // ===================== TMA producer (both CTAs): own input boxes, own half of the weight tiles
            if (elect_one_sync()) {
                const uint32_t lead_a_full = mapa_u32(a_full, 0), lead_b_full = mapa_u32(b_full, 0);
                int as = 0, bs = 0; uint32_t aph = 0, bph = 0;
                for (int tile = pair; tile < total_tiles; tile += n_pairs) {
                    int n_tile, b, h0, w0;
                    decode(tile, n_tile, b, h0, w0);
                    for (int kc = 0; kc < KC; ++kc) {
                        for (int dwi = 0; dwi < 3; ++dwi) {
                            mbar_wait(a_empty + 8 * as, aph ^ 1);
                            if (rank == 0) mbar_arrive_expect_tx(a_full + 8 * as, 2 * A2_SUB_BYTES);
                            tma_load_4d_2sm(base + L::A_OFFSET + as * A2_SUB_BYTES, &tmap_x, lead_a_full + 8 * as,
                                            kc * 64, w0 + dwi - 1, h0 - 1, b);
                            if (++as == A_STAGES) { as = 0; aph ^= 1; }
                            for (int dhi = 0; dhi < 3; dhi += TPS) {
                                mbar_wait(b_empty + 8 * bs, bph ^ 1);
                                if (rank == 0) mbar_arrive_expect_tx(b_full + 8 * bs, 2 * L::B_HALF_BYTES);
                                tma_load_3d_2sm(base + L::B_OFFSET + bs * L::B_HALF_BYTES, &tmap_w,
                                                lead_b_full + 8 * bs, kc * 64,
                                                n_tile * BLOCK_N + (int)rank * (BLOCK_N / 2), dwi * 3 + dhi);
                                if (++bs == B_STAGES) { bs = 0; bph ^= 1; }
                            }
                        }
                    }
                }
            }
        } else if (warp == 9 && rank == 0) {
            // ===================== MMA issuer (leader only): M = 256 across the pair
            constexpr uint32_t idesc = make_idesc(256, BLOCK_N, 0, 0);
            int as = 0, bs = 0; uint32_t aph = 0, bph = 0;
            int it = 0;
            for (int tile = pair; tile < total_tiles; tile += n_pairs, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(tmem_empty + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                uint32_t first = 1;
                for (int kc = 0; kc < KC; ++kc) {
                    for (int dwi = 0; dwi < 3; ++dwi) {
                        mbar_wait(a_full + 8 * as, aph);
                        const uint32_t sa = base + L::A_OFFSET + as * A2_SUB_BYTES;
                        for (int dhi = 0; dhi < 3; dhi += TPS) {
                            mbar_wait(b_full + 8 * bs, bph);
                            tc_fence_after();
                            if (elect_one_sync()) {
#pragma unroll
                                for (int tt = 0; tt < TPS; ++tt) {
                                    const uint64_t adesc = make_smem_desc(sa + (dhi + tt) * (TILE_W * 128), 16, 1024);
                                    const uint64_t bdesc = make_smem_desc(
                                        base + L::B_OFFSET + bs * L::B_HALF_BYTES + tt * ((BLOCK_N / 2) * 128), 16, 1024);
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        umma_bf16_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, first ? 0u : 1u);
                                        first = 0;
                                    }
                                }
                                umma_commit_2sm(b_empty + 8 * bs, 3);
                                if (dhi + TPS >= 3) umma_commit_2sm(a_empty + 8 * as, 3);
                                if (kc == KC - 1 && dwi == 2 && dhi + TPS >= 3) umma_commit_2sm(tmem_full + 8 * acc, 3);
                            }
                            __syncwarp();
                            if (++bs == B_STAGES) { bs = 0; bph ^= 1; }
                        }
                        if (++as == A_STAGES) { as = 0; aph ^= 1; }
                    }
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        const uint32_t lead_tmem_empty = mapa_u32(tmem_empty, 0);
        halo_epilogue<BLOCK_N, TO>(tmem_base, tmem_full, t_buf, y, stats, B, H, W, Cout, bn_y, pool_cnt, dbg, pair, n_pairs,
                                   total_tiles, decode,
                                   [&](int acc) { mbar_arrive_cluster(lead_tmem_empty + 8 * acc); });
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();               // neither CTA may retire while its pair still reads its shared memory / barriers
    tc_fence_after();
    if (warp == 9) tmem_dealloc_2sm(tmem_base, 512);
}

template <int BLOCK_N, typename TO>
int launch_halo2(const CUtensorMap& tx, const CUtensorMap& tw, void* y, double* stats, int B, int H, int W, int Cin,
                 int Cout, const void* bn_y, const void* pool_cnt, cudaStream_t stream) {
    constexpr int TPS = BLOCK_N == 256 ? 1 : 3;
    constexpr int A_STAGES = BLOCK_N == 256 ? 5 : (BLOCK_N == 128 ? 6 : 8);
    constexpr int B_STAGES = BLOCK_N == 256 ? 6 : (BLOCK_N == 128 ? 4 : 5);
    using L = Halo2Smem<BLOCK_N, A_STAGES, B_STAGES, TPS>;
    auto kern = conv_tc_fwd_halo2_kernel<BLOCK_N, TO, A_STAGES, B_STAGES, TPS>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int m_tiles = B * ((H + TILE_H - 1) / TILE_H) * (W / TILE_W);
    const int total_pairs = ((m_tiles + 1) / 2) * (Cout / BLOCK_N);
    const int max_pairs = sm_count() / 2;
    const int grid = 2 * (total_pairs < max_pairs ? total_pairs : max_pairs);
    static int dbg = getenv("TAG_HALO_DBG") ? atoi(getenv("TAG_HALO_DBG")) : 0;
    kern<<<grid, 384, L::TOTAL, stream>>>(tx, tw, (TO*)y, stats, B, H, W, Cin, Cout, (const bf16*)bn_y, (const uint8_t*)pool_cnt, dbg);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

}  // namespace

// CTA-pair launch of the halo convolution; called by tag_conv_tc_fwd_halo (conv_tc_halo.cu) for the layers whose weights
// stream (Cin >= 128).  tw: tap-major weights [9][Cout][Cin] with box (64, block_n / 2, block_n == 256 ? 1 : 3).
int tag_halo2_dispatch(const CUtensorMap& tx, const CUtensorMap& tw, void* y, int y_dtype, double* stats, int B, int H,
                       int W, int Cin, int Cout, int block_n, const void* bn_y, const void* pool_cnt, cudaStream_t stream) {
#define TAG_HALO2(BN_)                                                                                     \
    (y_dtype == TAG_DTYPE_BF16 ? launch_halo2<BN_, bf16>(tx, tw, y, stats, B, H, W, Cin, Cout, bn_y, pool_cnt, stream) \
                               : launch_halo2<BN_, float>(tx, tw, y, stats, B, H, W, Cin, Cout, bn_y, pool_cnt, stream))
    if (block_n == 256) return TAG_HALO2(256);
    if (block_n == 128) return TAG_HALO2(128);
    return TAG_HALO2(64);
#undef TAG_HALO2
}
