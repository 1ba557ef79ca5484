// 3x3 convolution forward / dgrad on tcgen05 with the input halo RE-USED across taps.
//
// conv_tc.cu loads one shifted 128-pixel box per (tap, 64-channel slab): every input element goes
// through the L2 -> SM path nine times.  Here an output tile is 16 time rows x 8 frequency bins; for
// each horizontal shift dw in {-1,0,+1} ONE box of (16+2) rows x 8 columns x 64 channels lands in
// shared memory (18 KB, rows = pixels, 128 B each, SWIZZLE_128B) and serves the three vertical taps:
// the tap (dh, dw) A operand is the same buffer with its start address advanced by (dh+1) * 8 rows
// = 1024 B — exactly one swizzle atom, so the descriptor stays canonical (SBO = 1024 B: 8-pixel groups
// = consecutive image rows).  A traffic drops from 144 KB to 54 KB per tile and slab; the weight
// tiles stream through their own ring.  Same warp roles / TMEM double buffering / epilogue as
// conv_tc_fwd_kernel.
#include "halo_epilogue.cuh"
#include <cstdlib>

namespace {

constexpr int A_SUB_BYTES = (TILE_H + 2) * TILE_W * 128;      // 18432

// TPS = taps per weight stage: for narrow tiles (BLOCK_N <= 128) the three vertical taps of one
// horizontal shift travel as ONE 3-D TMA box of the tap-major weight tensor [tap'][Cout][Cin]
// (tap' = dwi*3 + dhi), so the single MMA-issuing thread gets 12 MMAs per barrier round trip.
// A_STAGES: an input box feeds only 12 MMAs (384 cycles at N = 64), so the ring must be deep enough in
// TIME to cover the ~2000-cycle L2/HBM latency of a TMA load: 8 boxes in flight for the narrow tiles.
template <int BLOCK_N, int B_STAGES, int TPS, bool WRES>
struct HaloSmem {
    static constexpr int A_STAGES = BLOCK_N == 64 ? (WRES ? 6 : 4) : (BLOCK_N == 128 ? 3 : 4);
    static_assert(true, "");
    static constexpr int B_TILE_BYTES = TPS * BLOCK_N * 128;
    static constexpr int A_OFFSET = 0;
    static constexpr int B_OFFSET = A_STAGES * A_SUB_BYTES;
    static constexpr int BAR_OFFSET = B_OFFSET + B_STAGES * B_TILE_BYTES;
    static constexpr int STATS_OFFSET = BAR_OFFSET + 512;
    static constexpr int TBUF_OFFSET = STATS_OFFSET + 2 * BLOCK_N * 4;       // 8 warps x BLOCK_N floats (flush)
    static constexpr int TOTAL = TBUF_OFFSET + 8 * BLOCK_N * 4 + 1024;
    static_assert(TOTAL <= 227 * 1024, "shared memory budget");
};

// WRES (Cin == 64, Cout == BLOCK_N <= 128): the whole 9-tap weight tensor (<= 144 KB) is loaded ONCE
// per CTA into the three weight stages and stays resident; only the input boxes stream.
template <int BLOCK_N, typename TO, int B_STAGES, int TPS, bool WRES>
__global__ void __launch_bounds__(384, 1)
conv_tc_fwd_halo_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                        TO* __restrict__ y, double* __restrict__ stats, int B, int H, int W, int Cin, int Cout,
                        const bf16* __restrict__ bn_y, const uint8_t* __restrict__ pool_cnt, int dbg) {
    using L = HaloSmem<BLOCK_N, B_STAGES, TPS, WRES>;
    constexpr int A_STAGES = L::A_STAGES;
    constexpr bool TWO_ISSUERS = WRES && BLOCK_N == 64 && A_STAGES == 6;      // see the MMA issuer section
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

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_w = W / TILE_W;
    const int tiles_h = (H + TILE_H - 1) / TILE_H;
    const int tiles_img = tiles_w * tiles_h;
    const int m_tiles = B * tiles_img;
    const int n_tiles = Cout / BLOCK_N;
    const int total_tiles = m_tiles * n_tiles;
    const int KC = Cin / 64;

    if (threadIdx.x == 0) {
        for (int s = 0; s < A_STAGES; ++s) { mbar_init(a_full + 8 * s, 1); mbar_init(a_empty + 8 * s, 1); }
        for (int s = 0; s < B_STAGES; ++s) { mbar_init(b_full + 8 * s, 1); mbar_init(b_empty + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tmem_full + 8 * a, 1); mbar_init(tmem_empty + 8 * a, 8); }
        fence_barrier_init();
    }
    if (warp == 8 && lane == 0) { prefetch_tmap(&tmap_x); prefetch_tmap(&tmap_w); }
    if (warp == 9) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x < 256)
        for (int i = threadIdx.x; i < 2 * BLOCK_N; i += 256) s_stats[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // 12 warps = 3 warpgroups: the two epilogue warpgroups take the registers the producer / MMA / idle warps
    // (warpgroup 2) do not need — the epilogue keeps 64 running statistics per thread on top of a 32-column chunk

    auto decode = [&](int tile, int& n_tile, int& b, int& h0, int& w0) {
        n_tile = tile / m_tiles;
        int m_tile = tile - n_tile * m_tiles;
        b = m_tile / tiles_img;
        m_tile -= b * tiles_img;
        const int th = m_tile / tiles_w;
        h0 = th * TILE_H;
        w0 = (m_tile - th * tiles_w) * TILE_W;
    };

    if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 8) {
        // ===================== TMA producer =====================
        if (elect_one_sync()) {
            int as = 0, bs = 0; uint32_t aph = 0, bph = 0;
            if (WRES) {
                for (int dwi = 0; dwi < 3; ++dwi) {
                    mbar_arrive_expect_tx(b_full + 8 * dwi, L::B_TILE_BYTES);
                    tma_load_3d(base + L::B_OFFSET + dwi * L::B_TILE_BYTES, &tmap_w, b_full + 8 * dwi, 0, 0, dwi * 3);
                }
            }
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int n_tile, b, h0, w0;
                decode(tile, n_tile, b, h0, w0);
                for (int kc = 0; kc < KC; ++kc) {
                    for (int dwi = 0; dwi < 3; ++dwi) {
                        mbar_wait(a_empty + 8 * as, aph ^ 1);
                        mbar_arrive_expect_tx(a_full + 8 * as, A_SUB_BYTES);
                        tma_load_4d(base + L::A_OFFSET + as * A_SUB_BYTES, &tmap_x, a_full + 8 * as, kc * 64,
                                    w0 + dwi - 1, h0 - 1, b);
                        if (++as == A_STAGES) { as = 0; aph ^= 1; }
                        for (int dhi = 0; dhi < 3 && !WRES; dhi += TPS) {
                            mbar_wait(b_empty + 8 * bs, bph ^ 1);
                            mbar_arrive_expect_tx(b_full + 8 * bs, L::B_TILE_BYTES);
                            tma_load_3d(base + L::B_OFFSET + bs * L::B_TILE_BYTES, &tmap_w, b_full + 8 * bs,
                                        kc * 64, n_tile * BLOCK_N, dwi * 3 + dhi);
                            if (++bs == B_STAGES) { bs = 0; bph ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 9 || (TWO_ISSUERS && warp == 10)) {
        // ===================== MMA issuer(s) =====================
        // At BLOCK_N = 64 an MMA occupies the tensor pipe for 32 cycles, less than one warp needs to build the
        // descriptors and issue it (~65 cycles measured): a single issuer left the pipe 60 % idle.  The weights-resident
        // BLOCK_N = 64 variant therefore runs TWO issuing warps, one per TMEM accumulator: warp 9 takes the even tiles
        // of this CTA, warp 10 the odd ones.  This is only sound because there A_STAGES == 2 * (boxes per tile): each
        // warp owns a fixed half of the input ring, so it never waits on a barrier whose previous phase belongs to
        // the other warp (mbarrier parity waits cannot tell phases two apart).
        constexpr uint32_t idesc = make_idesc(128, BLOCK_N, 0, 0);
        constexpr int NI = TWO_ISSUERS ? 2 : 1;
        const int mw = warp - 9;
        const uint32_t a_per_tile = (uint32_t)KC * 3u;
        const uint32_t b_per_tile = a_per_tile * (3 / TPS);
        int it = mw;
        for (int tile = blockIdx.x + mw * gridDim.x; tile < total_tiles; tile += NI * gridDim.x, it += NI) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const uint32_t na = (uint32_t)it * a_per_tile, nb = (uint32_t)it * b_per_tile;
            int as = (int)(na % A_STAGES), bs = (int)(nb % B_STAGES);
            uint32_t aph = (na / A_STAGES) & 1u, bph = (nb / B_STAGES) & 1u;
            mbar_wait(tmem_empty + 8 * acc, acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
            uint32_t first = 1;
            for (int kc = 0; kc < KC; ++kc) {
                for (int dwi = 0; dwi < 3; ++dwi) {
                    mbar_wait(a_full + 8 * as, aph);
                    const uint32_t sa = base + L::A_OFFSET + as * A_SUB_BYTES;
                    for (int dhi = 0; dhi < 3; dhi += TPS) {
                        if (WRES) { bs = dwi; bph = 0; }           // resident stage: completed once, stays readable
                        mbar_wait(b_full + 8 * bs, bph);
                        tc_fence_after();
                        if (elect_one_sync()) {
#pragma unroll
                            for (int tt = 0; tt < TPS; ++tt) {
                                const uint64_t adesc = make_smem_desc(sa + (dhi + tt) * (TILE_W * 128), 16, 1024);
                                const uint64_t bdesc = make_smem_desc(
                                    base + L::B_OFFSET + bs * L::B_TILE_BYTES + tt * (BLOCK_N * 128), 16, 1024);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, first ? 0u : 1u);
                                    first = 0;
                                }
                            }
                            if (!WRES) umma_commit(b_empty + 8 * bs);
                            if (dhi + TPS >= 3) umma_commit(a_empty + 8 * as);
                            if (kc == KC - 1 && dwi == 2 && dhi + TPS >= 3) umma_commit(tmem_full + 8 * acc);
                        }
                        __syncwarp();
                        if (!WRES && ++bs == B_STAGES) { bs = 0; bph ^= 1; }
                    }
                    if (++as == A_STAGES) { as = 0; aph ^= 1; }
                }
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        halo_epilogue<BLOCK_N, TO>(tmem_base, tmem_full, t_buf, y, stats, B, H, W, Cout, bn_y, pool_cnt, dbg,
                                   (int)blockIdx.x, (int)gridDim.x, total_tiles, decode,
                                   [&](int acc) { mbar_arrive(tmem_empty + 8 * acc); });
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 9) tmem_dealloc(tmem_base, 512);
}

template <int BLOCK_N, typename TO, bool WRES>
int launch_halo(const CUtensorMap& tx, const CUtensorMap& tw, void* y, double* stats, int B, int H, int W,
                int Cin, int Cout, const void* bn_y, const void* pool_cnt, cudaStream_t stream) {
    constexpr int TPS = BLOCK_N == 256 ? 1 : 3;
    constexpr int B_STAGES = WRES ? 3 : (BLOCK_N == 256 ? 4 : (BLOCK_N == 128 ? 3 : 5));
    using L = HaloSmem<BLOCK_N, B_STAGES, TPS, WRES>;
    auto kern = conv_tc_fwd_halo_kernel<BLOCK_N, TO, B_STAGES, TPS, WRES>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int total_tiles = B * ((H + TILE_H - 1) / TILE_H) * (W / TILE_W) * (Cout / BLOCK_N);
    const int grid = total_tiles < sm_count() ? total_tiles : sm_count();
    static int dbg = getenv("TAG_HALO_DBG") ? atoi(getenv("TAG_HALO_DBG")) : 0;
    kern<<<grid, 384, L::TOTAL, stream>>>(tx, tw, (TO*)y, stats, B, H, W, Cin, Cout, (const bf16*)bn_y, (const uint8_t*)pool_cnt, dbg);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

}  // namespace

namespace {

// fp32 master [Co][tap][Ci] -> bf16 tap-major operand of the halo kernel, tap' = dwi*3 + dhi:
//   flip_transpose = 0 (forward): out[tap'][co][ci] = w[co][dhi*3+dwi][ci]
//   flip_transpose = 1 (dgrad)  : out[tap'][ci][co] = w[co][8 - (dhi*3+dwi)][ci]
__global__ void weight_prep_tapmajor_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Co, int Ci,
                                            int flip_transpose) {
    const long n = (long)Co * Ci * 9;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int rows = flip_transpose ? Ci : Co, cols = flip_transpose ? Co : Ci;
        const int c = (int)(i % cols);
        const long r = i / cols;
        const int row = (int)(r % rows);
        const int tp = (int)(r / rows);
        const int tap = (tp % 3) * 3 + tp / 3;
        float v;
        if (flip_transpose) v = w[((long)c * 9 + (8 - tap)) * Ci + row];      // row = ci, c = co
        else v = w[((long)row * 9 + tap) * Ci + c];                           // row = co, c = ci
        out[i] = __float2bfloat16_rn(v);
    }
}

// Split-bf16 ("bf16x3", csrc/split.cu) forward operand: out[tap'][co][3*Ci] = [w_hi | w_lo | w_hi]; with the activation
// split as [x_hi | x_hi | x_lo] along the channels the same halo kernel computes an fp32-accurate convolution
// (hi.hi + hi.lo + lo.hi, fp32 accumulation in TMEM) at three times the bf16 cost.
__global__ void weight_prep_tapmajor_x3_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Co, int Ci) {
    const long n = (long)Co * Ci * 9;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Ci);
        const long r = i / Ci;
        const int co = (int)(r % Co);
        const int tp = (int)(r / Co);
        const int tap = (tp % 3) * 3 + tp / 3;
        const float v = w[((long)co * 9 + tap) * Ci + c];
        const bf16 hi = __float2bfloat16_rn(v);
        const bf16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        bf16* o = out + ((long)tp * Co + co) * 3 * Ci + c;
        o[0] = hi; o[Ci] = lo; o[2 * Ci] = hi;
    }
}

// weights [9][Cout][Cin] bf16 viewed as (Cin, Cout, 9); box = (64, block_n, tps)
int make_w3_tmap(CUtensorMap* map, const void* ptr, int Cout, int Cin, int block_n, int tps) {
    EncodeTiledFn enc = get_encode_fn();
    if (enc == nullptr) return TAG_ERR_UNSUPPORTED;
    cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, 9};
    cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cout * Cin * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)block_n, (cuuint32_t)tps};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? TAG_OK : 20000 + (int)r;
}

}  // namespace

extern "C" int tag_weight_prep_tapmajor_bf16(const float* w, void* out, int Co, int Ci, int flip_transpose,
                                             cudaStream_t stream) {
    const long n = (long)Co * Ci * 9;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 4096) blocks = 4096;
    weight_prep_tapmajor_kernel<<<blocks, 256, 0, stream>>>(w, (bf16*)out, Co, Ci, flip_transpose);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_weight_prep_tapmajor_x3(const float* w, void* out, int Co, int Ci, cudaStream_t stream) {
    const long n = (long)Co * Ci * 9;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 4096) blocks = 4096;
    weight_prep_tapmajor_x3_kernel<<<blocks, 256, 0, stream>>>(w, (bf16*)out, Co, Ci);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

int tag_halo2_dispatch(const CUtensorMap& tx, const CUtensorMap& tw, void* y, int y_dtype, double* stats, int B, int H,
                       int W, int Cin, int Cout, int block_n, const void* bn_y, const void* pool_cnt,
                       cudaStream_t stream);                 // conv_tc_halo2.cu

namespace {
int g_pair_mode = getenv("TAG_B200_NO_PAIR") ? 0 : 1;
}

// 0: every layer on the one-CTA-per-tile kernel; 1 (default): CTA pairs (cta_group::2) where the weights stream
extern "C" int tag_conv_halo_set_pair_mode(int mode) {
    g_pair_mode = mode;
    return TAG_OK;
}

// 3x3 only, no bias / ReLU.  x: bf16 NHWC, W a multiple of 8; w: bf16 TAP-MAJOR [9][Cout][Cin]
// (tag_weight_prep_tapmajor_bf16); y bf16 or fp32.
extern "C" int tag_conv_tc_fwd_halo(const void* x, const void* w, void* y, int y_dtype, double* stats, int B,
                                    int H, int W, int Cin, int Cout, const void* bn_act, const void* pool_cnt,
                                    cudaStream_t stream) {
    const void* bn_y = bn_act;
    if (pool_cnt != nullptr && bn_act == nullptr) return TAG_ERR_BAD_ARG;
    if (Cin % 64 != 0 || Cout % 64 != 0 || W % TILE_W != 0 || B <= 0 || H <= 0) return TAG_ERR_BAD_ARG;
    if (bn_y != nullptr && stats == nullptr) return TAG_ERR_BAD_ARG;
    const int block_n = Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64);
    const bool wres = Cin == 64 && Cout == block_n && block_n <= 128;
    CUtensorMap tx, tw;
    int rc = make_act_tmap(&tx, x, B, H, W, Cin, TILE_W, TILE_H + 2);
    if (rc != TAG_OK) return rc;
    if (g_pair_mode != 0 && !wres) {
        // weights stream (Cin >= 128): CTA pairs with cta_group::2 MMAs, each CTA loads half of every weight tile
        rc = make_w3_tmap(&tw, w, Cout, Cin, block_n / 2, block_n == 256 ? 1 : 3);
        if (rc != TAG_OK) return rc;
        return tag_halo2_dispatch(tx, tw, y, y_dtype, stats, B, H, W, Cin, Cout, block_n, bn_y, pool_cnt, stream);
    }
    rc = make_w3_tmap(&tw, w, Cout, Cin, block_n, block_n == 256 ? 1 : 3);
    if (rc != TAG_OK) return rc;
#define TAG_HALO(BN_, WR_)                                                                               \
    (y_dtype == TAG_DTYPE_BF16 ? launch_halo<BN_, bf16, WR_>(tx, tw, y, stats, B, H, W, Cin, Cout, bn_y, pool_cnt, stream)  \
                               : launch_halo<BN_, float, WR_>(tx, tw, y, stats, B, H, W, Cin, Cout, bn_y, pool_cnt, stream))
    if (block_n == 256) return TAG_HALO(256, false);
    if (block_n == 128) return wres ? TAG_HALO(128, true) : TAG_HALO(128, false);
    return wres ? TAG_HALO(64, true) : TAG_HALO(64, false);
#undef TAG_HALO
}
