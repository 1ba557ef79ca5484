// Weight gradient of a 3x3 convolution with Cin = 64 on tcgen05: ALL nine taps in one CTA.
//
// The generic wgrad tile (conv_tc.cu) pairs two taps per 128-row accumulator and re-reads dy once
// per tap pair and x once per tap — for the block-1 layers (64 channels at full resolution) that
// makes the kernel L2-bandwidth-bound at ~30 % of the tensor peak.  Here one CTA keeps the whole
// dW[64 co][9 taps][64 ci] block in TMEM (five 128-row accumulators x 64 columns = 320 columns) and
// streams 16x8-pixel tiles: per tile ONE dy box and THREE x boxes (one per horizontal shift, 18 rows
// each) — the three vertical taps of a shift read the same box at row offsets 0 / 8 / 16 pixels
// (= 0 / 1024 / 2048 B, whole swizzle atoms).  Both operands are MN-major; the two 64-row halves of
// an accumulator are two taps, addressed through the descriptor's leading-dimension offset.
#include "tc_common.cuh"

namespace {

constexpr int XBOX = 18 * 8 * 128;      // 18432 B: (16+2) rows x 8 cols x 64 ch
constexpr int YBOX = 16 * 8 * 128;      // 16384 B
constexpr int STAGE = 3 * XBOX + YBOX;  // 71680 B
constexpr int STAGES = 3;
constexpr int SMEM_TOTAL = STAGES * STAGE + 256 + 1024;

__device__ __forceinline__ uint32_t tap_addr(int tp) { return (uint32_t)((tp / 3) * XBOX + (tp % 3) * 1024); }

__global__ void __launch_bounds__(192, 1)
conv_tc_wgrad64_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                       float* __restrict__ dw, int B, int H, int W, int Cout, int n_blocks, int tiles_per_split,
                       long slab_stride) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - raw);
    const uint32_t full_bar = base + STAGES * STAGE;
    const uint32_t empty_bar = full_bar + 8 * STAGES;
    const uint32_t tmem_full = empty_bar + 8 * STAGES;
    const uint32_t tmem_slot = tmem_full + 8;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES * STAGE + 16 * STAGES + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_w = W / 8, tiles_h = (H + 15) / 16;
    const int tiles_img = tiles_w * tiles_h;
    const int k_tiles = B * tiles_img;
    const int n_blk = blockIdx.x % n_blocks;
    const int split = blockIdx.x / n_blocks;
    const int kt_begin = split * tiles_per_split;
    const int kt_end = min(k_tiles, kt_begin + tiles_per_split);
    if (kt_begin >= kt_end) return;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 4 && lane == 0) { prefetch_tmap(&tmap_x); prefetch_tmap(&tmap_dy); }
    if (warp == 5) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 4) {
        if (elect_one_sync()) {
            int stage = 0; uint32_t phase = 0;
            for (int kt = kt_begin; kt < kt_end; ++kt) {
                const int b = kt / tiles_img;
                const int r = kt - b * tiles_img;
                const int h0 = (r / tiles_w) * 16, w0 = (r % tiles_w) * 8;
                mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                mbar_arrive_expect_tx(full_bar + 8 * stage, STAGE);
                const uint32_t sa = base + stage * STAGE;
#pragma unroll
                for (int dwi = 0; dwi < 3; ++dwi)
                    tma_load_4d(sa + dwi * XBOX, &tmap_x, full_bar + 8 * stage, 0, w0 + dwi - 1, h0 - 1, b);
                tma_load_4d(sa + 3 * XBOX, &tmap_dy, full_bar + 8 * stage, n_blk * 64, w0, h0, b);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 5) {
        constexpr uint32_t idesc = make_idesc(128, 64, 1, 1);
        int stage = 0; uint32_t phase = 0;
        for (int kt = kt_begin; kt < kt_end; ++kt) {
            mbar_wait(full_bar + 8 * stage, phase);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint32_t sa = base + stage * STAGE;
                const uint64_t bdesc = make_smem_desc(sa + 3 * XBOX, YBOX, 1024);
#pragma unroll
                for (int m = 0; m < 5; ++m) {
                    const uint32_t a0 = tap_addr(2 * m), a1 = tap_addr(2 * m + 1);
                    const uint64_t adesc = make_smem_desc(sa + a0, a1 - a0, 1024);
#pragma unroll
                    for (int k = 0; k < 8; ++k)       // 16 pixels (2 image rows of the tile) per K step
                        umma_bf16(tmem_base + m * 64, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128),
                                  idesc, (kt > kt_begin || k > 0) ? 1u : 0u);
                }
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
        const int ci = row & 63;
        const bool slab = slab_stride != 0;
        float* dwp = dw + (long)split * slab_stride;
#pragma unroll 1
        for (int m = 0; m < 5; ++m) {
            const int tp = 2 * m + (row >> 6);          // tap' = dwi*3 + dhi
            const int tap = (tp % 3) * 3 + tp / 3;      // kh*3 + kw
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + m * 64 + c * 32, r);
                tmem_ld_wait();
                if (tp < 9) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int co = n_blk * 64 + c * 32 + j;
                        wg_put(dwp + ((long)co * 9 + tap) * 64 + ci, __uint_as_float(r[j]), slab);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 5) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// dy: bf16 NHWC [B,H,W,Cout]; x: bf16 NHWC [B,H,W,64]; dw: fp32 [Cout][9][64], accumulated (+=)
extern "C" int tag_conv_tc_wgrad64(const void* dy, const void* x, float* dw, int B, int H, int W, int Cout,
                                   int splits, cudaStream_t stream) {
    if (Cout % 64 != 0 || W % 8 != 0 || B <= 0 || H <= 0 || splits <= 0) return TAG_ERR_BAD_ARG;
    CUtensorMap tx, tdy;
    int rc = make_act_tmap(&tx, x, B, H, W, 64, 8, 18);
    if (rc != TAG_OK) return rc;
    rc = make_act_tmap(&tdy, dy, B, H, W, Cout, 8, 16);
    if (rc != TAG_OK) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_wgrad64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             SMEM_TOTAL);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int k_tiles = B * ((H + 15) / 16) * (W / 8);
    const int n_blocks = Cout / 64;
    if (splits > k_tiles) splits = k_tiles;
    const int tps = (k_tiles + splits - 1) / splits;
    splits = (k_tiles + tps - 1) / tps;
    const long n = (long)Cout * 9 * 64;
    float* target; long slab;
    rc = splitk_target(dw, splits, n, stream, &target, &slab);
    if (rc != TAG_OK) return rc;
    conv_tc_wgrad64_kernel<<<n_blocks * splits, 192, SMEM_TOTAL, stream>>>(tx, tdy, target, B, H, W, Cout, n_blocks, tps, slab);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return slab ? tag_splitk_reduce(target, splits, n, dw, stream) : TAG_OK;
}
