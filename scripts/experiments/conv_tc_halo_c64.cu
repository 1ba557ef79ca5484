// EXPERIMENT, NOT BUILT INTO libtag_b200.so (round 2; DESIGN.md section 9, gap 3).  Parity-green on B200 (tests/test_gpu_tc.py conv3x3
// + fused BN-backward cases, 67 passed) but SLOWER than the nine-tap kernel it was meant to replace, so it was taken out again:
//                                   nine-tap (conv_tc_halo.cu)      this kernel
//   forward, stats                  0.301 ms                        0.397-0.426 ms
//   dgrad, fused BN reduce          0.373 ms                        0.50-0.63 ms
//   main loop alone (TAG_HALO_DBG=3 / 19: no stores, no TMEM loads, no hand-over)   0.258 ms        0.250 ms
// What it taught: with a third fewer MMAs (24 instead of 36 per tile) and 146 instead of 192 B/cycle of operand reads the main
// loop takes the SAME time, and so does the nine-tap kernel with an 8-deep instead of 6-deep input ring (0.262 ms) — the N = 64
// layers are not bound by MMA issue, shared-memory operand bandwidth or TMA latency but by what both forms share: three 18 KB
// input boxes per 128-pixel tile through L2 (1.8 GB per launch, ~7 TB/s).  The row hand-over between TMEM lane quarters
// (one bar.sync of the 8 epilogue warps per tile + 32 shuffles) costs another 0.1 ms on top.
// To try it again: copy this file to texttoaudiogrounding_b200/csrc/ and call tag_halo_c64_dispatch from tag_conv_tc_fwd_halo for
// Cin == Cout == 64, bf16 output, no pool_cnt (tx and tw as built there for the weights-resident case).
//
// 3x3 convolution forward / dgrad for the 64 -> 64 channel layer (conv_block1.conv2: 4.1 M pixels per batch, the
// largest activation of the model), "shift-accumulate" form of the halo kernel (conv_tc_halo.cu).
//
// With N = 64 output channels a tcgen05.mma (128 x 64 x 16) occupies the tensor pipe for 32 cycles, costs the issuing
// thread about twice that, and asks shared memory for 192 B per cycle of operands: the nine-tap form (36 such MMAs
// per 128-pixel tile) ran the pipe at 40-50 %.  Here the three vertical taps of one horizontal shift share ONE input
// box (as before), but two of them also share the MMA:
//     box rows j = 0 .. 17  <->  image rows h0 - 1 + j,      out[h] = sum_dh A[h + dh - 1] W[dh]
//     MMA-1 (N = 128): A rows j      x [W0 | W1]  ->  TMEM cols  0.. 63 += A[j] W0     (belongs to output row j + 1)
//                                                   TMEM cols 64..127 += A[j] W1     (belongs to output row j)
//     MMA-2 (N =  64): A rows j + 1  x  W2        ->  TMEM cols 64..127 += A[j+1] W2   (belongs to output row j)
// so a tile costs 24 MMAs (12 of 64 cycles, 12 of 32) instead of 36 of 32, 146 B per cycle of operand reads, and the
// epilogue adds the W0 partial of the row above: out[j] = cols64..127[j] + cols0..63[j - 1], a shift by 8 TMEM lanes
// (one image row of the 8-pixel-wide tile) done with a warp shuffle plus a 1 KB shared-memory hand-over of the last row
// of each 32-lane quarter.  Output rows per tile: j = 1 .. 15 (15 of the 16 accumulator rows, 94 %).
// Weights (72 KB, tap-major [dw][dh][co][ci]) stay resident in shared memory; two MMA-issuing warps alternate tiles,
// each owning one TMEM accumulator and one half of the input ring, as in the nine-tap kernel.
//
// Epilogue modes (bf16 output): plain store | BatchNorm statistics (sum v, sum v^2) | fused ReLU gate + BatchNorm
// backward sums in the activation domain (bn_y = saved activation a: v gated by a > 0; sum v, sum v * a) — see
// halo_epilogue.cuh for the algebra.  Replaces cudnn conv forward / dgrad for reference models/panns.py:47-58.
#include "tc_common.cuh"
#include <cstdlib>

namespace {

constexpr int C64_TILE_W = 8;
constexpr int C64_OUT_H = 15;                                   // output rows per tile
constexpr int C64_BOX_H = 18;                                   // rows of the input box (rows 0 .. 16 are read)
constexpr int C64_A_BYTES = C64_BOX_H * C64_TILE_W * 128;       // 18432 = 18 swizzle atoms
constexpr int C64_A_STAGES = 6;                                 // two tiles of three boxes: one half per MMA issuer
constexpr int C64_TAP_BYTES = 64 * 128;                         // one tap: 64 output channels x 64 input channels
constexpr int C64_W_BYTES = 3 * C64_TAP_BYTES;                  // one horizontal shift: its three vertical taps

struct C64Smem {
    static constexpr int A_OFFSET = 0;
    static constexpr int W_OFFSET = C64_A_STAGES * C64_A_BYTES;             // 110592
    static constexpr int X_OFFSET = W_OFFSET + 3 * C64_W_BYTES;             // 184320: row hand-over between quarters
    static constexpr int X_BYTES = 2 * 8 * 8 * 32 * 4;                      // [buffer][warp][row pixel][column]
    static constexpr int BAR_OFFSET = X_OFFSET + X_BYTES;                   // 200704
    static constexpr int TBUF_OFFSET = BAR_OFFSET + 512;
    static constexpr int TOTAL = TBUF_OFFSET + 8 * 64 * 4 + 1024;
    static_assert(TOTAL <= 227 * 1024, "shared memory budget");
};

__global__ void __launch_bounds__(384, 1)
conv_tc_halo_c64_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                        bf16* __restrict__ y, double* __restrict__ stats, int B, int H, int W,
                        const bf16* __restrict__ bn_y, int dbg) {
    using L = C64Smem;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - raw);
    const uint32_t a_full = base + L::BAR_OFFSET;
    const uint32_t a_empty = a_full + 8 * C64_A_STAGES;
    const uint32_t w_full = a_empty + 8 * C64_A_STAGES;
    const uint32_t tmem_full = w_full + 8 * 3;
    const uint32_t tmem_empty = tmem_full + 16;
    const uint32_t tmem_slot = tmem_empty + 16;
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(base_ptr + L::BAR_OFFSET + 16 * C64_A_STAGES + 24 + 32);
    float* xbuf = reinterpret_cast<float*>(base_ptr + L::X_OFFSET);
    float* t_buf = reinterpret_cast<float*>(base_ptr + L::TBUF_OFFSET);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_w = W / C64_TILE_W;
    const int tiles_h = (H + C64_OUT_H - 1) / C64_OUT_H;
    const int tiles_img = tiles_w * tiles_h;
    const int total_tiles = B * tiles_img;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C64_A_STAGES; ++s) { mbar_init(a_full + 8 * s, 1); mbar_init(a_empty + 8 * s, 1); }
        for (int s = 0; s < 3; ++s) mbar_init(w_full + 8 * s, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(tmem_full + 8 * a, 1); mbar_init(tmem_empty + 8 * a, 8); }
        fence_barrier_init();
    }
    if (warp == 8 && lane == 0) { prefetch_tmap(&tmap_x); prefetch_tmap(&tmap_w); }
    if (warp == 9) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    auto decode = [&](int tile, int& b, int& h0, int& w0) {      // h0 = first OUTPUT row of the tile
        b = tile / tiles_img;
        const int r = tile - b * tiles_img;
        const int th = r / tiles_w;
        h0 = th * C64_OUT_H;
        w0 = (r - th * tiles_w) * C64_TILE_W;
    };

    if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        if (warp == 8) {
            // ===================== TMA producer =====================
            if (elect_one_sync()) {
                for (int dwi = 0; dwi < 3; ++dwi) {
                    mbar_arrive_expect_tx(w_full + 8 * dwi, C64_W_BYTES);
                    tma_load_3d(base + L::W_OFFSET + dwi * C64_W_BYTES, &tmap_w, w_full + 8 * dwi, 0, 0, dwi * 3);
                }
                int as = 0; uint32_t aph = 0;
                for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                    int b, h0, w0;
                    decode(tile, b, h0, w0);
                    for (int dwi = 0; dwi < 3; ++dwi) {
                        mbar_wait(a_empty + 8 * as, aph ^ 1);
                        mbar_arrive_expect_tx(a_full + 8 * as, C64_A_BYTES);
                        tma_load_4d(base + L::A_OFFSET + as * C64_A_BYTES, &tmap_x, a_full + 8 * as, 0, w0 + dwi - 1,
                                    h0 - 1, b);
                        if (++as == C64_A_STAGES) { as = 0; aph ^= 1; }
                    }
                }
            }
        } else if (warp == 9 || warp == 10) {
            // ===================== MMA issuers: warp 9 the even tiles of this CTA (accumulator 0, ring stages 0-2),
            // warp 10 the odd ones (accumulator 1, stages 3-5) =====================
            constexpr uint32_t idesc128 = make_idesc(128, 128, 0, 0);
            constexpr uint32_t idesc64 = make_idesc(128, 64, 0, 0);
            const int mw = warp - 9;
            int it = mw;
            for (int tile = blockIdx.x + mw * gridDim.x; tile < total_tiles; tile += 2 * gridDim.x, it += 2) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                int as = (it * 3) % C64_A_STAGES;
                const uint32_t aph = ((uint32_t)(it * 3) / C64_A_STAGES) & 1u;
                mbar_wait(tmem_empty + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 192;
                for (int dwi = 0; dwi < 3; ++dwi, ++as) {
                    mbar_wait(a_full + 8 * as, aph);
                    mbar_wait(w_full + 8 * dwi, 0);                // resident: completed once, stays readable
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint32_t sa = base + L::A_OFFSET + as * C64_A_BYTES;
                        const uint32_t wb = base + L::W_OFFSET + dwi * C64_W_BYTES;
                        const uint64_t a_row0 = make_smem_desc(sa, 16, 1024);
                        const uint64_t a_row1 = make_smem_desc(sa + C64_TILE_W * 128, 16, 1024);
                        const uint64_t b_w01 = make_smem_desc(wb, 16, 1024);
                        const uint64_t b_w2 = make_smem_desc(wb + 2 * C64_TAP_BYTES, 16, 1024);
                        const uint32_t d2 = (dbg & 8) ? d_tmem + 128 : d_tmem + 64;
                        const uint32_t acc2 = (dbg & 8) ? ((dwi) != 0 ? 1u : 0u) : 1u;
                        if (dbg & 4) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16(d_tmem, a_row0 + 2 * k, b_w01 + 2 * k, idesc128, (dwi | k) != 0 ? 1u : 0u);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16(d2, a_row1 + 2 * k, b_w2 + 2 * k, idesc64, k != 0 ? 1u : acc2);
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                umma_bf16(d_tmem, a_row0 + 2 * k, b_w01 + 2 * k, idesc128, (dwi | k) != 0 ? 1u : 0u);
                                umma_bf16(d2, a_row1 + 2 * k, b_w2 + 2 * k, idesc64, k != 0 ? 1u : acc2);
                            }
                        }
                        umma_commit(a_empty + 8 * as);
                        if (dwi == 2) umma_commit(tmem_full + 8 * acc);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        // ===================== epilogue: 8 warps; warp w drains TMEM lanes 32 (w % 4) .., output columns 32 (w / 4) ..
        const int lq = warp & 3, chalf = warp >> 2;
        const int m = lq * 32 + lane;
        const int j = m >> 3, wc = m & 7;                  // accumulator row -> box row j, pixel column wc
        float run_s[32], run_q[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) { run_s[k] = 0.f; run_q[k] = 0.f; }
        const bool fused = bn_y != nullptr;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            int b, h0, w0;
            decode(tile, b, h0, w0);
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int h = h0 - 1 + j;
            const bool valid = j >= 1 && h < H;
            const long off = (((long)b * H + h) * W + w0 + wc) * 64 + chalf * 32;
            uint4 yraw[4];
            if (fused) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) yraw[q4] = make_uint4(0u, 0u, 0u, 0u);
                if (valid) {
                    ld32(bn_y + off, yraw[0], yraw[1]);
                    ld32(bn_y + off + 16, yraw[2], yraw[3]);
                }
            }
            mbar_wait(tmem_full + 8 * acc, acc_phase);
            tc_fence_after();
            uint32_t da[32], db[32];
            if (!(dbg & 2)) {
                tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + acc * 192 + chalf * 32, db);
                tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + acc * 192 + 64 + chalf * 32, da);
                tmem_ld_wait();
                if (dbg & 8) {
                    uint32_t dc[32];
                    tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + acc * 192 + 128 + chalf * 32, dc);
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 32; ++k) da[k] = __float_as_uint(__uint_as_float(da[k]) + __uint_as_float(dc[k]));
                }
            } else {
#pragma unroll
                for (int k = 0; k < 32; ++k) { da[k] = 0u; db[k] = 0u; }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + 8 * acc);          // accumulator drained
            // the W0 partial of box row j belongs to output row j + 1: lanes 24..31 (the last row of this quarter) hand
            // theirs to the next quarter through shared memory, everybody else shifts by 8 lanes inside the warp
            float* xw = xbuf + (((it & 1) * 8 + warp) * 8) * 32;
            if (!(dbg & 16)) {
            if (lane >= 24) {
                float4* dst = reinterpret_cast<float4*>(xw + (lane - 24) * 32);
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4)
                    dst[k4] = make_float4(__uint_as_float(db[4 * k4]), __uint_as_float(db[4 * k4 + 1]),
                                          __uint_as_float(db[4 * k4 + 2]), __uint_as_float(db[4 * k4 + 3]));
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            float v[32];
#pragma unroll
            for (int k = 0; k < 32; ++k)
                v[k] = __uint_as_float(da[k]) + ((dbg & 16) ? __uint_as_float(db[k]) : __shfl_up_sync(0xffffffffu, __uint_as_float(db[k]), 8));
            if (lane < 8 && lq > 0 && !(dbg & 16)) {
                const float4* src = reinterpret_cast<const float4*>(xbuf + (((it & 1) * 8 + warp - 1) * 8 + lane) * 32);
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const float4 t = src[k4];
                    v[4 * k4] = __uint_as_float(da[4 * k4]) + t.x;
                    v[4 * k4 + 1] = __uint_as_float(da[4 * k4 + 1]) + t.y;
                    v[4 * k4 + 2] = __uint_as_float(da[4 * k4 + 2]) + t.z;
                    v[4 * k4 + 3] = __uint_as_float(da[4 * k4 + 3]) + t.w;
                }
            }
            // per group of four columns: ReLU gate (fused mode), round + pack (one F2FP per pair), recover the rounded
            // floats from the packed words, accumulate the two per-channel sums
            uint4 packed[4];
            uint32_t* pw = reinterpret_cast<uint32_t*>(packed);
            const bool count = stats != nullptr && valid;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
                float yv[4] = {0.f, 0.f, 0.f, 0.f};
                if (fused) {
                    unpack4<bf16>(yraw[j4 >> 1], j4 & 1, yv);
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[j4 * 4 + e] = yv[e] > 0.f ? v[j4 * 4 + e] : 0.f;
                }
#pragma unroll
                for (int e = 0; e < 4; e += 2) {
                    const int k = j4 * 4 + e;
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[k], v[k + 1]);
                    const uint32_t u = *reinterpret_cast<const uint32_t*>(&h2);
                    pw[k >> 1] = u;
                    const float r0 = __uint_as_float(u << 16), r1 = __uint_as_float(u & 0xFFFF0000u);
                    if (count) {
                        run_s[k] += r0;
                        run_s[k + 1] += r1;
                        run_q[k] = fmaf(r0, fused ? yv[e] : r0, run_q[k]);
                        run_q[k + 1] = fmaf(r1, fused ? yv[e + 1] : r1, run_q[k + 1]);
                    }
                }
            }
            if (valid && !(dbg & 1)) {
                st32(y + off, packed[0], packed[1]);
                st32(y + off + 16, packed[2], packed[3]);
            }
        }
        if (stats != nullptr) {
            const float fs = warp_transpose_tail<32>(run_s, lane);      // lane l: column l summed over this warp's rows
            const float fq = warp_transpose_tail<32>(run_q, lane);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            t_buf[warp * 64 + lane] = fs;
            t_buf[warp * 64 + 32 + lane] = fq;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x < 64) {
                const int ch = threadIdx.x >> 5, local = threadIdx.x & 31;
                double ds = 0.0, dq = 0.0;
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    ds += t_buf[(ch * 4 + q4) * 64 + local];
                    dq += t_buf[(ch * 4 + q4) * 64 + 32 + local];
                }
                atomicAdd(stats + threadIdx.x, ds);
                atomicAdd(stats + 64 + threadIdx.x, dq);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 9) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// tx: activation map with box (64 ch, 8 px, 18 rows, 1); tw: tap-major weights [9][64][64] with box (64, 64, 3).
int tag_halo_c64_dispatch(const CUtensorMap& tx, const CUtensorMap& tw, void* y, double* stats, int B, int H, int W,
                          const void* bn_y, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_halo_c64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             C64Smem::TOTAL);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int total_tiles = B * ((H + C64_OUT_H - 1) / C64_OUT_H) * (W / C64_TILE_W);
    const int grid = total_tiles < sm_count() ? total_tiles : sm_count();
    static int dbg = getenv("TAG_HALO_DBG") ? atoi(getenv("TAG_HALO_DBG")) : 0;
    conv_tc_halo_c64_kernel<<<grid, 384, C64Smem::TOTAL, stream>>>(tx, tw, (bf16*)y, stats, B, H, W, (const bf16*)bn_y, dbg);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}
