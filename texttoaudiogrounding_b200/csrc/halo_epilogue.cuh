// Epilogue of the halo convolution kernels (conv_tc_halo.cu: one CTA per tile; conv_tc_halo2.cu: a CTA pair per
// 256-pixel tile with cta_group::2 MMAs): 8 warps drain the fp32 accumulator of one 128-pixel x BLOCK_N tile from
// TMEM, apply the optional fused "ReLU gate + BatchNorm backward reductions" (bn_y = saved activation), round to the storage
// type, accumulate per-channel statistics in registers across tiles and store the tile.
// Fused reductions (stats != nullptr), selected by the side inputs:
//   none                : sum v, sum v^2                      (BatchNorm statistics of a forward convolution)
//   bn_y                : v gated by a > 0; sum v, sum v * a   (a = saved activation of the NEXT op in backward order: this
//                         dgrad's output is d(relu(bn(.)))  ->  dbeta, dgamma in the activation domain)
//   bn_y + pool_cnt     : sum v * cnt, sum v * p   (this dgrad's output is d(pool(relu(bn(.)))): p = the saved pooled output,
//                         cnt = its open-gate code from tag_bn_relu_pool_fwd = 4 x the window's gradient weight; the same
//                         two sums for the pooled layer, without a pass over its full-resolution input)
//   decode(tile, n_tile, b, h0, w0): tile -> output-channel tile, image, tile origin (b == B marks a padding tile)
//   arrive_empty(acc): hand accumulator `acc` back to the MMA issuer (a local or a remote mbarrier arrive)
#pragma once
#include "tc_common.cuh"

namespace {

constexpr int TILE_H = 16, TILE_W = 8;

template <int BLOCK_N, typename TO, typename Decode, typename ArriveEmpty>
__device__ __forceinline__ void halo_epilogue(uint32_t tmem_base, uint32_t tmem_full, float* t_buf,
                                              TO* __restrict__ y, double* __restrict__ stats, int B, int H, int W,
                                              int Cout, const bf16* __restrict__ bn_y,
                                              const uint8_t* __restrict__ pool_cnt, int dbg, int first_tile,
                                              int tile_stride, int total_tiles, Decode decode,
                                              ArriveEmpty arrive_empty) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // ===================== epilogue: 8 warps; warp w reads TMEM lanes 32*(w%4).., column half w/4
    int it = 0;
    int cur_n_tile = -1;
    constexpr int CPW = BLOCK_N / 64;            // 32-column chunks per warp
    // Per-channel statistics: a thread owns one pixel row of the tile (its TMEM lane) and 32 columns per chunk.
    // Reducing over the 32 rows of every tile costs a 31-shuffle butterfly per chunk and quantity — that, not the
    // MMAs, paced the narrow layers.  Instead the butterfly is cut after STAGES stages (none for BLOCK_N = 64) and
    // the partial sums (64 registers in total) keep accumulating across tiles; the remaining stages run once per
    // flush.  The lane -> column mapping of a butterfly stage is fixed, so accumulating in between is exact.
    constexpr int STAGES = CPW == 1 ? 0 : (CPW == 2 ? 1 : 2);
    constexpr int KEEP = 32 >> STAGES;
    float run_s[CPW][KEEP], run_q[CPW][KEEP];
#pragma unroll
    for (int i = 0; i < CPW; ++i)
#pragma unroll
        for (int k = 0; k < KEEP; ++k) { run_s[i][k] = 0.f; run_q[i][k] = 0.f; }
    auto flush_stats = [&](int n_tile) {
        float fs[CPW], fq[CPW];
#pragma unroll
        for (int i = 0; i < CPW; ++i) {
            fs[i] = warp_transpose_tail<KEEP>(run_s[i], lane);
            fq[i] = warp_transpose_tail<KEEP>(run_q[i], lane);
#pragma unroll
            for (int k = 0; k < KEEP; ++k) { run_s[i][k] = 0.f; run_q[i][k] = 0.f; }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
        for (int i = 0; i < CPW; ++i) {
            t_buf[warp * BLOCK_N + i * 32 + lane] = fs[i];
            t_buf[warp * BLOCK_N + BLOCK_N / 2 + i * 32 + lane] = fq[i];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int i = threadIdx.x; i < BLOCK_N; i += 256) {
            const int ch = i / (BLOCK_N / 2), local = i % (BLOCK_N / 2);
            double ds = 0.0, dq = 0.0;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                ds += t_buf[(ch * 4 + q4) * BLOCK_N + local];
                dq += t_buf[(ch * 4 + q4) * BLOCK_N + BLOCK_N / 2 + local];
            }
            atomicAdd(stats + n_tile * BLOCK_N + i, ds);
            atomicAdd(stats + Cout + n_tile * BLOCK_N + i, dq);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
    };
    // Side inputs of the fused reductions (saved activation / pooled output, open-gate codes) are prefetched into registers
    // one chunk ahead ACROSS tiles: the first chunk of the next tile is requested before this tile's accumulator is even
    // waited for.  With the request issued at the start of its own tile, a narrow tile (one chunk per warp) paid the whole
    // global-load latency per tile on top of its arithmetic and held the layer at 39 % tensor-pipe-active.
    constexpr bool CAN_FUSE = sizeof(TO) == 2;        // the fused reductions exist for bf16 gradients only
    if (!CAN_FUSE) { bn_y = nullptr; pool_cnt = nullptr; }
    uint4 ynext[4], cnext[2];
    const bf16* f_yrow = nullptr;
    const uint8_t* f_crow = nullptr;
    bool f_valid = false;
    const int e_lq = warp & 3, e_chalf = warp >> 2;
    auto setup_fetch = [&](int tile_) {               // row pointers of this thread in tile `tile_`
        int nt_, b_, h0_, w0_;
        decode(tile_, nt_, b_, h0_, w0_);
        const int row_ = e_lq * 32 + lane;
        const int h_ = h0_ + (row_ >> 3), w_ = w0_ + (row_ & 7);
        f_valid = h_ < H && b_ < B;
        const long off = (((long)b_ * H + h_) * W + w_) * Cout + nt_ * BLOCK_N;
        f_yrow = bn_y != nullptr ? bn_y + off : nullptr;
        f_crow = pool_cnt != nullptr ? pool_cnt + off : nullptr;
    };
    auto fetch_side = [&](int cc_) {
        if (bn_y != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) ynext[j] = make_uint4(0u, 0u, 0u, 0u);
            if (f_valid) {
                const bf16* p = f_yrow + (e_chalf * CPW + cc_) * 32;
                ld32(p, ynext[0], ynext[1]);
                ld32(p + 16, ynext[2], ynext[3]);
            }
        }
        if (pool_cnt != nullptr) {
            cnext[0] = cnext[1] = make_uint4(0u, 0u, 0u, 0u);
            if (f_valid) ld32(f_crow + (e_chalf * CPW + cc_) * 32, cnext[0], cnext[1]);
        }
    };
    // after the call the registers hold chunk (cc_ + 1) of this tile, or chunk 0 of the next tile of this CTA
    auto fetch_following = [&](int tile_, int cc_) {
        if (cc_ + 1 < CPW) { fetch_side(cc_ + 1); return; }
        if (tile_ + tile_stride < total_tiles) { setup_fetch(tile_ + tile_stride); fetch_side(0); }
    };
    if (bn_y != nullptr && first_tile < total_tiles) { setup_fetch(first_tile); fetch_side(0); }
    for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++it) {
        int n_tile, b, h0, w0;
        decode(tile, n_tile, b, h0, w0);
        if (stats != nullptr && cur_n_tile >= 0 && n_tile != cur_n_tile) flush_stats(cur_n_tile);
        cur_n_tile = n_tile;
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        const int lq = warp & 3, chalf = warp >> 2;
        const int row = lq * 32 + lane;
        const int h = h0 + (row >> 3), w = w0 + (row & 7);
        const bool valid = h < H && b < B;
        const bool edge_tile = h0 + TILE_H > H || b >= B;   // tile-uniform: some rows fall off the image
        TO* yrow = y + (((long)b * H + h) * W + w) * Cout + n_tile * BLOCK_N;
        // chunk 0 of this tile is in ynext / cnext already; take it and request what follows BEFORE waiting for the MMAs
        uint4 yraw[4], craw[2];
        if (bn_y != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) yraw[j] = ynext[j];
            craw[0] = cnext[0]; craw[1] = cnext[1];
            fetch_following(tile, 0);
        }
        mbar_wait(tmem_full + 8 * acc, acc_phase);
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < CPW; ++cc) {
            const int c = chalf * CPW + cc;
            uint32_t r[32];
            if (!(dbg & 2)) {
                tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + acc * BLOCK_N + c * 32, r);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = 0;
            }
            if (cc == CPW - 1) {                     // accumulator drained: hand it back to the MMA warp now
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_empty(acc);
            }
            float v[32];                             // the values as stored (rounded to TO, gated)
            float q[32];                             // second statistic's factor: v (plain) or the activation (fused BN bwd)
            const bool fused = bn_y != nullptr;
            const bool pooled = pool_cnt != nullptr;
            if (fused && cc > 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) yraw[j] = ynext[j];
                craw[0] = cnext[0]; craw[1] = cnext[1];
                fetch_following(tile, cc);
            }
            if (fused) {
                // fused "ReLU + BatchNorm backward, reduce pass": this kernel is the dgrad producing d(relu(bn(.))).
                // bn_y is the SAVED ACTIVATION a = relu(gamma * xhat + beta) of that layer: the ReLU gate is a > 0 and the
                // two reductions are taken in the activation domain, sum g and sum g * a — no per-channel parameter is
                // touched per element (32 LDS.128 per chunk paced the narrow layers); tag_bn_red_act_to_xhat turns them
                // into dbeta = sum g, dgamma = sum g * xhat = (sum g * a - beta * sum g) / gamma afterwards.
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    float yv[4];
                    unpack4<bf16>(yraw[j4 >> 1], j4 & 1, yv);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int j = j4 * 4 + e;
                        if (!pooled) r[j] = yv[e] > 0.f ? r[j] : 0u;
                        q[j] = yv[e];
                    }
                }
            }
            // round + pack first (one F2FP per pair), recover the rounded floats from the packed words
            uint4 packed[sizeof(TO) == 2 ? 4 : 8];
            if (sizeof(TO) == 2) {
                uint32_t* pw = reinterpret_cast<uint32_t*>(packed);
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(r[j]), __uint_as_float(r[j + 1]));
                    const uint32_t u = *reinterpret_cast<const uint32_t*>(&h2);
                    pw[j >> 1] = u;
                    v[j] = __uint_as_float(u << 16);
                    v[j + 1] = __uint_as_float(u & 0xFFFF0000u);
                }
            } else {
                uint32_t* pw = reinterpret_cast<uint32_t*>(packed);
#pragma unroll
                for (int j = 0; j < 32; ++j) { pw[j] = r[j]; v[j] = __uint_as_float(r[j]); }
            }
            if (stats != nullptr) {
                // the two terms per element: (v, v * v) | (v, v * a) | (v * cnt, v * p)
#pragma unroll
                for (int j = 0; j < 32; ++j) q[j] = v[j] * (fused ? q[j] : v[j]);
                if (pooled) {
                    const uint32_t* cw = reinterpret_cast<const uint32_t*>(craw);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= (float)((cw[j >> 2] >> (8 * (j & 3))) & 0xFFu);
                }
                if (STAGES == 0) {
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) { run_s[cc][j] += v[j]; run_q[cc][j] += q[j]; }
                    }
                } else {
                    if (edge_tile) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) { v[j] = valid ? v[j] : 0.f; q[j] = valid ? q[j] : 0.f; }
                    }
                    warp_transpose_head<STAGES>(v, lane);
                    warp_transpose_head<STAGES>(q, lane);
#pragma unroll
                    for (int k = 0; k < KEEP; ++k) { run_s[cc][k] += v[k]; run_q[cc][k] += q[k]; }
                }
            }
            if (valid && !(dbg & 1)) {
#pragma unroll
                for (int j = 0; j < (sizeof(TO) == 2 ? 4 : 8); j += 2)
                    st32(reinterpret_cast<uint8_t*>(yrow + c * 32) + 16 * j, packed[j], packed[j + 1]);
            }
        }
    }
    if (stats != nullptr && cur_n_tile >= 0) flush_stats(cur_n_tile);
}

}  // namespace
