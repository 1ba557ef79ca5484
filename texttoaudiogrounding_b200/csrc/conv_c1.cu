// First convolution of the stack (Cin = 1, Cout = 64, 3x3, pad 1) — a 9-tap stencil on CUDA cores.
// HBM-bound: the forward writes 64 channels per pixel (8.2 MB bf16 per 10 s clip), the backward
// reads them once.  Reference: conv_block1.conv1, models/panns.py:25-28,49 (+ its autograd).
//
// One CTA owns a tile of 8 time rows x 64 mel bins of one clip; the (8+2) x (64+2) input halo sits
// in shared memory; a thread owns (pixel, group of 8 channels) so every global access is 16 bytes.
//   fwd : y[p, co] = sum_tap x[p + d(tap)] w[co][tap]         (+ per-channel sum / sum-of-squares)
//   bwd : dw[co][tap] = sum_p dy[p, co] x[p + d(tap)]
//         dx[p] = sum_tap S[p - d(tap)][tap],  S[q][tap] = sum_co dy[q, co] w[co][tap]
//         (S is built once per halo pixel in shared memory: dy is read from HBM exactly once.)
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32_c1(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int CO = 64;
constexpr int TH = 8;            // tile rows
constexpr int TW = 64;           // tile width == W
constexpr int XS_W = TW + 2;

template <typename T>
__device__ __forceinline__ void load_x_tile(float (*xs)[XS_W], const T* x, int b, int h0, int H, int W) {
    for (int i = threadIdx.x; i < (TH + 2) * XS_W; i += blockDim.x) {
        const int r = i / XS_W, c = i - r * XS_W;
        const int h = h0 - 1 + r, w = c - 1;
        float v = 0.f;
        if (h >= 0 && h < H && w >= 0 && w < W) v = to_f<T>(x[((long)b * H + h) * W + w]);
        xs[r][c] = v;
    }
}

// ACT: the BatchNorm scale / shift of the layer are already known (statistics from tag_c1_moments, or running statistics
// in eval mode) and the kernel writes relu(scale * conv + shift) directly: the raw convolution output never exists in HBM.
template <typename T, bool ACT>
__global__ void __launch_bounds__(256)
conv_c1_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, T* __restrict__ y,
                   double* __restrict__ stats, int B, int H, int W, const float* __restrict__ bn_scale,
                   const float* __restrict__ bn_shift) {
    __shared__ float xs[TH + 2][XS_W];
    __shared__ float s_sum[8][CO], s_sq[8][CO];                   // per-warp totals, summed in warp order (no fp32 atomics)
    const int tiles_h = (H + TH - 1) / TH;
    const int cg = threadIdx.x & 7, pl = threadIdx.x >> 3;        // 8 channel groups x 32 pixels
    float wr[8][9];
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int t = 0; t < 9; ++t) wr[c][t] = __ldg(w + (cg * 8 + c) * 9 + t);
    float cs[8], cq[8], bsc[8], bsh[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        cs[c] = 0.f; cq[c] = 0.f;
        bsc[c] = ACT ? __ldg(bn_scale + cg * 8 + c) : 1.f;
        bsh[c] = ACT ? __ldg(bn_shift + cg * 8 + c) : 0.f;
    }
    // persistent over tiles: the per-channel statistics stay in registers and are flushed ONCE per CTA (one
    // flush per tile meant ~1 M double atomics on 8 cache lines, which serialise in L2)
    // the halo tile of the NEXT tile is fetched into registers before this tile's arithmetic and committed to shared
    // memory after it: the global-load latency (~1 us against ~1.5 us of work per tile) is off the critical path
    constexpr int XN = ((TH + 2) * XS_W + 255) / 256;
    float xnext[XN];
    auto fetch_tile = [&](int tile) {
        const bool ok = tile < B * tiles_h;
        const int b = ok ? tile / tiles_h : 0, h0 = ok ? (tile % tiles_h) * TH : 0;
#pragma unroll
        for (int j = 0; j < XN; ++j) {
            const int i = threadIdx.x + 256 * j;
            const int r = i / XS_W, c = i - r * XS_W;
            const int h = h0 - 1 + r, wc = c - 1;
            float v = 0.f;
            if (ok && i < (TH + 2) * XS_W && h >= 0 && h < H && wc >= 0 && wc < W) v = to_f<T>(x[((long)b * H + h) * W + wc]);
            xnext[j] = v;
        }
    };
    auto commit_tile = [&]() {
#pragma unroll
        for (int j = 0; j < XN; ++j) {
            const int i = threadIdx.x + 256 * j;
            if (i < (TH + 2) * XS_W) (&xs[0][0])[i] = xnext[j];
        }
    };
    fetch_tile(blockIdx.x);
    for (int tile = blockIdx.x; tile < B * tiles_h; tile += gridDim.x) {
    const int b = tile / tiles_h, h0 = (tile % tiles_h) * TH;
    __syncthreads();
    commit_tile();
    __syncthreads();
    fetch_tile(tile + gridDim.x);
#pragma unroll 2
    for (int it = 0; it < TH * TW / 32; ++it) {
        const int pix = it * 32 + pl;
        const int r = pix / TW, c0 = pix - r * TW;
        const int h = h0 + r;
        if (h >= H) break;
        float xn[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) xn[t] = xs[r + t / 3][c0 + t % 3];
        float o[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float a = 0.f;
#pragma unroll
            for (int t = 0; t < 9; ++t) a = fmaf(xn[t], wr[c][t], a);
            if (ACT) a = fmaxf(fmaf(a, bsc[c], bsh[c]), 0.f);
            a = round_to<T>(a);
            o[c] = a;
            if (!ACT) { cs[c] += a; cq[c] += a * a; }
        }
        store8<T>(y + (((long)b * H + h) * W + c0) * CO + cg * 8, o);
    }
    }
    if (!ACT && stats != nullptr) {
        // lanes with equal (lane & 7) share the channel group: reduce over lane bits 3,4
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float sv = cs[c], qv = cq[c];
            sv += __shfl_xor_sync(0xffffffffu, sv, 8);  qv += __shfl_xor_sync(0xffffffffu, qv, 8);
            sv += __shfl_xor_sync(0xffffffffu, sv, 16); qv += __shfl_xor_sync(0xffffffffu, qv, 16);
            if ((threadIdx.x & 31) < 8) {
                s_sum[threadIdx.x >> 5][cg * 8 + c] = sv;
                s_sq[threadIdx.x >> 5][cg * 8 + c] = qv;
            }
        }
        __syncthreads();
        if (threadIdx.x < CO) {
            double ts = 0.0, tq = 0.0;
#pragma unroll
            for (int wv = 0; wv < 8; ++wv) { ts += (double)s_sum[wv][threadIdx.x]; tq += (double)s_sq[wv][threadIdx.x]; }
            atomicAdd(stats + threadIdx.x, ts);
            atomicAdd(stats + CO + threadIdx.x, tq);
        }
    }
}

// WGRAD / DGRAD select the half of the backward this instantiation computes: the two halves run
// as two launches so that neither needs more than 128 registers (72 wgrad accumulators vs the
// projection state); dy is then read twice from HBM (2 x 8.2 MB per clip), still far from the bound.
template <typename T, bool WGRAD, bool DGRAD>
__global__ void __launch_bounds__(256, 2)
conv_c1_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ w,
                   float* __restrict__ dw, float* __restrict__ dx, int B, int H, int W) {
    __shared__ float xs[TH + 2][XS_W];
    __shared__ __align__(16) float wsm[8][76];            // [cg][tap*8 + c], padded: conflict-free LDS.128
    __shared__ float S[TH + 2][TW][9];                    // projected gradients of the halo tile
    __shared__ float s_dw[CO * 9];
    const int tiles_h = (H + TH - 1) / TH;
    const int b = blockIdx.x / tiles_h, h0 = (blockIdx.x % tiles_h) * TH;
    const int cg = threadIdx.x & 7, pl = threadIdx.x >> 3;
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < CO * 9; i += blockDim.x) {
        const int co = i / 9, t = i - co * 9;
        wsm[co >> 3][t * 8 + (co & 7)] = w[i];
        s_dw[i] = 0.f;
    }
    load_x_tile<T>(xs, x, b, h0, H, W);
    __syncthreads();

    float acc[8][9];
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[c][t] = 0.f;

    // ---- phase A: every halo pixel q = (h0 - 1 + r, c0): project dy[q, :] onto the 9 taps
    // the wgrad-only instantiation skips the halo rows; loads are software-pipelined one pass ahead
    constexpr int IT0 = DGRAD ? 0 : TW / 32;
    constexpr int IT1 = DGRAD ? (TH + 2) * TW / 32 : (TH + 1) * TW / 32;
    // dy is prefetched PF passes ahead as raw 16-byte vectors: with one 16-byte load in flight per thread
    // the kernel was latency-bound at ~1 TB/s
    constexpr int PF = 4;
    constexpr int RAWN = (int)sizeof(T) * 8 / 16;          // uint4 per 8 channels (1 for bf16, 2 for fp32)
    uint4 ring[PF][RAWN];
    auto fetch = [&](int it, uint4* raw) {
        const int pix = it * 32 + pl;
        const int r = pix / TW, c0 = pix - r * TW;
        const int h = h0 - 1 + r;
        const bool ok = it < IT1 && h >= 0 && h < H;
        const T* src = dy + (((long)b * H + h) * W + c0) * CO + cg * 8;
#pragma unroll
        for (int j = 0; j < RAWN; ++j) raw[j] = ok ? ld16(reinterpret_cast<const uint8_t*>(src) + 16 * j) : make_uint4(0u, 0u, 0u, 0u);
    };
#pragma unroll
    for (int i = 0; i < PF; ++i) fetch(IT0 + i, ring[i]);
#pragma unroll 1
    for (int it0 = IT0; it0 < IT1; it0 += PF) {
#pragma unroll
    for (int ii = 0; ii < PF; ++ii) {
        const int it = it0 + ii;
        float g[8];
        if (RAWN == 1) { unpack4<T>(ring[ii][0], 0, g); unpack4<T>(ring[ii][0], 1, g + 4); }
        else { unpack4<T>(ring[ii][0], 0, g); unpack4<T>(ring[ii][RAWN - 1], 0, g + 4); }
        fetch(it + PF, ring[ii]);
        if (it >= IT1) continue;
        const int pix = it * 32 + pl;
        const int r = pix / TW, c0 = pix - r * TW;
        const int h = h0 - 1 + r;
        const bool row_ok = h >= 0 && h < H;
        if (DGRAD) {
        float s[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const float4 w0 = *reinterpret_cast<const float4*>(&wsm[cg][t * 8]);
            const float4 w1 = *reinterpret_cast<const float4*>(&wsm[cg][t * 8 + 4]);
            s[t] = g[0] * w0.x + g[1] * w0.y + g[2] * w0.z + g[3] * w0.w +
                   g[4] * w1.x + g[5] * w1.y + g[6] * w1.z + g[7] * w1.w;
        }
        // reduce over the 8 channel-group lanes (lane bits 0..2)
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            s[t] += __shfl_xor_sync(0xffffffffu, s[t], 1);
            s[t] += __shfl_xor_sync(0xffffffffu, s[t], 2);
            s[t] += __shfl_xor_sync(0xffffffffu, s[t], 4);
        }
        if ((lane & 7) == 0) {
#pragma unroll
            for (int t = 0; t < 9; ++t) S[r][c0][t] = s[t];
        }
        }
        // wgrad for interior pixels
        if (WGRAD && r >= 1 && r <= TH && row_ok) {
            float xn[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) xn[t] = xs[r - 1 + t / 3][c0 + t % 3];
#pragma unroll
            for (int c = 0; c < 8; ++c)
#pragma unroll
                for (int t = 0; t < 9; ++t) acc[c][t] = fmaf(g[c], xn[t], acc[c][t]);
        }
    }
    }
    // ---- wgrad: reduce the per-thread accumulators (lanes with equal cg, then across warps)
    if (WGRAD) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            float v = acc[c][t];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (lane < 8) atomicAdd(&s_dw[(cg * 8 + c) * 9 + t], v);
        }
    }
    __syncthreads();
    if (WGRAD)
        for (int i = threadIdx.x; i < CO * 9; i += blockDim.x) atomicAdd(dw + i, s_dw[i]);
    // ---- phase B: dx[p] = sum_tap S[p - d(tap)][tap]
    if (DGRAD && dx != nullptr) {
        for (int pix = threadIdx.x; pix < TH * TW; pix += blockDim.x) {
            const int r = pix / TW, c0 = pix - r * TW;
            const int h = h0 + r;
            if (h >= H) continue;
            float a = 0.f;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int dh = t / 3 - 1, dwv = t % 3 - 1;
                const int cc = c0 - dwv;
                if (cc >= 0 && cc < TW) a += S[r + 1 - dh][cc][t];
            }
            dx[((long)b * H + h) * W + c0] = a;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// bf16 backward on mma.sync (m16n8k16): both halves of the Cin = 1 backward are thin GEMMs over dy,
//   S[p][tap]    = sum_co dy[p][co] * w[co][tap]        (M = 16 pixels, N = 9 -> 16 taps, K = 64 channels)
//   dw^T[tap][co] = sum_p xcol[p][tap] * dy[p][co]       (M = 9 -> 16 taps, N = 64 channels, K = 16 pixels)
// which the CUDA-core version above spends ~200 instructions per 4 pixels on.  Here a warp streams units of
// 16 consecutive pixels x 64 channels (2 KB) through a private cp.async ring in shared memory (XOR-swizzled
// 16-byte chunks); `ldmatrix` of a unit yields the A fragments of the first product and `ldmatrix.trans` of the
// SAME addresses the B fragments of the second, so dy is read from HBM once and never touches the register
// file as scalars.  dx[p] = sum_tap S[p - d(tap)][tap] is then assembled from the S tile (halo rows recomputed).
constexpr int MTH = 16;                       // tile rows
constexpr int MROWS = MTH + 2;
constexpr int MNST = 3;                       // cp.async stages per warp
constexpr int MXS_W = 68;                     // bf16 halo row: 66 used
constexpr int MMA_SMEM = 8 * MNST * 2048 + MROWS * 64 * 9 * 4 + MROWS * MXS_W * 2 + CO * 9 * 4;
constexpr int MMA_SMEM_FUSE = 8 * MNST * 2048 + MROWS * 64 * 9 * 4 + (MROWS + 2) * MXS_W * 2 + CO * 9 * 4 + 3 * CO * 4;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
    uint32_t d;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
    return d;
}

// FUSE_BN: `dy` is g = d(relu(bn1(y))) already gated by the ReLU mask (the fused epilogue of the conv2 dgrad), and the
// kernel applies the BatchNorm backward itself before the two products:
//     dy1 = A[c] * g + Bc[c] * y + C[c],   y = conv(x, w) RECOMPUTED here (one more thin MMA: 16 pixels x 9 taps x 64),
//     A = gamma * invstd,  Bc = -A * dgamma_mean * invstd,  C = -A * dbeta_mean + A * dgamma_mean * mean * invstd
// (tag_bn_relu_pool_bwd mode 1), so neither the raw convolution output nor dy1 ever exists in HBM.  The transform runs on
// the A fragments in registers (the accumulator layout of y = the A layout of g); the wgrad B fragments are their 8x8
// transposes (movmatrix) instead of ldmatrix.trans of the untransformed shared-memory unit.
template <bool FUSE_BN>
__global__ void __launch_bounds__(256, FUSE_BN ? 1 : 2)
conv_c1_bwd_mma_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ w,
                       float* __restrict__ dw, float* __restrict__ dx, int B, int H,
                       const float* __restrict__ bn_scale, const float* __restrict__ bn_mean,
                       const float* __restrict__ bn_invstd, const double* __restrict__ bn_red, float inv_count,
                       int bn_training) {
    constexpr int XOFF = FUSE_BN ? 2 : 1;                 // the recomputation needs one more halo row of x each side
    constexpr int XROWS = MROWS + 2 * (XOFF - 1);
    extern __shared__ __align__(128) uint8_t msm[];
    uint8_t* stage = msm;                                                     // [8 warps][MNST][2048]
    float* S = reinterpret_cast<float*>(msm + 8 * MNST * 2048);               // [MROWS][64][9]
    uint16_t* xs = reinterpret_cast<uint16_t*>(S + MROWS * 64 * 9);           // [XROWS][MXS_W] bf16 bits
    float* s_dw = reinterpret_cast<float*>(xs + XROWS * MXS_W);               // [64][9]
    float* s_bn = s_dw + CO * 9;                                              // FUSE_BN: [3][64] = A | Bc | C
    const int tiles_h = (H + MTH - 1) / MTH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    constexpr int W = TW;

    for (int i = threadIdx.x; i < CO * 9; i += 256) s_dw[i] = 0.f;
    uint32_t wyh[8][2], wyl[8][2];                        // FUSE_BN: B fragments of y = xcol * w^T (k = tap, n = channel)
    if (FUSE_BN) {
        if (threadIdx.x < CO) {
            const int c = threadIdx.x;
            const float sc = bn_scale[c], mu = bn_mean[c], is = bn_invstd[c];
            const float dbe = bn_training ? (float)bn_red[c] * inv_count : 0.f;
            const float dga = bn_training ? (float)bn_red[CO + c] * inv_count : 0.f;
            s_bn[c] = sc;
            s_bn[CO + c] = -sc * dga * is;
            s_bn[2 * CO + c] = -sc * dbe + sc * dga * mu * is;
        }
        auto lo_y = [](float v) { return v - __bfloat162float(__float2bfloat16_rn(v)); };
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const float* wc = w + (n * 8 + g) * 9;
            const float w0 = wc[2 * t], w1 = wc[2 * t + 1], w8 = wc[8];
            wyh[n][0] = pack2(w0, w1);
            wyl[n][0] = pack2(lo_y(w0), lo_y(w1));
            wyh[n][1] = t == 0 ? pack2(w8, 0.f) : 0u;
            wyl[n][1] = t == 0 ? pack2(lo_y(w8), 0.f) : 0u;
        }
    }
    // B fragments of the dgrad product: w[co][tap], k = co, n = tap (second n-tile: tap 8 only).  The fp32 weights
    // enter as bf16 hi + lo pairs (two MMAs): a once-rounded weight is a SYSTEMATIC error that does not average out
    // over the 4 M pixels summed into the bn0 gradients (dgamma0 / dbeta0 cancel heavily: bn1 removes scale and shift)
    uint32_t wb[4][2][2], wl[4][2][2];
    auto lo_of = [](float v) { return v - __bfloat162float(__float2bfloat16_rn(v)); };
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int co = ks * 16 + hf * 8 + 2 * t;
            const float w0 = w[co * 9 + g], w1 = w[(co + 1) * 9 + g];
            const float v0 = w[co * 9 + 8], v1 = w[(co + 1) * 9 + 8];
            wb[ks][0][hf] = pack2(w0, w1);
            wl[ks][0][hf] = pack2(lo_of(w0), lo_of(w1));
            wb[ks][1][hf] = g == 0 ? pack2(v0, v1) : 0u;
            wl[ks][1][hf] = g == 0 ? pack2(lo_of(v0), lo_of(v1)) : 0u;
        }
    float acc[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[n][k] = 0.f;
    const uint32_t my_stage = smem_u32_c1(stage + warp * MNST * 2048);
    // persistent over tiles: the weight-gradient accumulators stay in registers and are flushed once per CTA
#pragma unroll 1
    for (int tile = blockIdx.x; tile < B * tiles_h; tile += gridDim.x) {
    const int b = tile / tiles_h, h0 = (tile % tiles_h) * MTH;
    __syncthreads();                                      // previous tile's dx phase is done with S / xs
    for (int i = threadIdx.x; i < XROWS * MXS_W; i += 256) {
        const int r = i / MXS_W, c = i - r * MXS_W;
        const int h = h0 - XOFF + r, wc = c - 1;
        uint16_t v = 0;
        if (c < 66 && h >= 0 && h < H && wc >= 0 && wc < W)
            v = reinterpret_cast<const uint16_t*>(x)[((long)b * H + h) * W + wc];
        xs[i] = v;
    }
    __syncthreads();

    constexpr int NUNITS = MROWS * 4;                     // units of 16 pixels
    constexpr int PER_WARP = (NUNITS + 7) / 8;
    auto unit_row_ok = [&](int u, int& r, int& c0) {
        r = u >> 2; c0 = (u & 3) * 16;
        const int h = h0 - 1 + r;
        return u < NUNITS && h >= 0 && h < H;
    };
    auto issue = [&](int i) {
        int r, c0;
        const int u = warp + 8 * i;
        if (i < PER_WARP && unit_row_ok(u, r, c0)) {
            const bf16* src = dy + (((long)b * H + (h0 - 1 + r)) * W + c0) * CO;
            const uint32_t dst = my_stage + (i % MNST) * 2048;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = j * 32 + lane;
                const int px = idx >> 3, ch = idx & 7;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + px * 128 + ((ch ^ (px & 7)) << 4)),
                             "l"(src + px * CO + ch * 8) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int i = 0; i < MNST - 1; ++i) issue(i);
    // ldmatrix lane address inside a unit: row = (lane & 7) + 8 * ((lane >> 3) & 1), 16-byte chunk = 2 * q + (lane >> 4)
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8;
    const int lchk = lane >> 4;
#pragma unroll 1
    for (int i = 0; i < PER_WARP; ++i) {
        issue(i + MNST - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(MNST - 1) : "memory");
        __syncwarp();
        int r, c0;
        const int u = warp + 8 * i;
        if (u >= NUNITS) break;
        const bool ok = unit_row_ok(u, r, c0);
        float* Srow = S + ((r * 64) + c0) * 9;
        if (!ok) {
            for (int k = lane; k < 16 * 9; k += 32) Srow[k] = 0.f;
            continue;
        }
        const uint32_t sbase = my_stage + (i % MNST) * 2048 + lrow * 128;
        uint32_t a[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) ldsm_x4(a[q], sbase + (((2 * q + lchk) ^ (lrow & 7)) << 4));
        if (FUSE_BN) {
            // y[16 pixels][64] = xcol[16][9 -> 16] * w^T: A rows = the unit's pixels, k = taps (x from the halo tile)
            uint32_t ay[4];
            {
                auto xat = [&](int m, int tap) -> uint32_t {
                    return xs[(r + tap / 3 - 2 + XOFF) * MXS_W + c0 + m + tap % 3];
                };
                ay[0] = xat(g, 2 * t) | (xat(g, 2 * t + 1) << 16);
                ay[1] = xat(g + 8, 2 * t) | (xat(g + 8, 2 * t + 1) << 16);
                ay[2] = t == 0 ? xat(g, 8) : 0u;
                ay[3] = t == 0 ? xat(g + 8, 8) : 0u;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const int n = 2 * q + hf;
                    float yv[4] = {0.f, 0.f, 0.f, 0.f};
                    mma16816(yv, ay, wyh[n][0], wyh[n][1]);
                    mma16816(yv, ay, wyl[n][0], wyl[n][1]);
                    const int ch = 16 * q + 8 * hf + 2 * t;
                    const float2 cA = *reinterpret_cast<const float2*>(s_bn + ch);
                    const float2 cB = *reinterpret_cast<const float2*>(s_bn + CO + ch);
                    const float2 cC = *reinterpret_cast<const float2*>(s_bn + 2 * CO + ch);
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {                 // rr = 0: pixel g, rr = 1: pixel g + 8
                        const uint32_t u = a[q][2 * hf + rr];
                        const float g0 = __uint_as_float(u << 16), g1 = __uint_as_float(u & 0xFFFF0000u);
                        const float d0 = fmaf(cA.x, g0, fmaf(cB.x, yv[2 * rr], cC.x));
                        const float d1 = fmaf(cA.y, g1, fmaf(cB.y, yv[2 * rr + 1], cC.y));
                        a[q][2 * hf + rr] = pack2(d0, d1);
                    }
                }
            }
        }
        float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            mma16816(s0, a[q], wb[q][0][0], wb[q][0][1]);
            mma16816(s1, a[q], wb[q][1][0], wb[q][1][1]);
            mma16816(s0, a[q], wl[q][0][0], wl[q][0][1]);
            mma16816(s1, a[q], wl[q][1][0], wl[q][1][1]);
        }
        Srow[g * 9 + 2 * t] = s0[0];
        Srow[g * 9 + 2 * t + 1] = s0[1];
        Srow[(g + 8) * 9 + 2 * t] = s0[2];
        Srow[(g + 8) * 9 + 2 * t + 1] = s0[3];
        if (t == 0) { Srow[g * 9 + 8] = s1[0]; Srow[(g + 8) * 9 + 8] = s1[2]; }
        if (r >= 1 && r <= MTH) {
            // A fragments of the wgrad product: rows = taps, k = the unit's 16 pixels; x[h + dh][c + dw] from the halo tile
            uint32_t ax[4];
            {
                const int dh = g / 3, dwv = g - dh * 3;                       // tap g (g <= 7): x row h - 1 + dh, col c + dwv
                const uint16_t* xr = xs + (r - 2 + XOFF + dh) * MXS_W + c0 + dwv + 2 * t;
                ax[0] = (uint32_t)xr[0] | ((uint32_t)xr[1] << 16);
                ax[2] = (uint32_t)xr[8] | ((uint32_t)xr[9] << 16);
                const uint16_t* x8 = xs + (r + XOFF) * MXS_W + c0 + 2 + 2 * t;  // tap 8 = (dh, dw) = (+1, +1)
                ax[1] = g == 0 ? ((uint32_t)x8[0] | ((uint32_t)x8[1] << 16)) : 0u;
                ax[3] = g == 0 ? ((uint32_t)x8[8] | ((uint32_t)x8[9] << 16)) : 0u;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t bt[4];
                if (FUSE_BN) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) bt[j] = movmatrix_trans(a[q][j]);
                } else {
                    ldsm_x4_trans(bt, sbase + (((2 * q + lchk) ^ (lrow & 7)) << 4));
                }
                mma16816(acc[2 * q], ax, bt[0], bt[1]);
                mma16816(acc[2 * q + 1], ax, bt[2], bt[3]);
            }
        }
        __syncwarp();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // ---- dx[p] = sum_tap S[p - d(tap)][tap]
    if (dx != nullptr) {
        for (int pix = threadIdx.x; pix < MTH * W; pix += 256) {
            const int r = pix / W, c = pix - r * W;
            const int h = h0 + r;
            if (h >= H) break;
            float v = 0.f;
#pragma unroll
            for (int tp = 0; tp < 9; ++tp) {
                const int dh = tp / 3 - 1, dwv = tp % 3 - 1;
                const int cc = c - dwv;
                if (cc >= 0 && cc < W) v += S[(((r + 1 - dh) * 64) + cc) * 9 + tp];
            }
            dx[((long)b * H + h) * W + c] = v;
        }
    }
    }
    // ---- wgrad: acc[n][0..1] = (tap g, co 8n + 2t, +1), acc[n][2..3] = (tap 8 when g == 0)
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        const int co = n * 8 + 2 * t;
        atomicAdd(&s_dw[co * 9 + g], acc[n][0]);
        atomicAdd(&s_dw[(co + 1) * 9 + g], acc[n][1]);
        if (g == 0) {
            atomicAdd(&s_dw[co * 9 + 8], acc[n][2]);
            atomicAdd(&s_dw[(co + 1) * 9 + 8], acc[n][3]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < CO * 9; i += 256) atomicAdd(dw + i, s_dw[i]);
}

// ---------------------------------------------------------------------------------------------
// bf16 forward of the one-pass layer on mma.sync: y[16 pixels][64] = xcol[16][9 -> 16] * w^T (weights as bf16 hi + lo:
// 16 HMMA per 16 pixels instead of 288 FFMA per lane — the CUDA-core stencil above is issue-bound at a quarter of the FMA
// rate), then relu(scale * y + shift) on the accumulators and a shared-memory transpose (XOR-swizzled 16-byte chunks) so
// that every global store is a full 128-byte pixel row.  HBM-bound target: 8.2 MB of bf16 output per 10 s clip.
constexpr int FWD_MMA_SMEM = 8 * 2048 + MROWS * MXS_W * 2;

__global__ void __launch_bounds__(256, 2)
conv_c1_fwd_act_mma_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bn_scale,
                           const float* __restrict__ bn_shift, bf16* __restrict__ y, int B, int H) {
    extern __shared__ __align__(128) uint8_t fsm[];
    uint8_t* stage = fsm;                                                     // [8 warps][16 rows][128 B]
    uint16_t* xs = reinterpret_cast<uint16_t*>(fsm + 8 * 2048);               // [MROWS][MXS_W] bf16 bits
    const int tiles_h = (H + MTH - 1) / MTH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    constexpr int W = TW;

    // B fragments of y = xcol * w^T (k = tap, n = channel) and the BatchNorm affine of this lane's 16 channels
    uint32_t wyh[8][2], wyl[8][2];
    float2 bsc[8], bsh[8];
    auto lo_y = [](float v) { return v - __bfloat162float(__float2bfloat16_rn(v)); };
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        const float* wc = w + (n * 8 + g) * 9;
        const float w0 = wc[2 * t], w1 = wc[2 * t + 1], w8 = wc[8];
        wyh[n][0] = pack2(w0, w1);
        wyl[n][0] = pack2(lo_y(w0), lo_y(w1));
        wyh[n][1] = t == 0 ? pack2(w8, 0.f) : 0u;
        wyl[n][1] = t == 0 ? pack2(lo_y(w8), 0.f) : 0u;
        bsc[n] = *reinterpret_cast<const float2*>(bn_scale + n * 8 + 2 * t);
        bsh[n] = *reinterpret_cast<const float2*>(bn_shift + n * 8 + 2 * t);
    }
    uint8_t* my_stage = stage + warp * 2048;
    // the halo tile of the NEXT tile is fetched into registers before this tile's arithmetic (x is L2-resident, 16 MB)
    constexpr int XN = (MROWS * MXS_W + 255) / 256;
    uint16_t xnext[XN];
    auto fetch_tile = [&](int tile) {
        const bool ok = tile < B * tiles_h;
        const int b = ok ? tile / tiles_h : 0, h0 = ok ? (tile % tiles_h) * MTH : 0;
#pragma unroll
        for (int j = 0; j < XN; ++j) {
            const int i = threadIdx.x + 256 * j;
            const int r = i / MXS_W, c = i - r * MXS_W;
            const int h = h0 - 1 + r, wc = c - 1;
            uint16_t v = 0;
            if (ok && i < MROWS * MXS_W && c < 66 && h >= 0 && h < H && wc >= 0 && wc < W)
                v = reinterpret_cast<const uint16_t*>(x)[((long)b * H + h) * W + wc];
            xnext[j] = v;
        }
    };
    fetch_tile(blockIdx.x);
#pragma unroll 1
    for (int tile = blockIdx.x; tile < B * tiles_h; tile += gridDim.x) {
        const int b = tile / tiles_h, h0 = (tile % tiles_h) * MTH;
        __syncthreads();                                  // the previous tile's readers of xs are done
#pragma unroll
        for (int j = 0; j < XN; ++j) {
            const int i = threadIdx.x + 256 * j;
            if (i < MROWS * MXS_W) xs[i] = xnext[j];
        }
        __syncthreads();
        fetch_tile(tile + gridDim.x);
#pragma unroll 1
        for (int u = warp; u < MTH * 4; u += 8) {         // units of 16 consecutive pixels of one row
            const int r = u >> 2, c0 = (u & 3) * 16;
            const int h = h0 + r;
            if (h >= H) break;
            uint32_t ay[4];
            {
                auto xat = [&](int m, int tap) -> uint32_t { return xs[(r + tap / 3) * MXS_W + c0 + m + tap % 3]; };
                ay[0] = xat(g, 2 * t) | (xat(g, 2 * t + 1) << 16);
                ay[1] = xat(g + 8, 2 * t) | (xat(g + 8, 2 * t + 1) << 16);
                ay[2] = t == 0 ? xat(g, 8) : 0u;
                ay[3] = t == 0 ? xat(g + 8, 8) : 0u;
            }
            __syncwarp();                                 // the previous unit's staging reads are done
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                float yv[4] = {0.f, 0.f, 0.f, 0.f};
                mma16816(yv, ay, wyh[n][0], wyh[n][1]);
                mma16816(yv, ay, wyl[n][0], wyl[n][1]);
                // accumulator: (pixel g, channels 8n + 2t, +1), (pixel g + 8, same channels)
                const uint32_t p0 = pack2(fmaxf(fmaf(yv[0], bsc[n].x, bsh[n].x), 0.f), fmaxf(fmaf(yv[1], bsc[n].y, bsh[n].y), 0.f));
                const uint32_t p1 = pack2(fmaxf(fmaf(yv[2], bsc[n].x, bsh[n].x), 0.f), fmaxf(fmaf(yv[3], bsc[n].y, bsh[n].y), 0.f));
                *reinterpret_cast<uint32_t*>(my_stage + g * 128 + ((n ^ g) << 4) + 4 * t) = p0;
                *reinterpret_cast<uint32_t*>(my_stage + (g + 8) * 128 + ((n ^ g) << 4) + 4 * t) = p1;
            }
            __syncwarp();
            bf16* dst = y + (((long)b * H + h) * W + c0) * CO;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = (lane >> 3) + 4 * i, ch = lane & 7;
                const uint4 v = *reinterpret_cast<const uint4*>(my_stage + row * 128 + ((ch ^ (row & 7)) << 4));
                st16(reinterpret_cast<uint8_t*>(dst + row * CO) + 16 * ch, v);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm statistics of the Cin = 1 convolution WITHOUT running it: y[p, co] = sum_k w[co][k] x[p + d_k], hence
//   sum_p y[p, co]   = sum_k w[co][k] * M1[k],            M1[k]     = sum_p x[p + d_k]
//   sum_p y[p, co]^2 = sum_{k,k'} w[co][k] w[co][k'] * M2[k][k'],  M2[k][k'] = sum_p x[p + d_k] x[p + d_k']
// over all output pixels p (zero padding).  The 9 + 45 moments of the 16 MB input cost one pass over it; the layer can
// then write relu(bn(conv)) in a single kernel (conv_c1_fwd_kernel<ACT>).  mom: [45 upper-triangular M2 | 9 M1] doubles.
constexpr int N_MOM = 54;

// After the call lane l holds the sum over the 32 lanes of a[l] (butterfly transpose-reduce, 31 shuffles).
__device__ __forceinline__ float c1_transpose_sum(float (&a)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = upper ? a[i] : a[i + off];
            const float keep = upper ? a[i + off] : a[i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return a[0];
}

// W == 64 (the model's mel axis): a thread owns one column and walks MOM_ROWS rows with a rolling 3x3 window, so every
// input element is loaded three times (coalesced 128-byte rows) instead of nine, with no index divisions.
constexpr int MOM_ROWS = 16;
template <typename T>
__global__ void __launch_bounds__(256)
c1_moments_w64_kernel(const T* __restrict__ x, int B, int H, double* __restrict__ mom) {
    constexpr int W = 64;
    __shared__ float s_m[8][64];               // per-warp totals, summed in warp order: no order-dependent fp32 atomics
    const int chunks = (H + 4 * MOM_ROWS - 1) / (4 * MOM_ROWS);
    const int b = blockIdx.x / chunks;
    const int h_begin = (blockIdx.x % chunks) * 4 * MOM_ROWS + (threadIdx.x >> 6) * MOM_ROWS;
    const int wq = threadIdx.x & 63;
    const T* xb = x + (long)b * H * W;
    auto load_row = [&](int hh, float (&r)[3]) {
        const bool ok = hh >= 0 && hh < H;
        const T* p = xb + (long)hh * W + wq;
        r[0] = (ok && wq > 0) ? to_f<T>(p[-1]) : 0.f;
        r[1] = ok ? to_f<T>(p[0]) : 0.f;
        r[2] = (ok && wq < W - 1) ? to_f<T>(p[1]) : 0.f;
    };
    float m[N_MOM];
#pragma unroll
    for (int i = 0; i < N_MOM; ++i) m[i] = 0.f;
    float xn[9];
    {
        float r0[3], r1[3];
        load_row(h_begin - 1, r0);
        load_row(h_begin, r1);
#pragma unroll
        for (int k = 0; k < 3; ++k) { xn[k] = 0.f; xn[3 + k] = r0[k]; xn[6 + k] = r1[k]; }
    }
#pragma unroll 2
    for (int h = h_begin; h < h_begin + MOM_ROWS; ++h) {
        float r2[3];
        load_row(h + 1, r2);
#pragma unroll
        for (int k = 0; k < 3; ++k) { xn[k] = xn[3 + k]; xn[3 + k] = xn[6 + k]; xn[6 + k] = r2[k]; }
        if (h < H) {
            int idx = 0;
#pragma unroll
            for (int k = 0; k < 9; ++k)
#pragma unroll
                for (int k2 = k; k2 < 9; ++k2) { m[idx] = fmaf(xn[k], xn[k2], m[idx]); ++idx; }
#pragma unroll
            for (int k = 0; k < 9; ++k) m[45 + k] += xn[k];
        }
    }
    // 54 sums over the warp as two 32-wide butterfly transpose-reductions (62 shuffles instead of 270): afterwards lane l
    // holds the warp totals of moments l and 32 + l
    float lo32[32], hi32[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { lo32[i] = m[i]; hi32[i] = 32 + i < N_MOM ? m[32 + i] : 0.f; }
    const int lane = threadIdx.x & 31;
    const float t0 = c1_transpose_sum(lo32, lane), t1 = c1_transpose_sum(hi32, lane);
    s_m[threadIdx.x >> 5][lane] = t0;
    s_m[threadIdx.x >> 5][32 + lane] = t1;
    __syncthreads();
    if (threadIdx.x < N_MOM) {
        double tot = 0.0;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) tot += (double)s_m[wv][threadIdx.x];
        atomicAdd(mom + threadIdx.x, tot);       // double: the order of the blocks matters only below fp32 resolution
    }
}

// stats[co] = sum y, stats[64 + co] = sum y^2 (the layout tag_bn_finalize reads) from the moments, in double
__global__ void c1_stats_from_moments_kernel(const double* __restrict__ mom, const float* __restrict__ w,
                                             double* __restrict__ stats) {
    const int co = threadIdx.x;
    if (co >= CO) return;
    double wk[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wk[k] = (double)w[co * 9 + k];
    double s1 = 0.0, s2 = 0.0;
    int idx = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        s1 += wk[k] * mom[45 + k];
#pragma unroll
        for (int k2 = k; k2 < 9; ++k2) { s2 += (k2 == k ? 1.0 : 2.0) * wk[k] * wk[k2] * mom[idx]; ++idx; }
    }
    stats[co] = s1;
    stats[CO + co] = s2;
}

// Reductions of the fused "ReLU gate + BatchNorm backward" epilogue (tag_conv_tc_fwd_halo with bn_act) are taken in the
// ACTIVATION domain: red[c] = sum g, red[C + c] = sum g * a with a = relu(gamma * xhat + beta).  g is zero wherever the
// gate is closed, and a = gamma * xhat + beta wherever it is open, so
//     sum g * xhat = (sum g * a - beta * sum g) / gamma        (the dgamma of the layer; red[c] is already its dbeta).
// gamma == 0 leaves no trace of xhat in a: the term is set to 0.
// sum_scale: the pooled form of the reduce delivers sum g / keep_scale of the dropout (a constant), applied here first.
// dgamma / dbeta (optional): the parameter gradients are accumulated in the same launch (tag_bn_param_grads).
__global__ void bn_red_act_to_xhat_kernel(double* __restrict__ red, const float* __restrict__ gamma,
                                          const float* __restrict__ beta, int C, float sum_scale,
                                          float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double gm = (double)gamma[c];
    const double sg = red[c] * (double)sum_scale;
    const double sx = gm != 0.0 ? (red[C + c] - (double)beta[c] * sg) / gm : 0.0;
    red[c] = sg;
    red[C + c] = sx;
    if (dbeta != nullptr) dbeta[c] += (float)sg;
    if (dgamma != nullptr) dgamma[c] += (float)sx;
}

}  // namespace

extern "C" int tag_c1_moments(const void* x, int dtype, int B, int H, int W, double* mom, cudaStream_t stream) {
    // the Cin = 1 layer only exists on the 64-bin mel axis (tag_conv_c1_* require W == 64)
    if (B <= 0 || H <= 0 || W != 64) return TAG_ERR_BAD_ARG;
    cudaError_t e = cudaMemsetAsync(mom, 0, N_MOM * sizeof(double), stream);
    if (e != cudaSuccess) return (int)e;
    {
        const int blocks64 = B * ((H + 4 * MOM_ROWS - 1) / (4 * MOM_ROWS));
        if (dtype == TAG_DTYPE_F32) c1_moments_w64_kernel<float><<<blocks64, 256, 0, stream>>>((const float*)x, B, H, mom);
        else c1_moments_w64_kernel<bf16><<<blocks64, 256, 0, stream>>>((const bf16*)x, B, H, mom);
    }
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_c1_stats_from_moments(const double* mom, const float* w, double* stats, cudaStream_t stream) {
    c1_stats_from_moments_kernel<<<1, 64, 0, stream>>>(mom, w, stats);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_bn_red_act_to_xhat(double* red, const float* gamma, const float* beta, int C, float sum_scale,
                                      float* dgamma, float* dbeta, cudaStream_t stream) {
    if (C <= 0) return TAG_ERR_BAD_ARG;
    bn_red_act_to_xhat_kernel<<<(C + 127) / 128, 128, 0, stream>>>(red, gamma, beta, C, sum_scale, dgamma, dbeta);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_conv_c1_fwd_act(const void* x, const float* w, const float* scale, const float* shift, void* y,
                                   int dtype, int B, int H, int W, cudaStream_t stream) {
    if (W != TW || B <= 0 || H <= 0) return TAG_ERR_BAD_ARG;
    int blocks = B * ((H + TH - 1) / TH);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (dtype == TAG_DTYPE_F32)
        conv_c1_fwd_kernel<float, true><<<blocks, 256, 0, stream>>>((const float*)x, w, (float*)y, nullptr, B, H, W, scale, shift);
    else {
        int mblocks = B * ((H + MTH - 1) / MTH);
        if (mblocks > 148 * 2) mblocks = 148 * 2;
        conv_c1_fwd_act_mma_kernel<<<mblocks, 256, FWD_MMA_SMEM, stream>>>((const bf16*)x, w, scale, shift, (bf16*)y, B, H);
    }
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_conv_c1_bwd_bn(const void* g, const void* x, const float* w, const float* scale, const float* mean,
                                  const float* invstd, const double* red, int bn_training, float* dw, float* dx, int B,
                                  int H, int W, cudaStream_t stream) {
    if (W != TW || B <= 0 || H <= 0) return TAG_ERR_BAD_ARG;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_c1_bwd_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MMA_SMEM_FUSE);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    int mblocks = B * ((H + MTH - 1) / MTH);
    if (mblocks > 148) mblocks = 148;
    const float inv_count = 1.0f / (float)((double)B * H * W);
    conv_c1_bwd_mma_kernel<true><<<mblocks, 256, MMA_SMEM_FUSE, stream>>>((const bf16*)g, (const bf16*)x, w, dw, dx, B, H, scale,
                                                                          mean, invstd, red, inv_count, bn_training);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_conv_c1_fwd(const void* x, const float* w, void* y, int dtype, double* stats, int B,
                               int H, int W, cudaStream_t stream) {
    if (W != TW || B <= 0 || H <= 0) return TAG_ERR_BAD_ARG;
    int blocks = B * ((H + TH - 1) / TH);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (dtype == TAG_DTYPE_F32)
        conv_c1_fwd_kernel<float, false><<<blocks, 256, 0, stream>>>((const float*)x, w, (float*)y, stats, B, H, W, nullptr, nullptr);
    else
        conv_c1_fwd_kernel<bf16, false><<<blocks, 256, 0, stream>>>((const bf16*)x, w, (bf16*)y, stats, B, H, W, nullptr, nullptr);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_conv_c1_bwd(const void* dy, const void* x, const float* w, int dtype, float* dw,
                               float* dx, int B, int H, int W, cudaStream_t stream) {
    if (W != TW || B <= 0 || H <= 0) return TAG_ERR_BAD_ARG;
    const int blocks = B * ((H + TH - 1) / TH);
    if (dtype == TAG_DTYPE_F32) {
        conv_c1_bwd_kernel<float, true, false><<<blocks, 256, 0, stream>>>((const float*)dy, (const float*)x, w, dw, dx, B, H, W);
        if (dx != nullptr)
            conv_c1_bwd_kernel<float, false, true><<<blocks, 256, 0, stream>>>((const float*)dy, (const float*)x, w, dw, dx, B, H, W);
    } else {
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(conv_c1_bwd_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MMA_SMEM);
            if (e != cudaSuccess) return (int)e;
            attr_set = true;
        }
        int mblocks = B * ((H + MTH - 1) / MTH);
        if (mblocks > 148 * 2) mblocks = 148 * 2;
        conv_c1_bwd_mma_kernel<false><<<mblocks, 256, MMA_SMEM, stream>>>((const bf16*)dy, (const bf16*)x, w, dw, dx, B, H,
                                                                          nullptr, nullptr, nullptr, nullptr, 0.f, 0);
    }
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}
