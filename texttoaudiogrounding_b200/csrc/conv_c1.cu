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

template <typename T>
__global__ void __launch_bounds__(256)
conv_c1_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, T* __restrict__ y,
                   double* __restrict__ stats, int B, int H, int W) {
    __shared__ float xs[TH + 2][XS_W];
    __shared__ float s_sum[CO], s_sq[CO];
    const int tiles_h = (H + TH - 1) / TH;
    const int cg = threadIdx.x & 7, pl = threadIdx.x >> 3;        // 8 channel groups x 32 pixels
    if (threadIdx.x < CO) { s_sum[threadIdx.x] = 0.f; s_sq[threadIdx.x] = 0.f; }
    float wr[8][9];
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int t = 0; t < 9; ++t) wr[c][t] = __ldg(w + (cg * 8 + c) * 9 + t);
    float cs[8], cq[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { cs[c] = 0.f; cq[c] = 0.f; }
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
            a = round_to<T>(a);
            o[c] = a;
            cs[c] += a;
            cq[c] += a * a;
        }
        store8<T>(y + (((long)b * H + h) * W + c0) * CO + cg * 8, o);
    }
    }
    if (stats != nullptr) {
        // lanes with equal (lane & 7) share the channel group: reduce over lane bits 3,4
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float sv = cs[c], qv = cq[c];
            sv += __shfl_xor_sync(0xffffffffu, sv, 8);  qv += __shfl_xor_sync(0xffffffffu, qv, 8);
            sv += __shfl_xor_sync(0xffffffffu, sv, 16); qv += __shfl_xor_sync(0xffffffffu, qv, 16);
            if ((threadIdx.x & 31) < 8) {
                atomicAdd(&s_sum[cg * 8 + c], sv);
                atomicAdd(&s_sq[cg * 8 + c], qv);
            }
        }
        __syncthreads();
        if (threadIdx.x < CO) {
            atomicAdd(stats + threadIdx.x, (double)s_sum[threadIdx.x]);
            atomicAdd(stats + CO + threadIdx.x, (double)s_sq[threadIdx.x]);
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

__global__ void __launch_bounds__(256, 2)
conv_c1_bwd_mma_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ w,
                       float* __restrict__ dw, float* __restrict__ dx, int B, int H) {
    extern __shared__ __align__(128) uint8_t msm[];
    uint8_t* stage = msm;                                                     // [8 warps][MNST][2048]
    float* S = reinterpret_cast<float*>(msm + 8 * MNST * 2048);               // [MROWS][64][9]
    uint16_t* xs = reinterpret_cast<uint16_t*>(S + MROWS * 64 * 9);           // [MROWS][MXS_W] bf16 bits
    float* s_dw = reinterpret_cast<float*>(xs + MROWS * MXS_W);               // [64][9]
    const int tiles_h = (H + MTH - 1) / MTH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    constexpr int W = TW;

    for (int i = threadIdx.x; i < CO * 9; i += 256) s_dw[i] = 0.f;
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
    for (int i = threadIdx.x; i < MROWS * MXS_W; i += 256) {
        const int r = i / MXS_W, c = i - r * MXS_W;
        const int h = h0 - 1 + r, wc = c - 1;
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
                const int dh = g / 3, dwv = g - dh * 3;                       // tap g (g <= 7): xs row r - 1 + dh, col c + dwv
                const uint16_t* xr = xs + (r - 1 + dh) * MXS_W + c0 + dwv + 2 * t;
                ax[0] = (uint32_t)xr[0] | ((uint32_t)xr[1] << 16);
                ax[2] = (uint32_t)xr[8] | ((uint32_t)xr[9] << 16);
                const uint16_t* x8 = xs + (r + 1) * MXS_W + c0 + 2 + 2 * t;  // tap 8 = (dh, dw) = (+1, +1)
                ax[1] = g == 0 ? ((uint32_t)x8[0] | ((uint32_t)x8[1] << 16)) : 0u;
                ax[3] = g == 0 ? ((uint32_t)x8[8] | ((uint32_t)x8[9] << 16)) : 0u;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t bt[4];
                ldsm_x4_trans(bt, sbase + (((2 * q + lchk) ^ (lrow & 7)) << 4));
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

}  // namespace

extern "C" int tag_conv_c1_fwd(const void* x, const float* w, void* y, int dtype, double* stats, int B,
                               int H, int W, cudaStream_t stream) {
    if (W != TW || B <= 0 || H <= 0) return TAG_ERR_BAD_ARG;
    int blocks = B * ((H + TH - 1) / TH);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (dtype == TAG_DTYPE_F32)
        conv_c1_fwd_kernel<float><<<blocks, 256, 0, stream>>>((const float*)x, w, (float*)y, stats, B, H, W);
    else
        conv_c1_fwd_kernel<bf16><<<blocks, 256, 0, stream>>>((const bf16*)x, w, (bf16*)y, stats, B, H, W);
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
            cudaError_t e = cudaFuncSetAttribute(conv_c1_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MMA_SMEM);
            if (e != cudaSuccess) return (int)e;
            attr_set = true;
        }
        int mblocks = B * ((H + MTH - 1) / MTH);
        if (mblocks > 148 * 2) mblocks = 148 * 2;
        conv_c1_bwd_mma_kernel<<<mblocks, 256, MMA_SMEM, stream>>>((const bf16*)dy, (const bf16*)x, w, dw, dx, B, H);
    }
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}
