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
    const int b = blockIdx.x / tiles_h, h0 = (blockIdx.x % tiles_h) * TH;
    const int cg = threadIdx.x & 7, pl = threadIdx.x >> 3;        // 8 channel groups x 32 pixels
    if (threadIdx.x < CO) { s_sum[threadIdx.x] = 0.f; s_sq[threadIdx.x] = 0.f; }
    load_x_tile<T>(xs, x, b, h0, H, W);
    float wr[8][9];
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int t = 0; t < 9; ++t) wr[c][t] = __ldg(w + (cg * 8 + c) * 9 + t);
    __syncthreads();
    float cs[8], cq[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { cs[c] = 0.f; cq[c] = 0.f; }
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

}  // namespace

extern "C" int tag_conv_c1_fwd(const void* x, const float* w, void* y, int dtype, double* stats, int B,
                               int H, int W, cudaStream_t stream) {
    if (W != TW || B <= 0 || H <= 0) return TAG_ERR_BAD_ARG;
    const int blocks = B * ((H + TH - 1) / TH);
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
        conv_c1_bwd_kernel<bf16, true, false><<<blocks, 256, 0, stream>>>((const bf16*)dy, (const bf16*)x, w, dw, dx, B, H, W);
        if (dx != nullptr)
            conv_c1_bwd_kernel<bf16, false, true><<<blocks, 256, 0, stream>>>((const bf16*)dy, (const bf16*)x, w, dw, dx, B, H, W);
    }
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}
