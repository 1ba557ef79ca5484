// fp32-accumulate CUDA-core implicit-GEMM kernels over NHWC activations.
//
// These are the exact-arithmetic ("fp32" precision mode) implementation of the dense
// contractions of the hot path and the small GEMMs (fc1, GRU projections):
//   * tag_conv_fwd   : y[p, co] = sum_{tap,ci} x[p + d(tap), ci] * w[co][tap][ci]
//                      (F.conv2d 3x3 pad 1, reference models/panns.py:49-50; with taps=1 it is
//                      the plain x @ W^T of fc1 / GRU input projection, audio_encoder.py:216-217)
//                      The same kernel computes dgrad when fed dy and the flipped+transposed
//                      weights produced by tag_weight_flip_transpose.
//   * tag_conv_wgrad : dw[co][tap][ci] += sum_p dy[p, co] * x[p + d(tap), ci]
// The bf16 tcgen05 path lives in conv_tc.cu; both share layouts: activations NHWC,
// weights [Cout][tap][Cin] (== a channels_last view of the reference's [Cout,Cin,3,3]).
#include "common.cuh"

namespace {

constexpr int BK = 16;
constexpr int NTHREADS = 256;

// ---------------------------------------------------------------------------------------
// forward / dgrad
// ---------------------------------------------------------------------------------------
template <typename TI, typename TO, int BN, int TAPS>
__global__ void __launch_bounds__(NTHREADS)
conv_fwd_kernel(const TI* __restrict__ x, const float* __restrict__ w, TO* __restrict__ y,
                const float* __restrict__ bias, int relu, double* __restrict__ stats,
                int B, int H, int W, int Cin, int Cout) {
    constexpr int BM = 128;
    constexpr int TN = BN / 16;
    constexpr int B_LOADS = (BN * BK / 4) / NTHREADS;   // float4 loads per thread for B
    __shared__ __align__(16) float As[BK][BM];
    __shared__ __align__(16) float Bs[BK][BN];
    __shared__ float s_sum[NTHREADS / 32][BN], s_sq[NTHREADS / 32][BN];     // per-warp totals (every warp covers all BN columns)

    const int tid = threadIdx.x;
    const long M = (long)B * H * W;
    const long m0 = (long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int K = TAPS * Cin;

    // A loader: this thread always serves tile row a_row, k-quads {a_kq, a_kq + 2}
    const int a_row = tid & (BM - 1);
    const int a_kq = tid >> 7;
    const long pm = m0 + a_row;
    const bool row_ok = pm < M;
    int pb = 0, ph = 0, pw = 0;
    if (row_ok) {
        pw = (int)(pm % W);
        long t = pm / W;
        ph = (int)(t % H);
        pb = (int)(t / H);
    }
    // B loader
    const int b_row = tid % BN;
    const int b_kq0 = tid / BN;

    float a_reg[2][4], b_reg[B_LOADS][4];
    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    auto load_global = [&](int k0) {
        const int tap = (TAPS == 1) ? 0 : k0 / Cin;
        const int c0 = k0 - tap * Cin;
        const int dh = (TAPS == 1) ? 0 : tap / 3 - 1;
        const int dw = (TAPS == 1) ? 0 : tap % 3 - 1;
        const int hh = ph + dh, ww = pw + dw;
        const bool ok = row_ok && hh >= 0 && hh < H && ww >= 0 && ww < W;
        const TI* src = x + (((long)pb * H + hh) * W + ww) * Cin + c0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int kq = a_kq + 2 * i;
            if (ok) {
                Vec4<TI>::load(src + kq * 4, a_reg[i]);
            } else {
                a_reg[i][0] = a_reg[i][1] = a_reg[i][2] = a_reg[i][3] = 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < B_LOADS; ++i) {
            const int kq = b_kq0 + (NTHREADS / BN) * i;
            Vec4<float>::load(w + (long)(n0 + b_row) * K + k0 + kq * 4, b_reg[i]);
        }
    };
    auto store_smem = [&]() {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int kq = a_kq + 2 * i;
#pragma unroll
            for (int e = 0; e < 4; ++e) As[kq * 4 + e][a_row] = a_reg[i][e];
        }
#pragma unroll
        for (int i = 0; i < B_LOADS; ++i) {
            const int kq = b_kq0 + (NTHREADS / BN) * i;
#pragma unroll
            for (int e = 0; e < 4; ++e) Bs[kq * 4 + e][b_row] = b_reg[i][e];
        }
    };

    const int tn = tid & 15, tm = tid >> 4;
    const int nk = K / BK;
    load_global(0);
    store_smem();
    __syncthreads();
    for (int kb = 0; kb < nk; ++kb) {
        if (kb + 1 < nk) load_global((kb + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[8], bv[TN];
            *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[k][tm * 8]);
            *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[k][tm * 8 + 4]);
#pragma unroll
            for (int jj = 0; jj < TN / 4; ++jj)
                *reinterpret_cast<float4*>(bv + 4 * jj) =
                    *reinterpret_cast<const float4*>(&Bs[k][jj * 64 + tn * 4]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
        }
        __syncthreads();
        if (kb + 1 < nk) {
            store_smem();
            __syncthreads();
        }
    }

    // ---- epilogue: bias / relu / round / per-channel statistics / store
#pragma unroll
    for (int jj = 0; jj < TN / 4; ++jj) {
        const int nl = jj * 64 + tn * 4;
        float bsv[4] = {0.f, 0.f, 0.f, 0.f};
        if (bias != nullptr) Vec4<float>::load(bias + n0 + nl, bsv);
        float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long m = m0 + tm * 8 + i;
            if (m < M) {
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float v = acc[i][jj * 4 + e] + bsv[e];
                    if (relu) v = fmaxf(v, 0.f);
                    v = round_to<TO>(v);
                    o[e] = v;
                    cs[e] += v;
                    cq[e] += v * v;
                }
                Vec4<TO>::store(y + m * Cout + n0 + nl, o);
            }
        }
        if (stats != nullptr) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float sv = cs[e] + __shfl_xor_sync(0xffffffffu, cs[e], 16);
                float qv = cq[e] + __shfl_xor_sync(0xffffffffu, cq[e], 16);
                if ((tid & 31) < 16) {                 // written, not accumulated: summed in warp order below
                    s_sum[tid >> 5][nl + e] = sv;
                    s_sq[tid >> 5][nl + e] = qv;
                }
            }
        }
    }
    if (stats != nullptr) {
        __syncthreads();
        if (tid < BN) {
            double ts = 0.0, tq = 0.0;
#pragma unroll
            for (int wv = 0; wv < NTHREADS / 32; ++wv) { ts += (double)s_sum[wv][tid]; tq += (double)s_sq[wv][tid]; }
            atomicAdd(stats + n0 + tid, ts);
            atomicAdd(stats + Cout + n0 + tid, tq);
        }
    }
}

// ---------------------------------------------------------------------------------------
// wgrad (split over pixels; fp32 atomics into a zero-initialised dw)
// ---------------------------------------------------------------------------------------
template <typename TA, typename TB, int BN, int TAPS>
__global__ void __launch_bounds__(NTHREADS)
conv_wgrad_kernel(const TA* __restrict__ dy, const TB* __restrict__ x, float* __restrict__ dw,
                  int B, int H, int W, int Cin, int Cout, long pix_per_split) {
    constexpr int BM = 64;
    constexpr int TM = 4;
    constexpr int TN = BN / 16;
    constexpr int B_LOADS = (BN * BK / 4) / NTHREADS;
    __shared__ __align__(16) float As[BK][BM];
    __shared__ __align__(16) float Bs[BK][BN];

    const int tid = threadIdx.x;
    const long P = (long)B * H * W;
    const int co0 = blockIdx.x * BM;
    const int n_ci_tiles = Cin / BN;
    const int tap = blockIdx.y / n_ci_tiles;
    const int ci0 = (blockIdx.y % n_ci_tiles) * BN;
    const int dh = (TAPS == 1) ? 0 : tap / 3 - 1;
    const int dw_ = (TAPS == 1) ? 0 : tap % 3 - 1;
    const long p_begin = (long)blockIdx.z * pix_per_split;
    const long p_end = min(P, p_begin + pix_per_split);
    if (p_begin >= p_end) return;

    // A: BK rows of BM/4 = 16 quads -> 256 quads, one per thread
    const int a_k = tid >> 4, a_q = tid & 15;
    // B: BK rows of BN/4 quads
    constexpr int BQ = BN / 4;

    float a_reg[4], b_reg[B_LOADS][4];
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    auto load_global = [&](long p0) {
        {
            const long p = p0 + a_k;
            if (p < p_end) {
                Vec4<TA>::load(dy + p * Cout + co0 + a_q * 4, a_reg);
            } else {
                a_reg[0] = a_reg[1] = a_reg[2] = a_reg[3] = 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < B_LOADS; ++i) {
            const int q = tid + NTHREADS * i;
            const int k = q / BQ, nq = q % BQ;
            const long p = p0 + k;
            bool ok = p < p_end;
            long src_pix = p;
            if (ok && TAPS != 1) {
                int pw = (int)(p % W);
                long t = p / W;
                int ph = (int)(t % H);
                int hh = ph + dh, ww = pw + dw_;
                ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
                src_pix = p + (long)dh * W + dw_;
            }
            if (ok) {
                Vec4<TB>::load(x + src_pix * Cin + ci0 + nq * 4, b_reg[i]);
            } else {
                b_reg[i][0] = b_reg[i][1] = b_reg[i][2] = b_reg[i][3] = 0.f;
            }
        }
    };
    auto store_smem = [&]() {
        *reinterpret_cast<float4*>(&As[a_k][a_q * 4]) = make_float4(a_reg[0], a_reg[1], a_reg[2], a_reg[3]);
#pragma unroll
        for (int i = 0; i < B_LOADS; ++i) {
            const int q = tid + NTHREADS * i;
            const int k = q / BQ, nq = q % BQ;
            *reinterpret_cast<float4*>(&Bs[k][nq * 4]) =
                make_float4(b_reg[i][0], b_reg[i][1], b_reg[i][2], b_reg[i][3]);
        }
    };

    const int tn = tid & 15, tm = tid >> 4;
    load_global(p_begin);
    store_smem();
    __syncthreads();
    for (long p0 = p_begin; p0 < p_end; p0 += BK) {
        const bool more = p0 + BK < p_end;
        if (more) load_global(p0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], bv[TN];
            *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[k][tm * 4]);
#pragma unroll
            for (int jj = 0; jj < TN / 4; ++jj)
                *reinterpret_cast<float4*>(bv + 4 * jj) =
                    *reinterpret_cast<const float4*>(&Bs[k][jj * 64 + tn * 4]);
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
        }
        __syncthreads();
        if (more) {
            store_smem();
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int co = co0 + tm * 4 + i;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int ci = ci0 + (j / 4) * 64 + tn * 4 + (j % 4);
            atomicAdd(dw + ((long)co * TAPS + tap) * Cin + ci, acc[i][j]);
        }
    }
}

__global__ void weight_flip_transpose_kernel(const float* __restrict__ w, float* __restrict__ wt,
                                             int Co, int Ci, int T) {
    const long n = (long)Co * Ci * T;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        // i indexes wt[ci][t][co]
        int co = (int)(i % Co);
        long r = i / Co;
        int t = (int)(r % T);
        int ci = (int)(r / T);
        wt[i] = w[((long)co * T + (T - 1 - t)) * Ci + ci];
    }
}

template <typename TI, typename TO, int BN, int TAPS>
int launch_fwd(const void* x, const float* w, void* y, const float* bias, int relu, double* stats,
               int B, int H, int W, int Cin, int Cout, cudaStream_t stream) {
    const long M = (long)B * H * W;
    dim3 grid((unsigned)((M + 127) / 128), Cout / BN);
    conv_fwd_kernel<TI, TO, BN, TAPS><<<grid, NTHREADS, 0, stream>>>(
        (const TI*)x, w, (TO*)y, bias, relu, stats, B, H, W, Cin, Cout);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

template <typename TI, typename TO>
int dispatch_fwd(const void* x, const float* w, void* y, const float* bias, int relu, double* stats,
                 int B, int H, int W, int Cin, int Cout, int taps, cudaStream_t stream) {
    const bool wide = (Cout % 128 == 0);
    if (taps == 9) {
        return wide ? launch_fwd<TI, TO, 128, 9>(x, w, y, bias, relu, stats, B, H, W, Cin, Cout, stream)
                    : launch_fwd<TI, TO, 64, 9>(x, w, y, bias, relu, stats, B, H, W, Cin, Cout, stream);
    }
    return wide ? launch_fwd<TI, TO, 128, 1>(x, w, y, bias, relu, stats, B, H, W, Cin, Cout, stream)
                : launch_fwd<TI, TO, 64, 1>(x, w, y, bias, relu, stats, B, H, W, Cin, Cout, stream);
}

template <typename TA, typename TB, int BN, int TAPS>
int launch_wgrad(const void* dy, const void* x, float* dw, int B, int H, int W, int Cin, int Cout,
                 int splits, cudaStream_t stream) {
    const long P = (long)B * H * W;
    long pps = (P + splits - 1) / splits;
    pps = (pps + BK - 1) / BK * BK;
    dim3 grid(Cout / 64, (Cin / BN) * TAPS, splits);
    conv_wgrad_kernel<TA, TB, BN, TAPS><<<grid, NTHREADS, 0, stream>>>(
        (const TA*)dy, (const TB*)x, dw, B, H, W, Cin, Cout, pps);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

template <typename TA, typename TB>
int dispatch_wgrad(const void* dy, const void* x, float* dw, int B, int H, int W, int Cin, int Cout,
                   int taps, int splits, cudaStream_t stream) {
    const bool wide = (Cin % 128 == 0);
    if (taps == 9) {
        return wide ? launch_wgrad<TA, TB, 128, 9>(dy, x, dw, B, H, W, Cin, Cout, splits, stream)
                    : launch_wgrad<TA, TB, 64, 9>(dy, x, dw, B, H, W, Cin, Cout, splits, stream);
    }
    return wide ? launch_wgrad<TA, TB, 128, 1>(dy, x, dw, B, H, W, Cin, Cout, splits, stream)
                : launch_wgrad<TA, TB, 64, 1>(dy, x, dw, B, H, W, Cin, Cout, splits, stream);
}

}  // namespace

extern "C" int tag_conv_fwd(const void* x, int x_dtype, const float* w, void* y, int y_dtype,
                            const float* bias, int relu, double* stats, int B, int H, int W,
                            int Cin, int Cout, int taps, cudaStream_t stream) {
    if ((taps != 1 && taps != 9) || Cin % 16 != 0 || Cout % 64 != 0 || B <= 0 || H <= 0 || W <= 0)
        return TAG_ERR_BAD_ARG;
    if (x_dtype == TAG_DTYPE_F32 && y_dtype == TAG_DTYPE_F32)
        return dispatch_fwd<float, float>(x, w, y, bias, relu, stats, B, H, W, Cin, Cout, taps, stream);
    if (x_dtype == TAG_DTYPE_BF16 && y_dtype == TAG_DTYPE_BF16)
        return dispatch_fwd<bf16, bf16>(x, w, y, bias, relu, stats, B, H, W, Cin, Cout, taps, stream);
    if (x_dtype == TAG_DTYPE_BF16 && y_dtype == TAG_DTYPE_F32)
        return dispatch_fwd<bf16, float>(x, w, y, bias, relu, stats, B, H, W, Cin, Cout, taps, stream);
    if (x_dtype == TAG_DTYPE_F32 && y_dtype == TAG_DTYPE_BF16)
        return dispatch_fwd<float, bf16>(x, w, y, bias, relu, stats, B, H, W, Cin, Cout, taps, stream);
    return TAG_ERR_BAD_ARG;
}

extern "C" int tag_conv_wgrad(const void* dy, int dy_dtype, const void* x, int x_dtype, float* dw,
                              int B, int H, int W, int Cin, int Cout, int taps, int splits,
                              cudaStream_t stream) {
    if ((taps != 1 && taps != 9) || Cin % 64 != 0 || Cout % 64 != 0 || splits <= 0 || splits > 65535)
        return TAG_ERR_BAD_ARG;
    if (dy_dtype == TAG_DTYPE_F32 && x_dtype == TAG_DTYPE_F32)
        return dispatch_wgrad<float, float>(dy, x, dw, B, H, W, Cin, Cout, taps, splits, stream);
    if (dy_dtype == TAG_DTYPE_BF16 && x_dtype == TAG_DTYPE_BF16)
        return dispatch_wgrad<bf16, bf16>(dy, x, dw, B, H, W, Cin, Cout, taps, splits, stream);
    if (dy_dtype == TAG_DTYPE_F32 && x_dtype == TAG_DTYPE_BF16)
        return dispatch_wgrad<float, bf16>(dy, x, dw, B, H, W, Cin, Cout, taps, splits, stream);
    if (dy_dtype == TAG_DTYPE_BF16 && x_dtype == TAG_DTYPE_F32)
        return dispatch_wgrad<bf16, float>(dy, x, dw, B, H, W, Cin, Cout, taps, splits, stream);
    return TAG_ERR_BAD_ARG;
}

extern "C" int tag_weight_flip_transpose(const float* w, float* wt, int Co, int Ci, int taps,
                                         cudaStream_t stream) {
    const long n = (long)Co * Ci * taps;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 4096) blocks = 4096;
    weight_flip_transpose_kernel<<<blocks, 256, 0, stream>>>(w, wt, Co, Ci, taps);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}
