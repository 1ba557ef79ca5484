// fp32 -> split-bf16 GEMM operands ("bf16x3"): x = hi + lo with hi = bf16(x), lo = bf16(x - hi) keeps 16 mantissa
// bits, and  a.b ~= a_hi.b_hi + a_hi.b_lo + a_lo.b_hi  (the dropped lo.lo term is 2^-16 relative).  Concatenating the
// three products along K turns them into ONE bf16 tensor-core GEMM with K' = 3K and fp32 accumulation in TMEM:
//     A' = [a_hi | a_hi | a_lo]   (mode 0)        B' = [b_hi | b_lo | b_hi]   (mode 1)
// so the fp32 heads (all-pairs alignment scores, attention projections) run on tcgen05 (tag_conv_tc_fwd, taps = 1)
// with ~1e-5 relative error instead of CUDA-core FMA.  Mode 2 writes the planes [hi; lo] separately for the
// weight-gradient form (three tag_conv_tc_wgrad calls accumulate hi.hi + hi.lo + lo.hi).
#include "common.cuh"

namespace {

__device__ __forceinline__ void split1(float x, bf16& hi, bf16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// in [rows][K] (transpose = 0) or [K][rows] (transpose = 1); out rows-major
__global__ void split_bf16x3_kernel(const float* __restrict__ in, bf16* __restrict__ out, long rows, int K, int mode,
                                    int transpose) {
    const long n = rows * K;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        long r; int k;
        float x;
        if (transpose) { k = (int)(i / rows); r = i - (long)k * rows; x = in[i]; }     // coalesced read along rows
        else { r = i / K; k = (int)(i - r * K); x = in[i]; }
        bf16 hi, lo;
        split1(x, hi, lo);
        if (mode == 2) {
            out[r * K + k] = hi;
            out[n + r * K + k] = lo;
        } else {
            bf16* o = out + r * 3 * K + k;
            o[0] = hi;
            o[K] = mode == 0 ? hi : lo;
            o[2 * K] = mode == 0 ? lo : hi;
        }
    }
}

}  // namespace

extern "C" int tag_split_bf16x3(const float* in, void* out, long rows, int K, int mode, int transpose,
                                cudaStream_t stream) {
    if (rows <= 0 || K <= 0 || mode < 0 || mode > 2) return TAG_ERR_BAD_ARG;
    long blocks = (rows * K + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    split_bf16x3_kernel<<<(int)blocks, 256, 0, stream>>>(in, (bf16*)out, rows, K, mode, transpose);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}
