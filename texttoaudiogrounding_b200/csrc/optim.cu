// Flat-buffer optimizer step: global L2 norm -> clip_grad_norm_ -> Adam, plus small utilities.
//
// Reference: torch.nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm) followed by
// torch.optim.Adam.step() (reference python_scripts/training/run_strong.py:143-145).  All
// parameters live in one contiguous fp32 buffer, so the whole step is two HBM-bound passes;
// the 1/world_size of the data-parallel gradient mean is folded in as grad_mult.
#include "common.cuh"

namespace {

__global__ void sumsq_kernel(const float* __restrict__ g, long n, double* __restrict__ out) {
    float acc = 0.f;
    const long n4 = n / 4;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(g)[i];
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (long i = n4 * 4; i < n; ++i) acc += g[i] * g[i];
    __shared__ float red[32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < (blockDim.x >> 5); ++i) s += red[i];
        atomicAdd(out, s);
    }
}

__global__ void step_inc_kernel(long long* step) { *step += 1; }

__global__ void clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long n, const double* __restrict__ sumsq,
                                 const long long* __restrict__ step_ptr, float grad_mult, float max_norm,
                                 float lr, const float* __restrict__ lr_dev, float beta1, float beta2, float eps,
                                 float* __restrict__ norm_out) {
    if (lr_dev != nullptr) lr = *lr_dev;      // device-resident learning rate: schedulers act on a replayed graph
    const double total_norm = sqrt(*sumsq) * (double)grad_mult;
    float coef = 1.0f;
    if (max_norm > 0.f) coef = fminf((float)((double)max_norm / (total_norm + 1e-6)), 1.0f);
    const float gm = grad_mult * coef;
    const double step = (double)(*step_ptr);
    const float bc1 = (float)(1.0 - pow((double)beta1, step));
    const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, step));
    const float step_size = lr / bc1;
    if (norm_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = (float)total_norm;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float gi = g[i] * gm;
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] -= step_size * mi / denom;
    }
}

__global__ void cast_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, long n) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        y[i] = __float2bfloat16_rn(x[i]);
}

}  // namespace

extern "C" int tag_sumsq(const float* g, long n, double* out, cudaStream_t stream) {
    long blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    sumsq_kernel<<<(int)blocks, 256, 0, stream>>>(g, n, out);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

// step_ptr is incremented first (device-side counter, so the step replays inside a CUDA graph)
extern "C" int tag_clip_adam(float* p, const float* g, float* m, float* v, long n, const double* sumsq,
                             long long* step_ptr, float grad_mult, float max_norm, float lr, const float* lr_dev,
                             float beta1, float beta2, float eps, float* norm_out, cudaStream_t stream) {
    step_inc_kernel<<<1, 1, 0, stream>>>(step_ptr);
    TAG_RETURN_IF_LAUNCH_FAILED();
    long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    clip_adam_kernel<<<(int)blocks, 256, 0, stream>>>(p, g, m, v, n, sumsq, step_ptr, grad_mult, max_norm, lr,
                                                      lr_dev, beta1, beta2, eps, norm_out);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_cast_f32_to_bf16(const float* x, void* y, long n, cudaStream_t stream) {
    long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    cast_bf16_kernel<<<(int)blocks, 256, 0, stream>>>(x, (bf16*)y, n);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_version(void) { return 100; }

// SMs the persistent tensor-core kernels leave free (tc_common.cuh sm_count()): the data-parallel step sets this while
// an all-reduce runs beside the last part of its backward pass, and back to 0 afterwards.  Read at launch time.
int g_tag_sm_reserve = 0;
extern "C" int tag_set_sm_reserve(int sms) {
    if (sms < 0 || sms > 64) return TAG_ERR_BAD_ARG;
    g_tag_sm_reserve = sms;
    return TAG_OK;
}

// ---------------------------------------------------------------------------------------------
// Deterministic split-K: with a workspace set, the tensor-core weight-gradient kernels store every split's partial tile
// into that split's slab (plain stores) and this reduce adds the slabs to dw in split order — bit-identical results from
// run to run, at the price of the slab traffic (default: fp32 atomics straight into dw, order not fixed).
float* g_tag_det_ws = nullptr;
long g_tag_det_ws_bytes = 0;

namespace {
__global__ void splitk_reduce_kernel(const float4* __restrict__ ws, int splits, long n4, float4* __restrict__ dw) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 a = ws[i];
    for (int s = 1; s < splits; ++s) {
        const float4 b = ws[(long)s * n4 + i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    float4 d = dw[i];
    d.x += a.x; d.y += a.y; d.z += a.z; d.w += a.w;
    dw[i] = d;
}
}  // namespace

int tag_splitk_reduce(const float* ws, int splits, long n, float* dw, cudaStream_t stream) {
    if (n % 4 != 0 || (reinterpret_cast<uintptr_t>(dw) & 15) != 0) return TAG_ERR_UNSUPPORTED;
    const long n4 = n / 4;
    splitk_reduce_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(ws), splits, n4,
                                                                          reinterpret_cast<float4*>(dw));
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_set_splitk_workspace(void* ws, long bytes) {
    if ((ws == nullptr) != (bytes == 0) || bytes < 0 || (reinterpret_cast<uintptr_t>(ws) & 15) != 0) return TAG_ERR_BAD_ARG;
    g_tag_det_ws = static_cast<float*>(ws);
    g_tag_det_ws_bytes = bytes;
    return TAG_OK;
}

// ---------------------------------------------------------------------------------------------
// Batched weight preparation: every bf16 GEMM operand of a step (plain casts, tap-major conv
// operands, flipped+transposed dgrad operands) from the fp32 master weights in ONE launch.
// Table entry (5 x int64): src, dst, (rows << 32 | cols), (taps << 32 | mode), first block.
//   mode 0: dst[i] = bf16(src[i])                                   (n = rows * cols * taps)
//   mode 1: tap-major forward   dst[tp][co][ci] = src[co][tap][ci]  (rows = Co, cols = Ci)
//   mode 2: tap-major dgrad     dst[tp][ci][co] = src[co][8-tap][ci]
//   mode 3: transpose (taps=1)  dst[ci][co]     = src[co][ci]
// with tap' = dwi*3 + dhi, tap = dhi*3 + dwi.
namespace {
constexpr int PREP_ELEMS_PER_BLOCK = 2048;

__global__ void __launch_bounds__(256) weight_prep_batch_kernel(const long long* __restrict__ table, int n_entries) {
    __shared__ float tile[64][33];
    int e = 0;
    while (e + 1 < n_entries && (long long)blockIdx.x >= table[(e + 1) * 5 + 4]) ++e;
    const float* src = reinterpret_cast<const float*>(table[e * 5 + 0]);
    bf16* dst = reinterpret_cast<bf16*>(table[e * 5 + 1]);
    const int Co = (int)(table[e * 5 + 2] >> 32), Ci = (int)(table[e * 5 + 2] & 0xFFFFFFFF);
    const int taps = (int)(table[e * 5 + 3] >> 32), mode = (int)(table[e * 5 + 3] & 0xFFFFFFFF);
    const unsigned n = (unsigned)Co * Ci * taps;             // < 2^31 for every operand of this model
    const unsigned blk = (unsigned)((long long)blockIdx.x - table[e * 5 + 4]);
    const int t = threadIdx.x;
    if (mode >= 2 && Ci % 32 == 0 && Co % 64 == 0) {
        // transposing modes through a 64(co) x 32(ci) shared-memory tile: 128-byte runs on both sides
        // (2048 elements per block, like the linear modes, so the host's block count holds)
        const unsigned tiles_ci = Ci / 32, tiles_co = Co / 64;
        const unsigned tco = blk % tiles_co, r1 = blk / tiles_co;
        const unsigned tci = r1 % tiles_ci, tp = r1 / tiles_ci;
        const int stap = (mode == 2) ? 8 - ((tp % 3) * 3 + tp / 3) : 0;
        const long row = (long)taps * Ci;                    // floats between consecutive co in src
        const float* s0 = src + (long)(tco * 64) * row + (long)stap * Ci + tci * 32;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int co = r * 8 + (t >> 5), ci = t & 31;
            tile[co][ci] = s0[(long)co * row + ci];
        }
        __syncthreads();
        bf16* d0 = dst + ((long)tp * Ci + tci * 32) * Co + tco * 64;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int ci = r * 4 + (t >> 6), co = t & 63;
            d0[(long)ci * Co + co] = __float2bfloat16_rn(tile[co][ci]);
        }
        return;
    }
    const unsigned base = blk * PREP_ELEMS_PER_BLOCK;
    for (unsigned i = base + t; i < base + PREP_ELEMS_PER_BLOCK && i < n; i += 256) {
        float v;
        if (mode == 0) {
            v = src[i];
        } else if (mode == 1) {            // i -> [tp][co][ci]
            const unsigned ci = i % Ci, r = i / Ci;
            const unsigned co = r % Co, tp = r / Co;
            v = src[((long)co * 9 + ((tp % 3) * 3 + tp / 3)) * Ci + ci];
        } else if (mode == 2) {            // i -> [tp][ci][co]
            const unsigned co = i % Co, r = i / Co;
            const unsigned ci = r % Ci, tp = r / Ci;
            v = src[((long)co * 9 + (8 - ((tp % 3) * 3 + tp / 3))) * Ci + ci];
        } else {                           // i -> [ci][co]
            const unsigned co = i % Co, ci = i / Co;
            v = src[(long)co * Ci + ci];
        }
        dst[i] = __float2bfloat16_rn(v);
    }
}
}  // namespace

extern "C" int tag_weight_prep_batch(const long long* table, int n_entries, int total_blocks,
                                     cudaStream_t stream) {
    if (n_entries <= 0 || total_blocks <= 0) return TAG_ERR_BAD_ARG;
    weight_prep_batch_kernel<<<total_blocks, 256, 0, stream>>>(table, n_entries);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}
