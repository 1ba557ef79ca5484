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

__global__ void weight_prep_batch_kernel(const long long* __restrict__ table, int n_entries) {
    int e = 0;
    while (e + 1 < n_entries && (long long)blockIdx.x >= table[(e + 1) * 5 + 4]) ++e;
    const float* src = reinterpret_cast<const float*>(table[e * 5 + 0]);
    bf16* dst = reinterpret_cast<bf16*>(table[e * 5 + 1]);
    const int Co = (int)(table[e * 5 + 2] >> 32), Ci = (int)(table[e * 5 + 2] & 0xFFFFFFFF);
    const int taps = (int)(table[e * 5 + 3] >> 32), mode = (int)(table[e * 5 + 3] & 0xFFFFFFFF);
    const long n = (long)Co * Ci * taps;
    const long base = ((long)blockIdx.x - table[e * 5 + 4]) * PREP_ELEMS_PER_BLOCK;
    for (long i = base + threadIdx.x; i < base + PREP_ELEMS_PER_BLOCK && i < n; i += blockDim.x) {
        float v;
        if (mode == 0) {
            v = src[i];
        } else if (mode == 1) {            // i -> [tp][co][ci]
            const int ci = (int)(i % Ci); const long r = i / Ci;
            const int co = (int)(r % Co); const int tp = (int)(r / Co);
            v = src[((long)co * 9 + ((tp % 3) * 3 + tp / 3)) * Ci + ci];
        } else if (mode == 2) {            // i -> [tp][ci][co]
            const int co = (int)(i % Co); const long r = i / Co;
            const int ci = (int)(r % Ci); const int tp = (int)(r / Ci);
            v = src[((long)co * 9 + (8 - ((tp % 3) * 3 + tp / 3))) * Ci + ci];
        } else {                           // i -> [ci][co]
            const int co = (int)(i % Co); const int ci = (int)(i / Co);
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
