// Multi-phrase (weakly supervised) head: every clip is matched against n phrases.
//
// Reference: MultiTextBiEncoder.forward (models/audio_text_model.py:147-229) expands the audio embedding to
// [B*n, T, D] (a 32x copy) and calls DotProduct (models/match.py:43-60) on the pairs, then pools the frame
// probabilities over time with one of the *_with_lens functions (models/utils.py:33-95).  Here the audio embedding
// is read in place: sim[b,t,j] = clamp(sigmoid(scale * <audio[b,t,:], seq[b,j,:]>), 1e-7, 1) for all j of a clip
// from one staging of the clip's phrase embeddings in shared memory; the pooling is a per-(b, j) scan over time.
#include "common.cuh"

namespace {

constexpr int MT_MAX_N = 64;          // phrases per clip handled by one launch
constexpr int MT_D = 512;
constexpr int MT_SROW = MT_D + 4;     // padded phrase row: float4 reads of 8 consecutive lanes cover all 32 banks
constexpr int MT_TT = 32;             // frames per CTA

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); return fmaf(a.w, b.w, acc);
}

// grid (ceil(T / MT_TT), B), 256 threads; lane j owns phrases j and j + 32
__global__ void __launch_bounds__(256)
multi_dot_fwd_kernel(const float* __restrict__ audio, const float* __restrict__ seq, float* __restrict__ sim,
                     int T, int n, float scale) {
    extern __shared__ __align__(16) float sm[];
    float* s_seq = sm;                                  // [n][MT_SROW]
    float* s_a = sm + (size_t)MT_MAX_N * MT_SROW;       // [8 warps][MT_D]
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < n * (MT_D / 4); i += 256) {
        const int j = i / (MT_D / 4), d4 = i - j * (MT_D / 4);
        *reinterpret_cast<float4*>(s_seq + j * MT_SROW + d4 * 4) =
            *reinterpret_cast<const float4*>(seq + ((long)b * n + j) * MT_D + d4 * 4);
    }
    __syncthreads();
    float* a_row = s_a + warp * MT_D;
    const int t_end = min(T, (int)(blockIdx.x + 1) * MT_TT);
    for (int t = blockIdx.x * MT_TT + warp; t < t_end; t += 8) {
        const float* a = audio + ((long)b * T + t) * MT_D;
#pragma unroll
        for (int k = 0; k < MT_D / 128; ++k)
            *reinterpret_cast<float4*>(a_row + k * 128 + lane * 4) = *reinterpret_cast<const float4*>(a + k * 128 + lane * 4);
        __syncwarp();
        const int j0 = lane < n ? lane : 0, j1 = lane + 32 < n ? lane + 32 : 0;
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 4
        for (int d = 0; d < MT_D; d += 4) {
            const float4 av = *reinterpret_cast<const float4*>(a_row + d);
            acc0 = dot4(av, *reinterpret_cast<const float4*>(s_seq + j0 * MT_SROW + d), acc0);
            if (n > 32) acc1 = dot4(av, *reinterpret_cast<const float4*>(s_seq + j1 * MT_SROW + d), acc1);
        }
        float* o = sim + ((long)b * T + t) * n;
        if (lane < n) o[lane] = fminf(fmaxf(1.0f / (1.0f + expf(-acc0 * scale)), 1e-7f), 1.0f);
        if (lane + 32 < n) o[lane + 32] = fminf(fmaxf(1.0f / (1.0f + expf(-acc1 * scale)), 1e-7f), 1.0f);
        __syncwarp();
    }
}

// d_logit[b,t,j] = d_sim * p (1 - p) (zero where the clamp is active);  d_audio[b,t,:] = scale * sum_j d_logit * seq[b,j,:]
__global__ void __launch_bounds__(256)
multi_dot_bwd_audio_kernel(const float* __restrict__ d_sim, const float* __restrict__ sim,
                           const float* __restrict__ seq, float* __restrict__ d_audio, float* __restrict__ d_logit,
                           int T, int n, float scale) {
    extern __shared__ __align__(16) float sm[];
    float* s_seq = sm;
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < n * (MT_D / 4); i += 256) {
        const int j = i / (MT_D / 4), d4 = i - j * (MT_D / 4);
        *reinterpret_cast<float4*>(s_seq + j * MT_SROW + d4 * 4) =
            *reinterpret_cast<const float4*>(seq + ((long)b * n + j) * MT_D + d4 * 4);
    }
    __syncthreads();
    const int t_end = min(T, (int)(blockIdx.x + 1) * MT_TT);
    for (int t = blockIdx.x * MT_TT + warp; t < t_end; t += 8) {
        const long row = ((long)b * T + t) * n;
        float g0 = 0.f, g1 = 0.f;
        if (lane < n) { const float p = sim[row + lane]; g0 = p > 1e-7f ? d_sim[row + lane] * p * (1.0f - p) : 0.f; d_logit[row + lane] = g0; }
        if (lane + 32 < n) { const float p = sim[row + lane + 32]; g1 = p > 1e-7f ? d_sim[row + lane + 32] * p * (1.0f - p) : 0.f; d_logit[row + lane + 32] = g1; }
        float4 acc[MT_D / 128];
#pragma unroll
        for (int k = 0; k < MT_D / 128; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < n; ++j) {
            const float g = __shfl_sync(0xffffffffu, j < 32 ? g0 : g1, j & 31);
#pragma unroll
            for (int k = 0; k < MT_D / 128; ++k) {
                const float4 sv = *reinterpret_cast<const float4*>(s_seq + j * MT_SROW + k * 128 + lane * 4);
                acc[k].x = fmaf(g, sv.x, acc[k].x); acc[k].y = fmaf(g, sv.y, acc[k].y);
                acc[k].z = fmaf(g, sv.z, acc[k].z); acc[k].w = fmaf(g, sv.w, acc[k].w);
            }
        }
        float* o = d_audio + ((long)b * T + t) * MT_D;
#pragma unroll
        for (int k = 0; k < MT_D / 128; ++k)
            *reinterpret_cast<float4*>(o + k * 128 + lane * 4) =
                make_float4(acc[k].x * scale, acc[k].y * scale, acc[k].z * scale, acc[k].w * scale);
    }
}

// d_seq[b,j,:] = scale * sum_t d_logit[b,t,j] * audio[b,t,:]; grid (MT_D / 128, B), warp w owns phrases w, w+8, ...
__global__ void __launch_bounds__(256)
multi_dot_bwd_seq_kernel(const float* __restrict__ d_logit, const float* __restrict__ audio,
                         float* __restrict__ d_seq, int T, int n, float scale) {
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d0 = blockIdx.x * 128 + lane * 4;
    float4 acc[MT_MAX_N / 8];
#pragma unroll
    for (int i = 0; i < MT_MAX_N / 8; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < T; ++t) {
        const float4 av = *reinterpret_cast<const float4*>(audio + ((long)b * T + t) * MT_D + d0);
        const float* gl = d_logit + ((long)b * T + t) * n;
#pragma unroll
        for (int i = 0; i < MT_MAX_N / 8; ++i) {
            const int j = warp + 8 * i;
            if (j < n) {
                const float g = __ldg(gl + j);
                acc[i].x = fmaf(g, av.x, acc[i].x); acc[i].y = fmaf(g, av.y, acc[i].y);
                acc[i].z = fmaf(g, av.z, acc[i].z); acc[i].w = fmaf(g, av.w, acc[i].w);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < MT_MAX_N / 8; ++i) {
        const int j = warp + 8 * i;
        if (j < n)
            *reinterpret_cast<float4*>(d_seq + ((long)b * n + j) * MT_D + d0) =
                make_float4(acc[i].x * scale, acc[i].y * scale, acc[i].z * scale, acc[i].w * scale);
    }
}

// ---- pooling over time with lengths (models/utils.py:33-95); one thread per (b, j)
enum { POOL_LINEAR_SOFTMAX = 0, POOL_MAX = 1, POOL_MEAN = 2, POOL_EXP_SOFTMAX = 3 };

__global__ void pool_with_lens_fwd_kernel(const float* __restrict__ sim, const long long* __restrict__ length,
                                          int mode, float* __restrict__ clip, int B, int T, int n) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * n) return;
    const int b = idx / n, j = idx - b * n;
    long long len = length[b];
    if (len > T) len = T;
    const float* f = sim + (long)b * T * n + j;
    float out;
    if (mode == POOL_LINEAR_SOFTMAX) {
        float s1 = 0.f, s2 = 0.f;
        for (int t = 0; t < len; ++t) { const float v = f[(long)t * n]; s1 += v; s2 = fmaf(v, v, s2); }
        out = s2 / s1;
    } else if (mode == POOL_MAX) {
        float m = -INFINITY;
        for (int t = 0; t < len; ++t) m = fmaxf(m, f[(long)t * n]);
        out = m;
    } else if (mode == POOL_MEAN) {
        float s1 = 0.f;
        for (int t = 0; t < len; ++t) s1 += f[(long)t * n];
        out = s1 / (float)length[b];
    } else {
        float m = -INFINITY;                       // the reference shifts by the max over ALL frames (utils.py:83)
        for (int t = 0; t < T; ++t) m = fmaxf(m, f[(long)t * n]);
        float se = 0.f, sf = 0.f;
        for (int t = 0; t < len; ++t) { const float v = f[(long)t * n]; const float e = expf(v - m); se += e; sf = fmaf(e, v, sf); }
        out = sf / se;
    }
    clip[idx] = out;
}

__global__ void pool_with_lens_bwd_kernel(const float* __restrict__ d_clip, const float* __restrict__ sim,
                                          const float* __restrict__ clip, const long long* __restrict__ length,
                                          int mode, float* __restrict__ d_sim, int B, int T, int n) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * n) return;
    const int b = idx / n, j = idx - b * n;
    long long len = length[b];
    if (len > T) len = T;
    const float* f = sim + (long)b * T * n + j;
    float* df = d_sim + (long)b * T * n + j;
    const float g = d_clip[idx], c = clip[idx];
    if (mode == POOL_LINEAR_SOFTMAX) {
        float s1 = 0.f;
        for (int t = 0; t < len; ++t) s1 += f[(long)t * n];
        const float inv = g / s1;                  // d(S2/S1)/df_t = (2 f_t - S2/S1) / S1
        for (int t = 0; t < T; ++t) df[(long)t * n] = t < len ? (2.0f * f[(long)t * n] - c) * inv : 0.f;
    } else if (mode == POOL_MAX) {
        int arg = -1;
        for (int t = 0; t < len; ++t) if (arg < 0 && f[(long)t * n] == c) arg = t;
        for (int t = 0; t < T; ++t) df[(long)t * n] = t == arg ? g : 0.f;
    } else if (mode == POOL_MEAN) {
        const float v = g / (float)length[b];
        for (int t = 0; t < T; ++t) df[(long)t * n] = t < len ? v : 0.f;
    } else {
        float m = -INFINITY;
        for (int t = 0; t < T; ++t) m = fmaxf(m, f[(long)t * n]);
        float se = 0.f;
        for (int t = 0; t < len; ++t) se += expf(f[(long)t * n] - m);
        const float inv = g / se;                  // d clip / d f_t = w_t (1 + f_t - clip)
        for (int t = 0; t < T; ++t) {
            const float v = f[(long)t * n];
            df[(long)t * n] = t < len ? expf(v - m) * (1.0f + v - c) * inv : 0.f;
        }
    }
}

constexpr size_t MT_SMEM_FWD = ((size_t)MT_MAX_N * MT_SROW + 8 * MT_D) * sizeof(float);
constexpr size_t MT_SMEM_BWD = (size_t)MT_MAX_N * MT_SROW * sizeof(float);

}  // namespace

extern "C" int tag_multi_dot_sigmoid_fwd(const float* audio, const float* seq, float* sim, int B, int T, int n,
                                         int D, float scale, cudaStream_t stream) {
    if (B <= 0 || T <= 0 || n <= 0) return TAG_ERR_BAD_ARG;
    if (D != MT_D || n > MT_MAX_N) return TAG_ERR_UNSUPPORTED;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(multi_dot_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MT_SMEM_FWD);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    dim3 grid((T + MT_TT - 1) / MT_TT, B);
    multi_dot_fwd_kernel<<<grid, 256, MT_SMEM_FWD, stream>>>(audio, seq, sim, T, n, scale);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_multi_dot_sigmoid_bwd(const float* d_sim, const float* sim, const float* audio, const float* seq,
                                         float* d_audio, float* d_seq, float* d_logit_ws, int B, int T, int n, int D,
                                         float scale, cudaStream_t stream) {
    if (B <= 0 || T <= 0 || n <= 0) return TAG_ERR_BAD_ARG;
    if (D != MT_D || n > MT_MAX_N) return TAG_ERR_UNSUPPORTED;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(multi_dot_bwd_audio_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MT_SMEM_BWD);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    dim3 grid((T + MT_TT - 1) / MT_TT, B);
    multi_dot_bwd_audio_kernel<<<grid, 256, MT_SMEM_BWD, stream>>>(d_sim, sim, seq, d_audio, d_logit_ws, T, n, scale);
    TAG_RETURN_IF_LAUNCH_FAILED();
    dim3 grid2(MT_D / 128, B);
    multi_dot_bwd_seq_kernel<<<grid2, 256, 0, stream>>>(d_logit_ws, audio, d_seq, T, n, scale);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_pool_with_lens_fwd(const float* sim, const long long* length, int mode, float* clip, int B, int T,
                                      int n, cudaStream_t stream) {
    if (B <= 0 || T <= 0 || n <= 0 || mode < 0 || mode > 3) return TAG_ERR_BAD_ARG;
    pool_with_lens_fwd_kernel<<<(B * n + 127) / 128, 128, 0, stream>>>(sim, length, mode, clip, B, T, n);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_pool_with_lens_bwd(const float* d_clip, const float* sim, const float* clip, const long long* length,
                                      int mode, float* d_sim, int B, int T, int n, cudaStream_t stream) {
    if (B <= 0 || T <= 0 || n <= 0 || mode < 0 || mode > 3) return TAG_ERR_BAD_ARG;
    pool_with_lens_bwd_kernel<<<(B * n + 127) / 128, 128, 0, stream>>>(d_clip, sim, clip, length, mode, d_sim, B, T, n);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}
