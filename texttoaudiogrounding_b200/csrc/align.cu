// Sentence-level audio<->text alignment: all-pairs frame/token similarity fused with its pooling.
//
// Reference: align.DotProduct (models/align.py:7-31) multiplies audio [Ba*T, D] with text [Bt*N, D]^T, applies
// sigmoid + clamp and returns the 4-D matrix sim[i, j, t, n] (Ba*Bt*T*N floats, 8.2 M at B=64); the sim_pooling
// classes (models/sim_pooling.py:6-204) then reduce it over the frames t < audio_len[i] (mean / max / linear /
// exp softmax) and over the tokens n < text_len[j] (mean / sum / max / mean+sum) to sim[i, j].  Here the frame
// pooling happens in the registers of the kernel that forms the dot products, so the 4-D matrix is never written
// (unless the caller asks for it, ``output_matrix`` of AudioTextAlignBy{Word,Phrase}.forward,
// models/audio_text_model.py:886-903,957-976); backward recomputes the probabilities and emits the gradient of the
// logits as one [Ba*T, C] matrix that feeds the two fp32 GEMMs (tag_conv_fwd / tag_conv_wgrad with taps = 1).
#include "common.cuh"

namespace {

constexpr int AL_D = 512;
constexpr int AL_COLS = 64;            // text rows (columns of the score matrix) per CTA
constexpr int AL_SROW = AL_D + 4;      // padded row: float4 reads of 8 consecutive lanes cover all 32 banks
constexpr int AL_FR = 4;               // frames per warp iteration (register tile 4 frames x 2 columns per lane)
constexpr int AL_WARPS = 8;

enum { A_MEAN = 0, A_MAX = 1, A_LINEAR = 2, A_EXP = 3 };
enum { T_MEAN = 0, T_SUM = 1, T_MAX = 2, T_MEANSUM = 3 };

__device__ __forceinline__ float fma4(const float4& a, const float4& b, float acc) {
    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); return fmaf(a.w, b.w, acc);
}

__device__ __forceinline__ float prob_of(float logit) {
    return fminf(fmaxf(1.0f / (1.0f + expf(-logit)), 1e-7f), 1.0f);
}

struct AlignSmem {
    float text[AL_COLS][AL_SROW];
    float a[AL_WARPS][AL_FR][AL_D];
    float red[AL_WARPS][AL_COLS][2];
};

__device__ __forceinline__ void stage_text(AlignSmem& s, const float* __restrict__ text, int c0) {
    for (int i = threadIdx.x; i < AL_COLS * (AL_D / 4); i += AL_WARPS * 32) {
        const int r = i / (AL_D / 4), d4 = i - r * (AL_D / 4);
        *reinterpret_cast<float4*>(&s.text[r][d4 * 4]) =
            *reinterpret_cast<const float4*>(text + (long)(c0 + r) * AL_D + d4 * 4);
    }
}

// probabilities of AL_FR frames x the lane's two columns; frames >= T read as zero rows (masked by the caller)
__device__ __forceinline__ void tile_probs(AlignSmem& s, const float* __restrict__ audio_i, int t0, int T, int warp,
                                           int lane, float scale, float (&p)[AL_FR][2]) {
#pragma unroll
    for (int f = 0; f < AL_FR; ++f) {
        const int t = t0 + f;
#pragma unroll
        for (int k = 0; k < AL_D / 128; ++k) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < T) v = *reinterpret_cast<const float4*>(audio_i + (long)t * AL_D + k * 128 + lane * 4);
            *reinterpret_cast<float4*>(&s.a[warp][f][k * 128 + lane * 4]) = v;
        }
    }
    __syncwarp();
    float acc[AL_FR][2];
#pragma unroll
    for (int f = 0; f < AL_FR; ++f) { acc[f][0] = 0.f; acc[f][1] = 0.f; }
#pragma unroll 2
    for (int d = 0; d < AL_D; d += 4) {
        const float4 x0 = *reinterpret_cast<const float4*>(&s.text[lane][d]);
        const float4 x1 = *reinterpret_cast<const float4*>(&s.text[lane + 32][d]);
#pragma unroll
        for (int f = 0; f < AL_FR; ++f) {
            const float4 av = *reinterpret_cast<const float4*>(&s.a[warp][f][d]);
            acc[f][0] = fma4(av, x0, acc[f][0]);
            acc[f][1] = fma4(av, x1, acc[f][1]);
        }
    }
    __syncwarp();
#pragma unroll
    for (int f = 0; f < AL_FR; ++f) { p[f][0] = prob_of(acc[f][0] * scale); p[f][1] = prob_of(acc[f][1] * scale); }
}

// grid (Cpad / 64, Ba).  colpool[i, c] = frame pooling of sim[i, c, :]; aux[i, c] = what backward needs besides it
// (linear: sum of p; exp: sum of exp(p); max: arg max frame).
template <int MODE>
__global__ void __launch_bounds__(AL_WARPS * 32)
align_fwd_kernel(const float* __restrict__ audio, const float* __restrict__ text,
                 const long long* __restrict__ audio_len, float* __restrict__ sim_matrix,
                 float* __restrict__ colpool, float* __restrict__ aux, int T, int C, int Cpad, int N, int Bt,
                 float scale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AlignSmem& s = *reinterpret_cast<AlignSmem*>(smem_raw);
    const int i = blockIdx.y, c0 = blockIdx.x * AL_COLS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    stage_text(s, text, c0);
    __syncthreads();
    const long long len_raw = audio_len[i];
    const int len = (int)(len_raw < T ? len_raw : T);
    const float* audio_i = audio + (long)i * T * AL_D;
    // the reference's exp softmax shifts by the max over all frames (models/utils.py:79-84); the weights are
    // shift invariant and p <= 1, so no shift is needed here
    float r0[2], r1[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) { r0[q] = MODE == A_MAX ? -INFINITY : 0.f; r1[q] = MODE == A_MAX ? 1e9f : 0.f; }
    const int t_stop = sim_matrix != nullptr ? T : len;
    for (int t0 = warp * AL_FR; t0 < t_stop; t0 += AL_WARPS * AL_FR) {
        float p[AL_FR][2];
        tile_probs(s, audio_i, t0, T, warp, lane, scale, p);
#pragma unroll
        for (int f = 0; f < AL_FR; ++f) {
            const int t = t0 + f;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int c = c0 + lane + 32 * q;
                const float v = p[f][q];
                if (sim_matrix != nullptr && t < T && c < C) {
                    const int j = c / N, n = c - j * N;
                    sim_matrix[(((long)i * Bt + j) * T + t) * N + n] = v;
                }
                if (t < len) {
                    if (MODE == A_MEAN) { r0[q] += v; }
                    else if (MODE == A_LINEAR) { r0[q] += v; r1[q] = fmaf(v, v, r1[q]); }
                    else if (MODE == A_MAX) { if (v > r0[q]) { r0[q] = v; r1[q] = (float)t; } }
                    else { const float e = expf(v); r0[q] += e; r1[q] = fmaf(e, v, r1[q]); }
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) { s.red[warp][lane + 32 * q][0] = r0[q]; s.red[warp][lane + 32 * q][1] = r1[q]; }
    __syncthreads();
    if (threadIdx.x < AL_COLS) {
        const int cl = threadIdx.x;
        float a0 = s.red[0][cl][0], a1 = s.red[0][cl][1];
        for (int w = 1; w < AL_WARPS; ++w) {
            const float b0 = s.red[w][cl][0], b1 = s.red[w][cl][1];
            if (MODE == A_MAX) { if (b0 > a0 || (b0 == a0 && b1 < a1)) { a0 = b0; a1 = b1; } }
            else { a0 += b0; a1 += b1; }
        }
        float out, ax;
        if (MODE == A_MEAN) { out = a0 / (float)len_raw; ax = 0.f; }
        else if (MODE == A_LINEAR) { out = a1 / a0; ax = a0; }
        else if (MODE == A_MAX) { out = a0; ax = a1; }
        else { out = a1 / a0; ax = a0; }
        colpool[(long)i * Cpad + c0 + cl] = out;
        aux[(long)i * Cpad + c0 + cl] = ax;
    }
}

// G[i*T + t, c] = scale * d(loss)/d(logit[i, t, c]); zero for t >= audio_len[i].
template <int MODE>
__global__ void __launch_bounds__(AL_WARPS * 32)
align_bwd_kernel(const float* __restrict__ audio, const float* __restrict__ text,
                 const long long* __restrict__ audio_len, const float* __restrict__ d_colpool,
                 const float* __restrict__ colpool, const float* __restrict__ aux, float* __restrict__ G, int T,
                 int Cpad, float scale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AlignSmem& s = *reinterpret_cast<AlignSmem*>(smem_raw);
    const int i = blockIdx.y, c0 = blockIdx.x * AL_COLS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    stage_text(s, text, c0);
    __syncthreads();
    const long long len_raw = audio_len[i];
    const int len = (int)(len_raw < T ? len_raw : T);
    const float* audio_i = audio + (long)i * T * AL_D;
    float g[2], cp[2], ax[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const long o = (long)i * Cpad + c0 + lane + 32 * q;
        g[q] = d_colpool[o]; cp[q] = colpool[o]; ax[q] = aux[o];
    }
    for (int t0 = warp * AL_FR; t0 < T; t0 += AL_WARPS * AL_FR) {
        float p[AL_FR][2];
        if (t0 < len) tile_probs(s, audio_i, t0, T, warp, lane, scale, p);
#pragma unroll
        for (int f = 0; f < AL_FR; ++f) {
            const int t = t0 + f;
            if (t >= T) break;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                float dl = 0.f;
                if (t < len) {
                    const float v = p[f][q];
                    float dp;
                    if (MODE == A_MEAN) dp = g[q] / (float)len_raw;
                    else if (MODE == A_LINEAR) dp = g[q] * (2.0f * v - cp[q]) / ax[q];
                    else if (MODE == A_MAX) dp = (float)t == ax[q] ? g[q] : 0.f;
                    else dp = g[q] * expf(v) * (1.0f + v - cp[q]) / ax[q];
                    dl = v > 1e-7f ? dp * v * (1.0f - v) * scale : 0.f;     // the clamp has zero gradient
                }
                G[((long)i * T + t) * Cpad + c0 + lane + 32 * q] = dl;
            }
        }
    }
}


// ---- tensor-core route (large batches): the logits [Ba*T, Cpad] come from ONE split-bf16 tcgen05 GEMM (split.cu);
// these kernels do the sigmoid / clamp / frame pooling (fwd) and the logit gradient (bwd) per (clip, text row)
// column, coalesced across the columns.  grid (Cpad / 128, Ba), 128 threads.
template <int MODE>
__global__ void __launch_bounds__(128)
align_logits_fwd_kernel(const float* __restrict__ logits, const long long* __restrict__ audio_len,
                        float* __restrict__ sim_matrix, float* __restrict__ colpool, float* __restrict__ aux, int T,
                        int C, int Cpad, int N, int Bt, float scale) {
    const int i = blockIdx.y, c = blockIdx.x * 128 + threadIdx.x;
    const long long len_raw = audio_len[i];
    const int len = (int)(len_raw < T ? len_raw : T);
    const float* l = logits + (long)i * T * Cpad + c;
    float r0 = MODE == A_MAX ? -INFINITY : 0.f, r1 = 0.f;
    const bool write = sim_matrix != nullptr && c < C;
    const int j = c / N, n = c - j * N;
    const int t_stop = write ? T : len;
    for (int t = 0; t < t_stop; ++t) {
        const float v = prob_of(l[(long)t * Cpad] * scale);
        if (write) sim_matrix[(((long)i * Bt + j) * T + t) * N + n] = v;
        if (t < len) {
            if (MODE == A_MEAN) r0 += v;
            else if (MODE == A_LINEAR) { r0 += v; r1 = fmaf(v, v, r1); }
            else if (MODE == A_MAX) { if (v > r0) { r0 = v; r1 = (float)t; } }
            else { const float e = expf(v); r0 += e; r1 = fmaf(e, v, r1); }
        }
    }
    float out, ax;
    if (MODE == A_MEAN) { out = r0 / (float)len_raw; ax = 0.f; }
    else if (MODE == A_LINEAR) { out = r1 / r0; ax = r0; }
    else if (MODE == A_MAX) { out = r0; ax = r1; }
    else { out = r1 / r0; ax = r0; }
    colpool[(long)i * Cpad + c] = out;
    aux[(long)i * Cpad + c] = ax;
}

// G = scale * d(loss)/d(logit) written as the split-bf16 operands of the two gradient GEMMs:
// g_kcat [Ba*T][3*Cpad] = [hi | hi | lo] (d_audio = G x text) and g_planes [2][Ba*T][Cpad] = hi, lo (d_text = G^T x audio)
template <int MODE>
__global__ void __launch_bounds__(128)
align_logits_bwd_kernel(const float* __restrict__ logits, const long long* __restrict__ audio_len,
                        const float* __restrict__ d_colpool, const float* __restrict__ colpool,
                        const float* __restrict__ aux, bf16* __restrict__ g_kcat, bf16* __restrict__ g_planes, int Ba,
                        int T, int Cpad, float scale) {
    const int i = blockIdx.y, c = blockIdx.x * 128 + threadIdx.x;
    const long long len_raw = audio_len[i];
    const int len = (int)(len_raw < T ? len_raw : T);
    const float* l = logits + (long)i * T * Cpad + c;
    const long o = (long)i * Cpad + c;
    const float g = d_colpool[o], cp = colpool[o], ax = aux[o];
    const long plane = (long)Ba * T * Cpad;
    for (int t = 0; t < T; ++t) {
        float dl = 0.f;
        if (t < len && g != 0.f) {
            const float v = prob_of(l[(long)t * Cpad] * scale);
            float dp;
            if (MODE == A_MEAN) dp = g / (float)len_raw;
            else if (MODE == A_LINEAR) dp = g * (2.0f * v - cp) / ax;
            else if (MODE == A_MAX) dp = (float)t == ax ? g : 0.f;
            else dp = g * expf(v) * (1.0f + v - cp) / ax;
            dl = v > 1e-7f ? dp * v * (1.0f - v) * scale : 0.f;
        }
        const bf16 hi = __float2bfloat16_rn(dl);
        const bf16 lo = __float2bfloat16_rn(dl - __bfloat162float(hi));
        const long row = (long)i * T + t;
        bf16* k = g_kcat + row * 3 * Cpad + c;
        k[0] = hi; k[Cpad] = hi; k[2 * Cpad] = lo;
        g_planes[row * Cpad + c] = hi;
        g_planes[plane + row * Cpad + c] = lo;
    }
}

// ---- token pooling: out[i, j] from colpool[i, j*N + n], n < text_len[j] (models/sim_pooling.py); one thread per pair
__global__ void align_text_pool_fwd_kernel(const float* __restrict__ colpool, const long long* __restrict__ text_len,
                                           int mode, float* __restrict__ out, int Ba, int Bt, int N, int Cpad) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Ba * Bt) return;
    const int i = idx / Bt, j = idx - i * Bt;
    const long long lr = text_len[j];
    const int len = (int)(lr < N ? lr : N);
    const float* f = colpool + (long)i * Cpad + (long)j * N;
    float sum = 0.f, mx = -INFINITY;
    for (int n = 0; n < len; ++n) { sum += f[n]; mx = fmaxf(mx, f[n]); }
    float o;
    if (mode == T_MEAN) o = sum / (float)lr;
    else if (mode == T_SUM) o = sum;
    else if (mode == T_MAX) o = mx;
    else o = sum + sum / (float)lr;
    out[idx] = o;
}

__global__ void align_text_pool_bwd_kernel(const float* __restrict__ d_out, const float* __restrict__ colpool,
                                           const long long* __restrict__ text_len, int mode,
                                           float* __restrict__ d_colpool, int Ba, int Bt, int N, int Cpad) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Ba * Bt) return;
    const int i = idx / Bt, j = idx - i * Bt;
    const long long lr = text_len[j];
    const int len = (int)(lr < N ? lr : N);
    const float* f = colpool + (long)i * Cpad + (long)j * N;
    float* df = d_colpool + (long)i * Cpad + (long)j * N;
    const float g = d_out[idx];
    int arg = -1;
    if (mode == T_MAX) {
        float mx = -INFINITY;
        for (int n = 0; n < len; ++n) if (f[n] > mx) { mx = f[n]; arg = n; }
    }
    for (int n = 0; n < N; ++n) {
        float v = 0.f;
        if (n < len) {
            if (mode == T_MEAN) v = g / (float)lr;
            else if (mode == T_SUM) v = g;
            else if (mode == T_MAX) v = n == arg ? g : 0.f;
            else v = g * (1.0f + 1.0f / (float)lr);
        }
        df[n] = v;
    }
}

// ---- MaxMarginRankingLoss (losses.py:226-264) on sim [n, n]: loss and d(loss)/d(sim), one CTA
__global__ void __launch_bounds__(256)
max_margin_rank_kernel(const float* __restrict__ x, int n, float margin, float lamda1, int fix_norm,
                       float* __restrict__ loss, float* __restrict__ dx) {
    __shared__ float s_part[8];
    const float inv = 1.0f / (fix_norm ? 2.0f * n * (n - 1) : 2.0f * n * n);
    for (int k = threadIdx.x; k < n * n; k += 256) dx[k] = 0.f;
    __syncthreads();
    float acc = 0.f;
    for (int k = threadIdx.x; k < n * n; k += 256) {
        const int i = k / n, j = k - i * n;
        if (fix_norm && i == j) continue;
        const float d = x[i * n + i];
        // first half: margin - (x[i,i] - x[i,j]);  second half: margin - (x[i,i] - lamda1 * x[j,i])
        const float m1 = margin - (d - x[i * n + j]);
        const float m2 = margin - (d - lamda1 * x[j * n + i]);
        if (m1 > 0.f) { acc += m1; atomicAdd(dx + i * n + i, -inv); atomicAdd(dx + i * n + j, inv); }
        if (m2 > 0.f) { acc += m2; atomicAdd(dx + i * n + i, -inv); atomicAdd(dx + j * n + i, lamda1 * inv); }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += s_part[w];
        *loss = t * inv;
    }
}

template <int MODE>
int launch_fwd(const float* audio, const float* text, const long long* audio_len, float* sim_matrix, float* colpool,
               float* aux, int Ba, int T, int C, int Cpad, int N, int Bt, float scale, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(align_fwd_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(AlignSmem));
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    align_fwd_kernel<MODE><<<dim3(Cpad / AL_COLS, Ba), AL_WARPS * 32, sizeof(AlignSmem), stream>>>(
        audio, text, audio_len, sim_matrix, colpool, aux, T, C, Cpad, N, Bt, scale);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

template <int MODE>
int launch_bwd(const float* audio, const float* text, const long long* audio_len, const float* d_colpool,
               const float* colpool, const float* aux, float* G, int Ba, int T, int Cpad, float scale,
               cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(align_bwd_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(AlignSmem));
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    align_bwd_kernel<MODE><<<dim3(Cpad / AL_COLS, Ba), AL_WARPS * 32, sizeof(AlignSmem), stream>>>(
        audio, text, audio_len, d_colpool, colpool, aux, G, T, Cpad, scale);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

}  // namespace

extern "C" int tag_align_pool_fwd(const float* audio, const float* text, const long long* audio_len, int a_mode,
                                  float* sim_matrix, float* colpool, float* aux, int Ba, int T, int Bt, int N,
                                  int Cpad, int D, float scale, cudaStream_t stream) {
    if (Ba <= 0 || T <= 0 || Bt <= 0 || N <= 0 || a_mode < 0 || a_mode > 3) return TAG_ERR_BAD_ARG;
    if (D != AL_D || Cpad % AL_COLS != 0 || Cpad < Bt * N) return TAG_ERR_UNSUPPORTED;
    const int C = Bt * N;
    switch (a_mode) {
        case A_MEAN: return launch_fwd<A_MEAN>(audio, text, audio_len, sim_matrix, colpool, aux, Ba, T, C, Cpad, N, Bt, scale, stream);
        case A_MAX: return launch_fwd<A_MAX>(audio, text, audio_len, sim_matrix, colpool, aux, Ba, T, C, Cpad, N, Bt, scale, stream);
        case A_LINEAR: return launch_fwd<A_LINEAR>(audio, text, audio_len, sim_matrix, colpool, aux, Ba, T, C, Cpad, N, Bt, scale, stream);
        default: return launch_fwd<A_EXP>(audio, text, audio_len, sim_matrix, colpool, aux, Ba, T, C, Cpad, N, Bt, scale, stream);
    }
}

extern "C" int tag_align_pool_bwd(const float* audio, const float* text, const long long* audio_len, int a_mode,
                                  const float* d_colpool, const float* colpool, const float* aux, float* G, int Ba,
                                  int T, int Cpad, int D, float scale, cudaStream_t stream) {
    if (Ba <= 0 || T <= 0 || a_mode < 0 || a_mode > 3) return TAG_ERR_BAD_ARG;
    if (D != AL_D || Cpad % AL_COLS != 0) return TAG_ERR_UNSUPPORTED;
    switch (a_mode) {
        case A_MEAN: return launch_bwd<A_MEAN>(audio, text, audio_len, d_colpool, colpool, aux, G, Ba, T, Cpad, scale, stream);
        case A_MAX: return launch_bwd<A_MAX>(audio, text, audio_len, d_colpool, colpool, aux, G, Ba, T, Cpad, scale, stream);
        case A_LINEAR: return launch_bwd<A_LINEAR>(audio, text, audio_len, d_colpool, colpool, aux, G, Ba, T, Cpad, scale, stream);
        default: return launch_bwd<A_EXP>(audio, text, audio_len, d_colpool, colpool, aux, G, Ba, T, Cpad, scale, stream);
    }
}

extern "C" int tag_align_text_pool_fwd(const float* colpool, const long long* text_len, int t_mode, float* out, int Ba,
                                       int Bt, int N, int Cpad, cudaStream_t stream) {
    if (Ba <= 0 || Bt <= 0 || N <= 0 || t_mode < 0 || t_mode > 3) return TAG_ERR_BAD_ARG;
    align_text_pool_fwd_kernel<<<(Ba * Bt + 127) / 128, 128, 0, stream>>>(colpool, text_len, t_mode, out, Ba, Bt, N, Cpad);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_align_text_pool_bwd(const float* d_out, const float* colpool, const long long* text_len, int t_mode,
                                       float* d_colpool, int Ba, int Bt, int N, int Cpad, cudaStream_t stream) {
    if (Ba <= 0 || Bt <= 0 || N <= 0 || t_mode < 0 || t_mode > 3) return TAG_ERR_BAD_ARG;
    align_text_pool_bwd_kernel<<<(Ba * Bt + 127) / 128, 128, 0, stream>>>(d_out, colpool, text_len, t_mode, d_colpool,
                                                                        Ba, Bt, N, Cpad);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_max_margin_rank(const float* sim, int n, float margin, float lamda1, int fix_norm, float* loss,
                                   float* d_sim, cudaStream_t stream) {
    if (n <= 0 || (fix_norm && n < 2)) return TAG_ERR_BAD_ARG;
    max_margin_rank_kernel<<<1, 256, 0, stream>>>(sim, n, margin, lamda1, fix_norm, loss, d_sim);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_align_logits_fwd(const float* logits, const long long* audio_len, int a_mode, float* sim_matrix,
                                    float* colpool, float* aux, int Ba, int T, int Bt, int N, int Cpad, float scale,
                                    cudaStream_t stream) {
    if (Ba <= 0 || T <= 0 || Bt <= 0 || N <= 0 || a_mode < 0 || a_mode > 3) return TAG_ERR_BAD_ARG;
    if (Cpad % 128 != 0 || Cpad < Bt * N) return TAG_ERR_UNSUPPORTED;
    const int C = Bt * N;
    const dim3 grid(Cpad / 128, Ba);
#define TAG_ALF(M_) align_logits_fwd_kernel<M_><<<grid, 128, 0, stream>>>(logits, audio_len, sim_matrix, colpool, aux, T, C, Cpad, N, Bt, scale)
    if (a_mode == A_MEAN) TAG_ALF(A_MEAN); else if (a_mode == A_MAX) TAG_ALF(A_MAX);
    else if (a_mode == A_LINEAR) TAG_ALF(A_LINEAR); else TAG_ALF(A_EXP);
#undef TAG_ALF
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_align_logits_bwd(const float* logits, const long long* audio_len, int a_mode,
                                    const float* d_colpool, const float* colpool, const float* aux, void* g_kcat,
                                    void* g_planes, int Ba, int T, int Cpad, float scale, cudaStream_t stream) {
    if (Ba <= 0 || T <= 0 || a_mode < 0 || a_mode > 3) return TAG_ERR_BAD_ARG;
    if (Cpad % 128 != 0) return TAG_ERR_UNSUPPORTED;
    const dim3 grid(Cpad / 128, Ba);
#define TAG_ALB(M_) align_logits_bwd_kernel<M_><<<grid, 128, 0, stream>>>(logits, audio_len, d_colpool, colpool, aux, (bf16*)g_kcat, (bf16*)g_planes, Ba, T, Cpad, scale)
    if (a_mode == A_MEAN) TAG_ALB(A_MEAN); else if (a_mode == A_MAX) TAG_ALB(A_MAX);
    else if (a_mode == A_LINEAR) TAG_ALB(A_LINEAR); else TAG_ALB(A_EXP);
#undef TAG_ALB
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}
