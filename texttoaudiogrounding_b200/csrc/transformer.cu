// Row-wise pieces of the CLAP text tower (RoBERTa-base layout; transformers ClapTextModel as wired by the reference's
// LaionClapEncoder, models/text_encoder.py:311-327 == models/hf_modeling_grounding.py:183-199; BASELINE.json
// configs[4], inference): embedding sum + LayerNorm, residual + LayerNorm, GELU / tanh, L2 normalisation.  The dense
// layers are bf16 tcgen05 GEMMs (tag_conv_tc_fwd, taps = 1) and the attention core is tag_mha_core_fwd.
// One warp per row, fp32 math, E = 32 * EL with EL in {16, 24} (512 / 768).
#include "common.cuh"

namespace {

template <int EL>
__device__ __forceinline__ void layernorm_row(float (&x)[EL], const float* __restrict__ gamma,
                                              const float* __restrict__ beta, float eps, int lane,
                                              float* __restrict__ out) {
    constexpr int E = EL * 32;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < EL; ++i) s += x[i];
    const float mean = warp_sum(s) * (1.0f / E);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < EL; ++i) { const float d = x[i] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / E) + eps);
#pragma unroll
    for (int i = 0; i < EL; ++i) {
        const int e = lane + 32 * i;
        out[e] = fmaf((x[i] - mean) * rstd, gamma[e], beta[e]);
    }
}

// ClapTextEmbeddings.forward: word + token_type(0) + position, position id = (number of non-pad ids up to and
// including l) + pad_idx for non-pad tokens, pad_idx for pads; then LayerNorm.
template <int EL>
__global__ void __launch_bounds__(256)
roberta_embed_ln_kernel(const long long* __restrict__ ids, const float* __restrict__ word,
                        const float* __restrict__ pos, const float* __restrict__ type0,
                        const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ out,
                        long rows, int L, int vocab, int max_pos, int pad_idx, float eps) {
    constexpr int E = EL * 32;
    const long r = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const long b = r / L;
    const int l = (int)(r - b * L);
    int cnt = 0;
    for (int l0 = 0; l0 <= l; l0 += 32) {
        const int li = l0 + lane;
        const bool nz = li <= l && ids[b * L + li] != pad_idx;
        cnt += __popc(__ballot_sync(0xffffffffu, nz));
    }
    long long id = ids[r];
    const bool is_pad = id == pad_idx;
    if (id < 0) id = 0;
    if (id >= vocab) id = vocab - 1;
    int p = is_pad ? pad_idx : cnt + pad_idx;
    if (p >= max_pos) p = max_pos - 1;
    float x[EL];
#pragma unroll
    for (int i = 0; i < EL; ++i) {
        const int e = lane + 32 * i;
        x[i] = word[id * E + e] + type0[e] + pos[(long)p * E + e];
    }
    layernorm_row<EL>(x, gamma, beta, eps, lane, out + r * E);
}

template <int EL>
__global__ void __launch_bounds__(256)
add_layernorm_kernel(const float* __restrict__ a, const float* __restrict__ res, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float* __restrict__ out, long rows, float eps) {
    constexpr int E = EL * 32;
    const long r = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    float x[EL];
#pragma unroll
    for (int i = 0; i < EL; ++i) {
        const long o = r * E + lane + 32 * i;
        x[i] = a[o] + (res != nullptr ? res[o] : 0.f);
    }
    layernorm_row<EL>(x, gamma, beta, eps, lane, out + r * E);
}

__global__ void unary_kernel(const float* __restrict__ in, float* __restrict__ out, long n, int op) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float v = in[i];
        out[i] = op == 0 ? 0.5f * v * (1.0f + erff(v * 0.70710678118654752f)) : tanhf(v);
    }
}

// F.normalize(x, dim=-1): x / max(||x||_2, eps)
__global__ void l2_normalize_kernel(const float* __restrict__ in, float* __restrict__ out, long rows, int E,
                                    float eps) {
    const long r = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    float q = 0.f;
    for (int e = lane; e < E; e += 32) { const float v = in[r * E + e]; q = fmaf(v, v, q); }
    const float inv = 1.0f / fmaxf(sqrtf(warp_sum(q)), eps);
    for (int e = lane; e < E; e += 32) out[r * E + e] = in[r * E + e] * inv;
}

}  // namespace

extern "C" int tag_roberta_embed_ln(const long long* ids, const float* word, const float* pos, const float* type0,
                                    const float* gamma, const float* beta, float* out, int B, int L, int E, int vocab,
                                    int max_pos, int pad_idx, float eps, cudaStream_t stream) {
    if (B <= 0 || L <= 0) return TAG_ERR_BAD_ARG;
    const long rows = (long)B * L;
    const int blocks = (int)((rows * 32 + 255) / 256);
    if (E == 768)
        roberta_embed_ln_kernel<24><<<blocks, 256, 0, stream>>>(ids, word, pos, type0, gamma, beta, out, rows, L, vocab, max_pos, pad_idx, eps);
    else if (E == 512)
        roberta_embed_ln_kernel<16><<<blocks, 256, 0, stream>>>(ids, word, pos, type0, gamma, beta, out, rows, L, vocab, max_pos, pad_idx, eps);
    else
        return TAG_ERR_UNSUPPORTED;
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_add_layernorm(const float* a, const float* res, const float* gamma, const float* beta, float* out,
                                 long rows, int E, float eps, cudaStream_t stream) {
    if (rows <= 0) return TAG_ERR_BAD_ARG;
    const int blocks = (int)((rows * 32 + 255) / 256);
    if (E == 768) add_layernorm_kernel<24><<<blocks, 256, 0, stream>>>(a, res, gamma, beta, out, rows, eps);
    else if (E == 512) add_layernorm_kernel<16><<<blocks, 256, 0, stream>>>(a, res, gamma, beta, out, rows, eps);
    else return TAG_ERR_UNSUPPORTED;
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_unary_f32(const float* in, float* out, long n, int op, cudaStream_t stream) {
    if (n <= 0 || op < 0 || op > 1) return TAG_ERR_BAD_ARG;
    long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    unary_kernel<<<(int)blocks, 256, 0, stream>>>(in, out, n, op);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_l2_normalize(const float* in, float* out, long rows, int E, float eps, cudaStream_t stream) {
    if (rows <= 0 || E <= 0) return TAG_ERR_BAD_ARG;
    l2_normalize_kernel<<<(int)((rows * 32 + 255) / 256), 256, 0, stream>>>(in, out, rows, E, eps);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}
