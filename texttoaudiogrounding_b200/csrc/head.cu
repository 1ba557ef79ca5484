// Text side and matching head: embedding gather + masked mean, frame-wise scaled dot product ->
// sigmoid -> clamp, masked frame BCE; forward and analytic backward.  Warp-shuffle reductions.
//
// Reference: models/text_encoder.py:39-43,79-88 + models/utils.py:33-58 (EmbeddingAgg 'mean'),
// models/match.py:43-60 (DotProduct, l2norm=False), losses.py:12-24 (FrameBceLoss).
#include "common.cuh"

namespace {

// one CTA per sequence; seq[b,:] = sum_{n < len} E[text[b,n], :] / len
__global__ void embed_mean_fwd_kernel(const long long* __restrict__ text, const long long* __restrict__ text_len,
                                      const float* __restrict__ emb, float* __restrict__ token_emb,
                                      float* __restrict__ seq_emb, int N, int D, int vocab) {
    const int b = blockIdx.x;
    const long long len = text_len[b];
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float acc = 0.f;
        for (int n = 0; n < N; ++n) {
            long long id = text[(long)b * N + n];
            if (id < 0) id = 0;
            if (id >= vocab) id = vocab - 1;
            const float v = emb[id * D + d];
            if (token_emb != nullptr) token_emb[((long)b * N + n) * D + d] = v;
            if (n < len) acc += v;
        }
        seq_emb[(long)b * D + d] = acc / (float)len;
    }
}

__global__ void embed_mean_bwd_kernel(const long long* __restrict__ text, const long long* __restrict__ text_len,
                                      const float* __restrict__ d_seq, float* __restrict__ d_emb,
                                      int N, int D, int vocab) {
    const int b = blockIdx.x;
    const long long len = text_len[b];
    const float inv = 1.0f / (float)len;
    for (int n = 0; n < N && n < len; ++n) {
        long long id = text[(long)b * N + n];
        if (id < 0) id = 0;
        if (id >= vocab) id = vocab - 1;
        for (int d = threadIdx.x; d < D; d += blockDim.x)
            atomicAdd(d_emb + id * D + d, d_seq[(long)b * D + d] * inv);
    }
}

// d_emb[text[b,n], :] += d_token[b,n,:] for every position (nn.Embedding backward; pads included)
__global__ void embed_token_bwd_kernel(const long long* __restrict__ text, const float* __restrict__ d_token,
                                       float* __restrict__ d_emb, int D, int vocab) {
    const long row = blockIdx.x;
    long long id = text[row];
    if (id < 0) id = 0;
    if (id >= vocab) id = vocab - 1;
    for (int d = threadIdx.x; d < D; d += blockDim.x) atomicAdd(d_emb + id * D + d, d_token[row * D + d]);
}

// one warp per (b, t): sim = clamp(sigmoid(scale * <a[b,t,:], s[b,:]>), 1e-7, 1)
__global__ void dot_sigmoid_fwd_kernel(const float* __restrict__ audio, const float* __restrict__ seq,
                                       float* __restrict__ sim, float* __restrict__ logits,
                                       long BT, int T, int D, float scale) {
    const long wid = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= BT) return;
    const long b = wid / T;
    const float* a = audio + wid * D;
    const float* s = seq + b * D;
    float acc = 0.f;
    for (int d = lane * 4; d < D; d += 128) {
        const float4 av = *reinterpret_cast<const float4*>(a + d);
        const float4 sv = *reinterpret_cast<const float4*>(s + d);
        acc += av.x * sv.x + av.y * sv.y + av.z * sv.z + av.w * sv.w;
    }
    acc = warp_sum(acc) * scale;
    if (lane == 0) {
        const float p = 1.0f / (1.0f + expf(-acc));
        sim[wid] = fminf(fmaxf(p, 1e-7f), 1.0f);
        if (logits != nullptr) logits[wid] = acc;
    }
}

// d_audio[b,t,:] = dlogit * seq[b,:] * scale ; dlogit = d_sim * p (1-p), zero where clamped
__global__ void dot_sigmoid_bwd_audio_kernel(const float* __restrict__ d_sim, const float* __restrict__ sim,
                                             const float* __restrict__ seq, float* __restrict__ d_audio,
                                             float* __restrict__ d_logit, long BT, int T, int D, float scale) {
    const long wid = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= BT) return;
    const long b = wid / T;
    const float p = sim[wid];
    float g = (p > 1e-7f) ? d_sim[wid] * p * (1.0f - p) : 0.f;
    if (lane == 0) d_logit[wid] = g;
    g *= scale;
    const float* s = seq + b * D;
    float* o = d_audio + wid * D;
    for (int d = lane * 4; d < D; d += 128) {
        const float4 sv = *reinterpret_cast<const float4*>(s + d);
        *reinterpret_cast<float4*>(o + d) = make_float4(g * sv.x, g * sv.y, g * sv.z, g * sv.w);
    }
}

// d_seq[b,:] = scale * sum_t dlogit[b,t] * audio[b,t,:]   (one CTA per b, thread per 2 features)
__global__ void dot_sigmoid_bwd_seq_kernel(const float* __restrict__ d_logit, const float* __restrict__ audio,
                                           float* __restrict__ d_seq, int T, int D, float scale) {
    const int b = blockIdx.x;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float acc = 0.f;
        for (int t = 0; t < T; ++t) acc = fmaf(d_logit[(long)b * T + t], audio[((long)b * T + t) * D + d], acc);
        d_seq[(long)b * D + d] = acc * scale;
    }
}

// loss = sum_{b, t < min(len_b, Tt)} bce(sim[b,t], label[b,t]) / sum_b min(len_b, Tt)
// (single CTA; B*T is at most a few 10^4).  Also writes d_sim (d loss / d sim) when asked.
__global__ void frame_bce_kernel(const float* __restrict__ sim, long sim_stride,
                                 const float* __restrict__ label, long label_stride,
                                 const long long* __restrict__ length, int B, int Tt,
                                 float* __restrict__ loss_out, float* __restrict__ d_sim, long dsim_stride,
                                 float grad_scale) {
    __shared__ float red[32];
    __shared__ float s_cnt;
    float cnt = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        long long l = length[b];
        if (l > Tt) l = Tt;
        if (l < 0) l = 0;
        cnt += (float)l;
    }
    cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        float c = 0.f;
        for (int i = 0; i < (blockDim.x >> 5); ++i) c += red[i];
        s_cnt = c;
    }
    __syncthreads();
    const float total = s_cnt;
    float acc = 0.f;
    for (long i = threadIdx.x; i < (long)B * Tt; i += blockDim.x) {
        const int b = (int)(i / Tt), t = (int)(i % Tt);
        const bool on = t < length[b];
        const float p = sim[b * sim_stride + t];
        const float y = label[b * label_stride + t];
        if (on) {
            const float lp = fmaxf(logf(p), -100.f);
            const float lq = fmaxf(log1pf(-p), -100.f);
            acc -= y * lp + (1.f - y) * lq;
        }
        if (d_sim != nullptr) {
            float g = 0.f;
            if (on) {
                // torch's binary_cross_entropy_backward: (p - y) / max((1 - p) * p, 1e-12)
                g = (p - y) / fmaxf((1.f - p) * p, 1e-12f) * grad_scale / total;
            }
            d_sim[b * dsim_stride + t] = g;
        }
    }
    __syncthreads();
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0 && loss_out != nullptr) {
        float c = 0.f;
        for (int i = 0; i < (blockDim.x >> 5); ++i) c += red[i];
        *loss_out = c / total;
    }
}


// ---- normalised match heads (seq-level text): one warp per (b, t) row, D = 512
//   mode 1: DotProduct(l2norm=True)  sim = clamp(sigmoid(scale * <a^, s^>), 1e-7, 1)     (models/match.py:43-60; cosine)
//   mode 2: ExpNegL2(l2norm=True)    sim = exp(-|a^ - s^|)                                (models/match.py:10-33)
//   mode 3: ExpNegL2(l2norm=False)   sim = exp(-|a - s|)
// a^ = a / max(|a|, 1e-12) (F.normalize).  Backward: d_audio overwritten, d_seq accumulated (pre-zeroed).
constexpr int MN_EL = 16;       // 512 / 32

__device__ __forceinline__ void mn_load(const float* __restrict__ p, int lane, float (&v)[MN_EL]) {
#pragma unroll
    for (int i = 0; i < MN_EL; ++i) v[i] = p[lane + 32 * i];
}

__global__ void match_norm_fwd_kernel(const float* __restrict__ audio, const float* __restrict__ seq,
                                      float* __restrict__ sim, long BT, int T, int mode, float scale) {
    const long wid = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= BT) return;
    float a[MN_EL], s[MN_EL];
    mn_load(audio + wid * 512, lane, a);
    mn_load(seq + (wid / T) * 512, lane, s);
    float ia = 1.f, is = 1.f;
    if (mode != 3) {
        float aa = 0.f, ss = 0.f;
#pragma unroll
        for (int i = 0; i < MN_EL; ++i) { aa = fmaf(a[i], a[i], aa); ss = fmaf(s[i], s[i], ss); }
        ia = 1.0f / fmaxf(sqrtf(warp_sum(aa)), 1e-12f);
        is = 1.0f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < MN_EL; ++i) {
        const float x = a[i] * ia, y = s[i] * is;
        if (mode == 1) acc = fmaf(x, y, acc);
        else { const float d = x - y; acc = fmaf(d, d, acc); }
    }
    acc = warp_sum(acc);
    if (lane == 0)
        sim[wid] = mode == 1 ? fminf(fmaxf(1.0f / (1.0f + expf(-acc * scale)), 1e-7f), 1.0f) : expf(-sqrtf(acc));
}

__global__ void match_norm_bwd_kernel(const float* __restrict__ d_sim, const float* __restrict__ sim,
                                      const float* __restrict__ audio, const float* __restrict__ seq,
                                      float* __restrict__ d_audio, float* __restrict__ d_seq, long BT, int T, int mode,
                                      float scale) {
    const long wid = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= BT) return;
    const long b = wid / T;
    float a[MN_EL], s[MN_EL];
    mn_load(audio + wid * 512, lane, a);
    mn_load(seq + b * 512, lane, s);
    float ia = 1.f, is = 1.f;
    if (mode != 3) {
        float aa = 0.f, ss = 0.f;
#pragma unroll
        for (int i = 0; i < MN_EL; ++i) { aa = fmaf(a[i], a[i], aa); ss = fmaf(s[i], s[i], ss); }
        ia = 1.0f / fmaxf(sqrtf(warp_sum(aa)), 1e-12f);
        is = 1.0f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
    }
    // normalised rows, their dot product c and (modes 2, 3) the distance n
    float c = 0.f, n2 = 0.f;
#pragma unroll
    for (int i = 0; i < MN_EL; ++i) {
        a[i] *= ia; s[i] *= is;
        c = fmaf(a[i], s[i], c);
        const float d = a[i] - s[i];
        n2 = fmaf(d, d, n2);
    }
    c = warp_sum(c);
    n2 = warp_sum(n2);
    const float p = sim[wid], g = d_sim[wid];
    // d(loss)/d(a^) = ka * a^ + ks * s^ written per mode; then through the normalisation (modes 1, 2):
    //   da = (dA - a^ <a^, dA>) / |a|,  ds = (dS - s^ <s^, dS>) / |s|
    float da_a, da_s, ds_a, ds_s;          // coefficients of (a^, s^) in dA and dS
    if (mode == 1) {
        const float dc = p > 1e-7f ? g * p * (1.0f - p) * scale : 0.f;
        da_a = 0.f; da_s = dc; ds_a = dc; ds_s = 0.f;
    } else {
        const float n = sqrtf(n2);
        const float k = n > 1e-12f ? -g * p / n : 0.f;         // d(loss)/d(diff) = k * diff
        da_a = k; da_s = -k; ds_a = -k; ds_s = k;
    }
    if (mode != 3) {
        // <a^, dA> = da_a + da_s c ;  <s^, dS> = ds_a c + ds_s   (unit vectors)
        const float pa = da_a + da_s * c, ps = ds_a * c + ds_s;
        da_a = (da_a - pa) * ia; da_s = da_s * ia;
        ds_s = (ds_s - ps) * is; ds_a = ds_a * is;
    }
#pragma unroll
    for (int i = 0; i < MN_EL; ++i) {
        d_audio[wid * 512 + lane + 32 * i] = fmaf(da_a, a[i], da_s * s[i]);
        atomicAdd(d_seq + b * 512 + lane + 32 * i, fmaf(ds_a, a[i], ds_s * s[i]));
    }
}

// ---- EmbeddingAgg(aggregation="attention") (models/text_encoder.py:46-58,84-85): score[n] = <x[b,n,:], w> + bias,
// masked to n < len with -1e10, softmax over n, out = sum_n weight[n] x[b,n,:].  One CTA (128 threads) per sequence.
constexpr int AP_MAX_N = 128;

__global__ void attn_pool_fwd_kernel(const float* __restrict__ x, const long long* __restrict__ lens,
                                     const float* __restrict__ w, const float* __restrict__ bias,
                                     float* __restrict__ out, float* __restrict__ weight, int N, int D) {
    __shared__ float s_score[AP_MAX_N];
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const long long len = lens[b];
    const float* xb = x + (long)b * N * D;
    for (int n = warp; n < N; n += nwarp) {
        float acc = 0.f;
        for (int d = lane; d < D; d += 32) acc = fmaf(xb[(long)n * D + d], w[d], acc);
        acc = warp_sum(acc);
        if (lane == 0) s_score[n] = n < len ? acc + bias[0] : -1e10f;
    }
    __syncthreads();
    if (warp == 0) {
        float m = -INFINITY;
        for (int n = lane; n < N; n += 32) m = fmaxf(m, s_score[n]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float z = 0.f;
        for (int n = lane; n < N; n += 32) z += expf(s_score[n] - m);
        z = warp_sum(z);
        for (int n = lane; n < N; n += 32) {
            const float p = expf(s_score[n] - m) / z;
            s_score[n] = p;
            weight[(long)b * N + n] = p;
        }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float acc = 0.f;
        for (int n = 0; n < N; ++n) acc = fmaf(s_score[n], xb[(long)n * D + d], acc);
        out[(long)b * D + d] = acc;
    }
}

// d_x[b,n,:] = weight[n] d_out + d_score[n] w ; d_score[n] = weight[n] (g[n] - sum_m weight[m] g[m]), g[n] = <d_out, x[n]>
// d_w += sum_n d_score[n] x[n,:] ; d_bias += sum_n d_score[n]
__global__ void attn_pool_bwd_kernel(const float* __restrict__ d_out, const float* __restrict__ x,
                                     const float* __restrict__ w, const float* __restrict__ weight,
                                     float* __restrict__ d_x, float* __restrict__ d_w, float* __restrict__ d_bias,
                                     int N, int D) {
    __shared__ float s_g[AP_MAX_N];
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const float* xb = x + (long)b * N * D;
    const float* go = d_out + (long)b * D;
    const float* pw = weight + (long)b * N;
    for (int n = warp; n < N; n += nwarp) {
        float acc = 0.f;
        for (int d = lane; d < D; d += 32) acc = fmaf(go[d], xb[(long)n * D + d], acc);
        acc = warp_sum(acc);
        if (lane == 0) s_g[n] = acc;
    }
    __syncthreads();
    if (warp == 0) {
        float mean = 0.f;
        for (int n = lane; n < N; n += 32) mean = fmaf(pw[n], s_g[n], mean);
        mean = warp_sum(mean);
        float tot = 0.f;
        for (int n = lane; n < N; n += 32) {
            const float ds = pw[n] * (s_g[n] - mean);
            s_g[n] = ds;
            tot += ds;
        }
        tot = warp_sum(tot);
        if (lane == 0) atomicAdd(d_bias, tot);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float g = go[d], wd = w[d];
        float dw = 0.f;
        for (int n = 0; n < N; ++n) {
            const float ds = s_g[n];
            d_x[((long)b * N + n) * D + d] = fmaf(pw[n], g, ds * wd);
            dw = fmaf(ds, xb[(long)n * D + d], dw);
        }
        atomicAdd(d_w + d, dw);
    }
}

// ---- F.interpolate(mode="linear", align_corners=False) along T to `To` frames (BiEncoder upsample=True,
// models/audio_text_model.py:90-97 and :216-223).  x [outer, T, inner] -> y [outer, To, inner].
// ATen: scale = T / To (float), src = scale * (dst + 0.5) - 0.5 clamped at 0; i0 = min(floor(src), T-1); i1 = min(i0+1, T-1).
__device__ __forceinline__ void upsample_src(int dst, int T, float scale, int* i0, int* i1, float* lam) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    int a = (int)src;
    if (a > T - 1) a = T - 1;
    *i0 = a;
    *i1 = a + (a < T - 1 ? 1 : 0);
    *lam = fminf(fmaxf(src - (float)a, 0.f), 1.f);
}

__global__ void upsample_linear_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long total, int T,
                                           int To, int inner) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % inner);
    const int t = (int)((i / inner) % To);
    const long o = i / ((long)inner * To);
    int i0, i1; float lam;
    upsample_src(t, T, (float)T / (float)To, &i0, &i1, &lam);
    const float* xo = x + o * (long)T * inner + c;
    y[i] = (1.0f - lam) * xo[(long)i0 * inner] + lam * xo[(long)i1 * inner];
}

// transpose of the above: every output frame scatters into its two source frames (dx pre-zeroed by the caller)
__global__ void upsample_linear_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, long total, int T,
                                           int To, int inner) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % inner);
    const int t = (int)((i / inner) % To);
    const long o = i / ((long)inner * To);
    int i0, i1; float lam;
    upsample_src(t, T, (float)T / (float)To, &i0, &i1, &lam);
    float* xo = dx + o * (long)T * inner + c;
    const float g = dy[i];
    atomicAdd(xo + (long)i0 * inner, (1.0f - lam) * g);
    atomicAdd(xo + (long)i1 * inner, lam * g);
}

}  // namespace

extern "C" int tag_embed_mean_fwd(const long long* text, const long long* text_len, const float* emb,
                                  float* token_emb, float* seq_emb, int B, int N, int D, int vocab,
                                  cudaStream_t stream) {
    if (B <= 0) return TAG_ERR_BAD_ARG;
    embed_mean_fwd_kernel<<<B, 128, 0, stream>>>(text, text_len, emb, token_emb, seq_emb, N, D, vocab);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_embed_mean_bwd(const long long* text, const long long* text_len, const float* d_seq,
                                  float* d_emb, int B, int N, int D, int vocab, cudaStream_t stream) {
    if (B <= 0) return TAG_ERR_BAD_ARG;
    embed_mean_bwd_kernel<<<B, 128, 0, stream>>>(text, text_len, d_seq, d_emb, N, D, vocab);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_embed_token_bwd(const long long* text, const float* d_token, float* d_emb, int B, int N, int D,
                                   int vocab, cudaStream_t stream) {
    if (B <= 0 || N <= 0) return TAG_ERR_BAD_ARG;
    embed_token_bwd_kernel<<<B * N, 128, 0, stream>>>(text, d_token, d_emb, D, vocab);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_dot_sigmoid_fwd(const float* audio, const float* seq, float* sim, float* logits,
                                   int B, int T, int D, float scale, cudaStream_t stream) {
    if (D % 4 != 0) return TAG_ERR_BAD_ARG;
    const long BT = (long)B * T;
    const int blocks = (int)((BT * 32 + 255) / 256);
    dot_sigmoid_fwd_kernel<<<blocks, 256, 0, stream>>>(audio, seq, sim, logits, BT, T, D, scale);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_dot_sigmoid_bwd(const float* d_sim, const float* sim, const float* audio,
                                   const float* seq, float* d_audio, float* d_seq, float* d_logit_ws,
                                   int B, int T, int D, float scale, cudaStream_t stream) {
    if (D % 4 != 0) return TAG_ERR_BAD_ARG;
    const long BT = (long)B * T;
    const int blocks = (int)((BT * 32 + 255) / 256);
    dot_sigmoid_bwd_audio_kernel<<<blocks, 256, 0, stream>>>(d_sim, sim, seq, d_audio, d_logit_ws, BT, T, D, scale);
    TAG_RETURN_IF_LAUNCH_FAILED();
    dot_sigmoid_bwd_seq_kernel<<<B, 256, 0, stream>>>(d_logit_ws, audio, d_seq, T, D, scale);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_frame_bce(const float* sim, long sim_stride, const float* label, long label_stride,
                             const long long* length, int B, int Tt, float* loss_out, float* d_sim,
                             long dsim_stride, float grad_scale, cudaStream_t stream) {
    if (B <= 0 || Tt <= 0) return TAG_ERR_BAD_ARG;
    frame_bce_kernel<<<1, 1024, 0, stream>>>(sim, sim_stride, label, label_stride, length, B, Tt, loss_out,
                                             d_sim, dsim_stride, grad_scale);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_match_norm_fwd(const float* audio, const float* seq, float* sim, int B, int T, int D, int mode,
                                  float scale, cudaStream_t stream) {
    if (B <= 0 || T <= 0 || mode < 1 || mode > 3) return TAG_ERR_BAD_ARG;
    if (D != 512) return TAG_ERR_UNSUPPORTED;
    const long BT = (long)B * T;
    match_norm_fwd_kernel<<<(int)((BT * 32 + 255) / 256), 256, 0, stream>>>(audio, seq, sim, BT, T, mode, scale);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_match_norm_bwd(const float* d_sim, const float* sim, const float* audio, const float* seq,
                                  float* d_audio, float* d_seq, int B, int T, int D, int mode, float scale,
                                  cudaStream_t stream) {
    if (B <= 0 || T <= 0 || mode < 1 || mode > 3) return TAG_ERR_BAD_ARG;
    if (D != 512) return TAG_ERR_UNSUPPORTED;
    const long BT = (long)B * T;
    match_norm_bwd_kernel<<<(int)((BT * 32 + 255) / 256), 256, 0, stream>>>(d_sim, sim, audio, seq, d_audio, d_seq, BT, T,
                                                                          mode, scale);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_attn_pool_fwd(const float* x, const long long* lens, const float* w, const float* bias, float* out,
                                 float* weight, int B, int N, int D, cudaStream_t stream) {
    if (B <= 0 || N <= 0 || D <= 0) return TAG_ERR_BAD_ARG;
    if (N > AP_MAX_N) return TAG_ERR_UNSUPPORTED;
    attn_pool_fwd_kernel<<<B, 128, 0, stream>>>(x, lens, w, bias, out, weight, N, D);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_attn_pool_bwd(const float* d_out, const float* x, const float* w, const float* weight, float* d_x,
                                 float* d_w, float* d_bias, int B, int N, int D, cudaStream_t stream) {
    if (B <= 0 || N <= 0 || D <= 0) return TAG_ERR_BAD_ARG;
    if (N > AP_MAX_N) return TAG_ERR_UNSUPPORTED;
    attn_pool_bwd_kernel<<<B, 128, 0, stream>>>(d_out, x, w, weight, d_x, d_w, d_bias, N, D);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_upsample_linear_fwd(const float* x, float* y, long outer, int T, int To, int inner,
                                       cudaStream_t stream) {
    if (outer <= 0 || T <= 0 || inner <= 0 || To <= 0) return TAG_ERR_BAD_ARG;
    const long total = outer * To * inner;
    upsample_linear_fwd_kernel<<<(int)((total + 255) / 256), 256, 0, stream>>>(x, y, total, T, To, inner);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_upsample_linear_bwd(const float* dy, float* dx, long outer, int T, int To, int inner,
                                       cudaStream_t stream) {
    if (outer <= 0 || T <= 0 || inner <= 0 || To <= 0) return TAG_ERR_BAD_ARG;
    const long total = outer * To * inner;
    cudaError_t e = cudaMemsetAsync(dx, 0, sizeof(float) * outer * T * inner, stream);
    if (e != cudaSuccess) return (int)e;
    upsample_linear_bwd_kernel<<<(int)((total + 255) / 256), 256, 0, stream>>>(dy, dx, total, T, To, inner);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}
