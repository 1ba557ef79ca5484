// Attention-type heads of the later configurations (BASELINE.json configs[3]):
//   * text_encoder.SelfAttention (models/text_encoder.py:240-268): [cls; embedding] + positional encoding -> one
//     nn.MultiheadAttention layer with a key padding mask                      -> text_assemble_*, mha_core_*
//   * match.CrossAttention (models/match.py:63-88): MHA(audio, text, text) -> residual + dropout -> LayerNorm ->
//     Linear(E, 1) -> sigmoid                                                  -> mha_core_*, ln_linear_sigmoid_*
//   * cross_encoder.CrossAttentionGating (models/cross_encoder.py:5-79): additive (tanh) attention of every frame over
//     the tokens + sigmoid cross gating.  The reference materialises [B, T*N, 2E] (q_repeat / kv_repeat / cat) and
//     [B, T*N, E]; here h2attn is split into its query and key halves (two small GEMMs) and the tanh / v-dot /
//     softmax / weighted sum run per frame in registers                        -> additive_attn_*, sigmoid_gate_*
//   * match.DotProduct(text_level="token") on the cross-encoded per-frame text -> rowdot_sigmoid_*
// The dense projections (in_proj, out_proj, fc_u, fc_s, h2attn halves) are the fp32 GEMMs of conv_simt.cu (taps = 1).
// All kernels fp32, E = 512, one warp per row; sizes are tiny next to the audio encoder (T' = 250 frames, N <= 32).
#include "common.cuh"

namespace {

constexpr int AT_E = 512;
constexpr int AT_EL = AT_E / 32;        // elements of a row per lane (e = lane + 32 i)
constexpr int AT_MAXK = 128;            // keys per attention row (4 per lane)

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---------------------------------------------------------------- text assemble: [cls; emb[text]] + pe, dropout
__global__ void text_assemble_fwd_kernel(const long long* __restrict__ text, const float* __restrict__ emb,
                                         const float* __restrict__ cls, const float* __restrict__ pe,
                                         float* __restrict__ out, int N, int E, int vocab, uint32_t thresh,
                                         float keep, uint64_t seed, const uint64_t* __restrict__ seed_dev) {
    const int b = blockIdx.y, pos = blockIdx.x;              // pos 0 = cls, 1..N = tokens
    if (seed_dev != nullptr) seed += *seed_dev;
    const float* src = cls;
    if (pos > 0) {
        long long id = text[(long)b * N + pos - 1];
        if (id < 0) id = 0;
        if (id >= vocab) id = vocab - 1;
        src = emb + id * E;
    }
    const long row = (long)b * (N + 1) + pos;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        float v = src[e] + pe[(long)pos * E + e];
        if (thresh != 0u) v *= tag_dropout_scale(seed, (uint64_t)(row * E + e), thresh, keep);
        out[row * E + e] = v;
    }
}

__global__ void text_assemble_bwd_kernel(const long long* __restrict__ text, const float* __restrict__ d_out,
                                         float* __restrict__ d_emb, float* __restrict__ d_cls, int N, int E, int vocab,
                                         uint32_t thresh, float keep, uint64_t seed,
                                         const uint64_t* __restrict__ seed_dev) {
    const int b = blockIdx.y, pos = blockIdx.x;
    if (seed_dev != nullptr) seed += *seed_dev;
    float* dst = d_cls;
    if (pos > 0) {
        long long id = text[(long)b * N + pos - 1];
        if (id < 0) id = 0;
        if (id >= vocab) id = vocab - 1;
        dst = d_emb + id * E;
    }
    const long row = (long)b * (N + 1) + pos;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        float g = d_out[row * E + e];
        if (thresh != 0u) g *= tag_dropout_scale(seed, (uint64_t)(row * E + e), thresh, keep);
        atomicAdd(dst + e, g);
    }
}

// ---------------------------------------------------------------- multi-head attention core
// grid (heads, B), 128 threads.  q [B,Lq,E], k/v [B,Lk,E] are the projected tensors; head h owns columns
// [h*dh, (h+1)*dh).  probs [B,heads,Lq,Lk] keeps the softmax (before dropout) for backward.
template <int DHL>      // dh / 32
__global__ void __launch_bounds__(128)
mha_core_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                    const long long* __restrict__ key_len, float* __restrict__ out, float* __restrict__ probs,
                    int Lq, int Lk, int E, float scale, uint32_t thresh, float keep, uint64_t seed,
                    const uint64_t* __restrict__ seed_dev) {
    constexpr int DH = DHL * 32;
    extern __shared__ __align__(16) float sm[];
    float* sK = sm;                     // [Lk][DH]
    float* sV = sm + (size_t)Lk * DH;
    const int h = blockIdx.x, b = blockIdx.y, heads = gridDim.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (seed_dev != nullptr) seed += *seed_dev;
    for (int i = threadIdx.x; i < Lk * DH; i += 128) {
        const int n = i / DH, d = i - n * DH;
        sK[i] = k[((long)b * Lk + n) * E + h * DH + d];
        sV[i] = v[((long)b * Lk + n) * E + h * DH + d];
    }
    __syncthreads();
    const int klen = key_len != nullptr ? (int)min((long long)Lk, key_len[b]) : Lk;
    for (int t = warp; t < Lq; t += 4) {
        float qr[DHL];
#pragma unroll
        for (int i = 0; i < DHL; ++i) qr[i] = q[((long)b * Lq + t) * E + h * DH + lane + 32 * i] * scale;
        float sc[AT_MAXK / 32];
#pragma unroll
        for (int s = 0; s < AT_MAXK / 32; ++s) sc[s] = -INFINITY;
        for (int n = 0; n < klen; ++n) {
            float p = 0.f;
#pragma unroll
            for (int i = 0; i < DHL; ++i) p = fmaf(qr[i], sK[n * DH + lane + 32 * i], p);
            p = warp_sum(p);
#pragma unroll
            for (int s = 0; s < AT_MAXK / 32; ++s) if ((n >> 5) == s && (n & 31) == lane) sc[s] = p;
        }
        float m = -INFINITY;
#pragma unroll
        for (int s = 0; s < AT_MAXK / 32; ++s) m = fmaxf(m, sc[s]);
        m = warp_max(m);
        float sum = 0.f;
#pragma unroll
        for (int s = 0; s < AT_MAXK / 32; ++s) { sc[s] = sc[s] == -INFINITY ? 0.f : expf(sc[s] - m); sum += sc[s]; }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        const long prow = (((long)b * heads + h) * Lq + t) * Lk;
#pragma unroll
        for (int s = 0; s < AT_MAXK / 32; ++s) {
            const int n = lane + 32 * s;
            sc[s] *= inv;
            if (n < Lk) {
                probs[prow + n] = sc[s];
                if (thresh != 0u) sc[s] *= tag_dropout_scale(seed, (uint64_t)(prow + n), thresh, keep);
            }
        }
        float acc[DHL];
#pragma unroll
        for (int i = 0; i < DHL; ++i) acc[i] = 0.f;
        for (int n = 0; n < klen; ++n) {
            float pn = 0.f;
#pragma unroll
            for (int s = 0; s < AT_MAXK / 32; ++s) { const float c = __shfl_sync(0xffffffffu, sc[s], n & 31); if ((n >> 5) == s) pn = c; }
#pragma unroll
            for (int i = 0; i < DHL; ++i) acc[i] = fmaf(pn, sV[n * DH + lane + 32 * i], acc[i]);
        }
#pragma unroll
        for (int i = 0; i < DHL; ++i) out[((long)b * Lq + t) * E + h * DH + lane + 32 * i] = acc[i];
    }
}

template <int DHL>
__global__ void __launch_bounds__(128)
mha_core_bwd_kernel(const float* __restrict__ d_out, const float* __restrict__ q, const float* __restrict__ k,
                    const float* __restrict__ v, const float* __restrict__ probs,
                    const long long* __restrict__ key_len, float* __restrict__ dq, float* __restrict__ dk,
                    float* __restrict__ dv, int Lq, int Lk, int E, float scale, uint32_t thresh, float keep,
                    uint64_t seed, const uint64_t* __restrict__ seed_dev) {
    constexpr int DH = DHL * 32;
    extern __shared__ __align__(16) float sm[];
    float* sK = sm;
    float* sV = sm + (size_t)Lk * DH;
    float* sdK = sm + (size_t)2 * Lk * DH;
    float* sdV = sm + (size_t)3 * Lk * DH;
    const int h = blockIdx.x, b = blockIdx.y, heads = gridDim.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (seed_dev != nullptr) seed += *seed_dev;
    for (int i = threadIdx.x; i < Lk * DH; i += 128) {
        const int n = i / DH, d = i - n * DH;
        sK[i] = k[((long)b * Lk + n) * E + h * DH + d];
        sV[i] = v[((long)b * Lk + n) * E + h * DH + d];
        sdK[i] = 0.f;
        sdV[i] = 0.f;
    }
    __syncthreads();
    const int klen = key_len != nullptr ? (int)min((long long)Lk, key_len[b]) : Lk;
    for (int t = warp; t < Lq; t += 4) {
        float qr[DHL], go[DHL];
#pragma unroll
        for (int i = 0; i < DHL; ++i) {
            qr[i] = q[((long)b * Lq + t) * E + h * DH + lane + 32 * i];
            go[i] = d_out[((long)b * Lq + t) * E + h * DH + lane + 32 * i];
        }
        const long prow = (((long)b * heads + h) * Lq + t) * Lk;
        float p[AT_MAXK / 32], ks[AT_MAXK / 32], dp[AT_MAXK / 32];
#pragma unroll
        for (int s = 0; s < AT_MAXK / 32; ++s) {
            const int n = lane + 32 * s;
            p[s] = n < Lk ? probs[prow + n] : 0.f;
            ks[s] = (thresh != 0u && n < Lk) ? tag_dropout_scale(seed, (uint64_t)(prow + n), thresh, keep) : 1.f;
            dp[s] = 0.f;
        }
        for (int n = 0; n < klen; ++n) {
            float a = 0.f;
#pragma unroll
            for (int i = 0; i < DHL; ++i) a = fmaf(go[i], sV[n * DH + lane + 32 * i], a);
            a = warp_sum(a);                                    // d(loss)/d(dropped prob n)
            float pdn = 0.f;
#pragma unroll
            for (int s = 0; s < AT_MAXK / 32; ++s) {
                const float c = __shfl_sync(0xffffffffu, p[s] * ks[s], n & 31);
                if ((n >> 5) == s) { pdn = c; if ((n & 31) == lane) dp[s] = a * ks[s]; }
            }
#pragma unroll
            for (int i = 0; i < DHL; ++i) atomicAdd(sdV + n * DH + lane + 32 * i, pdn * go[i]);
        }
        float dot = 0.f;
#pragma unroll
        for (int s = 0; s < AT_MAXK / 32; ++s) dot = fmaf(p[s], dp[s], dot);
        dot = warp_sum(dot);
        float ds[AT_MAXK / 32];
#pragma unroll
        for (int s = 0; s < AT_MAXK / 32; ++s) ds[s] = p[s] * (dp[s] - dot);
        float dqr[DHL];
#pragma unroll
        for (int i = 0; i < DHL; ++i) dqr[i] = 0.f;
        for (int n = 0; n < klen; ++n) {
            float dsn = 0.f;
#pragma unroll
            for (int s = 0; s < AT_MAXK / 32; ++s) { const float c = __shfl_sync(0xffffffffu, ds[s], n & 31); if ((n >> 5) == s) dsn = c; }
#pragma unroll
            for (int i = 0; i < DHL; ++i) {
                dqr[i] = fmaf(dsn, sK[n * DH + lane + 32 * i], dqr[i]);
                atomicAdd(sdK + n * DH + lane + 32 * i, dsn * scale * qr[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < DHL; ++i) dq[((long)b * Lq + t) * E + h * DH + lane + 32 * i] = dqr[i] * scale;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Lk * DH; i += 128) {
        const int n = i / DH, d = i - n * DH;
        dk[((long)b * Lk + n) * E + h * DH + d] = sdK[i];
        dv[((long)b * Lk + n) * E + h * DH + d] = sdV[i];
    }
}

// ---------------------------------------------------------------- additive attention (Seq2SeqAttention)
// score[b,t,n] = sum_a v[a] tanh(hq[b,t,a] + hk[b,n,a]); masked (t >= q_len[b] or n >= kv_len[b]) entries are -1e10;
// attn = softmax_n(score); out[b,t,:] = sum_n attn[n] kv[b,n,:].     grid (ceil(T / 32), B), 256 threads.
constexpr int AA_TT = 32;

__global__ void __launch_bounds__(256)
additive_attn_fwd_kernel(const float* __restrict__ hq, const float* __restrict__ hk, const float* __restrict__ vvec,
                         const float* __restrict__ kv, const long long* __restrict__ q_len,
                         const long long* __restrict__ kv_len, float* __restrict__ attn, float* __restrict__ out,
                         int T, int N) {
    extern __shared__ __align__(16) float sm[];
    float* sHk = sm;                          // [N][E]
    float* sKv = sm + (size_t)N * AT_E;       // [N][E]
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < N * AT_E; i += 256) {
        sHk[i] = hk[(long)b * N * AT_E + i];
        sKv[i] = kv[(long)b * N * AT_E + i];
    }
    __syncthreads();
    float vr[AT_EL];
#pragma unroll
    for (int i = 0; i < AT_EL; ++i) vr[i] = vvec[lane + 32 * i];
    const long long ql = q_len[b], kl = kv_len[b];
    const int t_end = min(T, (int)(blockIdx.x + 1) * AA_TT);
    for (int t = blockIdx.x * AA_TT + warp; t < t_end; t += 8) {
        float hr[AT_EL];
#pragma unroll
        for (int i = 0; i < AT_EL; ++i) hr[i] = hq[((long)b * T + t) * AT_E + lane + 32 * i];
        float sc = -INFINITY;                 // lane n holds score n
        for (int n = 0; n < N; ++n) {
            float p = 0.f;
#pragma unroll
            for (int i = 0; i < AT_EL; ++i) p = fmaf(vr[i], tanhf(hr[i] + sHk[n * AT_E + lane + 32 * i]), p);
            p = warp_sum(p);
            if (t >= ql || n >= kl) p = -1e10f;
            if (lane == n) sc = p;
        }
        const float m = warp_max(sc);
        float e = lane < N ? expf(sc - m) : 0.f;
        const float inv = 1.0f / warp_sum(e);
        e *= inv;
        if (lane < N) attn[((long)b * T + t) * N + lane] = e;
        float acc[AT_EL];
#pragma unroll
        for (int i = 0; i < AT_EL; ++i) acc[i] = 0.f;
        for (int n = 0; n < N; ++n) {
            const float an = __shfl_sync(0xffffffffu, e, n);
#pragma unroll
            for (int i = 0; i < AT_EL; ++i) acc[i] = fmaf(an, sKv[n * AT_E + lane + 32 * i], acc[i]);
        }
#pragma unroll
        for (int i = 0; i < AT_EL; ++i) out[((long)b * T + t) * AT_E + lane + 32 * i] = acc[i];
    }
}

// d_hq [B,T,E] overwritten; d_hk, d_kv [B,N,E] and d_v [E] accumulated (pre-zeroed by the caller)
__global__ void __launch_bounds__(256)
additive_attn_bwd_kernel(const float* __restrict__ d_out, const float* __restrict__ hq, const float* __restrict__ hk,
                         const float* __restrict__ vvec, const float* __restrict__ kv,
                         const float* __restrict__ attn, const long long* __restrict__ q_len,
                         const long long* __restrict__ kv_len, float* __restrict__ d_hq, float* __restrict__ d_hk,
                         float* __restrict__ d_v, float* __restrict__ d_kv, int T, int N) {
    extern __shared__ __align__(16) float sm[];
    float* sHk = sm;
    float* sKv = sm + (size_t)N * AT_E;
    float* sdHk = sm + (size_t)2 * N * AT_E;
    float* sdKv = sm + (size_t)3 * N * AT_E;
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < N * AT_E; i += 256) {
        sHk[i] = hk[(long)b * N * AT_E + i];
        sKv[i] = kv[(long)b * N * AT_E + i];
        sdHk[i] = 0.f;
        sdKv[i] = 0.f;
    }
    __syncthreads();
    float vr[AT_EL], dvr[AT_EL];
#pragma unroll
    for (int i = 0; i < AT_EL; ++i) { vr[i] = vvec[lane + 32 * i]; dvr[i] = 0.f; }
    const long long ql = q_len[b], kl = kv_len[b];
    const int t_end = min(T, (int)(blockIdx.x + 1) * AA_TT);
    for (int t = blockIdx.x * AA_TT + warp; t < t_end; t += 8) {
        float hr[AT_EL], go[AT_EL], dh[AT_EL];
#pragma unroll
        for (int i = 0; i < AT_EL; ++i) {
            hr[i] = hq[((long)b * T + t) * AT_E + lane + 32 * i];
            go[i] = d_out[((long)b * T + t) * AT_E + lane + 32 * i];
            dh[i] = 0.f;
        }
        const float a = lane < N ? attn[((long)b * T + t) * N + lane] : 0.f;
        float da = 0.f;                       // lane n: d(loss)/d(attn n)
        for (int n = 0; n < N; ++n) {
            const float an = __shfl_sync(0xffffffffu, a, n);
            float p = 0.f;
#pragma unroll
            for (int i = 0; i < AT_EL; ++i) {
                p = fmaf(go[i], sKv[n * AT_E + lane + 32 * i], p);
                atomicAdd(sdKv + n * AT_E + lane + 32 * i, an * go[i]);
            }
            p = warp_sum(p);
            if (lane == n) da = p;
        }
        const float dot = warp_sum(a * da);
        float dsc = a * (da - dot);
        if (t >= ql || lane >= kl) dsc = 0.f;          // masked_fill cuts the gradient of masked scores
        for (int n = 0; n < N; ++n) {
            const float dn = __shfl_sync(0xffffffffu, dsc, n);
            if (dn == 0.f) continue;                   // warp-uniform
#pragma unroll
            for (int i = 0; i < AT_EL; ++i) {
                const float th = tanhf(hr[i] + sHk[n * AT_E + lane + 32 * i]);
                const float dpre = dn * vr[i] * (1.0f - th * th);
                dh[i] += dpre;
                atomicAdd(sdHk + n * AT_E + lane + 32 * i, dpre);
                dvr[i] = fmaf(dn, th, dvr[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < AT_EL; ++i) d_hq[((long)b * T + t) * AT_E + lane + 32 * i] = dh[i];
    }
#pragma unroll
    for (int i = 0; i < AT_EL; ++i) atomicAdd(d_v + lane + 32 * i, dvr[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < N * AT_E; i += 256) {
        atomicAdd(d_hk + (long)b * N * AT_E + i, sdHk[i]);
        atomicAdd(d_kv + (long)b * N * AT_E + i, sdKv[i]);
    }
}

// ---------------------------------------------------------------- out = x * sigmoid(z)  (CrossGating)
__global__ void sigmoid_gate_fwd_kernel(const float* __restrict__ x, const float* __restrict__ z,
                                        float* __restrict__ out, long n) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        out[i] = x[i] * sigmoidf_(z[i]);
}
__global__ void sigmoid_gate_bwd_kernel(const float* __restrict__ d_out, const float* __restrict__ x,
                                        const float* __restrict__ z, float* __restrict__ dx, float* __restrict__ dz,
                                        long n) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float s = sigmoidf_(z[i]), g = d_out[i];
        dx[i] = g * s;
        dz[i] = g * x[i] * s * (1.0f - s);
    }
}

// ---------------------------------------------------------------- sim[r] = clamp(sigmoid(scale <a[r,:], x[r,:]>), 1e-7, 1)
__global__ void rowdot_sigmoid_fwd_kernel(const float* __restrict__ a, const float* __restrict__ x,
                                          float* __restrict__ sim, long R, float scale) {
    const long r = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    float p = 0.f;
#pragma unroll
    for (int i = 0; i < AT_EL; ++i) p = fmaf(a[r * AT_E + lane + 32 * i], x[r * AT_E + lane + 32 * i], p);
    p = warp_sum(p) * scale;
    if (lane == 0) sim[r] = fminf(fmaxf(sigmoidf_(p), 1e-7f), 1.0f);
}
__global__ void rowdot_sigmoid_bwd_kernel(const float* __restrict__ d_sim, const float* __restrict__ sim,
                                          const float* __restrict__ a, const float* __restrict__ x,
                                          float* __restrict__ da, float* __restrict__ dx, long R, float scale) {
    const long r = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const float p = sim[r];
    const float g = p > 1e-7f ? d_sim[r] * p * (1.0f - p) * scale : 0.f;
#pragma unroll
    for (int i = 0; i < AT_EL; ++i) {
        const long o = r * AT_E + lane + 32 * i;
        da[o] = g * x[o];
        dx[o] = g * a[o];
    }
}

// ---------------------------------------------------------------- sigmoid(Linear(LayerNorm(audio + dropout(attn))))
__global__ void __launch_bounds__(256)
ln_linear_sigmoid_fwd_kernel(const float* __restrict__ audio, const float* __restrict__ attn_out,
                             const float* __restrict__ gamma, const float* __restrict__ beta,
                             const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ prob,
                             float* __restrict__ stat, long R, float eps, uint32_t thresh, float keep, uint64_t seed,
                             const uint64_t* __restrict__ seed_dev) {
    const long r = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    if (seed_dev != nullptr) seed += *seed_dev;
    float x[AT_EL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < AT_EL; ++i) {
        const long o = r * AT_E + lane + 32 * i;
        float av = attn_out[o];
        if (thresh != 0u) av *= tag_dropout_scale(seed, (uint64_t)o, thresh, keep);
        x[i] = audio[o] + av;
        s += x[i];
    }
    const float mean = warp_sum(s) * (1.0f / AT_E);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < AT_EL; ++i) { const float d = x[i] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / AT_E) + eps);
    float y = 0.f;
#pragma unroll
    for (int i = 0; i < AT_EL; ++i) {
        const int e = lane + 32 * i;
        y = fmaf(fmaf((x[i] - mean) * rstd, gamma[e], beta[e]), w[e], y);
    }
    y = warp_sum(y) + bias[0];
    if (lane == 0) { prob[r] = sigmoidf_(y); stat[2 * r] = mean; stat[2 * r + 1] = rstd; }
}

// d_gamma, d_beta, d_w [E], d_bias [1] accumulated (pre-zeroed); d_audio, d_attn [R,E] overwritten
__global__ void __launch_bounds__(256)
ln_linear_sigmoid_bwd_kernel(const float* __restrict__ d_prob, const float* __restrict__ prob,
                             const float* __restrict__ audio, const float* __restrict__ attn_out,
                             const float* __restrict__ gamma, const float* __restrict__ beta,
                             const float* __restrict__ w, const float* __restrict__ stat, float* __restrict__ d_audio,
                             float* __restrict__ d_attn, float* __restrict__ d_gamma, float* __restrict__ d_beta,
                             float* __restrict__ d_w, float* __restrict__ d_bias, long R, uint32_t thresh, float keep,
                             uint64_t seed, const uint64_t* __restrict__ seed_dev) {
    const int lane = threadIdx.x & 31;
    if (seed_dev != nullptr) seed += *seed_dev;
    float gg[AT_EL], gb[AT_EL], gw[AT_EL], gbias = 0.f;
#pragma unroll
    for (int i = 0; i < AT_EL; ++i) { gg[i] = 0.f; gb[i] = 0.f; gw[i] = 0.f; }
    const long warps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long r = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5; r < R; r += warps) {
        const float p = prob[r];
        const float dy = d_prob[r] * p * (1.0f - p);
        const float mean = stat[2 * r], rstd = stat[2 * r + 1];
        float xh[AT_EL], dxh[AT_EL], ksc[AT_EL];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < AT_EL; ++i) {
            const int e = lane + 32 * i;
            const long o = r * AT_E + e;
            ksc[i] = thresh != 0u ? tag_dropout_scale(seed, (uint64_t)o, thresh, keep) : 1.f;
            xh[i] = (audio[o] + attn_out[o] * ksc[i] - mean) * rstd;
            const float dln = dy * w[e];
            gw[i] = fmaf(dy, fmaf(xh[i], gamma[e], beta[e]), gw[i]);
            gg[i] = fmaf(dln, xh[i], gg[i]);
            gb[i] += dln;
            dxh[i] = dln * gamma[e];
            s1 += dxh[i];
            s2 = fmaf(dxh[i], xh[i], s2);
        }
        gbias += dy;
        s1 = warp_sum(s1) * (1.0f / AT_E);
        s2 = warp_sum(s2) * (1.0f / AT_E);
#pragma unroll
        for (int i = 0; i < AT_EL; ++i) {
            const long o = r * AT_E + lane + 32 * i;
            const float dx = rstd * (dxh[i] - s1 - xh[i] * s2);
            d_audio[o] = dx;
            d_attn[o] = dx * ksc[i];
        }
    }
#pragma unroll
    for (int i = 0; i < AT_EL; ++i) {
        atomicAdd(d_gamma + lane + 32 * i, gg[i]);
        atomicAdd(d_beta + lane + 32 * i, gb[i]);
        atomicAdd(d_w + lane + 32 * i, gw[i]);
    }
    if (lane == 0) atomicAdd(d_bias, gbias);
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
    }
    return TAG_OK;
}

}  // namespace

extern "C" int tag_text_assemble_fwd(const long long* text, const float* emb, const float* cls, const float* pe,
                                     float* out, int B, int N, int E, int vocab, float dropout_p, uint64_t seed,
                                     const uint64_t* seed_dev, cudaStream_t stream) {
    if (B <= 0 || N <= 0 || E <= 0) return TAG_ERR_BAD_ARG;
    uint32_t thresh; float keep;
    tag_dropout_params(dropout_p, &thresh, &keep);
    text_assemble_fwd_kernel<<<dim3(N + 1, B), 128, 0, stream>>>(text, emb, cls, pe, out, N, E, vocab, thresh, keep, seed, seed_dev);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_text_assemble_bwd(const long long* text, const float* d_out, float* d_emb, float* d_cls, int B,
                                     int N, int E, int vocab, float dropout_p, uint64_t seed,
                                     const uint64_t* seed_dev, cudaStream_t stream) {
    if (B <= 0 || N <= 0 || E <= 0) return TAG_ERR_BAD_ARG;
    uint32_t thresh; float keep;
    tag_dropout_params(dropout_p, &thresh, &keep);
    text_assemble_bwd_kernel<<<dim3(N + 1, B), 128, 0, stream>>>(text, d_out, d_emb, d_cls, N, E, vocab, thresh, keep, seed, seed_dev);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_mha_core_fwd(const float* q, const float* k, const float* v, const long long* key_len, float* out,
                                float* probs, int B, int Lq, int Lk, int E, int heads, float dropout_p, uint64_t seed,
                                const uint64_t* seed_dev, cudaStream_t stream) {
    if (B <= 0 || Lq <= 0 || Lk <= 0 || heads <= 0 || E % heads != 0) return TAG_ERR_BAD_ARG;
    const int dh = E / heads;
    if (Lk > AT_MAXK || (dh != 32 && dh != 64 && dh != 128)) return TAG_ERR_UNSUPPORTED;
    uint32_t thresh; float keep;
    tag_dropout_params(dropout_p, &thresh, &keep);
    const float scale = 1.0f / sqrtf((float)dh);
    const size_t smem = (size_t)2 * Lk * dh * sizeof(float);
#define TAG_MHA_FWD(DHL_)                                                                                      \
    do {                                                                                                       \
        int rc = set_smem(mha_core_fwd_kernel<DHL_>, smem);                                                    \
        if (rc != TAG_OK) return rc;                                                                           \
        mha_core_fwd_kernel<DHL_><<<dim3(heads, B), 128, smem, stream>>>(q, k, v, key_len, out, probs, Lq, Lk, E, \
                                                                          scale, thresh, keep, seed, seed_dev); \
    } while (0)
    if (dh == 32) TAG_MHA_FWD(1); else if (dh == 64) TAG_MHA_FWD(2); else TAG_MHA_FWD(4);
#undef TAG_MHA_FWD
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_mha_core_bwd(const float* d_out, const float* q, const float* k, const float* v, const float* probs,
                                const long long* key_len, float* dq, float* dk, float* dv, int B, int Lq, int Lk,
                                int E, int heads, float dropout_p, uint64_t seed, const uint64_t* seed_dev,
                                cudaStream_t stream) {
    if (B <= 0 || Lq <= 0 || Lk <= 0 || heads <= 0 || E % heads != 0) return TAG_ERR_BAD_ARG;
    const int dh = E / heads;
    if (Lk > AT_MAXK || (dh != 32 && dh != 64 && dh != 128)) return TAG_ERR_UNSUPPORTED;
    uint32_t thresh; float keep;
    tag_dropout_params(dropout_p, &thresh, &keep);
    const float scale = 1.0f / sqrtf((float)dh);
    const size_t smem = (size_t)4 * Lk * dh * sizeof(float);
    if (smem > 200 * 1024) return TAG_ERR_UNSUPPORTED;
#define TAG_MHA_BWD(DHL_)                                                                                      \
    do {                                                                                                       \
        int rc = set_smem(mha_core_bwd_kernel<DHL_>, smem);                                                    \
        if (rc != TAG_OK) return rc;                                                                           \
        mha_core_bwd_kernel<DHL_><<<dim3(heads, B), 128, smem, stream>>>(d_out, q, k, v, probs, key_len, dq, dk, dv, \
                                                                          Lq, Lk, E, scale, thresh, keep, seed, seed_dev); \
    } while (0)
    if (dh == 32) TAG_MHA_BWD(1); else if (dh == 64) TAG_MHA_BWD(2); else TAG_MHA_BWD(4);
#undef TAG_MHA_BWD
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_additive_attn_fwd(const float* hq, const float* hk, const float* v, const float* kv,
                                     const long long* q_len, const long long* kv_len, float* attn, float* out, int B,
                                     int T, int N, int E, cudaStream_t stream) {
    if (B <= 0 || T <= 0 || N <= 0) return TAG_ERR_BAD_ARG;
    if (E != AT_E || N > 32) return TAG_ERR_UNSUPPORTED;
    const size_t smem = (size_t)2 * N * AT_E * sizeof(float);
    int rc = set_smem(additive_attn_fwd_kernel, smem);
    if (rc != TAG_OK) return rc;
    additive_attn_fwd_kernel<<<dim3((T + AA_TT - 1) / AA_TT, B), 256, smem, stream>>>(hq, hk, v, kv, q_len, kv_len, attn, out, T, N);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_additive_attn_bwd(const float* d_out, const float* hq, const float* hk, const float* v,
                                     const float* kv, const float* attn, const long long* q_len,
                                     const long long* kv_len, float* d_hq, float* d_hk, float* d_v, float* d_kv, int B,
                                     int T, int N, int E, cudaStream_t stream) {
    if (B <= 0 || T <= 0 || N <= 0) return TAG_ERR_BAD_ARG;
    if (E != AT_E || N > 24) return TAG_ERR_UNSUPPORTED;       // 4 x N x 2 KB of shared memory
    const size_t smem = (size_t)4 * N * AT_E * sizeof(float);
    int rc = set_smem(additive_attn_bwd_kernel, smem);
    if (rc != TAG_OK) return rc;
    additive_attn_bwd_kernel<<<dim3((T + AA_TT - 1) / AA_TT, B), 256, smem, stream>>>(d_out, hq, hk, v, kv, attn, q_len, kv_len,
                                                                                  d_hq, d_hk, d_v, d_kv, T, N);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_sigmoid_gate_fwd(const float* x, const float* z, float* out, long n, cudaStream_t stream) {
    if (n <= 0) return TAG_ERR_BAD_ARG;
    long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    sigmoid_gate_fwd_kernel<<<(int)blocks, 256, 0, stream>>>(x, z, out, n);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_sigmoid_gate_bwd(const float* d_out, const float* x, const float* z, float* dx, float* dz, long n,
                                    cudaStream_t stream) {
    if (n <= 0) return TAG_ERR_BAD_ARG;
    long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    sigmoid_gate_bwd_kernel<<<(int)blocks, 256, 0, stream>>>(d_out, x, z, dx, dz, n);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_rowdot_sigmoid_fwd(const float* a, const float* x, float* sim, long R, int E, float scale,
                                      cudaStream_t stream) {
    if (R <= 0) return TAG_ERR_BAD_ARG;
    if (E != AT_E) return TAG_ERR_UNSUPPORTED;
    rowdot_sigmoid_fwd_kernel<<<(int)((R * 32 + 255) / 256), 256, 0, stream>>>(a, x, sim, R, scale);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_rowdot_sigmoid_bwd(const float* d_sim, const float* sim, const float* a, const float* x, float* da,
                                      float* dx, long R, int E, float scale, cudaStream_t stream) {
    if (R <= 0) return TAG_ERR_BAD_ARG;
    if (E != AT_E) return TAG_ERR_UNSUPPORTED;
    rowdot_sigmoid_bwd_kernel<<<(int)((R * 32 + 255) / 256), 256, 0, stream>>>(d_sim, sim, a, x, da, dx, R, scale);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_ln_linear_sigmoid_fwd(const float* audio, const float* attn_out, const float* gamma,
                                         const float* beta, const float* w, const float* bias, float* prob,
                                         float* stat, long R, int E, float eps, float dropout_p, uint64_t seed,
                                         const uint64_t* seed_dev, cudaStream_t stream) {
    if (R <= 0) return TAG_ERR_BAD_ARG;
    if (E != AT_E) return TAG_ERR_UNSUPPORTED;
    uint32_t thresh; float keep;
    tag_dropout_params(dropout_p, &thresh, &keep);
    ln_linear_sigmoid_fwd_kernel<<<(int)((R * 32 + 255) / 256), 256, 0, stream>>>(audio, attn_out, gamma, beta, w, bias, prob,
                                                                             stat, R, eps, thresh, keep, seed, seed_dev);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_ln_linear_sigmoid_bwd(const float* d_prob, const float* prob, const float* audio,
                                         const float* attn_out, const float* gamma, const float* beta, const float* w,
                                         const float* stat, float* d_audio, float* d_attn, float* d_gamma,
                                         float* d_beta, float* d_w, float* d_bias, long R, int E, float dropout_p,
                                         uint64_t seed, const uint64_t* seed_dev, cudaStream_t stream) {
    if (R <= 0) return TAG_ERR_BAD_ARG;
    if (E != AT_E) return TAG_ERR_UNSUPPORTED;
    uint32_t thresh; float keep;
    tag_dropout_params(dropout_p, &thresh, &keep);
    long blocks = (R * 32 + 255) / 256;
    if (blocks > 148 * 2) blocks = 148 * 2;
    ln_linear_sigmoid_bwd_kernel<<<(int)blocks, 256, 0, stream>>>(d_prob, prob, audio, attn_out, gamma, beta, w, stat, d_audio,
                                                               d_attn, d_gamma, d_beta, d_w, d_bias, R, thresh, keep, seed,
                                                               seed_dev);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}
