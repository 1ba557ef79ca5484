// Backward through dropout -> avg+max pool -> ReLU -> BatchNorm(train) in two HBM-bound passes.
//
//   mode 0 (reduce): red[c] += sum g,  red[C+c] += sum g * xhat          (dbeta, dgamma)
//   mode 1 (apply) : dy = scale * (g - dbeta/N - xhat * dgamma/N)
// with g = d(out) routed through dropout, the avg+max pooling window and the ReLU gate, recomputed
// from the raw conv output y and the BN (scale, shift) — neither relu(bn(y)) nor the dropout
// mask nor the max-pool indices are ever stored.  ph = pw = 0 selects "no pooling" (bn1 of a
// block: the incoming gradient already has the conv resolution).
//
// Autograd equivalent in the reference: models/panns.py:49-58 + audio_encoder.py:203-211.
//
// Register diet (the first version ran at 12 % occupancy with 165 registers): 16-byte raw vector
// loads stay packed (4 registers per window element) while all loads of a window are in flight;
// channels are unpacked 4 at a time; per-channel constants live in shared memory and the apply
// pass is folded to dy = scale*g + y*k1 + k0.
#include "common.cuh"

namespace {

template <typename T, int PH, int PW, bool POOL, int MODE>
__global__ void __launch_bounds__(256, 2)
bn_relu_pool_bwd_kernel(const T* __restrict__ y, const T* __restrict__ dout, T* __restrict__ dy,
                        const float* __restrict__ scale, const float* __restrict__ shift,
                        const float* __restrict__ mean, const float* __restrict__ invstd,
                        double* __restrict__ red, float inv_count, int bn_training,
                        int B, int H, int W, int C, uint64_t seed, const uint64_t* __restrict__ seed_dev,
                        uint32_t thresh, float keep_scale) {
    constexpr int VEC = Raw16<T>::VEC;
    constexpr int NSUB = Raw16<T>::NSUB;
    constexpr int NE = PH * PW;
    extern __shared__ float s_par[];            // [4][C]: scale, shift, pA, pB
    float* s_sc = s_par;
    float* s_sh = s_par + C;
    float* s_pa = s_par + 2 * C;                // mode 0: xs = invstd        mode 1: k1
    float* s_pb = s_par + 3 * C;                // mode 0: xo = -mean*invstd  mode 1: k0
    if (seed_dev != nullptr) seed += *seed_dev;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float sc = scale[c], mu = mean[c], is = invstd[c];
        s_sc[c] = sc;
        s_sh[c] = shift[c];
        if (MODE == 0 && POOL) {
            // pooled reduce pass: the ReLU gate is the forward's own test fma(y, scale, shift) > 0; sums run on
            // w = sign(gamma) * invstd * y (u = sign(gamma) * xhat = w - off is monotone in the activation, so the window
            // maximum of the activation sits at the maximum of w) and are moved to the u domain once per window:
            // 7 instructions per element instead of 12 (this pass is issue-bound, not HBM-bound)
            const float sg = sc < 0.f ? -1.f : 1.f;
            s_pa[c] = sg * is;
            s_pb[c] = sg * mu * is;
        } else if (MODE == 0) {
            s_pa[c] = is;
            s_pb[c] = -mu * is;
        } else {
            const float dbe = bn_training ? (float)red[c] * inv_count : 0.f;
            const float dga = bn_training ? (float)red[C + c] * inv_count : 0.f;
            s_pa[c] = -sc * dga * is;
            s_pb[c] = -sc * dbe + sc * dga * mu * is;
        }
    }
    __syncthreads();

    const int Ho = POOL ? H / PH : H, Wo = POOL ? W / PW : W;
    const int Hs = (H + PH - 1) / PH, Ws = (W + PW - 1) / PW;      // window slots incl. partial ones
    const int CV = C / VEC;
    const int cv = threadIdx.x % CV;
    const int slot_lane = threadIdx.x / CV;
    const int slots_per_block = blockDim.x / CV;
    const unsigned n_slots = (unsigned)B * Hs * Ws;          // < 2^31 (host-checked): 32-bit index math
    const int c_base = cv * VEC;

    float rs[VEC], rq[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) { rs[k] = 0.f; rq[k] = 0.f; }

    // Slot coordinates and raw loads are one slot AHEAD of the arithmetic (software prefetch): with ~500 instructions
    // of math per slot and only 16 warps per SM the loads of the next window must already be in flight.
    // Element offsets are 32-bit (host-checked: the tensors have fewer than 2^31 elements).
    struct Slot { int b, hs, ws; bool full; unsigned oidx; };
    auto coord = [&](unsigned s) {
        Slot c;
        const unsigned r = s / (unsigned)Ws;
        c.ws = (int)(s - r * Ws);
        c.b = (int)(r / (unsigned)Hs);
        c.hs = (int)(r - (unsigned)c.b * Hs);
        c.full = POOL ? (c.hs < Ho && c.ws < Wo) : true;
        c.oidx = ((((unsigned)c.b * Ho + c.hs) * Wo + c.ws) * CV + cv) * VEC;
        return c;
    };
    auto fetch = [&](const Slot& c, uint4& gq, uint4 (&yq)[NE]) {
        const bool want = !(POOL && MODE == 0 && !c.full);          // the pooled reduce pass skips edge windows
        gq = (c.full && want) ? ld16(dout + c.oidx) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            const int h = c.hs * PH + e / PW, w = c.ws * PW + e % PW;
            yq[e] = (want && h < H && w < W) ? ld16(y + ((((unsigned)c.b * H + h) * W + w) * (unsigned)C + c_base))
                                             : make_uint4(0u, 0u, 0u, 0u);
        }
    };
    const unsigned s_stride = gridDim.x * slots_per_block;
    unsigned s = blockIdx.x * slots_per_block + slot_lane;
    Slot cur = coord(s < n_slots ? s : 0u);
    uint4 raw_g, raw_y[NE];
    if (s < n_slots) fetch(cur, raw_g, raw_y);
    Slot nxt = cur;
    uint4 nxt_g = make_uint4(0u, 0u, 0u, 0u), nxt_y[NE];
    for (; s < n_slots; s += s_stride, cur = nxt, raw_g = nxt_g) {
        if (s + s_stride < n_slots) { nxt = coord(s + s_stride); fetch(nxt, nxt_g, nxt_y); }
        const int b = cur.b, hs = cur.hs, ws_ = cur.ws;
        const bool full = cur.full;
        const unsigned oidx = cur.oidx;
        bool inb[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) inb[e] = (hs * PH + e / PW < H) && (ws_ * PW + e % PW < W);
        if (!(POOL && MODE == 0 && !full)) {
        uint4 raw_o[NE];
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub) {
            const int c0 = c_base + sub * 4;
            const float4 sc4 = *reinterpret_cast<const float4*>(s_sc + c0);
            const float4 sh4 = *reinterpret_cast<const float4*>(s_sh + c0);
            const float4 pa4 = *reinterpret_cast<const float4*>(s_pa + c0);
            const float4 pb4 = *reinterpret_cast<const float4*>(s_pb + c0);
            const float sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w}, sh[4] = {sh4.x, sh4.y, sh4.z, sh4.w};
            const float pa[4] = {pa4.x, pa4.y, pa4.z, pa4.w}, pb[4] = {pb4.x, pb4.y, pb4.z, pb4.w};
            float go[4];
            unpack4<T>(raw_g, sub, go);
            if (POOL && thresh != 0u && full) {
                float ds[4];
                tag_dropout_scale4(seed, (uint64_t)(oidx + sub * 4), thresh, keep_scale, ds);
#pragma unroll
                for (int k = 0; k < 4; ++k) go[k] *= ds[k];
            }
            if (POOL && MODE == 0) {
                // ---- reduce pass, pooled: per window and channel
                //   sum_e g_e        = G * (cnt/NE + [cnt > 0])
                //   sum_e g_e * u_e  = G * (S/NE   + [cnt > 0] * max_e u_e),   S = sum of u_e over a_e > 0
                // (the max-pool gradient lands on a maximal element; tied elements have equal u).  Windows that
                // hang over the edge carry no gradient (floor-mode pooling) and were skipped above.
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float Sw = 0.f, cnt = 0.f, m = -INFINITY;
#pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        float ve[4];
                        unpack4<T>(raw_y[e], sub, ve);
                        const float w = ve[k] * pa[k];
                        if (fmaf(ve[k], sc[k], sh[k]) > 0.f) { Sw += w; cnt += 1.f; }
                        m = fmaxf(m, w);
                    }
                    const bool any = cnt > 0.f;
                    const float S = fmaf(-cnt, pb[k], Sw);                 // sum of u over the active elements
                    rs[sub * 4 + k] = fmaf(go[k], fmaf(cnt, 1.0f / NE, any ? 1.f : 0.f), rs[sub * 4 + k]);
                    rq[sub * 4 + k] = fmaf(go[k], fmaf(S, 1.0f / NE, any ? m - pb[k] : 0.f), rq[sub * 4 + k]);
                }
            } else if (POOL) {
                // ---- apply pass, pooled: first maximum in scan order takes the max-pool gradient (torch semantics)
                float o[NE][4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float ve[NE], ae[NE];
                    float m = -INFINITY;
                    int idx = 0;
#pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        float t4[4];
                        unpack4<T>(raw_y[e], sub, t4);
                        ve[e] = t4[k];
                        ae[e] = fmaf(ve[e], sc[k], sh[k]);
                        const bool better = inb[e] && ae[e] > m;
                        m = better ? ae[e] : m;
                        idx = better ? e : idx;
                    }
                    const float g1 = full ? go[k] * (1.0f / NE) : 0.f;
                    const float g2 = g1 + (full ? go[k] : 0.f);
#pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        const float g = ae[e] > 0.f ? (idx == e ? g2 : g1) : 0.f;
                        o[e][k] = fmaf(sc[k], g, fmaf(ve[e], pa[k], pb[k]));
                    }
                }
#pragma unroll
                for (int e = 0; e < NE; ++e) pack4<T>(raw_o[e], sub, o[e]);
            } else {
                // ---- no pooling (NE == 1)
                float v0[4], o[4];
                unpack4<T>(raw_y[0], sub, v0);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float g = fmaf(v0[k], sc[k], sh[k]) > 0.f ? go[k] : 0.f;
                    if (MODE == 0) {
                        rs[sub * 4 + k] += g;
                        rq[sub * 4 + k] = fmaf(g, fmaf(v0[k], pa[k], pb[k]), rq[sub * 4 + k]);
                    } else {
                        o[k] = fmaf(sc[k], g, fmaf(v0[k], pa[k], pb[k]));
                    }
                }
                if (MODE == 1) pack4<T>(raw_o[0], sub, o);
            }
        }
        if (MODE == 1) {
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                if (inb[e]) {
                    const int h = hs * PH + e / PW, w = ws_ * PW + e % PW;
                    st16(dy + ((((unsigned)b * H + h) * W + w) * (unsigned)C + c_base), raw_o[e]);
                }
            }
        }
        }
#pragma unroll
        for (int e = 0; e < NE; ++e) raw_y[e] = nxt_y[e];
    }
    if (MODE == 0) {
        __syncthreads();
        // reuse the parameter smem for the block reduction: [2][256][VEC] floats
        __shared__ float sm[2][256][VEC + 1];
#pragma unroll
        for (int k = 0; k < VEC; ++k) { sm[0][threadIdx.x][k] = rs[k]; sm[1][threadIdx.x][k] = rq[k]; }
        __syncthreads();
        for (int i = threadIdx.x; i < CV * VEC; i += blockDim.x) {
            const int cvi = i / VEC, k = i % VEC;
            double ds = 0.0, dq = 0.0;
            for (int l = 0; l < slots_per_block; ++l) {
                ds += sm[0][l * CV + cvi][k];
                dq += sm[1][l * CV + cvi][k];
            }
            if (POOL && scale[i] < 0.f) dq = -dq;           // the pooled pass accumulated sum g * sign(gamma) * xhat
            atomicAdd(red + i, ds);
            atomicAdd(red + C + i, dq);
        }
    }
}

template <typename T, int MODE>
int pool_bwd_dispatch(const void* y, const void* dout, void* dy, const float* scale, const float* shift,
                      const float* mean, const float* invstd, double* red, float inv_count, int bn_training,
                      int B, int H, int W, int C, int ph, int pw, uint64_t seed, const uint64_t* seed_dev,
                      uint32_t thresh, float ks, cudaStream_t stream) {
    constexpr int VEC = Raw16<T>::VEC;
    if (C % VEC != 0) return TAG_ERR_BAD_ARG;
    const int CV = C / VEC;
    if (CV > 256 || 256 % CV != 0) return TAG_ERR_BAD_ARG;
    const int spb = 256 / CV;
    const int ph_e = ph > 0 ? ph : 1, pw_e = pw > 0 ? pw : 1;
    const long n_slots = (long)B * ((H + ph_e - 1) / ph_e) * ((W + pw_e - 1) / pw_e);
    if (n_slots >= (1L << 31) || (long)B * H * W * C >= (1L << 31)) return TAG_ERR_BAD_ARG;
    long blocks = (n_slots + spb - 1) / spb;
    // two CTAs are resident per SM; the reduce pass ends in 2C double atomics per CTA on a handful of cache lines
    // (they serialise in L2), so it runs exactly one resident wave
    if (blocks > 148 * (MODE == 0 ? 2 : 12)) blocks = 148 * (MODE == 0 ? 2 : 12);
    if (blocks < 1) blocks = 1;
    const size_t smem = (size_t)4 * C * sizeof(float);
#define TAG_LAUNCH_POOL_BWD(PH_, PW_, POOL_)                                                              \
    bn_relu_pool_bwd_kernel<T, PH_, PW_, POOL_, MODE><<<(int)blocks, 256, smem, stream>>>(                \
        (const T*)y, (const T*)dout, (T*)dy, scale, shift, mean, invstd, red, inv_count, bn_training, B, \
        H, W, C, seed, seed_dev, thresh, ks)
    if (ph == 2 && pw == 2) TAG_LAUNCH_POOL_BWD(2, 2, true);
    else if (ph == 1 && pw == 2) TAG_LAUNCH_POOL_BWD(1, 2, true);
    else if (ph == 0 && pw == 0) TAG_LAUNCH_POOL_BWD(1, 1, false);
    else return TAG_ERR_UNSUPPORTED;
#undef TAG_LAUNCH_POOL_BWD
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

}  // namespace

// mode 0: accumulate red[0:C] += sum g, red[C:2C] += sum g*xhat.  mode 1: write dy.
// ph = pw = 0 selects the "no pooling" variant (dout has the conv resolution).
extern "C" int tag_bn_relu_pool_bwd(int mode, const void* y, const void* dout, void* dy, int dtype,
                                    const float* scale, const float* shift, const float* mean,
                                    const float* invstd, double* red, int bn_training, int B, int H,
                                    int W, int C, int ph, int pw, float dropout_p, uint64_t seed,
                                    const uint64_t* seed_dev, cudaStream_t stream) {
    uint32_t thresh; float ks;
    tag_dropout_params(dropout_p, &thresh, &ks);
    const float inv_count = 1.0f / (float)((double)B * H * W);
#define TAG_POOL_BWD(T_, MODE_)                                                                            \
    pool_bwd_dispatch<T_, MODE_>(y, dout, dy, scale, shift, mean, invstd, red, inv_count, bn_training, B, \
                                 H, W, C, ph, pw, seed, seed_dev, thresh, ks, stream)
    if (dtype == TAG_DTYPE_F32) return mode == 0 ? TAG_POOL_BWD(float, 0) : TAG_POOL_BWD(float, 1);
    return mode == 0 ? TAG_POOL_BWD(bf16, 0) : TAG_POOL_BWD(bf16, 1);
#undef TAG_POOL_BWD
}
