// BiGRU recurrence, bf16 tensor-core variant (compute dtype bf16): same cluster decomposition as
// gru.cu (8 CTAs x 32 hidden units, 8 sequences per cluster, DSMEM exchange, one cluster barrier
// per step) but the per-step product with W_hh runs on mma.sync.m16n8k16 with the CTA's W_hh
// slice held in REGISTERS as A fragments for all T steps (48 registers per thread), so a step
// costs 12 MMAs per warp instead of ~800 FFMAs, and the exchanged state is bf16 rows of 64 B.
// The hidden state itself, the gates and all gradients stay fp32.
//
// (tcgen05 needs M >= 64 and pays a TMEM round trip per step; for a 96 x 8 x 256 product on the
// critical path of a 250-step recurrence the register-resident mma.sync form has lower latency.)
//
// Replaces the cuDNN RNN behind nn.GRU (reference models/audio_encoder.py:141,217).
#include "common.cuh"
#include <cooperative_groups.h>
#include <utility>
namespace cg = cooperative_groups;

namespace {

constexpr int HID = 256;
constexpr int NCTA = 8;
constexpr int JS = HID / NCTA;        // 32
constexpr int BS = 8;
constexpr int G3 = 3 * HID;           // 768
constexpr int HPAD = HID + 8;         // bf16 row stride of the exchanged hidden state
constexpr int GPAD = G3 + 8;          // bf16 row stride of the exchanged gate gradients

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float fast_sigmoid(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 2.0f / (1.0f + __expf(-2.0f * x)) - 1.0f; }

struct FwdSmem {
    bf16 h[2][BS][HPAD];              // double-buffered hidden state, [b][k]
    float part[8][BS][100];           // per-warp partial gate pre-activations, [warp][b][row (96) + pad]
};

__global__ void __cluster_dims__(NCTA, 1, 1) __launch_bounds__(256, 1)
gru_fwd_tc_kernel(const float* __restrict__ gi, const float* __restrict__ w_hh, const float* __restrict__ b_hh,
                  float* __restrict__ out, float* __restrict__ gates, int B, int T) {
    __shared__ __align__(16) FwdSmem s;
    cg::cluster_group cluster = cg::this_cluster();
    const int cta = (int)cluster.block_rank();
    const int slice = blockIdx.y, dir = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* W = w_hh + (long)dir * G3 * HID;
    const float* bh = b_hh + dir * G3;

    // ---- A fragments: rows (g, j) of this CTA, k in [32 warp, 32 warp + 32)
    uint32_t afrag[6][2][4];
#pragma unroll
    for (int mt = 0; mt < 6; ++mt) {
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
            const int k0 = warp * 32 + kt * 16 + (lane & 3) * 2;
#pragma unroll
            for (int hr = 0; hr < 2; ++hr) {            // rows lane/4 and lane/4 + 8 of the M tile
                const int rl = mt * 16 + (lane >> 2) + hr * 8;      // local row = g*32 + j
                const int grow = (rl >> 5) * HID + cta * JS + (rl & 31);
                const float* wr = W + (long)grow * HID;
                afrag[mt][kt][hr] = pack_bf16(wr[k0], wr[k0 + 1]);
                afrag[mt][kt][hr + 2] = pack_bf16(wr[k0 + 8], wr[k0 + 9]);
            }
        }
    }
    for (int i = tid; i < 2 * BS * HPAD; i += 256) (&s.h[0][0][0])[i] = __float2bfloat16_rn(0.f);
    cluster.sync();

    // finalisation role: hidden unit jg, sequence bl
    const int jg = cta * JS + lane;
    const int bl = warp;
    const int bglob = slice * BS + bl;
    const bool b_ok = bglob < B;
    const float bias_r = bh[jg], bias_z = bh[HID + jg], bias_n = bh[2 * HID + jg];
    bf16* peer_h[NCTA];
#pragma unroll
    for (int r = 0; r < NCTA; ++r) peer_h[r] = cluster.map_shared_rank(&s.h[0][0][0], r);

    // input projections are fetched two steps ahead (their L2/HBM latency is off the critical path)
    auto fetch_gi = [&](int step, float (&g3)[3]) {
        g3[0] = g3[1] = g3[2] = 0.f;
        if (b_ok && step < T) {
            const int t = dir == 0 ? step : T - 1 - step;
            const float* g = gi + ((long)bglob * T + t) * (2 * G3) + dir * G3;
            g3[0] = __ldg(g + jg); g3[1] = __ldg(g + HID + jg); g3[2] = __ldg(g + 2 * HID + jg);
        }
    };
    float gi_cur[3], gi_nxt[3], gi_nn[3];
    fetch_gi(0, gi_cur);
    fetch_gi(1, gi_nxt);
    float hprev = 0.f;
    int cur = 0;
    for (int step = 0; step < T; ++step) {
        const int t = dir == 0 ? step : T - 1 - step;
        fetch_gi(step + 2, gi_nn);
        const float gi_r = gi_cur[0], gi_z = gi_cur[1], gi_n = gi_cur[2];
        // ---- partial product on the tensor cores
        uint32_t bfrag[2][2];
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
            const bf16* hp = &s.h[cur][lane >> 2][warp * 32 + kt * 16 + (lane & 3) * 2];
            bfrag[kt][0] = *reinterpret_cast<const uint32_t*>(hp);
            bfrag[kt][1] = *reinterpret_cast<const uint32_t*>(hp + 8);
        }
#pragma unroll
        for (int mt = 0; mt < 6; ++mt) {
            float d[4] = {0.f, 0.f, 0.f, 0.f};
            mma_bf16_16816(d, afrag[mt][0], bfrag[0][0], bfrag[0][1]);
            mma_bf16_16816(d, afrag[mt][1], bfrag[1][0], bfrag[1][1]);
            const int r0 = mt * 16 + (lane >> 2), c0 = (lane & 3) * 2;
            s.part[warp][c0][r0] = d[0];
            s.part[warp][c0 + 1][r0] = d[1];
            s.part[warp][c0][r0 + 8] = d[2];
            s.part[warp][c0 + 1][r0 + 8] = d[3];
        }
        __syncthreads();
        float gh_r = bias_r, gh_z = bias_z, gh_n = bias_n;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) {
            gh_r += s.part[wv][bl][lane];
            gh_z += s.part[wv][bl][JS + lane];
            gh_n += s.part[wv][bl][2 * JS + lane];
        }
        const float r = fast_sigmoid(gi_r + gh_r);
        const float z = fast_sigmoid(gi_z + gh_z);
        const float n = fast_tanh(gi_n + r * gh_n);
        const float hnew = (1.f - z) * n + z * hprev;
        hprev = hnew;
        const int nxt = cur ^ 1;
        const int off = (nxt * BS + bl) * HPAD + jg;
        const bf16 hb = __float2bfloat16_rn(hnew);
#pragma unroll
        for (int rnk = 0; rnk < NCTA; ++rnk) peer_h[rnk][off] = hb;
        // arrive first, then issue the HBM stores: the release of the NEXT arrive waits for them,
        // a whole step later, instead of this one
        auto token = cluster.barrier_arrive();
        if (b_ok) {
            out[((long)bglob * T + t) * (2 * HID) + dir * HID + jg] = hnew;
            if (gates != nullptr) {
                float* gp = gates + (((long)bglob * T + t) * 2 + dir) * 4 * HID;
                gp[jg] = r; gp[HID + jg] = z; gp[2 * HID + jg] = n; gp[3 * HID + jg] = gh_n;
            }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) { gi_cur[i] = gi_nxt[i]; gi_nxt[i] = gi_nn[i]; }
        cluster.barrier_wait(std::move(token));
        cur = nxt;
    }
}

struct BwdSmem {
    bf16 dgh[2][BS][GPAD];            // double-buffered gathered gate gradients, [b][row]
    float part[8][BS][36];            // [warp][b][k_local (32) + pad]
};

__global__ void __cluster_dims__(NCTA, 1, 1) __launch_bounds__(256, 1)
gru_bwd_tc_kernel(const float* __restrict__ d_out, const float* __restrict__ out, const float* __restrict__ gates,
                  const float* __restrict__ w_hh, bf16* __restrict__ dgi, bf16* __restrict__ dgh_out,
                  bf16* __restrict__ hprev_out, int B, int T) {
    __shared__ __align__(16) BwdSmem s;
    cg::cluster_group cluster = cg::this_cluster();
    const int cta = (int)cluster.block_rank();
    const int slice = blockIdx.y, dir = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* W = w_hh + (long)dir * G3 * HID;

    // ---- A fragments of W_hh^T: A[m = k_local][kk = row] = W_hh[row][32 cta + k_local];
    //      this warp owns rows [96 warp, 96 warp + 96) = 6 K tiles, both M tiles.
    uint32_t afrag[2][6][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int kt = 0; kt < 6; ++kt) {
            const int row0 = warp * 96 + kt * 16 + (lane & 3) * 2;
#pragma unroll
            for (int hr = 0; hr < 2; ++hr) {
                const int kcol = cta * JS + mt * 16 + (lane >> 2) + hr * 8;
                afrag[mt][kt][hr] = pack_bf16(W[(long)row0 * HID + kcol], W[(long)(row0 + 1) * HID + kcol]);
                afrag[mt][kt][hr + 2] = pack_bf16(W[(long)(row0 + 8) * HID + kcol], W[(long)(row0 + 9) * HID + kcol]);
            }
        }
    }
    for (int i = tid; i < 2 * BS * GPAD; i += 256) (&s.dgh[0][0][0])[i] = __float2bfloat16_rn(0.f);
    cluster.sync();

    const int jg = cta * JS + lane;
    const int bl = warp;
    const int bglob = slice * BS + bl;
    const bool b_ok = bglob < B;
    bf16* peer[NCTA];
#pragma unroll
    for (int r = 0; r < NCTA; ++r) peer[r] = cluster.map_shared_rank(&s.dgh[0][0][0], r);

    // software pipeline: the operands of step s+1 are fetched while step s runs its product
    float p_r = 0.f, p_z = 0.f, p_n = 0.f, p_hn = 0.f, p_do = 0.f, p_hp = 0.f;
    auto fetch = [&](int step) {
        if (!b_ok || step >= T) return;
        const int t = dir == 0 ? T - 1 - step : step;
        const int t_prev = dir == 0 ? t - 1 : t + 1;
        const long bt = (long)bglob * T + t;
        const float* gp = gates + (bt * 2 + dir) * 4 * HID;
        p_r = __ldg(gp + jg); p_z = __ldg(gp + HID + jg); p_n = __ldg(gp + 2 * HID + jg); p_hn = __ldg(gp + 3 * HID + jg);
        p_do = __ldg(d_out + bt * (2 * HID) + dir * HID + jg);
        p_hp = (t_prev >= 0 && t_prev < T) ? __ldg(out + ((long)bglob * T + t_prev) * (2 * HID) + dir * HID + jg) : 0.f;
    };
    fetch(0);

    float dh = 0.f;
    int cur = 0;
    for (int step = 0; step < T; ++step) {
        const int t = dir == 0 ? T - 1 - step : step;
        float g_r = 0.f, g_z = 0.f, g_n = 0.f, dh_direct = 0.f, s_dn = 0.f, s_hp = 0.f;
        if (b_ok) {
            const float r = p_r, z = p_z, n = p_n, hn = p_hn, hprev = p_hp;
            const float dht = dh + p_do;
            const float dn_pre = dht * (1.f - z) * (1.f - n * n);
            const float dz_pre = dht * (hprev - n) * z * (1.f - z);
            const float dr_pre = dn_pre * hn * r * (1.f - r);
            s_dn = dn_pre; s_hp = hprev;
            g_r = dr_pre; g_z = dz_pre; g_n = dn_pre * r;
            dh_direct = dht * z;
        }
        const bf16 v_r = __float2bfloat16_rn(g_r), v_z = __float2bfloat16_rn(g_z), v_n = __float2bfloat16_rn(g_n);
        {
            const int o = (cur * BS + bl) * GPAD + jg;
#pragma unroll
            for (int rnk = 0; rnk < NCTA; ++rnk) {
                peer[rnk][o] = v_r; peer[rnk][o + HID] = v_z; peer[rnk][o + 2 * HID] = v_n;
            }
        }
        auto token = cluster.barrier_arrive();
        if (b_ok) {          // HBM stores after the arrive (see the forward kernel)
            const long bt = (long)bglob * T + t;
            // bf16 outputs: they are operands of the tensor-core weight-gradient / dgrad GEMMs
            bf16* gi_p = dgi + bt * (2 * G3) + dir * G3;
            gi_p[jg] = v_r; gi_p[HID + jg] = v_z; gi_p[2 * HID + jg] = __float2bfloat16_rn(s_dn);
            bf16* gh_p = dgh_out + ((long)dir * B * T + bt) * G3;
            gh_p[jg] = v_r; gh_p[HID + jg] = v_z; gh_p[2 * HID + jg] = v_n;
            hprev_out[((long)dir * B * T + bt) * HID + jg] = __float2bfloat16_rn(s_hp);
        }
        fetch(step + 1);
        cluster.barrier_wait(std::move(token));
        // dh_prev[b][k_local] = sum_row dgh[b][row] * W_hh[row][32 cta + k_local]
        float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kt = 0; kt < 6; ++kt) {
            const bf16* gp = &s.dgh[cur][lane >> 2][warp * 96 + kt * 16 + (lane & 3) * 2];
            const uint32_t b0 = *reinterpret_cast<const uint32_t*>(gp);
            const uint32_t b1 = *reinterpret_cast<const uint32_t*>(gp + 8);
            mma_bf16_16816(d0, afrag[0][kt], b0, b1);
            mma_bf16_16816(d1, afrag[1][kt], b0, b1);
        }
        {
            const int r0 = lane >> 2, c0 = (lane & 3) * 2;
            s.part[warp][c0][r0] = d0[0];      s.part[warp][c0 + 1][r0] = d0[1];
            s.part[warp][c0][r0 + 8] = d0[2];  s.part[warp][c0 + 1][r0 + 8] = d0[3];
            s.part[warp][c0][16 + r0] = d1[0]; s.part[warp][c0 + 1][16 + r0] = d1[1];
            s.part[warp][c0][24 + r0] = d1[2]; s.part[warp][c0 + 1][24 + r0] = d1[3];
        }
        __syncthreads();
        float sum = dh_direct;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) sum += s.part[wv][bl][lane];
        dh = sum;
        cur ^= 1;
        // part[] is next written after the following cluster.sync(), which orders it after these reads
    }
}

}  // namespace

extern "C" int tag_gru_fwd_bf16(const float* gi, const float* w_hh, const float* b_hh, float* out,
                                float* gates, int B, int T, cudaStream_t stream) {
    if (B <= 0 || T <= 0) return TAG_ERR_BAD_ARG;
    dim3 grid(NCTA, (B + BS - 1) / BS, 2);
    gru_fwd_tc_kernel<<<grid, 256, 0, stream>>>(gi, w_hh, b_hh, out, gates, B, T);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_gru_bwd_bf16(const float* d_out, const float* out, const float* gates, const float* w_hh,
                                void* dgi, void* dgh, void* hprev, int B, int T, cudaStream_t stream) {
    if (B <= 0 || T <= 0) return TAG_ERR_BAD_ARG;
    dim3 grid(NCTA, (B + BS - 1) / BS, 2);
    gru_bwd_tc_kernel<<<grid, 256, 0, stream>>>(d_out, out, gates, w_hh, (bf16*)dgi, (bf16*)dgh, (bf16*)hprev, B, T);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}
