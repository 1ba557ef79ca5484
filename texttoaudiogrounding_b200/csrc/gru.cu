// Bidirectional GRU recurrence (hidden 256), forward and BPTT, as persistent thread-block
// clusters: one cluster of 8 CTAs per (direction, slice of 8 sequences).  Each CTA keeps its
// 96 x 256 slice of W_hh resident in shared memory for all T steps and exchanges the new
// hidden state (forward) / gate gradients (backward) with its 7 peers through distributed
// shared memory, one cluster barrier per step.
//
// Replaces the cuDNN RNN recurrence behind nn.GRU(512, 256, bidirectional, batch_first)
// (reference models/audio_encoder.py:141,217).  Cell (gate order r,z,n):
//   r = s(gi_r + gh_r), z = s(gi_z + gh_z), n = tanh(gi_n + r * gh_n), h' = (1-z) n + z h
// with gi = x W_ih^T + b_ih (computed by tag_conv_fwd, taps=1) and gh = h W_hh^T + b_hh.
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {

constexpr int HID = 256;
constexpr int NCTA = 8;               // CTAs per cluster
constexpr int JS = HID / NCTA;        // 32 hidden units per CTA
constexpr int BS = 8;                 // sequences per cluster
constexpr int G3 = 3 * HID;           // 768

struct GruFwdSmem {
    float w[HID][3 * JS];             // w[k][g*32 + j] = W_hh[g*256 + 32*cta + j][k]      (96 KB)
    float h[2][HID][BS];              // double-buffered full hidden state, [k][b]         (16 KB)
    float part[8][3 * BS][JS];        // per-warp partial dot products                     (24 KB)
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// gi: [B, T, 2*768] fp32 (dir-major columns); out: [B, T, 512]; gates (optional): [B, T, 2, 4, 256]
__global__ void __cluster_dims__(NCTA, 1, 1) __launch_bounds__(256, 1)
gru_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ w_hh,
               const float* __restrict__ b_hh, float* __restrict__ out, float* __restrict__ gates,
               int B, int T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GruFwdSmem& s = *reinterpret_cast<GruFwdSmem*>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const int cta = (int)cluster.block_rank();
    const int slice = blockIdx.y;
    const int dir = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* W = w_hh + (long)dir * G3 * HID;
    const float* bh = b_hh + dir * G3;

    for (int i = tid; i < 3 * JS * HID; i += 256) {
        const int k = i % HID, row = i / HID;          // row = g*32 + j
        const int g = row / JS, j = row % JS;
        s.w[k][row] = W[(long)(g * HID + cta * JS + j) * HID + k];
    }
    for (int i = tid; i < 2 * HID * BS; i += 256) (&s.h[0][0][0])[i] = 0.f;
    cluster.sync();

    // finalisation role of this thread: hidden unit j_ = lane (global jg), sequence b_ = warp
    const int jg = cta * JS + lane;
    const int bl = warp;
    const int bglob = slice * BS + bl;
    const bool b_ok = bglob < B;
    const float bias_r = bh[jg], bias_z = bh[HID + jg], bias_n = bh[2 * HID + jg];

    float* peer_h[NCTA];
#pragma unroll
    for (int r = 0; r < NCTA; ++r) peer_h[r] = cluster.map_shared_rank(&s.h[0][0][0], r);

    int cur = 0;
    for (int step = 0; step < T; ++step) {
        const int t = dir == 0 ? step : T - 1 - step;
        // prefetch this step's input projections (independent of h)
        float gi_r = 0.f, gi_z = 0.f, gi_n = 0.f;
        if (b_ok) {
            const float* g = gi + ((long)bglob * T + t) * (2 * G3) + dir * G3;
            gi_r = __ldg(g + jg); gi_z = __ldg(g + HID + jg); gi_n = __ldg(g + 2 * HID + jg);
        }
        // ---- partial matvec: warp owns k in [32*warp, 32*warp+32), lane owns hidden unit j
        float acc[3][BS];
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int b = 0; b < BS; ++b) acc[g][b] = 0.f;
        const int kbase = warp * 32;
#pragma unroll 4
        for (int kk = 0; kk < 32; ++kk) {
            const int k = kbase + kk;
            const float w0 = s.w[k][lane], w1 = s.w[k][JS + lane], w2 = s.w[k][2 * JS + lane];
            const float4 h0 = *reinterpret_cast<const float4*>(&s.h[cur][k][0]);
            const float4 h1 = *reinterpret_cast<const float4*>(&s.h[cur][k][4]);
            const float hv[BS] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
            for (int b = 0; b < BS; ++b) {
                acc[0][b] = fmaf(w0, hv[b], acc[0][b]);
                acc[1][b] = fmaf(w1, hv[b], acc[1][b]);
                acc[2][b] = fmaf(w2, hv[b], acc[2][b]);
            }
        }
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int b = 0; b < BS; ++b) s.part[warp][g * BS + b][lane] = acc[g][b];
        __syncthreads();
        // ---- finalise (j = lane, b = warp)
        float gh_r = bias_r, gh_z = bias_z, gh_n = bias_n;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) {
            gh_r += s.part[wv][0 * BS + bl][lane];
            gh_z += s.part[wv][1 * BS + bl][lane];
            gh_n += s.part[wv][2 * BS + bl][lane];
        }
        const float hprev = s.h[cur][jg][bl];
        const float r = sigmoidf_(gi_r + gh_r);
        const float z = sigmoidf_(gi_z + gh_z);
        const float n = tanhf(gi_n + r * gh_n);
        const float hnew = (1.f - z) * n + z * hprev;
        if (b_ok) {
            out[((long)bglob * T + t) * (2 * HID) + dir * HID + jg] = hnew;
            if (gates != nullptr) {
                float* gp = gates + (((long)bglob * T + t) * 2 + dir) * 4 * HID;
                gp[jg] = r; gp[HID + jg] = z; gp[2 * HID + jg] = n; gp[3 * HID + jg] = gh_n;
            }
        }
        const int nxt = cur ^ 1;
        const int off = (nxt * HID + jg) * BS + bl;
#pragma unroll
        for (int rnk = 0; rnk < NCTA; ++rnk) peer_h[rnk][off] = hnew;
        cluster.sync();
        cur = nxt;
    }
}

struct GruBwdSmem {
    float w[G3][JS];                  // w[row][j] = W_hh[row][32*cta + j]                 (96 KB)
    float dgh[2][G3][BS];             // double-buffered gathered gate gradients [row][b]  (48 KB)
    float part[8][BS][JS];            //                                                    (8 KB)
};

// d_out: [B,T,512]; out: forward hidden states [B,T,512]; gates: [B,T,2,4,256];
// dgi: [B,T,1536] (gradient wrt the input projections); dgh: [2,B*T,768] (wrt h W_hh^T + b_hh);
// hprev_out: [2,B*T,256] the hidden state each step consumed (operand of the W_hh wgrad)
__global__ void __cluster_dims__(NCTA, 1, 1) __launch_bounds__(256, 1)
gru_bwd_kernel(const float* __restrict__ d_out, const float* __restrict__ out,
               const float* __restrict__ gates, const float* __restrict__ w_hh,
               float* __restrict__ dgi, float* __restrict__ dgh_out, float* __restrict__ hprev_out,
               int B, int T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GruBwdSmem& s = *reinterpret_cast<GruBwdSmem*>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const int cta = (int)cluster.block_rank();
    const int slice = blockIdx.y;
    const int dir = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* W = w_hh + (long)dir * G3 * HID;

    for (int i = tid; i < G3 * JS; i += 256) {
        const int j = i % JS, row = i / JS;
        s.w[row][j] = W[(long)row * HID + cta * JS + j];
    }
    for (int i = tid; i < 2 * G3 * BS; i += 256) (&s.dgh[0][0][0])[i] = 0.f;
    cluster.sync();

    const int jg = cta * JS + lane;
    const int bl = warp;
    const int bglob = slice * BS + bl;
    const bool b_ok = bglob < B;
    float* peer[NCTA];
#pragma unroll
    for (int r = 0; r < NCTA; ++r) peer[r] = cluster.map_shared_rank(&s.dgh[0][0][0], r);

    float dh = 0.f;          // gradient flowing into h_t from later steps (this thread's (j, b))
    int cur = 0;
    for (int step = 0; step < T; ++step) {
        // reverse of the forward order
        const int t = dir == 0 ? T - 1 - step : step;
        const int t_prev = dir == 0 ? t - 1 : t + 1;           // where h_{prev} was produced
        float g_r = 0.f, g_z = 0.f, g_n = 0.f;                 // dgh components
        float dh_direct = 0.f;
        if (b_ok) {
            const long bt = (long)bglob * T + t;
            const float dht = dh + d_out[bt * (2 * HID) + dir * HID + jg];
            const float* gp = gates + (bt * 2 + dir) * 4 * HID;
            const float r = gp[jg], z = gp[HID + jg], n = gp[2 * HID + jg], hn = gp[3 * HID + jg];
            const bool has_prev = (t_prev >= 0 && t_prev < T);
            const float hprev = has_prev ? out[((long)bglob * T + t_prev) * (2 * HID) + dir * HID + jg] : 0.f;
            const float dn = dht * (1.f - z);
            const float dz = dht * (hprev - n);
            const float dn_pre = dn * (1.f - n * n);
            const float dz_pre = dz * z * (1.f - z);
            const float dr = dn_pre * hn;
            const float dr_pre = dr * r * (1.f - r);
            float* gi_p = dgi + bt * (2 * G3) + dir * G3;
            gi_p[jg] = dr_pre; gi_p[HID + jg] = dz_pre; gi_p[2 * HID + jg] = dn_pre;
            g_r = dr_pre; g_z = dz_pre; g_n = dn_pre * r;
            float* gh_p = dgh_out + ((long)dir * B * T + bt) * G3;
            gh_p[jg] = g_r; gh_p[HID + jg] = g_z; gh_p[2 * HID + jg] = g_n;
            hprev_out[((long)dir * B * T + bt) * HID + jg] = hprev;
            dh_direct = dht * z;
        }
        // all-gather the gate gradients of this step into every CTA of the cluster
        {
            const int o_r = (cur * G3 + jg) * BS + bl;
            const int o_z = (cur * G3 + HID + jg) * BS + bl;
            const int o_n = (cur * G3 + 2 * HID + jg) * BS + bl;
#pragma unroll
            for (int rnk = 0; rnk < NCTA; ++rnk) {
                peer[rnk][o_r] = g_r; peer[rnk][o_z] = g_z; peer[rnk][o_n] = g_n;
            }
        }
        cluster.sync();
        // dh_prev[b][j] = sum_row dgh[row][b] * W_hh[row][j]; warp owns 96 rows
        float acc[BS];
#pragma unroll
        for (int b = 0; b < BS; ++b) acc[b] = 0.f;
        const int rbase = warp * (G3 / 8);
#pragma unroll 4
        for (int rr = 0; rr < G3 / 8; ++rr) {
            const int row = rbase + rr;
            const float wv = s.w[row][lane];
            const float4 d0 = *reinterpret_cast<const float4*>(&s.dgh[cur][row][0]);
            const float4 d1 = *reinterpret_cast<const float4*>(&s.dgh[cur][row][4]);
            acc[0] = fmaf(wv, d0.x, acc[0]); acc[1] = fmaf(wv, d0.y, acc[1]);
            acc[2] = fmaf(wv, d0.z, acc[2]); acc[3] = fmaf(wv, d0.w, acc[3]);
            acc[4] = fmaf(wv, d1.x, acc[4]); acc[5] = fmaf(wv, d1.y, acc[5]);
            acc[6] = fmaf(wv, d1.z, acc[6]); acc[7] = fmaf(wv, d1.w, acc[7]);
        }
#pragma unroll
        for (int b = 0; b < BS; ++b) s.part[warp][b][lane] = acc[b];
        __syncthreads();
        float sum = dh_direct;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) sum += s.part[wv][bl][lane];
        dh = sum;
        __syncthreads();      // part[] is rewritten next step
        cur ^= 1;
    }
}

}  // namespace

extern "C" int tag_gru_fwd(const float* gi, const float* w_hh, const float* b_hh, float* out,
                           float* gates, int B, int T, cudaStream_t stream) {
    if (B <= 0 || T <= 0) return TAG_ERR_BAD_ARG;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gru_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(GruFwdSmem));
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    dim3 grid(NCTA, (B + BS - 1) / BS, 2);
    gru_fwd_kernel<<<grid, 256, sizeof(GruFwdSmem), stream>>>(gi, w_hh, b_hh, out, gates, B, T);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_gru_bwd(const float* d_out, const float* out, const float* gates, const float* w_hh,
                           float* dgi, float* dgh, float* hprev, int B, int T, cudaStream_t stream) {
    if (B <= 0 || T <= 0) return TAG_ERR_BAD_ARG;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gru_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(GruBwdSmem));
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    dim3 grid(NCTA, (B + BS - 1) / BS, 2);
    gru_bwd_kernel<<<grid, 256, sizeof(GruBwdSmem), stream>>>(d_out, out, gates, w_hh, dgi, dgh, hprev, B, T);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}
