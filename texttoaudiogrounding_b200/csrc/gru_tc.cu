// BiGRU recurrence, bf16 tensor-core variant (compute dtype bf16): a cluster of 8 CTAs per (direction, 8 sequences),
// CTA c owning hidden units 32 c .. 32 c + 31.  The per-step product with W_hh runs on mma.sync.m16n8k16 with the CTA's
// W_hh slice held in REGISTERS as A fragments for all T steps (48 registers per thread); the hidden state, the gates and
// all gradients stay fp32; what crosses CTAs is bf16, as 16-byte `st.async` pieces that complete on the receiver's
// mbarrier (no cluster barrier, no fence in the loop).
//   forward : every CTA needs the whole new state: it ships its 32 units x 8 sequences to all 8 CTAs (256 pieces arrive
//             per CTA and step), each CTA multiplies its 96 gate rows with the full state (K split over the 8 warps).
//   backward: a CTA multiplies its OWN 96 gate-gradient rows with its rows of W_hh for all 256 units and ships the eight
//             8 x 32 partial products to their owners (256 pieces per CTA and step instead of 768 gate-gradient pieces).
// Measured (scripts/gru_timing.py, B = 64, T = 250): 1.50 -> 1.17 us / step forward, 2.0 -> 1.19 us / step backward against
// the cluster-barrier version.  The hand-off alone (scripts/micro/dsmem_handoff.cu) is ~0.35 us per step however the 4 KB
// are cut (16-byte st.async pieces or bulk copies); the rest of a step is the dependent chain inside the CTA.
//
// (tcgen05 needs M >= 64 and pays a TMEM round trip per step; for a 96 x 8 x 256 product on the critical path of a
// 250-step recurrence the register-resident mma.sync form has lower latency.)
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

// ---------------------------------------------------------------------------------------------
// Barrier-free hand-off.  An exchange through plain DSMEM stores needs one cluster barrier per step: arrive.release has
// to drain every earlier store of the thread (the HBM stores of the previous step included) and the barrier itself costs
// ~400 cycles.  Here every 16-byte piece of a state row travels as ONE `st.async` that carries its own completion: the
// store lands in the peer's shared memory and adds 16 to the transaction count of the peer's mbarrier for that step.
// A CTA waits on its LOCAL mbarrier until all 8 producers (itself included) have delivered the step's bytes — no
// cluster barrier, no fence, and the HBM stores are never ordered against anything.  Buffers and barriers alternate
// between two slots: a CTA can only be one step ahead of the slowest peer (it needs that peer's bytes to advance), so
// slot s & 1 is never overwritten while someone still reads it.
__device__ __forceinline__ void g_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void g_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        if (++spins > (1u << 22)) __trap();           // a lost hand-off must not hang the GPU
    }
}
__device__ __forceinline__ uint32_t g_mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async16(uint32_t cluster_addr, const uint4& v, uint32_t cluster_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(cluster_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(cluster_bar) : "memory");
}

// sigmoid / tanh on the hardware tanh (MUFU.TANH, 2^-11 relative error): the IEEE divisions of the formulas above were
// ~350 of the ~1900 cycles of a step, and the state crosses CTAs in bf16 (2^-9) anyway.
__device__ __forceinline__ float hw_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float hw_sigmoid(float x) { return fmaf(0.5f, hw_tanh(0.5f * x), 0.5f); }

#ifdef TAG_GRU_TIMING
__device__ long long g_gru_timing[64 * 8];
#define GRU_T(slot) do { if (tid == 32 && cta == 1 && slice == 0 && dir == 0 && step >= 16 && step < 80) \
        g_gru_timing[(step - 16) * 8 + (slot)] = clock64(); } while (0)
#define GRU_MARK(row) do { if (tid == 32 && cta == 1 && slice == 0 && dir == 0) g_gru_timing[(row) * 8 + 7] = clock64(); } while (0)
#else
#define GRU_T(slot) do { } while (0)
#define GRU_MARK(row) do { } while (0)
#endif

struct FwdSmemA {
    bf16 h[2][BS][HPAD];              // receive buffers: hidden state of all 256 units, [slot][b][k]
    float part[2][8][BS][100];        // per-warp partial gate pre-activations, alternating by step (see below)
    bf16 stage[BS][JS];               // this CTA's 32 new units per sequence (64 B rows), source of the st.async pieces
    unsigned long long bar[2];
};

__global__ void __cluster_dims__(NCTA, 1, 1) __launch_bounds__(256, 1)
gru_fwd_tc_async_kernel(const float* __restrict__ gi, const float* __restrict__ w_hh, const float* __restrict__ b_hh,
                        float* __restrict__ out, float* __restrict__ gates, int B, int T) {
    extern __shared__ __align__(16) uint8_t gru_fwd_smem[];       // 59 KB: above the static limit
    FwdSmemA& s = *reinterpret_cast<FwdSmemA*>(gru_fwd_smem);
    cg::cluster_group cluster = cg::this_cluster();
    const int cta = (int)cluster.block_rank();
    const int slice = blockIdx.y, dir = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* W = w_hh + (long)dir * G3 * HID;
    const float* bh = b_hh + dir * G3;
    constexpr uint32_t STEP_BYTES = NCTA * BS * JS * 2;          // 4096: every producer delivers 8 rows of 64 B

    uint32_t afrag[6][2][4];
#pragma unroll
    for (int mt = 0; mt < 6; ++mt) {
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
            const int k0 = warp * 32 + kt * 16 + (lane & 3) * 2;
#pragma unroll
            for (int hr = 0; hr < 2; ++hr) {
                const int rl = mt * 16 + (lane >> 2) + hr * 8;
                const int grow = (rl >> 5) * HID + cta * JS + (rl & 31);
                const float* wr = W + (long)grow * HID;
                afrag[mt][kt][hr] = pack_bf16(wr[k0], wr[k0 + 1]);
                afrag[mt][kt][hr + 2] = pack_bf16(wr[k0 + 8], wr[k0 + 9]);
            }
        }
    }
    for (int i = tid; i < 2 * BS * HPAD; i += 256) (&s.h[0][0][0])[i] = __float2bfloat16_rn(0.f);
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&s.bar[0]);
    if (tid == 0) {
        g_mbar_init(bar0, 1); g_mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        g_mbar_expect_tx(bar0, STEP_BYTES);                      // slot 0 receives h_2, slot 1 receives h_1
        g_mbar_expect_tx(bar0 + 8, STEP_BYTES);
    }
    cluster.sync();

    const int jg = cta * JS + lane;
    const int bl = warp;
    const int bglob = slice * BS + bl;
    const bool b_ok = bglob < B;
    const float bias_r = bh[jg], bias_z = bh[HID + jg], bias_n = bh[2 * HID + jg];
    // lane l delivers piece (l & 3) of this warp's 64-byte row to peer l >> 2
    const uint32_t peer = lane >> 2, piece = lane & 3;
    const uint32_t h_local = (uint32_t)__cvta_generic_to_shared(&s.h[0][0][0]);
    const uint32_t peer_h = g_mapa(h_local, peer) + (uint32_t)((bl * HPAD + cta * JS + piece * 8) * 2);
    const uint32_t peer_bar = g_mapa(bar0, peer);

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
    for (int step = 0; step < T; ++step) {
        const int t = dir == 0 ? step : T - 1 - step;
        const int cur = step & 1, nxt = cur ^ 1;
        GRU_T(0);
        fetch_gi(step + 2, gi_nn);
        if (step > 0) {
            // h_step: slot cur, its (step - 1) / 2-th use
            g_mbar_wait(bar0 + 8 * cur, (uint32_t)(((step - 1) >> 1) & 1));
            if (tid == 0) g_mbar_expect_tx(bar0 + 8 * cur, STEP_BYTES);      // re-arm for h_{step + 2}
        }
        GRU_T(1);
        const float gi_r = gi_cur[0], gi_z = gi_cur[1], gi_n = gi_cur[2];
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
            s.part[cur][warp][c0][r0] = d[0];
            s.part[cur][warp][c0 + 1][r0] = d[1];
            s.part[cur][warp][c0][r0 + 8] = d[2];
            s.part[cur][warp][c0 + 1][r0 + 8] = d[3];
        }
        GRU_T(2);
        // The partials alternate between two buffers: a warp that is already through the next step's mbarrier wait writes
        // the OTHER buffer while a slower warp still sums this one, and the buffer is re-used two steps later, behind the
        // next step's __syncthreads.  (With one buffer the ordering held only through the mbarrier: the wait needs every
        // warp's st.async, issued after its reads — correct, but invisible to compute-sanitizer racecheck.)
        __syncthreads();
        GRU_T(3);
        float gh_r = bias_r, gh_z = bias_z, gh_n = bias_n;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) {
            gh_r += s.part[cur][wv][bl][lane];
            gh_z += s.part[cur][wv][bl][JS + lane];
            gh_n += s.part[cur][wv][bl][2 * JS + lane];
        }
        GRU_T(4);
        const float r = hw_sigmoid(gi_r + gh_r);
        const float z = hw_sigmoid(gi_z + gh_z);
        const float n = hw_tanh(gi_n + r * gh_n);
        const float hnew = (1.f - z) * n + z * hprev;
        hprev = hnew;
        if (step + 1 < T) {
            s.stage[bl][lane] = __float2bfloat16_rn(hnew);
            __syncwarp();
            const uint4 v = *reinterpret_cast<const uint4*>(&s.stage[bl][piece * 8]);
            GRU_T(5);
            st_async16(peer_h + (uint32_t)(nxt * BS * HPAD * 2), v, peer_bar + 8 * nxt);
        }
        GRU_T(6);
        if (b_ok) {
            out[((long)bglob * T + t) * (2 * HID) + dir * HID + jg] = hnew;
            if (gates != nullptr) {
                float* gp = gates + (((long)bglob * T + t) * 2 + dir) * 4 * HID;
                gp[jg] = r; gp[HID + jg] = z; gp[2 * HID + jg] = n; gp[3 * HID + jg] = gh_n;
            }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) { gi_cur[i] = gi_nxt[i]; gi_nxt[i] = gi_nn[i]; }
    }
    cluster.sync();                   // no CTA retires while a peer may still deliver into it
}

// ---------------------------------------------------------------------------------------------
// Backward with the PRODUCT exchanged instead of its operand.  The hand-off costs ~2 ns per st.async at the receiving
// CTA (measured: 256 pieces / step -> 0.55 us, 512 -> 1.1 us, scripts/gru_timing.py), and shipping
// every CTA's 96 gate-gradient rows to all 8 CTAs would be 768 pieces per CTA and step: 1.5 us.  Here a CTA keeps its
// gate gradients local and multiplies them with ITS 96 rows of W_hh for ALL 256 hidden units (warp w: units 32 w .. + 31,
// 12 MMAs, the same count as before); the result is one eighth of dh_prev for everybody, and warp w's 8 x 32 block goes
// to CTA w alone as 32 bf16 pieces.  A CTA receives 8 x 32 = 256 pieces per step (one third), sums the 8 partial blocks
// in fp32 and adds the direct term.
constexpr int GPL = 112;              // local gate-gradient row stride (96 rows + pad: 224 B, sequences skew by 24 banks)
constexpr int SPL = 40;               // staging row stride (80 B: conflict-free 2-byte scatter, 16-byte aligned pieces)
struct BwdSmemP {
    bf16 dgl[2][BS][GPL];             // this CTA's gate gradients, [step parity][b][g * 32 + j]: the B operand.  Two buffers:
                                      // this CTA's mbarrier completes on ONE warp of every CTA, so a fast warp may write the
                                      // next step's rows while a slow one still multiplies this step's
    bf16 recv[2][NCTA][BS][JS];       // partial dh_prev blocks, [slot][source CTA][b][k local]
    bf16 stage[8][BS][SPL];           // per warp: its outgoing 8 x 32 block
    unsigned long long bar[2];
};

__global__ void __cluster_dims__(NCTA, 1, 1) __launch_bounds__(256, 1)
gru_bwd_tc_prod_kernel(const float* __restrict__ d_out, const float* __restrict__ out, const float* __restrict__ gates,
                       const float* __restrict__ w_hh, bf16* __restrict__ dgi, bf16* __restrict__ dgh_out,
                       bf16* __restrict__ hprev_out, int B, int T) {
    __shared__ __align__(16) BwdSmemP s;
    cg::cluster_group cluster = cg::this_cluster();
    const int cta = (int)cluster.block_rank();
    const int slice = blockIdx.y, dir = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = lane & 3, rho = lane >> 2;
    const float* W = w_hh + (long)dir * G3 * HID;
    constexpr uint32_t STEP_BYTES = NCTA * BS * JS * 2;          // 4096

    // A[m = kk][kdim = r] = W_hh[global row of local row r][32 warp + kk]; local row r = 32 g + j <-> global g * 256 + 32 cta + j.
    // The k order inside an MMA is permuted (fragment column 2c + e <-> r = 16 kt + 4c + e, column 2c + 8 + e <-> r = 16 kt
    // + 4c + 2 + e) so that a B fragment is ONE 8-byte shared-memory load.
    uint32_t afrag[2][6][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int kt = 0; kt < 6; ++kt) {
#pragma unroll
            for (int hr = 0; hr < 2; ++hr) {
                const int kcol = warp * 32 + mt * 16 + rho + hr * 8;
                float wv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int r = kt * 16 + 4 * c + e;
                    wv[e] = W[(long)((r >> 5) * HID + cta * JS + (r & 31)) * HID + kcol];
                }
                afrag[mt][kt][hr] = pack_bf16(wv[0], wv[1]);
                afrag[mt][kt][hr + 2] = pack_bf16(wv[2], wv[3]);
            }
        }
    }
    for (int i = tid; i < 2 * BS * GPL; i += 256) (&s.dgl[0][0][0])[i] = __float2bfloat16_rn(0.f);
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&s.bar[0]);
    if (tid == 0) {
        g_mbar_init(bar0, 1); g_mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        g_mbar_expect_tx(bar0, STEP_BYTES);
        g_mbar_expect_tx(bar0 + 8, STEP_BYTES);
    }
    cluster.sync();

    const int jg = cta * JS + lane;
    const int bl = warp;
    const int bglob = slice * BS + bl;
    const bool b_ok = bglob < B;
    // outgoing: lane l delivers piece (l & 3) of sequence l >> 2 of this warp's block to CTA `warp`
    const int pseq = lane >> 2, ppiece = lane & 3;
    const uint32_t recv_local = (uint32_t)__cvta_generic_to_shared(&s.recv[0][0][0][0]);
    const uint32_t dst_recv = g_mapa(recv_local, (uint32_t)warp) + (uint32_t)(((cta * BS + pseq) * JS + ppiece * 8) * 2);
    const uint32_t dst_bar = g_mapa(bar0, (uint32_t)warp);

    // per-step operands, fetched TWO steps ahead into alternating register sets (the loop body is instantiated twice)
    const long t_stride = dir == 0 ? -1 : 1;                     // the backward pass walks time in reverse
    const int t_first = dir == 0 ? T - 1 : 0;
    const long bt0 = (long)bglob * T + t_first;
    const float* gates_p = gates + (bt0 * 2 + dir) * 4 * HID + jg;
    const float* dout_p = d_out + bt0 * (2 * HID) + dir * HID + jg;
    const float* outp_p = out + bt0 * (2 * HID) + dir * HID + jg;             // out[t]; the previous state is out[t_prev]
    bf16* dgi_p = dgi + bt0 * (2 * G3) + dir * G3 + jg;
    bf16* dgh_p = dgh_out + ((long)dir * B * T + bt0) * G3 + jg;
    bf16* hp_p = hprev_out + ((long)dir * B * T + bt0) * HID + jg;
    struct Ops { float r, z, n, hn, dout, hprev; };
    auto fetch = [&](int step, Ops& o) {
        o.r = o.z = o.n = o.hn = o.dout = o.hprev = 0.f;
        if (!b_ok || step >= T) return;
        const long off = (long)step * t_stride;
        const float* gp = gates_p + off * (8 * HID);
        o.r = __ldg(gp); o.z = __ldg(gp + HID); o.n = __ldg(gp + 2 * HID); o.hn = __ldg(gp + 3 * HID);
        o.dout = __ldg(dout_p + off * (2 * HID));
        if (step + 1 < T) o.hprev = __ldg(outp_p + (off + t_stride) * (2 * HID));   // t_prev = t -+ 1 inside the sequence
    };
    float dh = 0.f;
    auto do_step = [&](int step, Ops& ops) {
        const int cur = step & 1;
        const Ops o = ops;                        // this step's operands; the set is refilled for step + 2 right away
        fetch(step + 2, ops);
        float g_r = 0.f, g_z = 0.f, g_n = 0.f, dh_direct = 0.f, s_dn = 0.f, s_hp = 0.f;
        if (b_ok) {
            const float r = o.r, z = o.z, n = o.n, hn = o.hn, hprev = o.hprev;
            const float dht = dh + o.dout;
            const float dn_pre = dht * (1.f - z) * (1.f - n * n);
            const float dz_pre = dht * (hprev - n) * z * (1.f - z);
            const float dr_pre = dn_pre * hn * r * (1.f - r);
            s_dn = dn_pre; s_hp = hprev;
            g_r = dr_pre; g_z = dz_pre; g_n = dn_pre * r;
            dh_direct = dht * z;
        }
        const bf16 v_r = __float2bfloat16_rn(g_r), v_z = __float2bfloat16_rn(g_z), v_n = __float2bfloat16_rn(g_n);
        const bool exchange = step + 1 < T;       // the last step's dh_prev has no consumer
        if (exchange) { s.dgl[cur][bl][lane] = v_r; s.dgl[cur][bl][JS + lane] = v_z; s.dgl[cur][bl][2 * JS + lane] = v_n; }
        if (b_ok) {
            const long off = (long)step * t_stride;
            bf16* gi_o = dgi_p + off * (2 * G3);
            gi_o[0] = v_r; gi_o[HID] = v_z; gi_o[2 * HID] = __float2bfloat16_rn(s_dn);
            bf16* gh_o = dgh_p + off * G3;
            gh_o[0] = v_r; gh_o[HID] = v_z; gh_o[2 * HID] = v_n;
            hp_p[off * HID] = __float2bfloat16_rn(s_hp);
        }
        if (!exchange) return;
        __syncthreads();                           // all 96 x 8 gate gradients of this CTA are in dgl
        float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
        const bf16* brow = &s.dgl[cur][rho][4 * c];
#pragma unroll
        for (int kt = 0; kt < 6; ++kt) {
            const uint2 b = *reinterpret_cast<const uint2*>(brow + kt * 16);
            mma_bf16_16816(d0, afrag[0][kt], b.x, b.y);
            mma_bf16_16816(d1, afrag[1][kt], b.x, b.y);
        }
        {   // d0: units rho, rho + 8 (of this warp's 32), d1: 16 + rho, 24 + rho; sequences 2c, 2c + 1
            bf16* st = &s.stage[warp][0][0];
            st[(2 * c) * SPL + rho] = __float2bfloat16_rn(d0[0]);      st[(2 * c + 1) * SPL + rho] = __float2bfloat16_rn(d0[1]);
            st[(2 * c) * SPL + rho + 8] = __float2bfloat16_rn(d0[2]);  st[(2 * c + 1) * SPL + rho + 8] = __float2bfloat16_rn(d0[3]);
            st[(2 * c) * SPL + rho + 16] = __float2bfloat16_rn(d1[0]); st[(2 * c + 1) * SPL + rho + 16] = __float2bfloat16_rn(d1[1]);
            st[(2 * c) * SPL + rho + 24] = __float2bfloat16_rn(d1[2]); st[(2 * c + 1) * SPL + rho + 24] = __float2bfloat16_rn(d1[3]);
        }
        __syncwarp();
        const uint4 v = *reinterpret_cast<const uint4*>(&s.stage[warp][pseq][ppiece * 8]);
        st_async16(dst_recv + (uint32_t)(cur * NCTA * BS * JS * 2), v, dst_bar + 8 * cur);
        g_mbar_wait(bar0 + 8 * cur, (uint32_t)((step >> 1) & 1));
        if (tid == 0) g_mbar_expect_tx(bar0 + 8 * cur, STEP_BYTES);          // re-arm for step + 2
        float sum = dh_direct;
#pragma unroll
        for (int src = 0; src < NCTA; ++src) sum += __bfloat162float(s.recv[cur][src][bl][lane]);
        dh = sum;
    };
    Ops oa, ob;
    fetch(0, oa);
    fetch(1, ob);
    for (int step = 0; step < T; step += 2) {
        do_step(step, oa);
        if (step + 1 < T) do_step(step + 1, ob);
    }
    cluster.sync();
}

}  // namespace

#ifdef TAG_GRU_TIMING
extern "C" int tag_gru_timing_read(long long* host64x8) {
    return (int)cudaMemcpyFromSymbol(host64x8, g_gru_timing, sizeof(long long) * 64 * 8);
}
#endif

extern "C" int tag_gru_fwd_bf16(const float* gi, const float* w_hh, const float* b_hh, float* out,
                                float* gates, int B, int T, cudaStream_t stream) {
    if (B <= 0 || T <= 0) return TAG_ERR_BAD_ARG;
    dim3 grid(NCTA, (B + BS - 1) / BS, 2);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gru_fwd_tc_async_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(FwdSmemA));
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    gru_fwd_tc_async_kernel<<<grid, 256, sizeof(FwdSmemA), stream>>>(gi, w_hh, b_hh, out, gates, B, T);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_gru_bwd_bf16(const float* d_out, const float* out, const float* gates, const float* w_hh,
                                void* dgi, void* dgh, void* hprev, int B, int T, cudaStream_t stream) {
    if (B <= 0 || T <= 0) return TAG_ERR_BAD_ARG;
    dim3 grid(NCTA, (B + BS - 1) / BS, 2);
    gru_bwd_tc_prod_kernel<<<grid, 256, 0, stream>>>(d_out, out, gates, w_hh, (bf16*)dgi, (bf16*)dgh, (bf16*)hprev, B, T);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}
