// What does one all-to-all hand-off of 4 KB per CTA cost inside an 8-CTA cluster, as a function of how it is cut up?
// Every CTA sends 512 B (8 rows x 64 B) to each of the 8 CTAs (itself included) per step and waits on its local
// mbarrier for the 4096 B of the step — the exchange pattern of the GRU forward recurrence (csrc/gru_tc.cu), no compute.
//   mode 0: 256 x st.async 16 B            (256 arrivals per CTA and step; what the GRU kernels do)
//   mode 1:  64 x cp.async.bulk 64 B       (lane l < 8 of warp w ships row w to CTA l)
//   mode 2:   8 x cp.async.bulk 512 B      (after a __syncthreads; thread p < 8 ships the whole block to CTA p)
//   mode 3: 128 x st.async 16 B + 0        (half the pieces, half the bytes: is the cost per piece or per byte?)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 dsmem_handoff.cu -o dsmem_handoff
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        if (++spins > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async16(uint32_t dst, const uint4& v, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2c(uint32_t dst, uint32_t src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "r"(src), "r"(bytes), "r"(bar) : "memory");
}

struct Smem {
    uint32_t recv[2][8][128];      // [slot][source CTA][8 rows x 16 words]
    uint32_t stage[2][8][16];      // this CTA's 8 rows of 64 B, alternating: a bulk copy may still read the last one
    unsigned long long bar[2];
};

template <int MODE>
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(256, 1) handoff(int steps, unsigned* bad) {
    __shared__ __align__(128) Smem s;
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t cta = cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr uint32_t STEP_BYTES = MODE == 3 ? 2048 : 4096;
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&s.bar[0]);
    if (tid == 0) {
        mbar_init(bar0, 1); mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar0, STEP_BYTES);
        mbar_expect_tx(bar0 + 8, STEP_BYTES);
    }
    cluster.sync();
    const uint32_t recv_local = (uint32_t)__cvta_generic_to_shared(&s.recv[0][0][0]);
    const uint32_t stage_local = (uint32_t)__cvta_generic_to_shared(&s.stage[0][0][0]);
    // mode 0/3: lane l delivers piece (l & 3) of row `warp` to peer l >> 2
    const uint32_t peer = lane >> 2, piece = lane & 3;
    const uint32_t dst0 = mapa(recv_local, peer) + (cta * 128 + warp * 16 + piece * 4) * 4;
    const uint32_t bar_peer0 = mapa(bar0, peer);
    // mode 1: lane l < 8 of warp w ships row w to CTA l
    const uint32_t dst1 = mapa(recv_local, lane & 7) + (cta * 128 + warp * 16) * 4;
    const uint32_t bar_peer1 = mapa(bar0, lane & 7);
    // mode 2: thread p < 8 ships the block to CTA p
    const uint32_t dst2 = mapa(recv_local, tid & 7) + cta * 128 * 4;
    const uint32_t bar_peer2 = mapa(bar0, tid & 7);
    unsigned errors = 0;
    for (int step = 0; step < steps; ++step) {
        const int slot = step & 1;
        if (step > 0) {
            const int prev = slot ^ 1;      // data of step - 1 went to slot (step - 1) & 1
            mbar_wait(bar0 + 8 * prev, (uint32_t)(((step - 1) >> 1) & 1));
            if (tid == 0) mbar_expect_tx(bar0 + 8 * prev, STEP_BYTES);
            // every thread checks one word of what arrived (source CTA = warp)
            const uint32_t got = s.recv[prev][warp][lane];
            if (got != (uint32_t)(step - 1) * 8u + warp && !(MODE == 3 && (lane & 15) >= 8)) ++errors;
        }
        if (lane < 16) s.stage[slot][warp][lane] = (uint32_t)step * 8u + cta;
        const uint32_t off = (uint32_t)slot * 8 * 128 * 4;
        if (MODE == 0 || MODE == 3) {
            __syncwarp();
            if (MODE == 0 || piece < 2) {
                const uint4 v = *reinterpret_cast<const uint4*>(&s.stage[slot][warp][piece * 4]);
                st_async16(dst0 + off, v, bar_peer0 + 8 * slot);
            }
        } else if (MODE == 1) {
            __syncwarp();
            if (lane < 8) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                bulk_s2c(dst1 + off, stage_local + slot * 512 + warp * 64, 64, bar_peer1 + 8 * slot);
            }
        } else {
            __syncthreads();
            if (tid < 8) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                bulk_s2c(dst2 + off, stage_local + slot * 512, 512, bar_peer2 + 8 * slot);
            }
        }
    }
    mbar_wait(bar0 + 8 * ((steps - 1) & 1), (uint32_t)(((steps - 1) >> 1) & 1));
    if (errors) atomicAdd(bad, errors);
    cluster.sync();
}

template <int MODE>
void run(const char* what, unsigned* bad) {
    const int steps = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaMemset(bad, 0, 4);
    handoff<MODE><<<dim3(128), 256>>>(200, bad);
    cudaEventRecord(e0);
    handoff<MODE><<<dim3(128), 256>>>(steps, bad);
    cudaEventRecord(e1);
    cudaError_t err = cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    unsigned h = 0; cudaMemcpy(&h, bad, 4, cudaMemcpyDeviceToHost);
    printf("%-34s %7.1f ns / step   (%s, %u bad words)\n", what, ms * 1e6 / steps, cudaGetErrorString(err), h);
}

int main() {
    unsigned* bad; cudaMalloc(&bad, 4);
    run<0>("256 x st.async 16 B", bad);
    run<3>("128 x st.async 16 B (half)", bad);
    run<1>(" 64 x cp.async.bulk 64 B", bad);
    run<2>("  8 x cp.async.bulk 512 B (+sync)", bad);
    return 0;
}
