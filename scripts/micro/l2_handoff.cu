// Does a consumer that walks a just-written tensor BACKWARDS hit the 126 MB L2 where a forward walk thrashes it?
// producer: y[i] = f(x[i]) ascending grid-stride; consumer: z[j] = g(y[j]) ascending or descending.  bf16-sized traffic.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 l2_handoff.cu -o l2_handoff
#include <cstdio>
#include <cuda_runtime.h>
__global__ void pass(const uint4* __restrict__ in, uint4* __restrict__ out, long n, int reverse) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const long j = reverse ? n - 1 - i : i;
        uint4 v = in[j];
        v.x ^= 0x1u; v.y += 3u;
        out[j] = v;
    }
}
int main() {
    const long sizes_mb[] = {64, 100, 131, 262, 525};
    uint4 *a, *b, *c, *flush;
    cudaMalloc(&a, 600l << 20); cudaMalloc(&b, 600l << 20); cudaMalloc(&c, 600l << 20); cudaMalloc(&flush, 512l << 20);
    cudaMemset(a, 1, 600l << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (long mb : sizes_mb) {
        const long n = (mb << 20) / 16;
        for (int rev = 0; rev < 2; ++rev) {
            float best = 1e9f;
            for (int it = 0; it < 5; ++it) {
                cudaMemsetAsync(flush, it, 512l << 20);
                pass<<<148 * 16, 256>>>(a, b, n, 0);            // producer writes b ascending
                cudaEventRecord(e0);
                pass<<<148 * 16, 256>>>(b, c, n, rev);          // consumer reads b
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                best = ms < best ? ms : best;
            }
            printf("%4ld MB  consumer %s: %.3f ms  (%.0f GB/s of read+write)\n", mb, rev ? "descending" : "ascending ",
                   best, 2.0 * mb * 1.048576 / best);
        }
    }
    return 0;
}
