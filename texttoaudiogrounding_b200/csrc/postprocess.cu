// Post-processing of frame probabilities into event regions, for every (sample, threshold) pair at once.
//
// Reference: Runner.eval_inference (python_scripts/training/run_strong.py:222-247) loops on the HOST over samples and
// over n_thresholds (50) thresholds and calls, per pair, eval_util.median_filter (utils/eval_util.py:55-63: sklearn
// binarize `x > th` + scipy.ndimage.median_filter along time, mode="reflect"), eval_util.connect_clusters (:74-116: merge
// regions whose gap is <= n frames) and eval_util.find_contiguous_regions (:18-44).  Integer / boolean work: the
// results here are bit-exact.  One CTA per sample (row staged in shared memory), one thread per threshold.
#include "common.cuh"

namespace {

__global__ void frame_regions_kernel(const float* __restrict__ sim, long sim_stride, const double* __restrict__ thresholds,
                                     int T, int n_th, int window, int n_connect, int max_regions,
                                     int* __restrict__ regions, int* __restrict__ counts) {
    extern __shared__ float s_row[];
    const int b = blockIdx.x;
    for (int t = threadIdx.x; t < T; t += blockDim.x) s_row[t] = sim[(long)b * sim_stride + t];
    __syncthreads();
    const int half = window / 2;
    const int need = window - half;          // ones needed for the rank-(window/2) element of a 0/1 window to be 1
    for (int k = threadIdx.x; k < n_th; k += blockDim.x) {
        const double th = thresholds[k];
        int* out = regions + ((long)b * n_th + k) * max_regions * 2;
        int n_out = 0;
        int cur_s = -1, cur_e = -1;          // pending (merged) region
        int run_s = -1;                      // start of the run being scanned
        for (int t = 0; t <= T; ++t) {
            bool on = false;
            if (t < T) {
                int ones = 0;
                for (int i = 0; i < window; ++i) {
                    int idx = t - half + i;
                    if (idx < 0) idx = -idx - 1;               // scipy "reflect": d c b a | a b c d | d c b a
                    if (idx >= T) idx = 2 * T - idx - 1;
                    idx = idx < 0 ? 0 : (idx >= T ? T - 1 : idx);
                    ones += (double)s_row[idx] > th ? 1 : 0;   // sklearn binarize: strictly greater, float64 compare
                }
                on = ones >= need;
            }
            if (on && run_s < 0) run_s = t;
            if (!on && run_s >= 0) {                           // run [run_s, t) ended
                if (cur_s < 0) { cur_s = run_s; cur_e = t; }
                else if (run_s - cur_e <= n_connect) { cur_e = t; }
                else {
                    if (n_out < max_regions) { out[2 * n_out] = cur_s; out[2 * n_out + 1] = cur_e; }
                    ++n_out;
                    cur_s = run_s; cur_e = t;
                }
                run_s = -1;
            }
        }
        if (cur_s >= 0) {
            if (n_out < max_regions) { out[2 * n_out] = cur_s; out[2 * n_out + 1] = cur_e; }
            ++n_out;
        }
        counts[(long)b * n_th + k] = n_out;
    }
}

}  // namespace

extern "C" int tag_frame_regions(const float* sim, long sim_stride, const double* thresholds, int B, int T, int n_th,
                                 int window, int n_connect, int max_regions, int* regions, int* counts,
                                 cudaStream_t stream) {
    if (B <= 0 || T <= 0 || n_th <= 0 || window <= 0 || n_connect < 0 || max_regions <= 0) return TAG_ERR_BAD_ARG;
    if ((size_t)T * sizeof(float) > 48 * 1024) return TAG_ERR_UNSUPPORTED;
    frame_regions_kernel<<<B, 64, (size_t)T * sizeof(float), stream>>>(sim, sim_stride, thresholds, T, n_th, window,
                                                                        n_connect, max_regions, regions, counts);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}
