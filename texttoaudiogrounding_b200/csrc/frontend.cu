// Fused log-mel frontend: reflect-pad framing -> Hann -> 1024-pt FFT in shared memory ->
// |X|^2 -> slaney mel filterbank -> 10*log10(clamp) (+ per-mel sum / sum-of-squares for bn0).
//
// Replaces the reference's torchaudio MelSpectrogram + AmplitudeToDB call chain
// (reference models/audio_encoder.py:113-124,183-184), which materialises the complex
// spectrum, the magnitude and the power tensors in HBM (SURVEY.md §8a rows a1,a2).
// Here one CTA stages 8 frames worth of contiguous samples (3264 floats, coalesced),
// runs four 1024-point complex FFTs (Stockham radix-4; two real frames packed per FFT) in shared
// memory and writes only the [B, T0, 64] dB tensor.
#include "common.cuh"

namespace {

constexpr int N_FFT = 1024;
constexpr int HOP = 320;
constexpr int N_BINS = 513;
constexpr int N_MELS = 64;
constexpr int FRAMES_PER_CTA = 8;
constexpr int PAIRS = FRAMES_PER_CTA / 2;
constexpr int SPAN = N_FFT + (FRAMES_PER_CTA - 1) * HOP;  // 3264 samples
constexpr int P_STRIDE = 516;
constexpr int THREADS = 256;

struct FrontendSmem {
    float samples[SPAN];
    float2 z[PAIRS][N_FFT];
    float2 tw[3 * N_FFT / 4];
    float win[N_FFT];
    float power[FRAMES_PER_CTA][P_STRIDE];
    float db[FRAMES_PER_CTA][N_MELS];
};

__global__ void __launch_bounds__(THREADS)
logmel_kernel(const float* __restrict__ wav, int n_clips, int L, long wav_stride, int T0,
              const float* __restrict__ window, const float* __restrict__ fb,
              const int* __restrict__ mel_range, float* __restrict__ db_out,
              double* __restrict__ stats) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrontendSmem& s = *reinterpret_cast<FrontendSmem*>(smem_raw);
    const int tid = threadIdx.x;
    // persistent over 8-frame groups: twiddles / window are staged once per CTA and the bn0 statistics are
    // flushed once per CTA (one flush per group meant 1 M double atomics on 8 cache lines)
    for (int i = tid; i < 3 * N_FFT / 4; i += THREADS) {
        float sn, cs;
        sincospif(2.0f * (float)i / (float)N_FFT, &sn, &cs);
        s.tw[i] = make_float2(cs, -sn);
    }
    for (int i = tid; i < N_FFT; i += THREADS) s.win[i] = __ldg(window + i);
    float st_sum = 0.f, st_sq = 0.f;
    const int groups_per_clip = (T0 + FRAMES_PER_CTA - 1) / FRAMES_PER_CTA;
#pragma unroll 1
    for (int grp = blockIdx.x; grp < n_clips * groups_per_clip; grp += gridDim.x) {
    const int b = grp / groups_per_clip;
    const int t0 = (grp - b * groups_per_clip) * FRAMES_PER_CTA;
    const float* w = wav + (long)b * wav_stride;

    // ---- stage samples (reflect padding of 512 on both sides, torch.stft center=True)
    const int start = t0 * HOP - N_FFT / 2;
    __syncthreads();                       // previous group's readers of samples / db are done
    for (int i = tid; i < SPAN; i += THREADS) {
        int src = start + i;
        if (src < 0) src = -src;
        if (src >= L) src = 2 * (L - 1) - src;
        float v = 0.f;
        if (src >= 0 && src < L) v = __ldg(w + src);
        s.samples[i] = v;
    }
    __syncthreads();

    // ---- windowed load (natural order); frame 2p -> real part, 2p+1 -> imaginary part
    for (int i = tid; i < PAIRS * N_FFT; i += THREADS) {
        int p = i >> 10, n = i & (N_FFT - 1);
        float wn = s.win[n];
        float xa = s.samples[(2 * p) * HOP + n] * wn;
        float xb = s.samples[(2 * p + 1) * HOP + n] * wn;
        s.z[p][n] = make_float2(xa, xb);
    }
    __syncthreads();

    // ---- Stockham radix-4, 5 stages (1024 = 4^5); each thread owns 4 butterflies per stage.
    //      in-place on one buffer: all reads of a stage complete before its writes.
#pragma unroll 1
    for (int ns_log = 0; ns_log < 10; ns_log += 2) {
        const int Ns = 1 << ns_log;                 // size of the sub-transforms already formed
        const int tw_step = (N_FFT / 4) >> ns_log;  // N / (4 Ns)
        float2 u[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int q = tid + THREADS * i;
            const int p = q >> 8, j = q & 255;
            const int k = j & (Ns - 1);
            const int m1 = k * tw_step;
            u[i][0] = s.z[p][j];
#pragma unroll
            for (int r = 1; r < 4; ++r) {
                const float2 x = s.z[p][j + r * (N_FFT / 4)];
                const float2 wv = s.tw[r * m1];
                u[i][r] = make_float2(x.x * wv.x - x.y * wv.y, x.x * wv.y + x.y * wv.x);
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int q = tid + THREADS * i;
            const int p = q >> 8, j = q & 255;
            const int k = j & (Ns - 1);
            const int j0 = ((j - k) << 2) + k;
            const float2 v0 = make_float2(u[i][0].x + u[i][2].x, u[i][0].y + u[i][2].y);
            const float2 v1 = make_float2(u[i][0].x - u[i][2].x, u[i][0].y - u[i][2].y);
            const float2 v2 = make_float2(u[i][1].x + u[i][3].x, u[i][1].y + u[i][3].y);
            const float2 d = make_float2(u[i][1].x - u[i][3].x, u[i][1].y - u[i][3].y);
            const float2 v3 = make_float2(d.y, -d.x);                   // (u1 - u3) * (-i)
            s.z[p][j0] = make_float2(v0.x + v2.x, v0.y + v2.y);
            s.z[p][j0 + Ns] = make_float2(v1.x + v3.x, v1.y + v3.y);
            s.z[p][j0 + 2 * Ns] = make_float2(v0.x - v2.x, v0.y - v2.y);
            s.z[p][j0 + 3 * Ns] = make_float2(v1.x - v3.x, v1.y - v3.y);
        }
        __syncthreads();
    }

    // ---- separate the two real spectra and take |.|^2
    for (int i = tid; i < PAIRS * N_BINS; i += THREADS) {
        int p = i / N_BINS, k = i - p * N_BINS;
        float2 zk = s.z[p][k];
        float2 zn = s.z[p][(N_FFT - k) & (N_FFT - 1)];
        float ar = zk.x + zn.x, ai = zk.y - zn.y;   // 2 * X_a[k]
        float br = zk.x - zn.x, bi = zk.y + zn.y;   // 2i * X_b[k]
        s.power[2 * p][k] = 0.25f * (ar * ar + ai * ai);
        s.power[2 * p + 1][k] = 0.25f * (br * br + bi * bi);
    }
    __syncthreads();

    // ---- mel filterbank (each mel touches only bins [lo, hi)) + dB
    for (int i = tid; i < FRAMES_PER_CTA * N_MELS; i += THREADS) {
        int f = i >> 6, m = i & 63;
        int lo = 0, hi = N_BINS;
        if (mel_range != nullptr) { lo = mel_range[2 * m]; hi = mel_range[2 * m + 1]; }
        float acc = 0.f;
        for (int k = lo; k < hi; ++k) acc = fmaf(s.power[f][k], __ldg(fb + k * N_MELS + m), acc);
        float v = 10.0f * log10f(fmaxf(acc, 1e-10f));
        s.db[f][m] = v;
        int t = t0 + f;
        if (t < T0) db_out[((long)b * T0 + t) * N_MELS + m] = v;
    }
    if (stats != nullptr) {
        __syncthreads();
        if (tid < N_MELS) {
            for (int f = 0; f < FRAMES_PER_CTA; ++f) {
                if (t0 + f < T0) { float v = s.db[f][tid]; st_sum += v; st_sq += v * v; }
            }
        }
    }
    }
    if (stats != nullptr && tid < N_MELS) {
        atomicAdd(stats + tid, (double)st_sum);
        atomicAdd(stats + N_MELS + tid, (double)st_sq);
    }
}

}  // namespace

extern "C" int tag_logmel_fwd(const float* wav, int batch, int n_samples, long wav_stride,
                              const float* window, const float* fb, const int* mel_range,
                              float* db_out, double* stats, cudaStream_t stream) {
    if (batch <= 0 || n_samples <= N_FFT / 2) return TAG_ERR_BAD_ARG;
    const int T0 = n_samples / HOP + 1;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(FrontendSmem));
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    long groups = (long)((T0 + FRAMES_PER_CTA - 1) / FRAMES_PER_CTA) * batch;
    const int grid = (int)(groups < 148 * 3 ? groups : 148 * 3);          // 3 CTAs (74 KB each) per SM
    logmel_kernel<<<grid, THREADS, sizeof(FrontendSmem), stream>>>(
        wav, batch, n_samples, wav_stride, T0, window, fb, mel_range, db_out, stats);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}
