// Fused log-mel frontend: reflect-pad framing -> Hann -> 1024-pt FFT in shared memory ->
// |X|^2 -> slaney mel filterbank -> 10*log10(clamp) (+ per-mel sum / sum-of-squares for bn0).
//
// Replaces the reference's torchaudio MelSpectrogram + AmplitudeToDB call chain
// (reference models/audio_encoder.py:113-124,183-184), which materialises the complex
// spectrum, the magnitude and the power tensors in HBM (SURVEY.md §8a rows a1,a2).
// Here one CTA stages 8 frames worth of contiguous samples (3264 floats, coalesced),
// runs four 1024-point complex FFTs (Stockham radix-4; two real frames packed per FFT) in shared
// memory and writes only the [B, T0, 64] dB tensor.
//
// Two kernels: logmel_warp_kernel (the production path: one warp per pair of frames, the 1024-point FFT as two
// register-resident 32-point passes around ONE shared-memory transpose, compact slaney filterbank) and
// logmel_kernel (any dense filterbank; Stockham radix-4 in shared memory) as the general case.
#include "common.cuh"
#include <cuda_fp16.h>
#include "fft32_gen.cuh"

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


// ------------------------------------------------------------------------------------------------------------
// Warp-per-frame-pair kernel.  1024 = 32 x 32: lane j loads x[j + 32k] (k = 0..31; every k is one coalesced 128 B
// row of the waveform, windowed on the fly, frames 2p / 2p+1 packed as re / im), runs a 32-point FFT over k in
// registers, multiplies by W_1024^(j m) (table [m][lane], conflict free), transposes through a padded [32][33]
// float2 tile, and runs the second 32-point FFT over j: lane m then holds X[m + 32 n].  The two real spectra are
// separated with one shuffle per bin (the partner X[1024 - k] lives in lane (32 - m) & 31), |X|^2 goes to shared
// memory and lane l accumulates mel bins l and 63 - l (narrow + wide triangle: balanced) from a compact
// filterbank.  No __syncthreads in the main loop.
constexpr int W_WARPS = 8;
constexpr int FB_CAP = 2048;                 // compact filterbank weights (slaney 64 x 513 needs ~1030)
constexpr int P_ROW = 520;

struct WarpSmem {
    float2 tw[32][32];                       // tw[m][j] = exp(-2 pi i j m / 1024)
    float win[N_FFT];
    float fbw[FB_CAP];
    int fb_lo[N_MELS], fb_hi[N_MELS], fb_off[N_MELS];
    float red[W_WARPS][N_MELS][2];
    union {
        float2 tile[32][33];
        float power[2][P_ROW];
    } w[W_WARPS];
};

template <typename TIn> __device__ __forceinline__ float wav_ld(const TIn* p);
template <> __device__ __forceinline__ float wav_ld<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float wav_ld<__half>(const __half* p) { return __half2float(__ldg(p)); }

template <typename TIn>
__global__ void __launch_bounds__(W_WARPS * 32, 2)
logmel_warp_kernel(const TIn* __restrict__ wav, int n_clips, int L, long wav_stride, int T0,
                   const float* __restrict__ window, const float* __restrict__ fb,
                   const int* __restrict__ mel_range, float* __restrict__ db_out, double* __restrict__ stats) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpSmem& s = *reinterpret_cast<WarpSmem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 32 * 32; i += W_WARPS * 32) {
        const int m = i >> 5, j = i & 31;
        float sn, cs;
        sincospif(2.0f * (float)(j * m) / (float)N_FFT, &sn, &cs);
        s.tw[m][j] = make_float2(cs, -sn);
    }
    for (int i = tid; i < N_FFT; i += W_WARPS * 32) s.win[i] = __ldg(window + i);
    if (tid == 0) {
        int off = 0;
        for (int m = 0; m < N_MELS; ++m) {
            s.fb_lo[m] = mel_range[2 * m]; s.fb_hi[m] = mel_range[2 * m + 1]; s.fb_off[m] = off;
            off += mel_range[2 * m + 1] - mel_range[2 * m];
        }
    }
    __syncthreads();
    for (int m = warp; m < N_MELS; m += W_WARPS)
        for (int k = s.fb_lo[m] + lane; k < s.fb_hi[m]; k += 32) s.fbw[s.fb_off[m] + k - s.fb_lo[m]] = __ldg(fb + k * N_MELS + m);
    __syncthreads();

    const int mA = lane, mB = N_MELS - 1 - lane;
    const int loA = s.fb_lo[mA], hiA = s.fb_hi[mA], loB = s.fb_lo[mB], hiB = s.fb_hi[mB];
    const float* wA = s.fbw + s.fb_off[mA] - loA;
    const float* wB = s.fbw + s.fb_off[mB] - loB;
    float stA = 0.f, sqA = 0.f, stB = 0.f, sqB = 0.f;
    float2 (&tile)[32][33] = s.w[warp].tile;
    float (&power)[2][P_ROW] = s.w[warp].power;

    const int pairs_per_clip = (T0 + 1) / 2;
    const int total = n_clips * pairs_per_clip;
#pragma unroll 1
    for (int item = blockIdx.x * W_WARPS + warp; item < total; item += gridDim.x * W_WARPS) {
        const int b = item / pairs_per_clip;
        const int tA = (item - b * pairs_per_clip) * 2, tB = tA + 1;
        const TIn* w = wav + (long)b * wav_stride;
        float re[32], im[32];
        const int startA = tA * HOP - N_FFT / 2;
        if (startA >= 0 && startA + HOP + N_FFT <= L && tB < T0) {          // interior pair: no reflection
            const TIn* pa = w + startA + lane;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const float wn = s.win[lane + 32 * k];
                re[k] = wav_ld(pa + 32 * k) * wn;
                im[k] = wav_ld(pa + HOP + 32 * k) * wn;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const float wn = s.win[lane + 32 * k];
                int ia = startA + lane + 32 * k;
                int ib = ia + HOP;
                if (ia < 0) ia = -ia;
                if (ia >= L) ia = 2 * (L - 1) - ia;
                if (ib < 0) ib = -ib;
                if (ib >= L) ib = 2 * (L - 1) - ib;
                re[k] = (ia >= 0 && ia < L) ? wav_ld(w + ia) * wn : 0.f;
                im[k] = (tB < T0 && ib >= 0 && ib < L) ? wav_ld(w + ib) * wn : 0.f;
            }
        }
        fft32(re, im);                                  // over k; Y_j[m] in index bitrev(m)
#pragma unroll
        for (int m = 0; m < 32; ++m) {
            const int r = fft32_bitrev(m);
            const float2 t = s.tw[m][lane];
            tile[m][lane] = make_float2(re[r] * t.x - im[r] * t.y, re[r] * t.y + im[r] * t.x);
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) { const float2 v = tile[lane][j]; re[j] = v.x; im[j] = v.y; }
        __syncwarp();                                   // tile is dead: its memory becomes the power rows
        fft32(re, im);                                  // over j; X[lane + 32 n] in index bitrev(n)
        // bins k = lane + 32 n, n = 0..15 (and k = 512 from lane 0): X_a = (Z[k] + conj Z[N-k]) / 2,
        // X_b = (Z[k] - conj Z[N-k]) / (2i)
        const int src = (32 - lane) & 31;
#pragma unroll
        for (int n = 0; n <= 16; ++n) {
            const int r = fft32_bitrev(n);
            const int rp = fft32_bitrev(31 - n), rp0 = fft32_bitrev((32 - n) & 31);
            float pr = __shfl_sync(0xffffffffu, re[rp], src);
            float pi = __shfl_sync(0xffffffffu, im[rp], src);
            if (lane == 0) { pr = re[rp0]; pi = im[rp0]; }
            const float ar = re[r] + pr, ai = im[r] - pi;
            const float br = re[r] - pr, bi = im[r] + pi;
            if (n < 16 || lane == 0) {
                power[0][lane + 32 * n] = 0.25f * (ar * ar + ai * ai);
                power[1][lane + 32 * n] = 0.25f * (br * br + bi * bi);
            }
        }
        __syncwarp();
        float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
        for (int k = loA; k < hiA; ++k) { const float wv = wA[k]; a0 = fmaf(power[0][k], wv, a0); a1 = fmaf(power[1][k], wv, a1); }
        for (int k = loB; k < hiB; ++k) { const float wv = wB[k]; b0 = fmaf(power[0][k], wv, b0); b1 = fmaf(power[1][k], wv, b1); }
        __syncwarp();                                   // power rows are dead: next item's tile may overwrite them
        a0 = 10.0f * log10f(fmaxf(a0, 1e-10f)); b0 = 10.0f * log10f(fmaxf(b0, 1e-10f));
        float* oA = db_out + ((long)b * T0 + tA) * N_MELS;
        oA[mA] = a0; oA[mB] = b0;
        stA += a0; sqA = fmaf(a0, a0, sqA); stB += b0; sqB = fmaf(b0, b0, sqB);
        if (tB < T0) {
            a1 = 10.0f * log10f(fmaxf(a1, 1e-10f)); b1 = 10.0f * log10f(fmaxf(b1, 1e-10f));
            oA[N_MELS + mA] = a1; oA[N_MELS + mB] = b1;
            stA += a1; sqA = fmaf(a1, a1, sqA); stB += b1; sqB = fmaf(b1, b1, sqB);
        }
    }
    if (stats != nullptr) {
        s.red[warp][mA][0] = stA; s.red[warp][mA][1] = sqA;
        s.red[warp][mB][0] = stB; s.red[warp][mB][1] = sqB;
        __syncthreads();
        if (tid < N_MELS) {
            double a = 0.0, q = 0.0;
            for (int wv = 0; wv < W_WARPS; ++wv) { a += (double)s.red[wv][tid][0]; q += (double)s.red[wv][tid][1]; }
            atomicAdd(stats + tid, a);
            atomicAdd(stats + N_MELS + tid, q);
        }
    }
}

template <typename TIn>
int launch_warp(const TIn* wav, int batch, int n_samples, long wav_stride, int T0, const float* window,
                const float* fb, const int* mel_range, float* db_out, double* stats, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(logmel_warp_kernel<TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(WarpSmem));
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const long items = (long)batch * ((T0 + 1) / 2);
    long ctas = (items + W_WARPS - 1) / W_WARPS;
    if (ctas > 148 * 2) ctas = 148 * 2;                                       // 2 CTAs (87 KB each) per SM
    logmel_warp_kernel<TIn><<<(int)ctas, W_WARPS * 32, sizeof(WarpSmem), stream>>>(
        wav, batch, n_samples, wav_stride, T0, window, fb, mel_range, db_out, stats);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

}  // namespace

// mel_range_host (optional, host pointer, [2*64] ints = the same values as mel_range): lets the launcher check that
// the filterbank is compact enough for the warp kernel; without it the general kernel runs.
static int logmel_dense(const float* wav, int batch, int n_samples, long wav_stride, const float* window,
                        const float* fb, const int* mel_range, float* db_out, double* stats, cudaStream_t stream) {
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

extern "C" int tag_logmel_fwd(const float* wav, int batch, int n_samples, long wav_stride,
                              const float* window, const float* fb, const int* mel_range,
                              float* db_out, double* stats, cudaStream_t stream) {
    if (batch <= 0 || n_samples <= N_FFT / 2) return TAG_ERR_BAD_ARG;
    return logmel_dense(wav, batch, n_samples, wav_stride, window, fb, mel_range, db_out, stats, stream);
}

extern "C" int tag_logmel_fwd_v2(const void* wav, int wav_dtype, int batch, int n_samples, long wav_stride,
                                 const float* window, const float* fb, const int* mel_range, int fb_nnz,
                                 float* db_out, double* stats, cudaStream_t stream) {
    if (batch <= 0 || n_samples <= N_FFT / 2 || mel_range == nullptr) return TAG_ERR_BAD_ARG;
    if (fb_nnz <= 0 || fb_nnz > FB_CAP) return TAG_ERR_UNSUPPORTED;
    const int T0 = n_samples / HOP + 1;
    if (wav_dtype == 0)
        return launch_warp<float>((const float*)wav, batch, n_samples, wav_stride, T0, window, fb, mel_range, db_out,
                                  stats, stream);
    if (wav_dtype == 2)
        return launch_warp<__half>((const __half*)wav, batch, n_samples, wav_stride, T0, window, fb, mel_range,
                                   db_out, stats, stream);
    return TAG_ERR_BAD_ARG;
}
