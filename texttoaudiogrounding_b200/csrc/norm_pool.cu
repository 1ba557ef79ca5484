// BatchNorm (train/eval) + ReLU + avg+max pooling + dropout, forward and backward, NHWC.
//
// Reference semantics: nn.BatchNorm2d (eps 1e-5, momentum 0.1, biased var for
// normalisation, unbiased into running_var), F.relu_, F.avg_pool2d + F.max_pool2d with
// kernel = stride = pool_size and floor mode, F.dropout (reference models/panns.py:47-62,
// models/audio_encoder.py:188-215).  Training-mode BN needs a global reduction, so the
// producing conv kernel emits per-channel sum / sum-of-squares (double) in its epilogue,
// tag_bn_finalize turns them into a per-channel (scale, shift) pair, and the consumer
// applies scale/shift/ReLU on the fly — the normalised tensor of the second conv of a block
// is never written to HBM, only its pooled result is.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------ statistics
__global__ void channel_stats_f32_kernel(const float* __restrict__ x, long rows, int C,
                                         double* __restrict__ stats) {
    // block = 256 threads = (256 / C) row lanes x C channels (C divides 256)
    const int c = threadIdx.x % C;
    const int lane_rows = blockDim.x / C;
    const int r0 = threadIdx.x / C;
    float s = 0.f, q = 0.f;
    for (long r = (long)blockIdx.x * lane_rows + r0; r < rows; r += (long)gridDim.x * lane_rows) {
        float v = x[r * C + c];
        s += v;
        q += v * v;
    }
    __shared__ float ss[256], sq[256];
    ss[threadIdx.x] = s;
    sq[threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.x < C) {
        double ds = 0.0, dq = 0.0;
        for (int i = 0; i < lane_rows; ++i) { ds += ss[i * C + c]; dq += sq[i * C + c]; }
        atomicAdd(stats + c, ds);
        atomicAdd(stats + C + c, dq);
    }
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats, double count, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float momentum, float eps, int training, int update_running,
                                   float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ save_mean, float* __restrict__ save_invstd) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float mean, var;
    if (training) {
        double m = stats[c] / count;
        double v = stats[C + c] / count - m * m;
        if (v < 0.0) v = 0.0;
        mean = (float)m;
        var = (float)v;
        if (update_running && running_mean != nullptr) {
            double unbiased = count > 1.0 ? v * count / (count - 1.0) : v;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
        }
    } else {
        mean = running_mean[c];
        var = running_var[c];
    }
    const float invstd = rsqrtf(var + eps);
    const float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = beta[c] - mean * sc;
    if (save_mean != nullptr) { save_mean[c] = mean; save_invstd[c] = invstd; }
}

// ------------------------------------------------------------------ y = act(x * scale[c] + shift[c])
template <typename TI, typename TO>
__global__ void scale_shift_act_kernel(const TI* __restrict__ x, TO* __restrict__ y,
                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                       long n_vec, int C, int relu) {
    // n_vec counts 8-element vectors; two independent vectors per thread per iteration
    const unsigned stride = gridDim.x * blockDim.x;            // n_vec < 2^31 (host-checked)
    const unsigned CV = (unsigned)C / 8;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)n_vec; i += 2 * stride) {
        const unsigned i2 = i + stride;
        const bool has2 = i2 < (unsigned)n_vec;
        float v[8], u[8];
        load8<TI>(x + (long)i * 8, v);
        if (has2) load8<TI>(x + (long)i2 * 8, u);
        {
            const int c = (int)(i % CV) * 8;
            float sc[8], sh[8];
            load8<float>(scale + c, sc);
            load8<float>(shift + c, sh);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                v[k] = fmaf(v[k], sc[k], sh[k]);
                if (relu) v[k] = fmaxf(v[k], 0.f);
            }
            store8<TO>(y + (long)i * 8, v);
        }
        if (has2) {
            const int c = (int)(i2 % CV) * 8;
            float sc[8], sh[8];
            load8<float>(scale + c, sc);
            load8<float>(shift + c, sh);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                u[k] = fmaf(u[k], sc[k], sh[k]);
                if (relu) u[k] = fmaxf(u[k], 0.f);
            }
            store8<TO>(y + (long)i2 * 8, u);
        }
    }
}

// ------------------------------------------------------------------ BN + ReLU + avg+max pool + dropout
// cnt (optional, uint8 per output element): 4 * (n / (PH PW) + [n > 0]) with n the number of window elements whose ReLU
// gate is open (an integer: n + 4 [n > 0] for 2x2, 2 n + 4 [n > 0] for 1x2), or 0 when the output element was dropped —
// all the backward reduce pass needs besides the output itself: with g the gradient routed back through pool + ReLU,
//     sum_window g * a = dout * out      and      sum_window g = dout * keep_scale * cnt / 4.
template <typename T, int PH, int PW, bool CNT>
__global__ void bn_relu_pool_fwd_kernel(const T* __restrict__ y, T* __restrict__ out, uint8_t* __restrict__ cnt,
                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                        int B, int H, int W, int C, uint64_t seed, const uint64_t* __restrict__ seed_dev, uint32_t thresh,
                                        float keep_scale) {
    if (seed_dev != nullptr) seed += *seed_dev;
    const int Ho = H / PH, Wo = W / PW, CV = C / 8;
    const unsigned n = (unsigned)B * Ho * Wo * CV;             // < 2^31 (host-checked)
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned r = i / (unsigned)CV;
        const int cv = (int)(i - r * CV);
        const unsigned r2 = r / (unsigned)Wo;
        const int wo = (int)(r - r2 * Wo);
        const int b = (int)(r2 / (unsigned)Ho);
        const int ho = (int)(r2 - (unsigned)b * Ho);
        float sc[8], sh[8];
        load8<float>(scale + cv * 8, sc);
        load8<float>(shift + cv * 8, sh);
        float sum[8], mx[8];
        uint32_t nopen[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { sum[k] = 0.f; mx[k] = -INFINITY; nopen[k] = 0u; }
#pragma unroll
        for (int dh = 0; dh < PH; ++dh)
#pragma unroll
            for (int dw = 0; dw < PW; ++dw) {
                float v[8];
                load8<T>(y + (((long)b * H + ho * PH + dh) * W + wo * PW + dw) * C + cv * 8, v);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float a = fmaxf(fmaf(v[k], sc[k], sh[k]), 0.f);
                    sum[k] += a;
                    mx[k] = fmaxf(mx[k], a);
                    if (CNT) nopen[k] += a > 0.f ? 1u : 0u;
                }
            }
        float o[8];
        const long obase = (long)i * 8;
        float ds[8];
        if (thresh != 0u) {
            tag_dropout_scale4(seed, (uint64_t)obase, thresh, keep_scale, ds);
            tag_dropout_scale4(seed, (uint64_t)obase + 4, thresh, keep_scale, ds + 4);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float v = sum[k] * (1.0f / (PH * PW)) + mx[k];
            if (thresh != 0u) { v *= ds[k]; if (CNT && ds[k] == 0.f) nopen[k] = 0u; }
            o[k] = v;
        }
        store8<T>(out + obase, o);
        if (CNT) {
#pragma unroll
            for (int k = 0; k < 8; ++k) nopen[k] = nopen[k] != 0u ? nopen[k] * (4u / (PH * PW)) + 4u : 0u;
            uint2 c;
            c.x = nopen[0] | (nopen[1] << 8) | (nopen[2] << 16) | (nopen[3] << 24);
            c.y = nopen[4] | (nopen[5] << 8) | (nopen[6] << 16) | (nopen[7] << 24);
            *reinterpret_cast<uint2*>(cnt + obase) = c;
        }
    }
}


// ------------------------------------------------------------------ mean over the frequency axis (+dropout)
template <typename T>
__global__ void freq_mean_fwd_kernel(const T* __restrict__ x, T* __restrict__ out, long rows, int Wf,
                                     int C, uint64_t seed, const uint64_t* __restrict__ seed_dev, uint32_t thresh, float keep_scale) {
    if (seed_dev != nullptr) seed += *seed_dev;
    const int CV = C / 4;
    const long n = rows * CV;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % CV);
        const long r = i / CV;
        float acc[4] = {0, 0, 0, 0};
        for (int w = 0; w < Wf; ++w) {
            float v[4];
            Vec4<T>::load(x + (r * Wf + w) * C + cv * 4, v);
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] += v[k];
        }
        const long obase = i * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            acc[k] *= 1.0f / Wf;
            if (thresh != 0u) acc[k] *= tag_dropout_scale(seed, (uint64_t)(obase + k), thresh, keep_scale);
        }
        Vec4<T>::store(out + obase, acc);
    }
}

// POOLRED: dx is the gradient of block 4's pooled output; with that output p and its open-gate codes cnt (see
// bn_relu_pool_fwd_kernel) the two BatchNorm-backward sums of the block are accumulated on the way:
// red[c] += sum dx * cnt, red[C + c] += sum dx * p  (activation domain, converted by tag_bn_red_act_to_xhat) — no pass
// over the block's full-resolution BatchNorm input.  A thread keeps its 4 channels for the whole grid-stride loop.
template <typename TI, typename TO, bool POOLRED>
__global__ void freq_mean_bwd_kernel(const TI* __restrict__ dm, TO* __restrict__ dx, long rows, int Wf,
                                     int C, uint64_t seed, const uint64_t* __restrict__ seed_dev, uint32_t thresh, float keep_scale,
                                     const TO* __restrict__ pool_out, const uint8_t* __restrict__ pool_cnt,
                                     double* __restrict__ red) {
    if (seed_dev != nullptr) seed += *seed_dev;
    const int CV = C / 4;
    const long n = rows * CV;
    float rs[4] = {0.f, 0.f, 0.f, 0.f}, rq[4] = {0.f, 0.f, 0.f, 0.f};
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % CV);
        const long r = i / CV;
        float g[4];
        Vec4<TI>::load(dm + i * 4, g);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (thresh != 0u) g[k] *= tag_dropout_scale(seed, (uint64_t)(i * 4 + k), thresh, keep_scale);
            g[k] *= 1.0f / Wf;
        }
        for (int w = 0; w < Wf; ++w) {
            const long o = (r * Wf + w) * C + cv * 4;
            Vec4<TO>::store(dx + o, g);
            if (POOLRED) {
                float p[4];
                Vec4<TO>::load(pool_out + o, p);
                const uint32_t cw = *reinterpret_cast<const uint32_t*>(pool_cnt + o);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float v = round_to<TO>(g[k]);                 // the value as stored
                    rs[k] = fmaf(v, (float)((cw >> (8 * k)) & 0xFFu), rs[k]);
                    rq[k] = fmaf(v, p[k], rq[k]);
                }
            }
        }
    }
    if (POOLRED) {
        extern __shared__ float s_fr[];                                 // [2][C]
        for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) s_fr[c] = 0.f;
        __syncthreads();
        const int cv = (int)((blockIdx.x * (long)blockDim.x + threadIdx.x) % CV);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            atomicAdd(&s_fr[cv * 4 + k], rs[k]);
            atomicAdd(&s_fr[C + cv * 4 + k], rq[k]);
        }
        __syncthreads();
        for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) atomicAdd(red + c, (double)s_fr[c]);
    }
}

__global__ void dropout_mask_kernel(float* __restrict__ mask, long n, uint64_t seed, const uint64_t* __restrict__ seed_dev, uint32_t thresh,
                                    float keep_scale) {
    if (seed_dev != nullptr) seed += *seed_dev;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        mask[i] = thresh != 0u ? tag_dropout_scale(seed, (uint64_t)i, thresh, keep_scale) : 1.0f;
}

// out[c] (+)= sum_r x[r, c] over rows; x is [rows, C] of type T (bias gradients)
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, long rows, int C, float* __restrict__ out,
                              long rows_per_block) {
    const long r0 = (long)blockIdx.y * rows_per_block;
    const long r1 = min(rows, r0 + rows_per_block);
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (long r = r0; r < r1; ++r) s += to_f<T>(x[r * C + c]);
    atomicAdd(out + c, s);
}

// red[c] += sum_r dx[r,c]; red[C+c] += sum_r dx[r,c] * (x[r,c] - mean[c]) * invstd[c]   (bn0 backward)
__global__ void bn_bwd_reduce_f32_kernel(const float* __restrict__ dx, const float* __restrict__ x,
                                         const float* __restrict__ mean, const float* __restrict__ invstd,
                                         long rows, int C, double* __restrict__ red) {
    const int c = threadIdx.x % C;
    const int lane_rows = blockDim.x / C;
    const int r0 = threadIdx.x / C;
    const float mu = mean[c], is = invstd[c];
    float s = 0.f, q = 0.f;
    for (long r = (long)blockIdx.x * lane_rows + r0; r < rows; r += (long)gridDim.x * lane_rows) {
        const float g = dx[r * C + c];
        s += g;
        q += g * (x[r * C + c] - mu) * is;
    }
    __shared__ float ss[256], sq[256];
    ss[threadIdx.x] = s;
    sq[threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.x < C) {
        double ds = 0.0, dq = 0.0;
        for (int i = 0; i < lane_rows; ++i) { ds += ss[i * C + c]; dq += sq[i * C + c]; }
        atomicAdd(red + c, ds);
        atomicAdd(red + C + c, dq);
    }
}

// dbeta[c] += red[c]; dgamma[c] += red[C + c]
__global__ void bn_param_grads_kernel(const double* __restrict__ red, int C, float* __restrict__ dgamma,
                                      float* __restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (dbeta != nullptr) dbeta[c] += (float)red[c];
    if (dgamma != nullptr) dgamma[c] += (float)red[C + c];
}

// out = dy where act > 0 else 0 (ReLU backward from the saved activation)
template <typename TD, typename TA, typename TO>
__global__ void relu_bwd_kernel(const TD* __restrict__ dy, const TA* __restrict__ act, TO* __restrict__ out,
                                long n_vec) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n_vec; i += (long)gridDim.x * blockDim.x) {
        float g[4], a[4];
        Vec4<TD>::load(dy + i * 4, g);
        Vec4<TA>::load(act + i * 4, a);
#pragma unroll
        for (int k = 0; k < 4; ++k) g[k] = a[k] > 0.f ? g[k] : 0.f;
        Vec4<TO>::store(out + i * 4, g);
    }
}

inline int grid_for(long n, int threads, int max_blocks = 148 * 16) {
    long b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (int)b;
}

inline void dropout_params(float p, uint32_t* thresh, float* keep_scale) {
    if (p <= 0.f) { *thresh = 0u; *keep_scale = 1.f; return; }
    double t = (double)p * 4294967296.0;
    if (t > 4294967295.0) t = 4294967295.0;
    *thresh = (uint32_t)t;
    *keep_scale = 1.0f / (1.0f - p);
}

}  // namespace

extern "C" int tag_channel_stats_f32(const float* x, long rows, int C, double* stats, cudaStream_t stream) {
    if (C <= 0 || 256 % C != 0) return TAG_ERR_BAD_ARG;
    const int lane_rows = 256 / C;
    int blocks = grid_for(rows / lane_rows + 1, 32, 148 * 8);
    channel_stats_f32_kernel<<<blocks, 256, 0, stream>>>(x, rows, C, stats);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_bn_finalize(const double* stats, double count, int C, const float* gamma,
                               const float* beta, float* running_mean, float* running_var,
                               float momentum, float eps, int training, int update_running,
                               float* scale, float* shift, float* save_mean, float* save_invstd,
                               cudaStream_t stream) {
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(stats, count, C, gamma, beta, running_mean,
                                                            running_var, momentum, eps, training,
                                                            update_running, scale, shift, save_mean,
                                                            save_invstd);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_scale_shift_act(const void* x, int x_dtype, void* y, int y_dtype, const float* scale,
                                   const float* shift, long n, int C, int relu, cudaStream_t stream) {
    if (n % 8 != 0 || C % 8 != 0) return TAG_ERR_BAD_ARG;
    const long nv = n / 8;
    const int blocks = grid_for((nv + 1) / 2, 256, 148 * 24);
    if (x_dtype == TAG_DTYPE_F32 && y_dtype == TAG_DTYPE_F32)
        scale_shift_act_kernel<float, float><<<blocks, 256, 0, stream>>>((const float*)x, (float*)y, scale, shift, nv, C, relu);
    else if (x_dtype == TAG_DTYPE_F32 && y_dtype == TAG_DTYPE_BF16)
        scale_shift_act_kernel<float, bf16><<<blocks, 256, 0, stream>>>((const float*)x, (bf16*)y, scale, shift, nv, C, relu);
    else if (x_dtype == TAG_DTYPE_BF16 && y_dtype == TAG_DTYPE_BF16)
        scale_shift_act_kernel<bf16, bf16><<<blocks, 256, 0, stream>>>((const bf16*)x, (bf16*)y, scale, shift, nv, C, relu);
    else if (x_dtype == TAG_DTYPE_BF16 && y_dtype == TAG_DTYPE_F32)
        scale_shift_act_kernel<bf16, float><<<blocks, 256, 0, stream>>>((const bf16*)x, (float*)y, scale, shift, nv, C, relu);
    else
        return TAG_ERR_BAD_ARG;
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

template <typename T>
static int pool_fwd_dispatch(const void* y, void* out, uint8_t* cnt, const float* scale, const float* shift, int B,
                             int H, int W, int C, int ph, int pw, uint64_t seed, const uint64_t* __restrict__ seed_dev, uint32_t thresh,
                             float ks, cudaStream_t stream) {
    const long n = (long)B * (H / ph) * (W / pw) * (C / 8);
    const int blocks = grid_for(n, 256);
#define TAG_POOL_FWD(PH_, PW_)                                                                                              \
    do {                                                                                                                    \
        if (cnt != nullptr)                                                                                                 \
            bn_relu_pool_fwd_kernel<T, PH_, PW_, true><<<blocks, 256, 0, stream>>>((const T*)y, (T*)out, cnt, scale, shift, B, \
                                                                                   H, W, C, seed, seed_dev, thresh, ks);     \
        else                                                                                                                \
            bn_relu_pool_fwd_kernel<T, PH_, PW_, false><<<blocks, 256, 0, stream>>>((const T*)y, (T*)out, cnt, scale, shift, \
                                                                                    B, H, W, C, seed, seed_dev, thresh, ks);  \
    } while (0)
    if (ph == 2 && pw == 2) TAG_POOL_FWD(2, 2);
    else if (ph == 1 && pw == 2) TAG_POOL_FWD(1, 2);
#undef TAG_POOL_FWD
    else
        return TAG_ERR_UNSUPPORTED;
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_bn_relu_pool_fwd(const void* y, void* out, void* cnt, int dtype, const float* scale,
                                    const float* shift, int B, int H, int W, int C, int ph, int pw,
                                    float dropout_p, uint64_t seed, const uint64_t* seed_dev, cudaStream_t stream) {
    if (C % 8 != 0) return TAG_ERR_BAD_ARG;
    uint32_t thresh; float ks;
    dropout_params(dropout_p, &thresh, &ks);
    if (dtype == TAG_DTYPE_F32) return pool_fwd_dispatch<float>(y, out, (uint8_t*)cnt, scale, shift, B, H, W, C, ph, pw, seed, seed_dev, thresh, ks, stream);
    return pool_fwd_dispatch<bf16>(y, out, (uint8_t*)cnt, scale, shift, B, H, W, C, ph, pw, seed, seed_dev, thresh, ks, stream);
}


extern "C" int tag_freq_mean_fwd(const void* x, void* out, int dtype, long rows, int Wf, int C,
                                 float dropout_p, uint64_t seed, const uint64_t* seed_dev, cudaStream_t stream) {
    if (C % 4 != 0) return TAG_ERR_BAD_ARG;
    uint32_t thresh; float ks;
    dropout_params(dropout_p, &thresh, &ks);
    const int blocks = grid_for(rows * (C / 4), 256);
    if (dtype == TAG_DTYPE_F32)
        freq_mean_fwd_kernel<float><<<blocks, 256, 0, stream>>>((const float*)x, (float*)out, rows, Wf, C, seed, seed_dev, thresh, ks);
    else
        freq_mean_fwd_kernel<bf16><<<blocks, 256, 0, stream>>>((const bf16*)x, (bf16*)out, rows, Wf, C, seed, seed_dev, thresh, ks);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_freq_mean_bwd(const void* dm, int dm_dtype, void* dx, int dx_dtype, long rows, int Wf,
                                 int C, float dropout_p, uint64_t seed, const uint64_t* seed_dev, const void* pool_out,
                                 const void* pool_cnt, double* red, cudaStream_t stream) {
    if (C % 4 != 0) return TAG_ERR_BAD_ARG;
    uint32_t thresh; float ks;
    dropout_params(dropout_p, &thresh, &ks);
    const int CV = C / 4;
    if (pool_cnt != nullptr) {
        // fused BatchNorm-backward sums of the pooled block: bf16 gradient, channels fixed per thread (stride % CV == 0)
        if (dx_dtype != TAG_DTYPE_BF16 || pool_out == nullptr || red == nullptr || 256 % CV != 0) return TAG_ERR_BAD_ARG;
        const int blocks = grid_for(rows * CV, 256, 148 * 4);
        const size_t smem = (size_t)2 * C * sizeof(float);
        if (dm_dtype == TAG_DTYPE_F32)
            freq_mean_bwd_kernel<float, bf16, true><<<blocks, 256, smem, stream>>>(
                (const float*)dm, (bf16*)dx, rows, Wf, C, seed, seed_dev, thresh, ks, (const bf16*)pool_out,
                (const uint8_t*)pool_cnt, red);
        else if (dm_dtype == TAG_DTYPE_BF16)
            freq_mean_bwd_kernel<bf16, bf16, true><<<blocks, 256, smem, stream>>>(
                (const bf16*)dm, (bf16*)dx, rows, Wf, C, seed, seed_dev, thresh, ks, (const bf16*)pool_out,
                (const uint8_t*)pool_cnt, red);
        else
            return TAG_ERR_BAD_ARG;
        TAG_RETURN_IF_LAUNCH_FAILED();
        return TAG_OK;
    }
    const int blocks = grid_for(rows * CV, 256);
    if (dm_dtype == TAG_DTYPE_F32 && dx_dtype == TAG_DTYPE_F32)
        freq_mean_bwd_kernel<float, float, false><<<blocks, 256, 0, stream>>>((const float*)dm, (float*)dx, rows, Wf, C, seed, seed_dev, thresh, ks, nullptr, nullptr, nullptr);
    else if (dm_dtype == TAG_DTYPE_F32 && dx_dtype == TAG_DTYPE_BF16)
        freq_mean_bwd_kernel<float, bf16, false><<<blocks, 256, 0, stream>>>((const float*)dm, (bf16*)dx, rows, Wf, C, seed, seed_dev, thresh, ks, nullptr, nullptr, nullptr);
    else if (dm_dtype == TAG_DTYPE_BF16 && dx_dtype == TAG_DTYPE_BF16)
        freq_mean_bwd_kernel<bf16, bf16, false><<<blocks, 256, 0, stream>>>((const bf16*)dm, (bf16*)dx, rows, Wf, C, seed, seed_dev, thresh, ks, nullptr, nullptr, nullptr);
    else
        return TAG_ERR_BAD_ARG;
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_dropout_mask(float* mask, long n, float dropout_p, uint64_t seed, const uint64_t* seed_dev, cudaStream_t stream) {
    uint32_t thresh; float ks;
    dropout_params(dropout_p, &thresh, &ks);
    dropout_mask_kernel<<<grid_for(n, 256), 256, 0, stream>>>(mask, n, seed, seed_dev, thresh, ks);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_colsum(const void* x, int dtype, long rows, int C, float* out, cudaStream_t stream) {
    long rpb = (rows + 147) / 148;
    if (rpb < 64) rpb = 64;
    dim3 grid((C + 127) / 128, (unsigned)((rows + rpb - 1) / rpb));
    if (dtype == TAG_DTYPE_F32)
        colsum_kernel<float><<<grid, 128, 0, stream>>>((const float*)x, rows, C, out, rpb);
    else
        colsum_kernel<bf16><<<grid, 128, 0, stream>>>((const bf16*)x, rows, C, out, rpb);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_bn_bwd_reduce_f32(const float* dx, const float* x, const float* mean,
                                     const float* invstd, long rows, int C, double* red,
                                     cudaStream_t stream) {
    if (C <= 0 || 256 % C != 0) return TAG_ERR_BAD_ARG;
    const int lane_rows = 256 / C;
    int blocks = grid_for(rows / lane_rows + 1, 32, 148 * 8);
    bn_bwd_reduce_f32_kernel<<<blocks, 256, 0, stream>>>(dx, x, mean, invstd, rows, C, red);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_bn_param_grads(const double* red, int C, float* dgamma, float* dbeta,
                                  cudaStream_t stream) {
    bn_param_grads_kernel<<<(C + 127) / 128, 128, 0, stream>>>(red, C, dgamma, dbeta);
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}

extern "C" int tag_relu_bwd(const void* dy, int dy_dtype, const void* act, int act_dtype, void* out,
                            int out_dtype, long n, cudaStream_t stream) {
    if (n % 4 != 0) return TAG_ERR_BAD_ARG;
    const long nv = n / 4;
    const int blocks = grid_for(nv, 256);
    if (dy_dtype == TAG_DTYPE_F32 && act_dtype == TAG_DTYPE_F32 && out_dtype == TAG_DTYPE_F32)
        relu_bwd_kernel<float, float, float><<<blocks, 256, 0, stream>>>((const float*)dy, (const float*)act, (float*)out, nv);
    else if (dy_dtype == TAG_DTYPE_F32 && act_dtype == TAG_DTYPE_BF16 && out_dtype == TAG_DTYPE_BF16)
        relu_bwd_kernel<float, bf16, bf16><<<blocks, 256, 0, stream>>>((const float*)dy, (const bf16*)act, (bf16*)out, nv);
    else if (dy_dtype == TAG_DTYPE_F32 && act_dtype == TAG_DTYPE_BF16 && out_dtype == TAG_DTYPE_F32)
        relu_bwd_kernel<float, bf16, float><<<blocks, 256, 0, stream>>>((const float*)dy, (const bf16*)act, (float*)out, nv);
    else
        return TAG_ERR_BAD_ARG;
    TAG_RETURN_IF_LAUNCH_FAILED();
    return TAG_OK;
}
