/* tag_b200.h — C ABI of libtag_b200.so: the B200 (sm_100a) kernels of the cnn8rnn-w2vmean hot path.
 *
 * The reference (wsntxxn/TextToAudioGrounding) has no FFI: its "operator API" for this path is
 * the set of torch / torchaudio library calls listed below.  Every entry point replaces one of
 * those call sites (file:line relative to the reference root) and is bound from Python with
 * ctypes by texttoaudiogrounding_b200/_lib.py (see INTEGRATION.md for the stub a reference
 * maintainer would add).
 *
 * Conventions: every pointer is a DEVICE pointer unless stated; the caller owns all buffers
 * (no allocation inside); calls are stream-ordered on `stream` with no internal
 * synchronisation and are re-entrant; the return value is 0 on success, a cudaError_t value
 * for a failed launch, or TAG_ERR_* (>= 10001) for a rejected argument.  Activations are NHWC;
 * dtype codes: 0 = float32, 1 = bfloat16.  Conv weights are [Cout][tap][Cin] fp32, i.e. the
 * memory of a channels_last [Cout,Cin,3,3] tensor (tap = kh*3+kw).
 */
#ifndef TAG_B200_H
#define TAG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define TAG_DTYPE_F32 0
#define TAG_DTYPE_BF16 1
#define TAG_ERR_BAD_ARG 10001
#define TAG_ERR_UNSUPPORTED 10002

int tag_version(void);

/* ---- log-mel frontend -------------------------------------------------------------------
 * torchaudio MelSpectrogram(n_fft=win=1024, hop=320, hann, center, reflect, power=2, slaney
 * fb) + AmplitudeToDB — models/audio_encoder.py:113-124 (construction), :183-184 (call).
 * wav [batch, n_samples] (row stride wav_stride) -> db_out [batch, T0 = n_samples/320+1, 64].
 * fb is the [513,64] filterbank buffer, window the 1024 Hann buffer, mel_range (optional,
 * int[64][2]) the [lo,hi) bin support of each mel.  stats (optional, double[128], pre-zeroed)
 * receives per-mel sum and sum of squares for bn0 (models/audio_encoder.py:188-190). */
int tag_logmel_fwd(const float* wav, int batch, int n_samples, long wav_stride, const float* window,
                   const float* fb, const int* mel_range, float* db_out, double* stats,
                   cudaStream_t stream);
/* Same contract for a COMPACT filterbank (mel_range required; fb_nnz = sum over mels of hi - lo, <= 2048 — the
 * slaney triangles need ~1030): one warp per pair of frames, 1024-point FFT as two register-resident 32-point passes.
 * wav_dtype: 0 = float32, 2 = float16 (the reference stores waveforms as float16, utils/data/pack_waveform.py:46-52;
 * they are widened on load, as `.float()` does in datasets/single_phrase_dataset.py). */
int tag_logmel_fwd_v2(const void* wav, int wav_dtype, int batch, int n_samples, long wav_stride,
                      const float* window, const float* fb, const int* mel_range, int fb_nnz, float* db_out,
                      double* stats, cudaStream_t stream);

/* ---- BatchNorm2d pieces — models/audio_encoder.py:133,188-190; models/panns.py:35-36,49-50 */
int tag_channel_stats_f32(const float* x, long rows, int C, double* stats, cudaStream_t stream);
int tag_bn_finalize(const double* stats, double count, int C, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float momentum, float eps, int training,
                    int update_running, float* scale, float* shift, float* save_mean,
                    float* save_invstd, cudaStream_t stream);
int tag_scale_shift_act(const void* x, int x_dtype, void* y, int y_dtype, const float* scale,
                        const float* shift, long n, int C, int relu, cudaStream_t stream);

/* ---- convolutions / dense contractions --------------------------------------------------
 * F.conv2d(3x3, pad 1, no bias) — models/panns.py:25-33,49-50; with taps=1 the x @ W^T (+bias,
 * relu) of fc1 and the GRU input projection — models/audio_encoder.py:140-141,216-217.
 * tag_conv_fwd also computes dgrad when given dy and tag_weight_flip_transpose'd weights.
 * stats (optional, double[2*Cout], pre-zeroed): per-channel sum / sum of squares of y. */
int tag_conv_c1_fwd(const void* x, const float* w, void* y, int dtype, double* stats, int B, int H,
                    int W, cudaStream_t stream);
int tag_conv_c1_bwd(const void* dy, const void* x, const float* w, int dtype, float* dw, float* dx,
                    int B, int H, int W, cudaStream_t stream);
/* conv_block1.conv1 + bn1 + ReLU as ONE pass (models/panns.py:49 for Cin = 1, audio_encoder.py:202): the BatchNorm batch
 * statistics of a Cin = 1 convolution follow from 54 second-order moments of its 16 MB input (mom: double[54], written
 * here), so the raw convolution output is never stored:
 *   tag_c1_moments -> tag_c1_stats_from_moments (stats: double[128] = sum y | sum y^2, the tag_bn_finalize input)
 *   -> tag_bn_finalize -> tag_conv_c1_fwd_act (y = relu(scale * conv(x, w) + shift)).
 * Backward: the fused reduce of tag_conv_tc_fwd_halo reads the saved ACTIVATION (bn_act) and tag_bn_red_act_to_xhat
 * converts its sums in place (red[c] *= sum_scale; red[C + c] = sum g * a -> sum g * xhat = (sum g * a - beta * sum g) / gamma);
 * tag_conv_c1_bwd_bn then takes the gated gradient g, recomputes conv(x, w), applies the BatchNorm backward (red = those
 * double[128]) and produces dw / dx — autograd of conv2d + batch_norm + relu without the intermediate tensors. */
int tag_c1_moments(const void* x, int dtype, int B, int H, int W, double* mom, cudaStream_t stream);
int tag_c1_stats_from_moments(const double* mom, const float* w, double* stats, cudaStream_t stream);
int tag_conv_c1_fwd_act(const void* x, const float* w, const float* scale, const float* shift, void* y, int dtype,
                        int B, int H, int W, cudaStream_t stream);
int tag_bn_red_act_to_xhat(double* red, const float* gamma, const float* beta, int C, float sum_scale,
                           float* dgamma, float* dbeta, cudaStream_t stream);   /* dgamma / dbeta (optional): += */
int tag_conv_c1_bwd_bn(const void* g, const void* x, const float* w, const float* scale, const float* mean,
                       const float* invstd, const double* red, int bn_training, float* dw, float* dx, int B,
                       int H, int W, cudaStream_t stream);
int tag_conv_fwd(const void* x, int x_dtype, const float* w, void* y, int y_dtype, const float* bias,
                 int relu, double* stats, int B, int H, int W, int Cin, int Cout, int taps,
                 cudaStream_t stream);
int tag_conv_wgrad(const void* dy, int dy_dtype, const void* x, int x_dtype, float* dw, int B, int H,
                   int W, int Cin, int Cout, int taps, int splits, cudaStream_t stream);
int tag_weight_flip_transpose(const float* w, float* wt, int Co, int Ci, int taps, cudaStream_t stream);

/* bf16 tensor-core (tcgen05 + TMA) versions of the two contractions above.  x / dy are bf16 NHWC,
 * w is bf16 [Cout][taps*Cin] (tag_cast_f32_to_bf16 of the fp32 master, or
 * tag_weight_flip_transpose_bf16 for dgrad), y is bf16 or fp32, dw is fp32 (accumulated).
 * Requirements: Cin, Cout multiples of 64; W in {1,2,4,...,64}. */
int tag_conv_tc_fwd(const void* x, const void* w, void* y, int y_dtype, const float* bias, int relu,
                    double* stats, int B, int H, int W, int Cin, int Cout, int taps, cudaStream_t stream);
int tag_conv_tc_wgrad(const void* dy, const void* x, float* dw, int B, int H, int W, int Cin, int Cout,
                      int taps, int splits, cudaStream_t stream);
/* 3x3, Cin = 64 only: all nine taps accumulate in one CTA's TMEM (x boxes shared by the vertical taps). */
int tag_conv_tc_wgrad64(const void* dy, const void* x, float* dw, int B, int H, int W, int Cout, int splits,
                        cudaStream_t stream);
/* 3x3 only: same result as tag_conv_tc_fwd(taps=9, no bias/relu) with the input halo tile re-used
 * across the three vertical taps (16x8-pixel output tiles; W must be a multiple of 8).
 * Optional fused BatchNorm-backward reductions in the epilogue (dgrad use; `stats` receives two sums per channel):
 *   bn_act only        : this dgrad's output is d(relu(bn(.))), bn_act = that layer's SAVED ACTIVATION a (bf16 NHWC
 *                        [B,H,W,Cout]): the output is gated by a > 0; stats = sum g | sum g * a.
 *   bn_act + pool_cnt  : this dgrad's output is d(dropout(pool(relu(bn(.))))), bn_act = the saved POOLED output p of that
 *                        layer, pool_cnt = its open-gate code (uint8, tag_bn_relu_pool_fwd): stats = sum dout * cnt |
 *                        sum dout * p, i.e. the same two sums for the pooled layer (4 sum g / keep_scale | sum g * a)
 *                        without a pass over its full-resolution input.
 * tag_bn_red_act_to_xhat turns either pair into (dbeta, dgamma) (sum_scale = 1, or keep_scale / 4 for the pooled form). */
int tag_conv_tc_fwd_halo(const void* x, const void* w, void* y, int y_dtype, double* stats, int B, int H,
                         int W, int Cin, int Cout, const void* bn_act, const void* pool_cnt, cudaStream_t stream);
/* Scheduling knob of tag_conv_tc_fwd_halo (same results either way): 1 (default) = layers whose weights stream
 * (Cin >= 128) run on CTA PAIRS — a 2-CTA cluster computes a 256-pixel tile with cta_group::2 tcgen05 MMAs, each CTA
 * loading half of every weight tile; 0 = one CTA per 128-pixel tile everywhere. */
int tag_conv_halo_set_pair_mode(int mode);
/* Scheduling knob of every persistent tensor-core kernel (conv forward/dgrad/wgrad): launch them on `sms` fewer SMs
 * (0 .. 64, default 0) from now on.  The data-parallel train step sets it while the gradient all-reduce runs beside the
 * last part of its backward pass (a collective's CTAs cannot share an SM with these kernels' CTAs) and resets it to 0.
 * Same results either way.  No reference counterpart (torch DDP overlaps its buckets the same way, run_strong.py has
 * no data-parallel path). */
int tag_set_sm_reserve(int sms);
/* Deterministic split-K for the tensor-core weight-gradient kernels (tag_conv_tc_wgrad, tag_conv_tc_wgrad64): with a
 * workspace set (device memory, 16-byte aligned; 64 MB covers every layer of cnn8rnn at bs=64), each split stores its
 * partial tile into its own slab and a second pass adds the slabs to dw in split order, so the weight gradients are
 * bit-identical from run to run (torch.use_deterministic_algorithms for the reference's cudnn wgrad).  ws = NULL,
 * bytes = 0 (default) restores fp32 atomics into dw.  A call whose slabs do not fit returns 10002.  The workspace is
 * shared: issue weight gradients on ONE stream while it is set. */
int tag_set_splitk_workspace(void* ws, long bytes);
/* its weight operand: bf16 tap-major [9][Cout][Cin] (flip_transpose=0) or, for dgrad,
 * [9][Cin][Cout] of the 180-degree rotated kernel (flip_transpose=1), from the fp32 master. */
int tag_weight_prep_tapmajor_bf16(const float* w, void* out, int Co, int Ci, int flip_transpose,
                                  cudaStream_t stream);
/* fp32-accurate convolution on the bf16 tensor cores: weights as [9][Cout][3*Cin] = [hi | lo | hi] for activations split
 * with tag_split_bf16x3 mode 0 ([.., 3*Cin] = [hi | hi | lo]); run tag_conv_tc_fwd_halo with Cin' = 3*Cin, fp32 output */
int tag_weight_prep_tapmajor_x3(const float* w, void* out, int Co, int Ci, cudaStream_t stream);
int tag_weight_flip_transpose_bf16(const float* w, void* wt, int Co, int Ci, int taps, cudaStream_t stream);

/* ---- BN + ReLU + avg+max pool + dropout — models/panns.py:50-58, audio_encoder.py:202-211 */
/* cnt (optional, uint8 [B, H/ph, W/pw, C]): per output element 4 * (n / (ph pw) + [n > 0]), n = the number of window
 * elements with an open ReLU gate, 0 if the element was dropped — with `out` itself all that the backward reductions of
 * the layer need (see tag_conv_tc_fwd_halo, pool_cnt). */
int tag_bn_relu_pool_fwd(const void* y, void* out, void* cnt, int dtype, const float* scale, const float* shift,
                         int B, int H, int W, int C, int ph, int pw, float dropout_p, uint64_t seed,
                         const uint64_t* seed_dev, cudaStream_t stream);
int tag_bn_relu_pool_bwd(int mode, const void* y, const void* dout, void* dy, int dtype,
                         const float* scale, const float* shift, const float* mean,
                         const float* invstd, double* red, int bn_training, int B, int H, int W, int C,
                         int ph, int pw, float dropout_p, uint64_t seed, const uint64_t* seed_dev,
                         cudaStream_t stream);
/* torch.mean(x, dim=3) + transpose + dropout(0.5) — models/audio_encoder.py:212-215 */
int tag_freq_mean_fwd(const void* x, void* out, int dtype, long rows, int Wf, int C, float dropout_p,
                      uint64_t seed, const uint64_t* seed_dev, cudaStream_t stream);
/* pool_out / pool_cnt / red (optional, bf16 dx): dx is the gradient of a pooled block output; accumulate its BatchNorm-backward
 * sums in the activation domain on the way (red[c] += sum dx * cnt, red[C + c] += sum dx * pool_out), as the pooled form of
 * tag_conv_tc_fwd_halo does for the blocks whose gradient comes from a convolution. */
int tag_freq_mean_bwd(const void* dm, int dm_dtype, void* dx, int dx_dtype, long rows, int Wf, int C,
                      float dropout_p, uint64_t seed, const uint64_t* seed_dev, const void* pool_out,
                      const void* pool_cnt, double* red, cudaStream_t stream);
int tag_dropout_mask(float* mask, long n, float dropout_p, uint64_t seed, const uint64_t* seed_dev,
                         cudaStream_t stream);
int tag_colsum(const void* x, int dtype, long rows, int C, float* out, cudaStream_t stream);
int tag_bn_bwd_reduce_f32(const float* dx, const float* x, const float* mean, const float* invstd,
                          long rows, int C, double* red, cudaStream_t stream);
int tag_bn_param_grads(const double* red, int C, float* dgamma, float* dbeta, cudaStream_t stream);
int tag_relu_bwd(const void* dy, int dy_dtype, const void* act, int act_dtype, void* out, int out_dtype,
                 long n, cudaStream_t stream);

/* ---- BiGRU recurrence — nn.GRU(512,256,bidirectional) models/audio_encoder.py:141,217 ---- */
int tag_gru_fwd(const float* gi, const float* w_hh, const float* b_hh, float* out, float* gates, int B,
                int T, cudaStream_t stream);
int tag_gru_bwd(const float* d_out, const float* out, const float* gates, const float* w_hh, float* dgi,
                float* dgh, float* hprev, int B, int T, cudaStream_t stream);
/* bf16 compute mode: W_hh slice resident in registers as mma.sync fragments, bf16 state exchange;
 * same arguments (weights and state stay fp32 in HBM). */
int tag_gru_fwd_bf16(const float* gi, const float* w_hh, const float* b_hh, float* out, float* gates,
                     int B, int T, cudaStream_t stream);
int tag_gru_bwd_bf16(const float* d_out, const float* out, const float* gates, const float* w_hh,
                     void* dgi, void* dgh, void* hprev, int B, int T, cudaStream_t stream);   /* bf16 outputs */

/* ---- text encoder + match + loss — models/text_encoder.py:39-43,79-88; models/utils.py:33-58;
 * models/match.py:43-60; losses.py:12-24 (host pointers: none) */
int tag_embed_mean_fwd(const long long* text, const long long* text_len, const float* emb,
                       float* token_emb, float* seq_emb, int B, int N, int D, int vocab,
                       cudaStream_t stream);
int tag_embed_mean_bwd(const long long* text, const long long* text_len, const float* d_seq,
                       float* d_emb, int B, int N, int D, int vocab, cudaStream_t stream);
int tag_dot_sigmoid_fwd(const float* audio, const float* seq, float* sim, float* logits, int B, int T,
                        int D, float scale, cudaStream_t stream);
int tag_dot_sigmoid_bwd(const float* d_sim, const float* sim, const float* audio, const float* seq,
                        float* d_audio, float* d_seq, float* d_logit_ws, int B, int T, int D,
                        float scale, cudaStream_t stream);
int tag_frame_bce(const float* sim, long sim_stride, const float* label, long label_stride,
                  const long long* length, int B, int Tt, float* loss_out, float* d_sim,
                  long dsim_stride, float grad_scale, cudaStream_t stream);

/* Normalised match heads on seq-level text (audio [B,T,512], seq [B,512] -> sim [B,T]):
 * mode 1 DotProduct(l2norm=True) = cosine similarity -> scale -> sigmoid -> clamp (models/match.py:43-60);
 * mode 2 / 3 ExpNegL2 with / without l2norm: exp(-|a^ - s^|) (models/match.py:10-33).  F.normalize eps 1e-12.
 * bwd: d_audio overwritten, d_seq accumulated (pre-zeroed). */
int tag_match_norm_fwd(const float* audio, const float* seq, float* sim, int B, int T, int D, int mode, float scale,
                       cudaStream_t stream);
int tag_match_norm_bwd(const float* d_sim, const float* sim, const float* audio, const float* seq, float* d_audio,
                       float* d_seq, int B, int T, int D, int mode, float scale, cudaStream_t stream);

/* ---- multi-phrase (weakly supervised) head — SURVEY.md §8f rank 1.  MultiTextBiEncoder.forward
 * (models/audio_text_model.py:147-229) matches every clip against n phrases; the reference expands the audio
 * embedding to [B*n,T,D] and calls DotProduct (models/match.py:43-60) — here audio [B,T,D] is read in place:
 * sim[b,t,j] = clamp(sigmoid(scale * <audio[b,t,:], seq[b,j,:]>), 1e-7, 1).  D = 512, n <= 64. */
int tag_multi_dot_sigmoid_fwd(const float* audio, const float* seq, float* sim, int B, int T, int n, int D,
                              float scale, cudaStream_t stream);
/* d_logit_ws: workspace [B,T,n]; d_audio [B,T,D] and d_seq [B,n,D] are overwritten */
int tag_multi_dot_sigmoid_bwd(const float* d_sim, const float* sim, const float* audio, const float* seq,
                              float* d_audio, float* d_seq, float* d_logit_ws, int B, int T, int n, int D,
                              float scale, cudaStream_t stream);
/* clip[b,j] = pooling over t < length[b] of sim[b,t,j] — models/utils.py:33-95; mode 0 linear_softmax
 * (sum f^2 / sum f), 1 max, 2 mean, 3 exp_softmax (shifted by the max over ALL frames, as the reference) */
int tag_pool_with_lens_fwd(const float* sim, const long long* length, int mode, float* clip, int B, int T, int n,
                           cudaStream_t stream);
int tag_pool_with_lens_bwd(const float* d_clip, const float* sim, const float* clip, const long long* length,
                           int mode, float* d_sim, int B, int T, int n, cudaStream_t stream);

/* ---- sentence-level alignment — SURVEY.md §8f rank 3.  align.DotProduct (models/align.py:7-31) forms
 * sim[i,j,t,n] = clamp(sigmoid(scale * <audio[i,t,:], text[j,n,:]>), 1e-7, 1) for ALL (clip i, text j) pairs and
 * the sim_pooling classes (models/sim_pooling.py:6-204) reduce it over frames t < audio_len[i] and tokens
 * n < text_len[j].  The frame pooling is fused into the kernel that forms the dot products; the 4-D matrix is only
 * written when sim_matrix != NULL.  text is [Cpad, D] (the Bt*N token rows, zero padded to a multiple of 64 rows),
 * colpool / aux are [Ba, Cpad] (frame-pooled value per (clip, text row) and the term backward needs).
 * a_mode: 0 mean, 1 max, 2 linear_softmax, 3 exp_softmax;  t_mode: 0 mean, 1 sum, 2 max, 3 mean+sum.  D = 512. */
int tag_align_pool_fwd(const float* audio, const float* text, const long long* audio_len, int a_mode,
                       float* sim_matrix, float* colpool, float* aux, int Ba, int T, int Bt, int N, int Cpad,
                       int D, float scale, cudaStream_t stream);
/* G [Ba*T, Cpad] = scale * d(loss)/d(logit); d_audio = G x text and d_text = G^T x audio are then plain GEMMs
 * (tag_conv_fwd / tag_conv_wgrad with taps = 1) */
int tag_align_pool_bwd(const float* audio, const float* text, const long long* audio_len, int a_mode,
                       const float* d_colpool, const float* colpool, const float* aux, float* G, int Ba, int T,
                       int Cpad, int D, float scale, cudaStream_t stream);
int tag_align_text_pool_fwd(const float* colpool, const long long* text_len, int t_mode, float* out, int Ba, int Bt,
                            int N, int Cpad, cudaStream_t stream);
int tag_align_text_pool_bwd(const float* d_out, const float* colpool, const long long* text_len, int t_mode,
                            float* d_colpool, int Ba, int Bt, int N, int Cpad, cudaStream_t stream);
/* Tensor-core route of the same path for large batches: the logits [Ba*T, Cpad] (Cpad a multiple of 128) come from one
 * split-bf16 tcgen05 GEMM (tag_split_bf16x3 + tag_conv_tc_fwd); _fwd applies sigmoid / clamp / frame pooling, _bwd
 * writes the logit gradient as split-bf16 GEMM operands: g_kcat [Ba*T][3*Cpad] and g_planes [2][Ba*T][Cpad]. */
int tag_align_logits_fwd(const float* logits, const long long* audio_len, int a_mode, float* sim_matrix,
                         float* colpool, float* aux, int Ba, int T, int Bt, int N, int Cpad, float scale,
                         cudaStream_t stream);
int tag_align_logits_bwd(const float* logits, const long long* audio_len, int a_mode, const float* d_colpool,
                         const float* colpool, const float* aux, void* g_kcat, void* g_planes, int Ba, int T,
                         int Cpad, float scale, cudaStream_t stream);
/* fp32 -> split-bf16 GEMM operands (x = hi + lo, 16 mantissa bits): mode 0 -> [rows][hi|hi|lo] (3K per row),
 * mode 1 -> [rows][hi|lo|hi], mode 2 -> planes [2][rows][K].  transpose != 0: `in` is [K][rows].  A mode-0 operand times
 * a mode-1 operand in ONE bf16 tensor-core GEMM of depth 3K gives the fp32 product to ~2^-16 relative. */
int tag_split_bf16x3(const float* in, void* out, long rows, int K, int mode, int transpose, cudaStream_t stream);
/* MaxMarginRankingLoss (losses.py:226-264) on sim [n, n]: loss (scalar) and d(loss)/d(sim) in one launch */
int tag_max_margin_rank(const float* sim, int n, float margin, float lamda1, int fix_norm, float* loss,
                        float* d_sim, cudaStream_t stream);
/* gradient of the token embeddings (EmbeddingLayer.forward, models/text_encoder.py:39-43): d_emb[text[b,n]] += d_token[b,n,:] */
int tag_embed_token_bwd(const long long* text, const float* d_token, float* d_emb, int B, int N, int D, int vocab,
                        cudaStream_t stream);

/* EmbeddingAgg(aggregation="attention") = AttentionPooling (models/text_encoder.py:46-58, 84-85): score = x w + bias
 * masked with -1e10 beyond lens, softmax over the N tokens (N <= 128), out [B,D] = sum_n weight[n] x[b,n,:];
 * weight [B,N] is saved for backward.  bwd: d_x overwritten, d_w [D] and d_bias [1] accumulated. */
int tag_attn_pool_fwd(const float* x, const long long* lens, const float* w, const float* bias, float* out,
                      float* weight, int B, int N, int D, cudaStream_t stream);
int tag_attn_pool_bwd(const float* d_out, const float* x, const float* w, const float* weight, float* d_x,
                      float* d_w, float* d_bias, int B, int N, int D, cudaStream_t stream);
/* F.interpolate(size=To, mode="linear", align_corners=False) along T — BiEncoder / MultiTextBiEncoder(upsample=True),
 * models/audio_text_model.py:90-97, 216-223.  x [outer, T, inner] -> y [outer, To, inner]; bwd is its transpose
 * (dx is zeroed inside). */
int tag_upsample_linear_fwd(const float* x, float* y, long outer, int T, int To, int inner, cudaStream_t stream);
int tag_upsample_linear_bwd(const float* dy, float* dx, long outer, int T, int To, int inner, cudaStream_t stream);

/* ---- attention-type heads of the later configurations — SURVEY.md §8f rank 2 (BASELINE.json configs[3]).  All fp32,
 * E = 512; the dense projections around them are tag_conv_fwd / tag_conv_wgrad with taps = 1.
 * text_encoder.SelfAttention (models/text_encoder.py:240-268): out[b,0] = cls + pe[0], out[b,1+n] = emb[text[b,n]] +
 * pe[1+n], then dropout (PositionalEncoding, :128-146).  bwd accumulates into d_emb [vocab,E] and d_cls [E]. */
int tag_text_assemble_fwd(const long long* text, const float* emb, const float* cls, const float* pe, float* out,
                          int B, int N, int E, int vocab, float dropout_p, uint64_t seed, const uint64_t* seed_dev,
                          cudaStream_t stream);
int tag_text_assemble_bwd(const long long* text, const float* d_out, float* d_emb, float* d_cls, int B, int N, int E,
                          int vocab, float dropout_p, uint64_t seed, const uint64_t* seed_dev, cudaStream_t stream);
/* scaled-dot-product core of nn.MultiheadAttention (text_encoder.py:266, match.py:66-73,82): q [B,Lq,E], k/v [B,Lk,E]
 * already projected; keys n >= key_len[b] are masked (key_padding_mask; NULL = none); softmax; dropout on the
 * probabilities; out [B,Lq,E] (heads concatenated).  probs [B,heads,Lq,Lk] is kept for backward.  Lk <= 128,
 * E / heads in {32, 64, 128}. */
int tag_mha_core_fwd(const float* q, const float* k, const float* v, const long long* key_len, float* out,
                     float* probs, int B, int Lq, int Lk, int E, int heads, float dropout_p, uint64_t seed,
                     const uint64_t* seed_dev, cudaStream_t stream);
int tag_mha_core_bwd(const float* d_out, const float* q, const float* k, const float* v, const float* probs,
                     const long long* key_len, float* dq, float* dk, float* dv, int B, int Lq, int Lk, int E,
                     int heads, float dropout_p, uint64_t seed, const uint64_t* seed_dev, cudaStream_t stream);
/* cross_encoder.Seq2SeqAttention (models/cross_encoder.py:12-42) with h2attn split into its query / key halves:
 * hq = query W_q^T [B,T,E], hk = kv W_k^T + b [B,N,E]; score = sum_a v[a] tanh(hq + hk), -1e10 where t >= q_len[b] or
 * n >= kv_len[b], softmax over n, out = attn x kv.  N <= 32 (backward: N <= 24).  bwd: d_hq overwritten; d_hk, d_kv,
 * d_v accumulated (pre-zeroed). */
int tag_additive_attn_fwd(const float* hq, const float* hk, const float* v, const float* kv, const long long* q_len,
                          const long long* kv_len, float* attn, float* out, int B, int T, int N, int E,
                          cudaStream_t stream);
int tag_additive_attn_bwd(const float* d_out, const float* hq, const float* hk, const float* v, const float* kv,
                          const float* attn, const long long* q_len, const long long* kv_len, float* d_hq, float* d_hk,
                          float* d_v, float* d_kv, int B, int T, int N, int E, cudaStream_t stream);
/* cross_encoder.CrossGating (models/cross_encoder.py:45-57): out = x * sigmoid(z) */
int tag_sigmoid_gate_fwd(const float* x, const float* z, float* out, long n, cudaStream_t stream);
int tag_sigmoid_gate_bwd(const float* d_out, const float* x, const float* z, float* dx, float* dz, long n,
                         cudaStream_t stream);
/* match.DotProduct(text_level="token") on per-frame text embeddings (models/match.py:43-60): a, x [R,E] -> sim [R] */
int tag_rowdot_sigmoid_fwd(const float* a, const float* x, float* sim, long R, int E, float scale, cudaStream_t stream);
int tag_rowdot_sigmoid_bwd(const float* d_sim, const float* sim, const float* a, const float* x, float* da, float* dx,
                           long R, int E, float scale, cudaStream_t stream);
/* tail of match.CrossAttention.forward (models/match.py:84-88): sigmoid(Linear_{E->1}(LayerNorm(audio +
 * dropout(attn_out)))).  stat [R,2] keeps (mean, rstd).  bwd: d_audio, d_attn overwritten; d_gamma, d_beta, d_w [E],
 * d_bias [1] accumulated (pre-zeroed). */
int tag_ln_linear_sigmoid_fwd(const float* audio, const float* attn_out, const float* gamma, const float* beta,
                              const float* w, const float* bias, float* prob, float* stat, long R, int E, float eps,
                              float dropout_p, uint64_t seed, const uint64_t* seed_dev, cudaStream_t stream);
int tag_ln_linear_sigmoid_bwd(const float* d_prob, const float* prob, const float* audio, const float* attn_out,
                              const float* gamma, const float* beta, const float* w, const float* stat, float* d_audio,
                              float* d_attn, float* d_gamma, float* d_beta, float* d_w, float* d_bias, long R, int E,
                              float dropout_p, uint64_t seed, const uint64_t* seed_dev, cudaStream_t stream);

/* ---- CLAP text tower (inference; BASELINE.json configs[4]) — transformers ClapTextModel as wired by
 * LaionClapEncoder (models/text_encoder.py:311-327 == models/hf_modeling_grounding.py:183-199).  Row-wise pieces
 * only: the dense layers are bf16 tag_conv_tc_fwd GEMMs, the attention core is tag_mha_core_fwd.  E in {512, 768}.
 * tag_roberta_embed_ln: word[ids] + token_type[0] + position[(#non-pad ids up to l) + pad_idx] -> LayerNorm. */
int tag_roberta_embed_ln(const long long* ids, const float* word, const float* pos, const float* type0,
                         const float* gamma, const float* beta, float* out, int B, int L, int E, int vocab,
                         int max_pos, int pad_idx, float eps, cudaStream_t stream);
/* out = LayerNorm(a + res) (res may be NULL) */
int tag_add_layernorm(const float* a, const float* res, const float* gamma, const float* beta, float* out, long rows,
                      int E, float eps, cudaStream_t stream);
/* op 0: exact (erf) GELU, op 1: tanh */
int tag_unary_f32(const float* in, float* out, long n, int op, cudaStream_t stream);
/* F.normalize(x, dim=-1): rows / max(||row||_2, eps) */
int tag_l2_normalize(const float* in, float* out, long rows, int E, float eps, cudaStream_t stream);

/* ---- post-processing of frame probabilities (SURVEY.md §8f rank 4 iii) — Runner.eval_inference,
 * python_scripts/training/run_strong.py:222-247 with utils/eval_util.py:18-116: for every (sample, threshold) pair
 * binarize (sim > threshold, float64 compare), median-filter along time (window, scipy "reflect" boundary), merge
 * regions whose gap is <= n_connect frames, list the regions.  sim [B, T] (row stride sim_stride, floats),
 * thresholds double[n_th]; regions int[B][n_th][max_regions][2] = (onset, offset) frames, counts int[B][n_th] (a count
 * above max_regions means truncated output).  Bit-exact with the reference. */
int tag_frame_regions(const float* sim, long sim_stride, const double* thresholds, int B, int T, int n_th, int window,
                      int n_connect, int max_regions, int* regions, int* counts, cudaStream_t stream);

/* ---- optimizer step — clip_grad_norm_ + Adam, python_scripts/training/run_strong.py:143-145 */
int tag_sumsq(const float* g, long n, double* out, cudaStream_t stream);
/* lr_dev (device float, may be NULL) overrides lr: a scheduler (run_strong.py:136-137) then acts on a captured graph */
int tag_clip_adam(float* p, const float* g, float* m, float* v, long n, const double* sumsq,
                  long long* step_ptr, float grad_mult, float max_norm, float lr, const float* lr_dev, float beta1,
                  float beta2, float eps, float* norm_out, cudaStream_t stream);
int tag_cast_f32_to_bf16(const float* x, void* y, long n, cudaStream_t stream);
/* all bf16 GEMM operands of one step from the fp32 master weights in one launch; `table` (device) holds
 * 5 int64 per entry: src, dst, rows<<32|cols, taps<<32|mode, first block (2048 elements per block);
 * modes: 0 cast, 1 tap-major fwd, 2 tap-major dgrad (flip + transpose), 3 transpose. */
int tag_weight_prep_batch(const long long* table, int n_entries, int total_blocks, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TAG_B200_H */
