/* sos_b200.h -- C ABI of libsos_b200.so (sm_100a only).
 *
 * Drop-in boundary of the B200-native hot path for
 * henryxrl/Listening-to-Sound-of-Silence-for-Speech-Denoising.  The reference has no FFI of its own (it is
 * pure Python over PyTorch 1.3 + librosa 0.7.1); every entry point below replaces a library call the
 * reference makes at the cited file:line (paths relative to the reference root; M1 =
 * model_1_silent_interval_detection/audioonly_model, M2 = model_2_audio_denoising/audio_denoising_model).
 *
 * Conventions
 *   - every function returns 0 (SOS_OK) or a negative code; sos_last_error() gives the message
 *   - the caller owns all memory: device pointers, contiguous, fp32 unless stated; nothing is allocated
 *     per call (only the constant DFT tables at sos_init) and nothing synchronises the device
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*-compatible handle)
 *   - activations inside the networks are NHWC fp32 ("pixel-major"): (N, H=freq, W=time, C)
 *   - a "view" is 8 x int32: H, W, Hp, Wp, ph, pw, ld, coff = logical window H x W at offset (ph, pw) of a
 *     (N, Hp, Wp, ld) buffer, channels [coff, coff+C)
 */
#ifndef SOS_B200_H
#define SOS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define SOS_OK 0
#define SOS_ERR_ARG -1
#define SOS_ERR_CUDA -2
#define SOS_ERR_UNSUPPORTED -3

const char* sos_last_error(void);
int sos_version(void);
/* Builds the constant windowed-DFT tables on the current device (idempotent). */
int sos_init(void);

/* ------------------------------------------------------------------------------------------------ transforms
 * sos_stft_forward   M2/transform.py:188-193 fast_stft -> librosa.stft(y, 510, 158, 400), called from
 *                    M2/dataset.py:234-237, M1/dataset.py:288, M2/predict.py:320-326.
 *   wave (B, L) -> spec_out (B, 2, 256, T), T = 1 + L/158  (the Dataset item layout, M2/dataset.py:255-259).
 *   gate_mode 0: plain.  1: wave * mask (noise gate, M2/predict.py:317, M2/dataset.py:229).
 *   2: wave * (1 - mask) (M2/dataset.py:193).  The mask is the reference's bit string -> sample mask
 *   (M2/tools.py:340-362): bits (B, n_bits) uint8 with 0 = silent, frame_lo[i] = int(i * ratio) (n_bits + 1
 *   entries, computed by the caller with the reference's own float expression), ratio = sr / fps.
 */
int sos_stft_forward(const float* wave, int64_t batch, int64_t length, float* spec_out, const uint8_t* bits,
                     int64_t n_bits, const int32_t* frame_lo, double ratio, int gate_mode, cudaStream_t stream);
/* sos_istft_forward  M2/transform.py:196-202 fast_istft -> librosa.istft(S, 158, 400); M2/predict.py:424-447.
 *   spec (B, 2, 256, T) -> wave_out (B, 158 (T-1)).  frames_ws: workspace B*T*400 floats.
 *   crm_or_null != NULL fuses fast_icRM_sigmoid (M2/transform.py:141-153): spec is then the mixture Y. */
int sos_istft_forward(const float* spec, const float* crm_or_null, int64_t batch, int64_t n_frames, float* frames_ws,
                      float* wave_out, cudaStream_t stream);
/* sos_gate_wave      M2/tools.py:340-362 + the gating multiplies; mode as gate_mode above (1 or 2). */
/* The inverse transform's second stage alone: gather overlap-add of windowed inverse-DFT frames (batch * n_frames, 400) + division by
 * the window-sum-square envelope + trim (librosa.istft, M2/transform.py:196-202).  The frames come from sos_istft_forward's own
 * first stage or from a tensor-core GEMM (ops.istft: spectrogram rows x the windowed inverse-DFT table through ops.gemm3). */
int sos_istft_ola(const float* frames, int64_t batch, int64_t n_frames, float* wave_out, cudaStream_t stream);
int sos_gate_wave(const float* wave, int64_t batch, int64_t length, const uint8_t* bits, int64_t n_bits,
                  const int32_t* frame_lo, double ratio, int mode, float* out_or_null, float* mask_or_null,
                  cudaStream_t stream);
/* sos_icrm_*         M2/transform.py:156-169 batch_fast_icRM_sigmoid (and its gradient w.r.t. crm).
 *   Y, crm, rec: (B, 2, plane) with plane = 256*T. */
int sos_icrm_forward(const float* Y, const float* crm, float* rec, int64_t batch, int64_t plane, float a, float b,
                     cudaStream_t stream);
int sos_icrm_backward(const float* Y, const float* crm, const float* grad_rec, float* grad_crm, int64_t batch,
                      int64_t plane, float a, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------ training-item construction
 * sos_add_signals   M2/tools.py:217-303 add_signals / add_noise_to_audio (one noise track), called from M2/dataset.py:217:
 *   new_noise = noise / ratio with ratio = sqrt(P_noise) / sqrt(P_signal / 10^(snr/10)) (noise unchanged when P_signal == 0 or
 *   ratio == 0), mixed = signal + new_noise, then all three divided by max|mixed| / norm (norm == 0: no normalisation).
 *   signal, noise (the already cropped interval), outputs: (B, L); snr_db: (B) device floats.
 * sos_crm_forward   M2/transform.py:130-138 fast_cRM_sigmoid (the Dataset's "mask" entry, M2/dataset.py:239):
 *   crm = sigmoid(a * M - b), M = clean / mixed as a complex ratio; (B, 2, plane) layout, a = 0.1, b = 0 in the reference. */
int sos_add_signals(const float* signal, const float* noise, const float* snr_db, int64_t batch, int64_t length, float norm,
                    float* mixed, float* clean, float* full_noise, cudaStream_t stream);
int sos_crm_forward(const float* clean_spec, const float* mixed_spec, float* crm, int64_t batch, int64_t plane, float a, float b,
                    cudaStream_t stream);

/* sos_ssnr           M2/metrics.py:86-129 metrics_ssnr (shift = 0) and :132-175 metrics_ssnr_shift (shift = 1), the evaluation
 *   step's segmental SNR (30 ms Hann-windowed frames, quarter-frame hop, each clamped to [min_snr, max_snr]) and overall SNR,
 *   for a batch of equally long waveform pairs: ref, deg (B, L) -> overall_out, segmental_out (B). */
/* librosa.load's resampler (M2/predict.py:303, M1/dataset.py:226) = resampy.resample(filter='kaiser_best'): x (batch, n_in) ->
 * y (batch, n_out), n_out = int(n_in * ratio).  interp_win / interp_delta: the half filter table and its forward differences
 * (doubles, n_win entries, num_table samples per zero crossing; scaled by the ratio when down-sampling), time_register[t] = the
 * reference's repeatedly-added input time of output sample t.  Built by ops.resample. */
int sos_resample(const float* x, int64_t batch, int64_t n_in, float* y, int64_t n_out, const double* interp_win,
                 const double* interp_delta, int64_t n_win, int64_t num_table, double sample_ratio, const double* time_register,
                 cudaStream_t stream);
/* Frame-wise objective measures of M2/metrics.py on (batch, length) waveform pairs: 30 ms Hann frames every 7.5 ms,
 * sos_metric_frames(length, srate) of them per clip.  sos_wss: weighted spectral slope distance per frame (:404-558);
 * sos_llr: log-likelihood ratio per frame from order-16 (10 below 10 kHz) LPC (:561-681).  dist_out (batch, frames) doubles. */
int sos_metric_frames(int64_t length, int64_t srate);
int sos_wss(const float* ref, const float* deg, int64_t batch, int64_t length, int64_t srate, double eps, double* dist_out,
            cudaStream_t stream);
int sos_llr(const float* ref, const float* deg, int64_t batch, int64_t length, int64_t srate, double* dist_out, cudaStream_t stream);
int sos_ssnr(const float* ref, const float* deg, int64_t batch, int64_t length, int64_t srate, double win_len_ms, float min_snr,
             float max_snr, double eps, int shift, float* overall_out, float* segmental_out, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------ losses / optimiser
 * nn.MSELoss / nn.BCEWithLogitsLoss (M2/agent.py:172-190, M1/agent.py:185-202): *loss_sum += sum of
 * element losses (caller divides by n); grad = dloss/dpred * grad_scale when grad != NULL. */
int sos_mse_fwd_bwd(const float* pred, const float* target, int64_t n, float* loss_sum, float* grad_or_null,
                    float grad_scale, cudaStream_t stream);
int sos_bce_logits_fwd_bwd(const float* logits, const float* labels, int64_t n, float* loss_sum, float* grad_or_null,
                           float grad_scale, cudaStream_t stream);
/* optim.Adam(lr) defaults (M2/agent.py:167-170, M1/agent.py:175-183) over one flat buffer; step >= 1. */
int sos_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, int64_t step, float grad_scale, cudaStream_t stream);
/* The same update with the optimiser clock in device memory, so that a captured CUDA graph of the training step (agent.py:
 * GraphedTrainStep) replays correctly: state = float[4] {lr, step, lr / (1 - beta1^step), 1 / sqrt(1 - beta2^step)}; the call
 * advances state[1] by one and refreshes state[2..3] before the update.  The host sets state[0] (StepLR, M2/agent.py:108-111)
 * and state[1] (checkpoint restore).  n % 4 == 0, buffers 16-byte aligned. */
int sos_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float* state, float beta1,
                      float beta2, float eps, float grad_scale, cudaStream_t stream);

/* In-place round-to-nearest of fp32 values to TF32 (10-bit mantissa).  The tensor-core GEMMs below read TF32 operands by
 * TRUNCATING fp32; every producer of such an operand (packed weights, network inputs, BatchNorm outputs, gradient maps)
 * rounds instead, either through this call or through the SOS_ACT_ROUND_TF32 flag OR-ed into an `act` argument. */
#define SOS_ACT_ROUND_TF32 16
int sos_round_tf32(float* x, int64_t n, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------ BatchNorm + activation
 * nn.BatchNorm2d(eps 1e-5, momentum 0.1) + ReLU / PReLU of ConvBlock (M2/networks.py:28-51),
 * Conv2dBlock (M1/networks.py:28-51), Down/UpConvBlock (M2/networks.py:97-149) over NHWC rows.
 * act: 0 none, 1 ReLU, 2 PReLU (single slope), optionally | SOS_ACT_ROUND_TF32 (output feeds a tensor-core GEMM).  partial: sos_bn_partial_blocks(rows, C) * 3 * C floats.
 * sos_bn_act_backward ADDS the PReLU slope gradient into *dslope (the caller zeroes it). */
int sos_bn_partial_blocks(int64_t rows, int64_t channels);
int sos_bn_stats(const float* y, int64_t rows, int64_t channels, float* partial, cudaStream_t stream);
int sos_bn_finalize(const float* partial, int64_t rows, int64_t channels, const float* gamma, const float* beta, float eps,
                    float momentum, float* running_mean, float* running_var, float* mean, float* invstd, float* scale,
                    float* shift, cudaStream_t stream);
/* sos_bn_finalize on partial sums produced elsewhere (the conv epilogue): partial [g_rows][2][channels]. */
int sos_bn_finalize_partial(const float* partial, int64_t g_rows, int64_t rows, int64_t channels, const float* gamma, const float* beta,
                            float eps, float momentum, float* running_mean, float* running_var, float* mean, float* invstd, float* scale,
                            float* shift, cudaStream_t stream);
int sos_bn_eval_coeffs(int64_t channels, const float* gamma, const float* beta, const float* running_mean,
                       const float* running_var, float eps, float* scale, float* shift, cudaStream_t stream);
int sos_bn_act(const float* y, float* z, const int32_t* z_view, int64_t rows, int64_t channels, const float* scale,
               const float* shift, int act, const float* slope, cudaStream_t stream);
int sos_bn_act_backward(const float* dz, const int32_t* dz_view, const float* y, float* dy, int64_t rows, int64_t channels,
                        const float* scale, const float* shift, const float* mean, const float* invstd, int act,
                        const float* slope, float* partial, float* dgamma, float* dbeta, float* dslope, float* m1, float* m2,
                        cudaStream_t stream);
/* Half-precision operand producers (the tensor-core GEMMs run kind::f16 on them).
 * sos_bn_act_half: as sos_bn_act with a dense half output z (rows, channels), channels % 8 == 0.
 * sos_bn_act_backward_half: as sos_bn_act_backward (dense dz) with dy written as half(dy * s), s the power of two that brings
 *   the tensor's RMS to ~1 (fp16 keeps 11 significant bits from 6e-5 to 65504 only; gradient maps live near 1e-8).
 *   scal (3 floats; sos_bn_act_backward_half zeroes scal[2] itself, the _pre variant expects it zeroed by the caller): out scal[0] = s,
 *   scal[1] = 1/s (the out_scale of the consuming GEMMs), scal[2] = sum of dy^2.  partial: sos_bn_partial_blocks(rows, C) * 4 * C floats.  accumulate_param_grads: dgamma / dbeta
 *   (first real_channels entries; 0 = all) are ADDED to (they are the parameters' .grad) instead of written.
 *   y_dtype / dz_dtype (SOS_DTYPE_TF32 = fp32 storage, SOS_DTYPE_F16 = half storage): the raw conv output y may be the half
 *   array sos_conv2d_tc writes with y_dtype = F16 (its batch statistics still come from the fp32 accumulators), and dz may be the
 *   half array a data-gradient call writes WITHOUT out_scale, i.e. still multiplied by the operand scale of the layer above;
 *   dz_inv_scale (device scalar or NULL = 1) is that scale's inverse: parameter gradients are multiplied by it and the published
 *   scal[0..1] compose it, so the chain needs no pass to unscale.  Per element: 4 B forward, 10 B backward (fp32 storage: 6 / 18).
 * sos_to_half: out (rows, cd) half = x (rows, cs) fp32 zero-padded to cd channels; with scal != NULL (3 floats, scal[2]
 *   zeroed by the caller) the values are scaled as above. */
int sos_bn_act_half(const void* y, int y_dtype, void* z_half, int64_t rows, int64_t channels, const float* scale, const float* shift,
                    int act, const float* slope, cudaStream_t stream);
/* sos_bn_act_backward_half_pre: the same with pass 1 (the reduction over dz, y) already done by the producer of dz: `partial` is an
 * INPUT of partial_rows x 4 x channels floats (sos_conv_args::bnr_partial). */
int sos_bn_act_backward_half(const void* dz, int dz_dtype, const float* dz_inv_scale, const void* y, int y_dtype, void* dy_half,
                             int64_t rows, int64_t channels, const float* scale, const float* shift, const float* mean,
                             const float* invstd, int act, const float* slope, float* partial, float* dgamma, float* dbeta,
                             float* dslope, float* m1, float* m2, float* scal, int accumulate_param_grads, int64_t real_channels,
                             cudaStream_t stream);
int sos_bn_act_backward_half_pre(const void* dz, int dz_dtype, const float* dz_inv_scale, const void* y, int y_dtype, void* dy_half,
                             int64_t rows, int64_t channels, const float* scale, const float* shift, const float* mean,
                             const float* invstd, int act, const float* slope, const float* partial, int64_t partial_rows, float* dgamma, float* dbeta,
                             float* dslope, float* m1, float* m2, float* scal, int accumulate_param_grads, int64_t real_channels,
                             cudaStream_t stream);
int sos_to_half(const float* x, int64_t rows, int64_t cs, void* out_half, int64_t cd, float* scal, cudaStream_t stream);
/* eval-mode backward of z = act(y*scale+shift): dy = dz*act'(pre)*scale (no batch statistics). */
int sos_affine_act_backward(const float* dz, const int32_t* dz_view, const float* y, float* dy, int64_t rows,
                            int64_t channels, const float* scale, const float* shift, int act, const float* slope,
                            cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------ layout
 * NCHW <-> NHWC, torch.cat slices / F.interpolate(nearest) size fix-ups (M2/networks.py:198-204),
 * nn.ReflectionPad2d borders (M2/networks.py:104,129) and the encoder -> LSTM reshape + nearest resample
 * (M1/networks.py:131-135, M2/networks.py:83-86). */
int sos_nchw_to_nhwc(const float* x, int64_t batch, int64_t channels, float* out, const int32_t* out_view,
                     int64_t slice_channels, cudaStream_t stream);
int sos_nhwc_to_nchw(const float* in, const int32_t* in_view, int64_t batch, int64_t channels, float* out,
                     cudaStream_t stream);
/* (B, C, H, W) fp32 -> dense NHWC half (B, H, W, padded_channels), zero padded: the network inputs of the half path. */
int sos_nchw_to_nhwc_half(const float* x, int64_t batch, int64_t channels, int64_t H, int64_t W, void* out_half,
                          int64_t padded_channels, cudaStream_t stream);
int sos_copy_view(const float* src, const int32_t* src_view, float* dst, const int32_t* dst_view, int64_t batch,
                  int64_t channels, int accumulate, cudaStream_t stream);
int sos_copy_view_backward(const float* grad_dst, const int32_t* dst_view, float* grad_src, const int32_t* src_view,
                           int64_t batch, int64_t channels, cudaStream_t stream);
/* Backward of reflect padding + concatenation in one pass (backward of nn.ReflectionPad2d over a torch.cat, M2/networks.py:104,129,198-204):
 * dst(view) = interior of the padded gradient map + the border positions mirrored onto each pixel; the padded map is not modified. */
int sos_copy_view_fold(const float* grad_padded, const int32_t* padded_view, float* dst, const int32_t* dst_view,
                       int64_t batch, int64_t channels, cudaStream_t stream);
int sos_reflect_fill(float* buf, int64_t batch, int64_t H, int64_t W, int64_t pad, int64_t channels, cudaStream_t stream);
int sos_reflect_fold(float* grad_buf, int64_t batch, int64_t H, int64_t W, int64_t pad, int64_t channels,
                     cudaStream_t stream);
int sos_feat_to_seq(const float* in, int64_t B, int64_t F, int64_t T, int64_t C, float* out, int64_t V, int64_t ld,
                    int64_t coff, cudaStream_t stream);
int sos_feat_to_seq_backward(const float* grad_out, int64_t B, int64_t F, int64_t T, int64_t C, float* grad_in_zeroed,
                             int64_t V, int64_t ld, int64_t coff, cudaStream_t stream);
/* out (cols, rows) = in (rows, cols)^T */
int sos_transpose(const float* in, int64_t rows, int64_t cols, float* out, cudaStream_t stream);
/* y[r][c] = act(y[r][c] + bias[c]) in place; act 0 none, 1 ReLU, 3 sigmoid.  ld = row pitch (floats). */
/* fp32-grade GEMM operands for the tap GEMM (see the kernel comment in csrc/elementwise.cu): x = hi + lo with hi = tf32(x),
 * lo = tf32(x - hi); out[s * slot_stride + r * ld_out + k] = part_s(src[r * stride_r + (k + k_shift) * stride_k]) for r < R,
 * k < KP, zero where k >= K or the shifted index leaves [0, K).  n_slots = 1: {hi} (single-pass TF32 operand); 2: {hi, lo}
 * (activation stack); 3: {hi, hi, lo}
 * (weight slots along K).  One of the strides must be 1 (the kernel transposes through shared memory when stride_k != 1).
 * batch > 1: that many independent (R, K) operands src + i * src_batch_stride -> out + i * out_batch_stride (e.g. the spectrogram of
 * clip i, (512, T) with T contiguous, becoming rows i*T .. i*T + T - 1 of the inverse transform's activation operand). */
int sos_split_tf32(const float* src, int64_t R, int64_t K, int64_t KP, int64_t stride_r, int64_t stride_k, int64_t k_shift, float* out,
                   int64_t ld_out, int64_t slot_stride, int n_slots, int64_t batch, int64_t src_batch_stride, int64_t out_batch_stride,
                   cudaStream_t stream);
/* dst = (base ? base : dst) + alpha * src (parameter gradients of the LSTM / Linear layers added into the flat gradient buffer;
 * b_ih + b_hh of the LSTM). */
int sos_axpy(float* dst, const float* src, int64_t n, float alpha, const float* base_or_null, cudaStream_t stream);
/* (T, B, ld)[..., :C] sequence rows <-> (B, C, T) maps: `permute(0, 2, 1).view(B, 2, 256, T)` of M2/networks.py:92-93 and its adjoint. */
int sos_seq_map(const float* in, float* out, int64_t T, int64_t B, int64_t C, int64_t ld, int to_map, cudaStream_t stream);
int sos_bias_act(float* y, int64_t rows, int64_t cols, int64_t ld, const float* bias, int act, cudaStream_t stream);
/* dpre = dy * act'(y) (y = saved OUTPUT of the activation); dbias[c] += sum_r dpre (dbias zeroed by caller). */
int sos_bias_act_backward(const float* dy, const float* y, float* dpre, int64_t rows, int64_t cols, int64_t ld, int act,
                          float* dbias_or_null, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------ weights
 * PyTorch conv weight (Cout, Cin, kh, kw) -> GEMM operand.
 *   mode 0: forward   [Cout][ntaps * CinP], k = tap * CinP + ci
 *   mode 1: data grad [Cin][ntaps * CoutP], k = tap' * CoutP + co with tap' the flipped tap */
int sos_pack_conv_weight(const float* w, int64_t Cout, int64_t Cin, int64_t kh, int64_t kw, int64_t CinP, int64_t CoutP,
                         int mode, float* out, cudaStream_t stream);
/* Strided gather of selected taps (any of the three GEMM roles, sub-pixel phases of transposed / strided convs):
 *   out[r][t*KP + k] = w[r*row_stride + k*k_stride + tap_off[t]]  (elements; zero for K <= k < KP), rows x ntaps*KP, optionally
 *   rounded to TF32.  tap_off: host array, ntaps <= 49 entries. */
int sos_pack_taps(const float* w, int64_t rows, int64_t K, int64_t KP, int64_t row_stride, int64_t k_stride, int64_t ntaps,
                  const int32_t* tap_off, int round_tf32, float* out, cudaStream_t stream);
/* Same gather with a half output (operand of the kind::f16 GEMMs).  A NEGATIVE tap_off entry is a zero tap (row padding). */
int sos_pack_taps_half(const float* w, int64_t rows, int64_t K, int64_t KP, int64_t row_stride, int64_t k_stride, int64_t ntaps,
                       const int32_t* tap_off, void* out_half, cudaStream_t stream);
/* Taps folded into channels for the two-channel inputs (the real / imaginary spectrogram planes every network starts from,
 * M1/networks.py:124, M2/networks.py:76,159,166): out[n, oh, ow, 2 t + c] = x[n, oh + dh_t, ow + dw_t, c], c < 2, zero outside the
 * image and for the padding columns (out_channels >= 2 ntaps, a multiple of 8; ntaps <= 32).  The convolution then is a ONE-tap
 * GEMM over 128-byte pixel rows instead of `ntaps` GEMMs over 16- / 32-byte rows, with the weight packed by sos_pack_taps_half
 * (KP = K = 2, zero taps as padding). */
int sos_im2col_half(const void* x_half, int64_t batch, int64_t H, int64_t W, int64_t channels, int64_t ntaps, const int32_t* tap_dh,
                    const int32_t* tap_dw, int64_t OH, int64_t OW, void* out_half, int64_t out_channels, cudaStream_t stream);
/* The same gather for MANY weights in one launch (a training step re-packs every convolution weight in its forward and data-gradient
 * layouts once per optimiser step: 133 packs): `descs` is a DEVICE array of n descriptors with the arguments of sos_pack_taps_half. */
typedef struct sos_pack_desc {
  const void* w;           /* device, fp32 */
  void* out_half;          /* device, rows x ntaps*KP halves */
  int64_t rows, K, KP, row_stride, k_stride, ntaps;
  int32_t tap_off[50];     /* ntaps <= 49 entries used */
} sos_pack_desc;
int sos_pack_taps_half_multi(const sos_pack_desc* descs_device, int64_t n, cudaStream_t stream);
/* wgrad buffer [tap][CoutP][CinP] -> PyTorch (Cout, Cin, kh, kw) (transposed=0) or ConvTranspose (Cin, Cout, kh, kw)
 * (transposed=1, in which case src is [tap][CinP'][CoutP'] with the roles swapped by the caller). */
int sos_unpack_wgrad(const float* src, int64_t Cout, int64_t Cin, int64_t ntaps, int64_t CinP, float* dst, int accumulate,
                     cudaStream_t stream);
/* dst (rows, cols, taps) += src [tap][rows_padded][cols_padded]: the weight-gradient buffer added straight into a parameter's
 * .grad (Conv2d: rows = Cout, cols = Cin; ConvTranspose2d: rows = Cin, cols = Cout with the roles swapped by the caller). */
int sos_accumulate_wgrad(const float* src, int64_t ntaps, int64_t rows_padded, int64_t cols_padded, int64_t rows, int64_t cols,
                         float* dst, cudaStream_t stream);
/* Same, and the source elements it read are set back to zero: a persistent weight-gradient workspace is then ready for the next
 * sos_conv2d_wgrad (which accumulates into a zeroed buffer) without a fill launch. */
int sos_accumulate_wgrad_clear(float* src, int64_t ntaps, int64_t rows_padded, int64_t cols_padded, int64_t rows, int64_t cols,
                               float* dst, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------ tensor-core tap GEMM
 * One kernel family (tcgen05 kind::tf32, fp32 accumulate in TMEM, TMA-staged operands) serves
 *   nn.Conv2d of Conv2dBlock / ConvBlock / DownConvBlock (M1/networks.py:36, M2/networks.py:36,105-106),
 *   nn.ConvTranspose2d of UpConvBlock as 4 sub-pixel convs (M2/networks.py:130-131),
 *   their data gradients (flipped taps), and the nn.Linear / LSTM input projections as 1-tap GEMMs
 *   (M1/networks.py:95-98, M2/networks.py:64-70).
 *
 *   y[n, oh*osh+oph, ow*osw+opw, y_coff + co] = epi( sum_t sum_ci x[n, oh*stride + tap_dh[t], ow*stride + tap_dw[t], ci]
 *                                                      * wk[co][t*Cin + ci] )
 *   x  : (N, H, W, Cin) NHWC fp32, Cin % 8 == 0; out-of-range pixels read as zero
 *   wk : (Cout, ntaps*Cin) row-major (sos_pack_conv_weight layout)
 *   y  : (N, YH, YW, Cy) NHWC fp32; epi: v*epi_scale[co]+epi_shift[co] (optional) then act (0/1 relu/2 prelu)
 */
typedef struct sos_conv_args {
  const float* x;
  const float* wk;
  float* y;
  const int32_t* tap_dh;   /* host arrays, ntaps entries */
  const int32_t* tap_dw;
  int64_t N, H, W, Cin;
  int64_t Cout, OH, OW;
  int64_t ntaps, stride;
  int64_t YH, YW, Cy, y_coff;
  int64_t osh, osw, oph, opw;
  const float* epi_scale;  /* device, Cout entries, or NULL */
  const float* epi_shift;
  int64_t act;
  const float* slope;      /* device scalar for act == 2 */
  int64_t force_plan;      /* -1 = automatic */
  int32_t* plan_out;       /* host, 8 ints, or NULL */
  /* Optional fused BatchNorm statistics of the RAW outputs (training mode, M2/networks.py:37 nn.BatchNorm2d): the epilogue
   * writes per-warp partial sums [rows][2][stats_channels] (sum | sum of squares; rows = *stats_rows_out <= sos_conv_stats_rows())
   * for sos_bn_finalize_partial.  Needs epi_scale == NULL, act == 0, Cout <= 256. */
  float* stats_partial;    /* device, or NULL */
  int64_t stats_channels;
  int32_t* stats_rows_out; /* host */
  /* Operand type: SOS_DTYPE_TF32 (default, 0) = fp32 storage read as TF32 (kind::tf32); SOS_DTYPE_F16 = x and wk are IEEE
   * half arrays (same 11-bit significand as TF32, half the bytes, twice the tensor rate: kind::f16), Cin % 16 == 0.  The
   * accumulators and y stay fp32 unless y_dtype == SOS_DTYPE_F16 (y is then a half array; Cy and y_coff in halfs, % 8 == 0). */
  int64_t x_dtype, y_dtype;
  /* Optional device scalar multiplied into every output before the affine / activation: undoes the power-of-two scale a
   * half-precision gradient operand carries (sos_bn_act_backward_half, sos_to_half). */
  const float* out_scale;
  /* Optional fused BatchNorm-backward REDUCTION of the layer below (a data-gradient call inside an encoder chain, whose output is
   * that layer's dz): given that layer's raw conv output bnr_y (half map with this call's output geometry, Cy channels, y_coff = 0)
   * and its BatchNorm coefficients (scale = gamma*invstd, shift = beta - mean*scale, mean, invstd; >= Cout entries), the epilogue
   * also writes per-CTA partial sums [*bnr_rows_out][4][bnr_channels] -- sum g, sum g*xhat, 0, sum g^2 with g = the stored output
   * value where that layer's ReLU was active -- i.e. what pass 1 of sos_bn_act_backward_half computes from dz and y, ready for
   * sos_bn_act_backward_half_pre.  *bnr_rows_out = 0 when the kernel serving this call does not support it (then run the plain
   * sos_bn_act_backward_half). */
  const void* bnr_y;
  const float* bnr_scale;
  const float* bnr_shift;
  const float* bnr_mean;
  const float* bnr_invstd;
  float* bnr_partial;       /* device, sos_conv_stats_rows() * 4 * bnr_channels floats */
  int64_t bnr_channels;
  int32_t* bnr_rows_out;    /* host */
} sos_conv_args;
#define SOS_DTYPE_TF32 0
#define SOS_DTYPE_F16 1
int sos_conv_stats_rows(void);   /* upper bound of *stats_rows_out (one row per CTA: the SM count) */
/* Plan cache (SURVEY 8b sos_plan_*): the planner result, MMA program and tensor-map geometry of sos_conv2d_tc are computed once
 * per distinct (shapes, taps, types) key and reused; tensor maps are re-encoded only when a base pointer changes. */
void sos_plan_cache_stats(int64_t* hits, int64_t* misses, int64_t* entries);
int sos_conv2d_tc(const sos_conv_args* args, cudaStream_t stream);
/* Host-only planner query (no CUDA call, pointers other than the tap arrays are ignored): info[16] = {fast_is_w, share (+2: half outputs staged as
 * 128-byte rows of 64 channels, +4: a stage's weight tiles arrive with one TMA load), lattice g,
 * sub-tiles S, tap groups, pipeline stages, stage bytes, grid, K chunk (elements), K chunks, N, epilogue chunk, FB, SB, output staging buffers
 * (+100 when the call runs on CTA pairs), dynamic shared memory}; info[0] = 2 (rest zero but N): the row-streaming kernel serves it. */
int sos_conv2d_plan(const sos_conv_args* args, int32_t* info);

/* Weight gradient of the same operator (split over pixels, accumulated with fp32 atomics):
 *   dw[t][co][ci] += sum_{n,oh,ow} dy[n, oh, ow, dy_coff + co] * x[n, oh*stride + tap_dh[t], ow*stride + tap_dw[t], ci]
 *   dy : (N, OH, OW, Cdy) NHWC;  dw : [ntaps][Cout][Cin] fp32, zeroed by the caller.
 *   Tensor-core path needs Cin % 16 == 0, Cout % 16 == 0, Cout <= 128*8; otherwise a CUDA-core kernel runs. */
typedef struct sos_wgrad_args {
  const float* x;
  const float* dy;
  float* dw;
  const int32_t* tap_dh;
  const int32_t* tap_dw;
  int64_t N, H, W, Cin;
  int64_t Cout, OH, OW, Cdy, dy_coff;
  int64_t ntaps, stride;
  int64_t force_plan;
  int32_t* plan_out;
  int64_t dtype;            /* SOS_DTYPE_TF32 (x, dy fp32) or SOS_DTYPE_F16 (x, dy half; Cdy and dy_coff % 8 == 0) */
  const float* out_scale;   /* optional device scalar multiplied into the sums before they are added to dw */
} sos_wgrad_args;
int sos_conv2d_wgrad(const sos_wgrad_args* args, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------ LSTM recurrence
 * nn.LSTM(bidirectional, 1 layer) of M1/networks.py:95,147-148 and M2/networks.py:64,88-89 after the input
 * projection: gx (T, B, 2, 4H) = x W_ih^T + b_ih + b_hh per direction (gate order i, f, g, o).
 *   w_hh (2, 4H, H);  out (T, B, 2H) = [h_fwd | h_rev];  gates_ws (T, B, 2, 4H) activated gates, cell_ws (T, B, 2, H).
 * backward: dout (T, B, 2H) -> dgx (T, B, 2, 4H) (pre-activation gate grads), dw_hh (2, 4H, H) accumulated. */
int sos_lstm_forward(const float* gx, const float* w_hh, int64_t T, int64_t B, int64_t H, float* out, float* gates_ws,
                     float* cell_ws, cudaStream_t stream);
int sos_lstm_backward(const float* dout, const float* w_hh, const float* out, const float* gates_ws, const float* cell_ws,
                      int64_t T, int64_t B, int64_t H, float* dgx, float* dh_ws, float* dc_ws, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SOS_B200_H */
