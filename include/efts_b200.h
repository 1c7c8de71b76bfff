/*
 * efts_b200 -- C ABI of the B200-native EFTS-CNN forward path.
 *
 * The reference (liusongxiang/efficient_tts, package `nntts`) has no FFI of its own: the boundary it
 * offers is the Python `torch.nn.Module` contract of `nntts.models.EfficientTTSCNN` and the
 * `nntts.layers` modules.  This header is the C surface a binding for that contract
 * calls; every entry point names the reference interface it replaces (paths relative to
 * /root/reference/nntts).  `efficient_tts_b200/_lib.py` is the ctypes binding, INTEGRATION.md shows
 * the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; every tensor argument is a DEVICE pointer unless its name ends in `_host`.
 *   - activations are channels-last: [B, T, C] fp32, contiguous.
 *   - the caller owns every input, output and workspace buffer; the library owns only the
 *     prepacked weights inside `efts_ctx` and never allocates or synchronises inside a call
 *     (except `efts_inference`, which must read T2 back exactly like the reference does, and the
 *     weight upload in `efts_finalize_weights`).
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream).
 *   - one call in flight per context: the context owns the device error word, the launch counter and the
 *     profile hooks next to the (immutable) prepacked weights; use one context per concurrent stream.
 *   - return value 0 = OK, negative = `efts_status`; `efts_last_error()` gives the text.
 *     There is no CPU fallback: a missing device / wrong architecture is an error.
 */
#ifndef EFTS_B200_H_
#define EFTS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct efts_ctx efts_ctx;

typedef enum {
  EFTS_OK = 0,
  EFTS_ERR_ARG = -1,         /* bad shape / null pointer / unknown name            */
  EFTS_ERR_UNSUPPORTED = -2, /* config outside the production path (see efts_create) */
  EFTS_ERR_CUDA = -3,        /* CUDA runtime / driver error                         */
  EFTS_ERR_STATE = -4,       /* weights missing / not finalised                     */
  EFTS_ERR_WORKSPACE = -5,   /* workspace too small                                 */
  EFTS_ERR_DATA = -6         /* data-dependent failure the reference also raises    */
} efts_status;

/* Constructor arguments of EfficientTTSCNN (models/efficient_tts.py:26-49).  Values outside the
 * production configuration (egs/lj/conf/efficient_tts_cnn_phnseq_noDropout.v1.yaml:16-22 plus the
 * ctor defaults) are rejected with EFTS_ERR_UNSUPPORTED, never emulated on another path. */
typedef struct {
  int32_t num_symbols;          /* rows of text_embedding_table                                   */
  int32_t odim;                 /* mel bins (80)                                                  */
  int32_t n_channels;           /* 512; the kernels are specialised for it                        */
  int32_t k_size;               /* 5 (ResConv1d kernel size, layers/efts_modules.py:24)           */
  int32_t n_text_encoder_layer; /* 5 */
  int32_t n_mel_encoder_layer;  /* 3 */
  int32_t n_decoder_layer;      /* 6 */
  int32_t n_duration_layer;     /* 2 */
  int32_t duration_kernel_size; /* 3 (layers/duration_predictor.py:24)                            */
  float sigma;                  /* 0.01: Gaussian re-alignment width (models/...:369)             */
  float sigma_e;                /* 0.5 : aligned-position softmax width (models/...:338)          */
  float duration_offset;        /* 1.0 (models/...:215, layers/duration_predictor.py:81-83)       */
  float leaky_relu_slope;       /* 0.1 */
  int32_t use_masking;          /* FastSpeechLoss(use_masking) (losses/fastspeech_loss.py:54-61): 1 = means over
                                   valid frames / tokens only, 0 = over the whole padded batch          */
  int32_t device;               /* CUDA device ordinal                                            */
} efts_config;

/* ---- life cycle: replaces EfficientTTSCNN.__init__ / load_state_dict / remove_weight_norm ---- */
int efts_create(const efts_config* cfg, efts_ctx** out);
void efts_destroy(efts_ctx* ctx);

/* One call per tensor of the reference `state_dict` (key list: SURVEY.md 8b) with the weight-norm
 * pair already folded (`*.conv.0.weight` as left by remove_weight_norm(), models/...:400-409,
 * layers/efts_modules.py:92-99).  `data_host` is fp32, row-major, in the reference's own shape. */
int efts_set_weight(efts_ctx* ctx, const char* name, const float* data_host, const int64_t* shape,
                    int32_t ndim);
/* Checks that every tensor arrived, prepacks (tap-major fp16 hi/lo operand planes) and uploads. */
int efts_finalize_weights(efts_ctx* ctx);

/* Bytes of scratch a call with these padded sizes needs (max over all entry points). */
size_t efts_workspace_bytes(const efts_ctx* ctx, int32_t B, int32_t T1, int32_t T2);

/* ---- EfficientTTSCNN.forward (models/efficient_tts.py:120-228), eval mode ----
 * text int64 [B,T1]; text_lengths int64 [B]; speech fp32 [B,T2,odim]; speech_lengths int64 [B].
 * Outputs: imv fp32 [B,T2]; reconst_alpha fp32 [B,T1,T2]; mel_pred fp32 [B,T2,odim];
 * scalars fp32 [8] = {loss, mel_loss, duration_loss, sum_sq, n_mel, sum_abs, n_tok, flags}
 *   flags (as float-encoded int): bit0 max(text_lengths) != T1, bit1 max(speech_lengths) != T2,
 *   bit2 a text id outside [0, num_symbols), bit3 an activation left the fp16 operand range
 *   (|x| > 65504: the split-fp16 tensor-core scheme cannot represent it; nothing in the reference
 *   corresponds to this, it is reported instead of returning inf/NaN; NaN / inf inputs raise it too), bit4 a
 *   length outside [0, padded dim] (the int32 lengths the kernels use are clamped, so nothing is read or written
 *   out of bounds)  -- bits 0-2 and 4 are the conditions the reference raises on
 *   (utils/nets_utils.py:148-156 size mismatch, embedding IndexError); the binding checks them
 *   when it reads the scalars back (the read-back the reference does with .item(), :225-227). */
int efts_forward(efts_ctx* ctx, const int64_t* text, const int64_t* text_lengths, const float* speech,
                 const int64_t* speech_lengths, int32_t B, int32_t T1, int32_t T2, float* imv,
                 float* reconst_alpha, float* mel_pred, float* scalars, void* workspace,
                 size_t workspace_bytes, void* stream);

/* ---- EfficientTTSCNN.inference (models/efficient_tts.py:230-285), B = 1 like the reference ----
 * Phase 1 runs text encoder, value projection, duration predictor and the duration cumsum; it
 * leaves e[T1] in the workspace and writes T2 = round_half_even(e[T1-1]) (models/...:361) to
 * `t2_dev` (device int32[2]: T2, flags).  The caller reads T2 back (the reference's .item()),
 * allocates the outputs and calls phase 2: Gaussian reconstruction, expansion, decoder, mel head.
 * The workspace must be the same buffer for both phases, sized for (1, T1, T2max). */
int efts_inference_phase1(efts_ctx* ctx, const int64_t* text, int32_t T1, int32_t* t2_dev,
                          void* workspace, size_t workspace_bytes, void* stream);
int efts_inference_phase2(efts_ctx* ctx, int32_t T1, int32_t T2, float* mel_pred,
                          float* reconst_alpha, void* workspace, size_t workspace_bytes, void* stream);

/* ---- batched variable-length synthesis (no counterpart in the reference, whose
 * inference() is limited to B = 1 by the .item() at models/efficient_tts.py:361) ----
 * Utterance b of a padded batch (text int64 [B,T1], text_lengths int64 [B]) is computed exactly as
 * `inference(text[b:b+1, :text_lengths[b]])` would compute it: every layer writes zeros beyond the
 * utterance's own length, so neighbours and padding cannot leak in.
 * Phase 1 leaves e[B,T1] in the workspace and writes t2_dev int32 [B+1] = per-utterance frame counts
 * T2_b = round_half_even(e[b, L1_b - 1]) followed by a flags word (bit2: token id out of range).
 * The caller reads t2_dev back, allocates mel_pred [B,T2max,odim] and reconst_alpha [B,T1,T2max]
 * (T2max = max_b T2_b; a T2_b < 1 is the caller's error to raise) and calls phase 2 with the same
 * workspace (sized for (B, T1, T2max)); frames t >= T2_b are written as zeros. */
int efts_inference_batch_phase1(efts_ctx* ctx, const int64_t* text, const int64_t* text_lengths, int32_t B,
                                int32_t T1, int32_t* t2_dev, void* workspace, size_t workspace_bytes,
                                void* stream);
int efts_inference_batch_phase2(efts_ctx* ctx, int32_t B, int32_t T1, int32_t T2max, const int32_t* t2_dev,
                                float* mel_pred, float* reconst_alpha, void* workspace,
                                size_t workspace_bytes, void* stream);

/* ---- ResConvBlock.forward (layers/efts_modules.py:54-79) on channels-last data ----
 * stack: 0 = text_encoder, 1 = mel_encoder, 2 = decoder.  x, y fp32 [B,T,C]; may alias. */
int efts_conv_stack_fwd(efts_ctx* ctx, int32_t stack, const float* x, float* y, int32_t B, int32_t T,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- DurationPredictor.forward / .inference (layers/duration_predictor.py:66-113) ----
 * x fp32 [B,T,C]; lengths int32 [B] or NULL (no mask).  mode 0: log-domain fp32 out (forward);
 * mode 1: clamp(exp(x)-offset, 0) fp32 (inference, to_round=False); mode 2: clamp(round(exp(x)-offset),0)
 * int64 out (inference, to_round=True).  Positions t >= lengths[b] are written as 0. */
int efts_duration_predictor_fwd(efts_ctx* ctx, const float* x, const int32_t* lengths, int32_t B,
                                int32_t T, int32_t mode, void* out, void* workspace,
                                size_t workspace_bytes, void* stream);

/* ---- LengthRegulator.forward (layers/length_regulator.py:35-79; pad_list utils/nets_utils.py:28-55)
 * plan: ds int64 [B,T1] (rewritten in place only for the all-zero fix-up when alpha == 1, like the
 * reference's view semantics :52,76-78); ilens int64 [B].  Writes ds_eff int64 [B,T1] (rounded /
 * fixed-up durations actually used), out_lens int64 [B] and plan int64 [2] = {max_b out_lens, flags}
 * (flags bit0: a negative duration -- torch.repeat raises).  The caller reads `plan` back to size
 * the output, then calls fwd: xs fp32 [B,T1,D] -> out fp32 [B,Tout,D] (pad_value beyond out_lens),
 * idx int64 [B,Tout] (source token of every output frame, -1 on padding; may be NULL). */
int efts_length_regulator_plan(int64_t* ds, const int64_t* ilens, float alpha, int32_t B, int32_t T1,
                               int64_t* ds_eff, int64_t* out_lens, int64_t* plan, void* stream);
int efts_length_regulator_fwd(const float* xs, const int64_t* ds_eff, const int64_t* ilens,
                              const int64_t* out_lens, int32_t B, int32_t T1, int32_t D, int64_t Tout,
                              float pad_value, float* out, int64_t* idx, void* stream);

/* ---- building blocks exposed for parity tests (same kernels the calls above launch) ----
 * Tap-GEMM on the tensor cores: out[b,t,n] = sum_tap sum_k x[b,t+tap-pad,k] * w[z,n,k]
 * (z = tap, or z = b when `batched`), fp32 in / fp32 out through the split-fp16 operand planes.
 * x [B,T,K], w [Z,N,K], out [B,T,N]; K and N multiples of 8. */
int efts_tap_gemm(efts_ctx* ctx, const float* x, const float* w, float* out, int32_t B, int32_t T,
                  int32_t K, int32_t N, int32_t ntaps, int32_t pad, int32_t batched, void* workspace,
                  size_t workspace_bytes, void* stream);
/* The alignment block alone (models/efficient_tts.py:167-194): mel_h fp32 [B,T2,C] (mel-encoder
 * output), key/value fp32 [B,T1,C] (already zero at pad tokens), lengths int32.  Outputs imv
 * [B,T2], e [B,T1], reconst_alpha [B,T1,T2], expanded [B,T2,C] (value at frame rate, zero at pad). */
int efts_alignment_fwd(efts_ctx* ctx, const float* mel_h, const float* key, const float* value,
                       const int32_t* text_lengths, const int32_t* speech_lengths, int32_t B,
                       int32_t T1, int32_t T2, float* imv, float* e, float* reconst_alpha,
                       float* expanded, void* workspace, size_t workspace_bytes, void* stream);

/* ---- stand-alone helper methods of EfficientTTSCNN (models/efficient_tts.py:287-398) ----
 * The reference exposes the stages of its alignment block as public methods taking bool masks; masks
 * built by make_non_pad_mask are prefix masks, which travel here as int32 lengths (efts_mask_lengths
 * converts a bool [B,T] mask).  forward() uses fused kernels; these are for callers of the methods. */
int efts_mask_lengths(const uint8_t* mask, int32_t B, int32_t T, int32_t* lengths, void* stream);
/* generate_index_vector (:287-297): p fp32 [B,T1]. */
int efts_index_vector(const int32_t* text_lengths, int32_t B, int32_t T1, float* p, void* stream);
/* scaled_dot_product_attention (:377-398): query [B,T2,C], key [B,T1,C] -> alpha fp32 [B,T1,T2]. */
int efts_attention_alpha(efts_ctx* ctx, const float* query, const float* key, const int32_t* text_lengths,
                         int32_t B, int32_t T1, int32_t T2, float* alpha, void* workspace,
                         size_t workspace_bytes, void* stream);
/* imv_generator (:299-324): alpha [B,T1,T2], p [B,T1] -> imv [B,T2]. */
int efts_imv_generator(const float* alpha, const float* p, const int32_t* text_lengths,
                       const int32_t* speech_lengths, int32_t B, int32_t T1, int32_t T2, float* imv,
                       void* workspace, size_t workspace_bytes, void* stream);
/* get_aligned_positions (:326-345): imv [B,T2], p [B,T1] (NULL = index vector) -> e [B,T1]. */
int efts_aligned_positions(const float* imv, const float* p, const int32_t* text_lengths,
                           const int32_t* speech_lengths, int32_t B, int32_t T1, int32_t T2, float sigma_e,
                           float* e, void* stream);
/* reconstruct_align_from_aligned_position (:347-375): e [B,T1] -> fp32 [B,T1,T2]; NULL lengths = no mask. */
int efts_reconstruct_alignment(const float* e, const int32_t* text_lengths, const int32_t* speech_lengths,
                               int32_t B, int32_t T1, int32_t T2, float delta, float* reconst_alpha,
                               void* stream);

/* ---- HiFi-GAN V1 generator: the vocoder the reference runs right after inference(),
 * `y = voc_model(mel_pred.transpose(1, 2))` (bin/inference.py:108-109, vocoders/hifigan_model.py:95-136).
 * Same context type, same weight protocol: efts_vocoder_create, one efts_set_weight per tensor of the
 * Generator's state_dict with the weight-norm pairs folded (names as left by remove_weight_norm():
 * "conv_pre.weight" [C0,80,7], "ups.i.weight" [Cin,Cout,k] (ConvTranspose1d layout), "resblocks.n.convs1.m.weight"
 * [C,C,k], "resblocks.n.convs2.m.weight", "conv_post.weight" [1,C,7] and the matching ".bias"), then
 * efts_vocoder_finalize (ResBlock2 generators name their convs "resblocks.n.convs.m.weight").  Supported:
 * upsample_kernel_size == 2 * upsample_rate (rate even), resblock kernels <= 11 with dilation * (k - 1) <= 72,
 * channels multiples of 8 (the V1, V2 and V3 configurations of the HiFi-GAN release); anything else is
 * EFTS_ERR_UNSUPPORTED. */
typedef struct {
  int32_t num_mels;                    /* 80 (Conv1d(80, ...) at vocoders/hifigan_model.py:101)            */
  int32_t upsample_initial_channel;    /* 512                                                             */
  int32_t num_upsamples;               /* 4                                                               */
  int32_t upsample_rates[8];           /* 8, 8, 2, 2                                                      */
  int32_t upsample_kernel_sizes[8];    /* 16, 16, 4, 4                                                    */
  int32_t num_kernels;                 /* 3 (resblocks per stage)                                         */
  int32_t resblock_kernel_sizes[4];    /* 3, 7, 11                                                        */
  int32_t resblock_dilations[4][3];    /* {1, 3, 5} each                                                  */
  int32_t resblock_type;               /* 1: ResBlock1 (:31-63, convs1/convs2 pairs); 2: ResBlock2 (:71-88) */
  int32_t num_dilations;               /* dilations per resblock: 3 for ResBlock1, 2 for ResBlock2        */
  int32_t device;
} efts_vocoder_config;
int efts_vocoder_create(const efts_vocoder_config* cfg, efts_ctx** out);
int efts_vocoder_finalize(efts_ctx* ctx);
size_t efts_vocoder_workspace_bytes(const efts_ctx* ctx, int32_t B, int32_t T);
/* Generator.forward (vocoders/hifigan_model.py:120-136): mel fp32 [B, num_mels, T] (channels-first, exactly what
 * the reference passes) -> waveform fp32 [B, 1, T * prod(upsample_rates)].  Raises bit 3 of the error flags
 * (efts_error_flags) when an activation leaves the fp16 operand range. */
int efts_vocoder_forward(efts_ctx* ctx, const float* mel, int32_t B, int32_t T, float* audio, void* workspace,
                         size_t workspace_bytes, void* stream);

/* Host-only test hooks (no device): the fp32 weights the generator's packers hand to the tap-GEMM.
 * efts_host_map_transposed: ConvTranspose1d weight [Cin,Cout,k] (k = 2u) -> [3][u*Cout][Cin] (3-tap polyphase GEMM);
 * efts_host_map_grouped: Conv1d weight [C,C,k], dilation d, G time steps per GEMM row -> [taps][G*C][G*C]
 * (`out` may be NULL to query `taps`). */
int efts_host_map_transposed(const float* w, int32_t Cin, int32_t Cout, int32_t k, int32_t u, float* out);
int efts_host_map_grouped(const float* w, int32_t C, int32_t k, int32_t d, int32_t G, float* out, int32_t* taps);

/* ---- log-mel front-end: nntts.datasets.meldataset.mel_spectrogram, datasets/meldataset.py:49-82 ----
 * reflect padding of (n_fft - hop) / 2 -> STFT (Hann window, center = False, one-sided) -> sqrt(re^2 + im^2 + 1e-9)
 * -> mel filter bank -> log(clamp(x, 1e-5)).  The STFT runs as a tap-GEMM over hop-sized chunks of the padded
 * waveform, the mel projection as a GEMM with the log in its epilogue, both on the tcgen05 kernel of the path.
 * Weights (host fp32, set with efts_set_weight, then efts_frontend_finalize):
 *   "stft.weight" [n_fft, hop, n_fft / hop] -- output column n < n_fft/2 + 1: w[s] cos(2 pi n s / n_fft) (re_n),
 *   column n > n_fft / 2: -w[s] sin(2 pi (n - n_fft/2) s / n_fft) (im of bins 1 .. n_fft/2 - 1; im_0 and im_{n_fft/2}
 *   vanish), with s = tap * hop + k; "stft.bias" [n_fft] zeros; "mel_basis.weight" [num_mels, round8(n_fft/2 + 1)]
 *   (librosa.filters.mel, zero-padded columns); "mel_basis.bias" [num_mels] zeros.
 * efts_frontend_finalize keeps the bins up to the last one some mel filter weighs (the others are multiplied by zero at
 * :74) and repacks the basis as (re_f, im_f) column pairs, so that the STFT GEMM's epilogue writes the magnitude planes
 * itself; the fp32 spectrum is never stored. */
typedef struct efts_frontend_config {
  int32_t n_fft, hop_size, win_size, num_mels;
  int32_t device;
} efts_frontend_config;
int efts_frontend_create(const efts_frontend_config* cfg, efts_ctx** out);
int efts_frontend_finalize(efts_ctx* ctx);
/* Frames of an utterance of `length` samples: 1 + (length + 2 * ((n_fft - hop) / 2) - n_fft) / hop (reflect padding on
 * both sides, center = False); 0 when the padded utterance is shorter than one window. */
int32_t efts_frontend_frames(const efts_ctx* ctx, int64_t length);
size_t efts_frontend_workspace_bytes(const efts_ctx* ctx, int32_t B, int32_t Lmax);
/* audio fp32 [B, Lmax] in [-1, 1]; lengths int64 [B] or NULL (all Lmax): every utterance is reflect-padded at its OWN
 * length, exactly as if mel_spectrogram had been called on it alone (TextMelLoader.get_mel, datasets/taco2_data.py:
 * 72-78).  Outputs: mel fp32 [B, Tmax, num_mels] with Tmax = frames(Lmax), zero beyond each utterance's frames (the
 * layout and padding TextMelCollate hands to the model, datasets/taco2_data.py:125-139), mel_lengths int64 [B] or
 * NULL.  Error bits (efts_error_flags): 3 sample outside the fp16 operand range, 4 length outside [0, Lmax], 5 an
 * utterance not longer than the reflect padding (torch raises on such input). */
int efts_frontend_forward(efts_ctx* ctx, const float* audio, const int64_t* lengths, int32_t B, int32_t Lmax, float* mel,
                          int64_t* mel_lengths, void* workspace, size_t workspace_bytes, void* stream);

/* ---- training slice: ResConvBlock under autograd ----
 * The first piece of the reference's training step (trainers/efficient_tts_trainer.py:152-154, loss.backward()):
 * forward of a stack of residual conv layers y = x + lrelu_{0.1}(conv_k(x) + b) (layers/efts_modules.py:48-51,77-79)
 * that keeps what the backward needs, and the backward itself -- data gradient as the tap-GEMM with flipped,
 * transposed weights, weight gradient as one GEMM per tap that reduces over positions (split over the SMs), bias
 * gradient as a column sum.  Weights are the caller's CURRENT fp32 device tensors (they change every step), stacked:
 * weights [n_layers, 512, 512, k], biases [n_layers, 512]; activations are channels-last [B, T, 512].
 *   fwd: acts [n_layers + 1, B, T, 512] (acts[0] = x, acts[l + 1] = output of layer l; the last slice is the result)
 *        and us [n_layers, B, T, 512] (the activated conv outputs; their sign is LeakyReLU') are written.
 *   bwd: grad_out = dL/d acts[n_layers] -> grad_x [B, T, 512], grad_w [n_layers, 512, 512, k], grad_b [n_layers, 512].
 * efficient_tts_b200/layers.py wraps the pair in a torch.autograd.Function so that ResConvBlock trains; the
 * weight-norm reparametrisation (weight_g, weight_v) is differentiated by torch on top of grad_w. */
size_t efts_resconv_train_workspace_bytes(const efts_ctx* ctx, int32_t B, int32_t T, int32_t k);
int efts_resconv_train_fwd(efts_ctx* ctx, const float* x, const float* weights, const float* biases, int32_t n_layers,
                           int32_t k, int32_t B, int32_t T, float* acts, float* us, void* workspace,
                           size_t workspace_bytes, void* stream);
int efts_resconv_train_bwd(efts_ctx* ctx, const float* grad_out, const float* acts, const float* us, const float* weights,
                           int32_t n_layers, int32_t k, int32_t B, int32_t T, float* grad_x, float* grad_w, float* grad_b,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Second slice: DurationPredictor (layers/duration_predictor.py:57-88) under autograd -- n_layers x [Conv1d(k) -> ReLU ->
 * LayerNorm over channels -> Dropout], Linear(512 -> 1), masked positions 0; `forward` semantics (log domain).
 *   fwd: x fp32 [B,T,512]; conv_w [L,512,512,k], conv_b [L,512], ln_g / ln_b [L,512], head_w [512], head_b [1];
 *        mask uint8 [B,T] (non-zero = padded position, may be NULL); keep fp32 [L,B,T,512] = the train-mode dropout masks
 *        already scaled by 1/(1-p), drawn by the caller's generator (NULL = no dropout).
 *        Saves acts [L+1,B,T,512] (acts[0] = x, acts[l+1] = layer l's output) and us [L,B,T,512] (post-ReLU conv outputs,
 *        whose sign is ReLU'); out fp32 [B,T].
 *   bwd: grad_out = dL/dout [B,T] -> grad_x [B,T,512], grad_conv_w [L,512,512,k], grad_conv_b, grad_ln_g, grad_ln_b
 *        [L,512], grad_head_w [512], grad_head_b [1].  Convolutions: the data- and weight-gradient GEMMs of the first
 *        slice (ReLU slope 0, no residual); LayerNorm / head: one warp per row, column sums in double, deterministic. */
size_t efts_duration_train_workspace_bytes(const efts_ctx* ctx, int32_t B, int32_t T, int32_t k);
int efts_duration_train_fwd(efts_ctx* ctx, const float* x, const float* conv_w, const float* conv_b, const float* ln_g,
                            const float* ln_b, const float* head_w, const float* head_b, const uint8_t* mask,
                            const float* keep, int32_t n_layers, int32_t k, int32_t B, int32_t T, float* acts, float* us,
                            float* out, void* workspace, size_t workspace_bytes, void* stream);
int efts_duration_train_bwd(efts_ctx* ctx, const float* grad_out, const float* acts, const float* us, const float* conv_w,
                            const float* ln_g, const float* head_w, const uint8_t* mask, const float* keep,
                            int32_t n_layers, int32_t k, int32_t B, int32_t T, float* grad_x, float* grad_conv_w,
                            float* grad_conv_b, float* grad_ln_g, float* grad_ln_b, float* grad_head_w,
                            float* grad_head_b, void* workspace, size_t workspace_bytes, void* stream);

/* The criterion with gradients: FastSpeechLoss.forward (losses/fastspeech_loss.py:54-67) with after_outs = None and
 * use_weighted_masking = False -- what models/efficient_tts.py:220 calls.  losses[0] = mean over the selected elements of
 * (before_outs - ys)^2 (use_mse; |.| otherwise), losses[1] = mean over the selected tokens of |d_outs - ds|; "selected" =
 * inside olens / ilens when use_masking, everything otherwise.  grad_before [B,T2,odim] / grad_d [B,T1] (either may be
 * NULL) receive d losses[0] / d before_outs and d losses[1] / d d_outs.  Block partial sums in double, added in block
 * order (deterministic).  A length outside [0, padded dim] raises flag bit 4 (the reference's mask broadcast fails).
 * efts_scale_by_scalar: out = in * scalar[0] with the scalar on the device (chain rule of a loss term, no read-back). */
size_t efts_fastspeech_loss_workspace_bytes(const efts_ctx* ctx);
int efts_fastspeech_loss(efts_ctx* ctx, const float* before_outs, const float* d_outs, const float* ys, const float* ds,
                         const int64_t* ilens, const int64_t* olens, int32_t B, int32_t T1, int32_t T2, int32_t odim,
                         int32_t use_masking, int32_t use_mse, float* losses, float* grad_before, float* grad_d,
                         void* workspace, size_t workspace_bytes, void* stream);
int efts_scale_by_scalar(efts_ctx* ctx, const float* in, const float* scalar, size_t n, float* out, void* stream);

/* ---- data parallelism ----
 * The survey's sketch of this ABI listed efts_dp_init / efts_dp_allgather / efts_dp_allreduce_loss.  They are
 * deliberately NOT exported: the reference's only parallelism is torch DDP with a DistributedSampler
 * (bin/train.py:64-68,136-150,210-216) -- the process group, its NCCL communicator and its streams belong to
 * torch.distributed in the reference and in every caller of this library, and a second communicator owned by this
 * library would have to be bootstrapped out of band and ordered against torch's.  What the library provides for a
 * shard is everything the exchange needs, device-resident and without a read-back: efts_forward returns the four
 * loss partial sums and the error bits in scalars[3..7], and a shard is called with the GLOBAL padded (T1, T2) so
 * per-utterance outputs are bitwise independent of the shard count.  The exchange itself -- ONE all-reduce of a
 * 7-float vector per step, optionally an all-gather of the outputs -- is efficient_tts_b200/data_parallel.py on the
 * caller's process group (NCCL over NVLink; gloo in the CPU tests).  bench.py times it inside the step at N > 1. */

/* ---- introspection ---- */
/* Tuning / test switches, all of which keep results within the parity budget (most are bitwise neutral):
 * "skip_pad_tiles" (0/1: skip row tiles that cannot reach a valid output), "pair" (CTA pairs for the weight GEMMs),
 * "wide" (16-epilogue-warp variant for short reductions), "fuse_b" (Ahi x [Bhi|Blo] as one N = 256 MMA),
 * "chunk_kb" (k-blocks per main-accumulator flush), "split_k" (split reduction for small launches), "pdl"
 * (programmatic dependent launch), "imv_version" (2: block-per-utterance IMV kernels, 1: the warp-per-row kernels
 * that serve rows too long for shared memory -- bitwise equal), "voc_group" / "voc_wide" / "voc_narrow"
 * (vocoder layer packing), "stack" (resident layer-stack kernels for B = 1 synthesis; 0 = one launch per layer,
 * bitwise the same results), "stack_trace" (measurement hook of efts_profile_stack_trace).
 * No option selects a kernel outside the precision budget.  Returns EFTS_ERR_ARG for an unknown name. */
int efts_set_option(efts_ctx* ctx, const char* name, int32_t value);
/* Measurement hooks: while a tag's bit is set in `tag_mask`, every launch of that kind is bracketed
 * by a CUDA-event pair on the caller's stream (no extra synchronisation).  Tags: 0 text-encoder conv
 * layer, 1 mel-encoder conv layer, 2 decoder conv layer, 3 linear, 4 energy GEMM, 5 softmax-expectation,
 * 6 IMV scan, 7 aligned positions, 8 Gaussian reconstruction, 9 expansion GEMM, 10 duration predictor.
 * `efts_profile_read` waits for the recorded events and returns their summed duration and count;
 * `efts_profile_enable` also clears what was recorded. */
int efts_profile_enable(efts_ctx* ctx, uint32_t tag_mask);
int efts_profile_read(efts_ctx* ctx, int32_t tag, double* total_ms, int64_t* count);
/* Template instantiation of the last tensor-core GEMM launched under `tag` while profiling was enabled, e.g.
 * "gemm2_kernel<2, 0, 0, 1, 136, 128>" (empty if none): bench.py matches it against the kernel name of the
 * committed ncu capture before quoting that capture's numbers. */
int efts_profile_kernel_name(const efts_ctx* ctx, int32_t tag, char* buf, size_t n);
/* SM clock stamps (clock64 of CTA 0) the resident layer-stack kernel of B = 1 synthesis recorded at every phase
 * boundary of its last launch while option "stack_trace" was on: start, [after embedding, after its barrier,]
 * then per layer {GEMM done, barrier passed, reduce done, barrier passed}; entries 36..47 and 48..55 hold stamps of
 * CTA 0's MMA warp (operands of step i landed) and of its roles during layer 1 (tools/c1_phases.py decodes them).
 * Synchronising copy of `n` <= 64 values. */
int efts_profile_stack_trace(efts_ctx* ctx, int64_t* out, int32_t n);
/* Data-dependent error bits raised by the kernels of the calls issued on `stream` since the last
 * efts_forward / efts_inference_phase1 (bit 3: activation outside the fp16 operand range; bit 6: a grid barrier
 * of the resident layer-stack kernel timed out -- the call's results are invalid).
 * Synchronises the stream (4-byte read-back); efts_forward reports the same bits in scalars[7]. */
int efts_error_flags(efts_ctx* ctx, void* stream, int32_t* flags_host);
/* Synchronising read-back of `n` <= 64 device words through the context's pinned staging buffer (the T2 / flags
 * read of efts_inference_phase1 -- the reference's .item() at models/efficient_tts.py:361 -- without a pageable
 * copy).  One call in flight per context. */
int efts_read_words(efts_ctx* ctx, const int32_t* dev, int32_t* host, int32_t n, void* stream);
/* Kernels launched by this context since creation (bench.py's `gpu_launches`). */
int64_t efts_launch_count(const efts_ctx* ctx);
const char* efts_last_error(void);
/* "efts_b200 <ver> (...) src <sha16>": sha16 = hash of the sources this binary was built from (build.py). */
const char* efts_version(void);

#ifdef __cplusplus
}
#endif
#endif /* EFTS_B200_H_ */
