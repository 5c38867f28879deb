/* b200asr -- C ABI of the B200-native ASR engine (libb200asr.so).
 *
 * This is the drop-in boundary for the ONE hot path this repository replaces:
 * onnxruntime.InferenceSession.run()/run_with_iobinding() as called by the
 * reference's Whisper driver script.  Plain pointers and sizes only; no torch,
 * numpy or ORT types.  Every function returns 0 on success, a negative
 * B200ASR_E_* code otherwise; b200asr_last_error() gives the message.
 *
 * Reference interfaces replaced (paths relative to the reference checkout):
 *   b200asr_encode / _transcribe   PROBE_SESSION.run_with_iobinding -- encoder half
 *                                  Whisper/Inference_Whisper_ONNX.py:493-550 (call :549),
 *                                  math Whisper/Export_Whisper.py:422-447, Whisper/STFT_Process.py:224-246
 *   b200asr_prefill                PREFILL_SESSION.run_with_iobinding  :437-490 (call :489),
 *                                  graph = merge_prefill_greedy Whisper/Shared_Merged.py:864-874
 *   b200asr_decode_step / _decode  DECODE_SESSION.run_with_iobinding   :584-663 (call :640),
 *                                  graph = merge_decode_greedy Whisper/Shared_Merged.py:877-888
 *   b200asr_no_speech_prob         NO_SPEECH_SESSION.run_with_iobinding :691-699,
 *                                  math Whisper/Export_Whisper.py:334-348
 *   b200asr_set_tensor             SessionOptions.add_initializer via attach_shared_initializers
 *                                  Whisper/Shared_Merged.py:1713-1743 (one shared weight blob)
 *
 * Threading: one engine per GPU, one CUDA stream per engine, calls on one
 * engine are not re-entrant (the reference is single-threaded, ORT_SEQUENTIAL).
 * Ownership: the caller owns every host buffer and must keep it alive until the
 * call returns; the engine owns all device memory.
 */
#ifndef B200ASR_H_
#define B200ASR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200ASR_OK 0
#define B200ASR_E_INVALID (-1)   /* bad argument / shape / state */
#define B200ASR_E_CUDA (-2)      /* CUDA runtime or kernel failure */
#define B200ASR_E_MISSING (-3)   /* a required weight tensor was never set */
#define B200ASR_E_NOGPU (-4)     /* no usable sm_100 device */

#define B200ASR_PRECISION_F32 0  /* fp32 CUDA-core path: the parity mode (logits within 1e-3) */
#define B200ASR_PRECISION_BF16 1 /* bf16 operands on tcgen05 tensor cores, fp32 accumulate/residual */

#define B200ASR_PCM_I16 0        /* raw int16 PCM; the 1/32768 scale is applied on device */
#define B200ASR_PCM_F32 1        /* float PCM already divided by audio_pcm_scale */

typedef struct b200asr_engine b200asr_engine;

typedef struct b200asr_config {
  int32_t n_mels, d_model, n_heads, ffn, enc_layers, dec_layers, vocab;
  int32_t max_source;      /* encoder positions (1500) */
  int32_t max_target;      /* decoder positions = MAX_SEQ_LEN metadata (448) */
  int32_t n_fft, hop;      /* 400 / 160 */
  int32_t max_batch;       /* utterances resident at once */
  int32_t max_samples;     /* per-utterance PCM capacity (<= 480000) */
  int32_t precision;       /* B200ASR_PRECISION_* */
  int32_t device;          /* CUDA ordinal */
  int32_t use_tensor_cores;/* bf16 only: 1 = tcgen05 GEMMs (default), 0 = CUDA-core GEMMs (debug cross-check) */
} b200asr_config;

/* lifecycle ------------------------------------------------------------------*/
int b200asr_create(const b200asr_config* cfg, b200asr_engine** out);
void b200asr_destroy(b200asr_engine* e);
const char* b200asr_last_error(const b200asr_engine* e);   /* e may be NULL: last create() error */

/* weights: folded fp32 tensors by name (see DESIGN.md "Weight tensors"); the
 * engine converts to its storage dtype on device.  finalize() checks that the
 * set is complete. */
int b200asr_set_tensor(b200asr_engine* e, const char* name, const float* host_data, int64_t numel);
int b200asr_finalize_weights(b200asr_engine* e);

/* encoder: PCM [batch][n_samples] (row stride = n_samples) -> cross-KV resident in HBM */
int b200asr_encode(b200asr_engine* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples);
/* same, split so a benchmark can time with inputs already resident */
int b200asr_upload_pcm(b200asr_engine* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples);
int b200asr_encode_resident(b200asr_engine* e);
/* ragged batch: clip b has lens[b] samples (n_fft <= lens[b] <= n_samples), rows of pcm_host are n_samples apart and n_samples is
 * the longest clip.  Replaces running the reference's dynamic-length graph (audio axis: Whisper/Export_Whisper.py:743) clip by
 * clip: each clip gets its own reflect pad, log-mel maximum, conv zero padding and attention key range, so its tokens are what
 * it gets when encoded alone.  Rows >= (lens[b] / hop + 1) / 2 of that clip's "enc_out" / cross-KV stage are padding. */
int b200asr_encode_ragged(b200asr_engine* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples,
                          const int32_t* lens);
int b200asr_upload_pcm_ragged(b200asr_engine* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples,
                              const int32_t* lens);

/* decoder: prefill resets the self-KV cache and the token bookkeeping.
 * prompt_ids [batch][n_prompt]; logits_out (optional) [batch][vocab] raw logits
 * (suppress bias applied, begin-suppress NOT applied = the graph's "logits"
 * output); first_token_out (optional) [batch] = argmax after begin-suppress. */
int b200asr_set_decode_options(b200asr_engine* e, const int32_t* stop_ids, int32_t n_stop, int32_t generate_limit,
                               float repeat_penalty, int32_t penalty_range);
/* sampling head = TOPK_TOPP_SAMPLING of Whisper/Export_Whisper.py:263-307 (strategy "sampling" of
 * Whisper/Inference_Whisper_ONNX.py:294-304, controls bound at :571-582): HF-style repetition penalty over every
 * id selected so far, / temperature, sorted top-k (<= 64), top-p on the sorted softmax, Gumbel-max.
 * temperature <= 0 switches back to the argmax heads.  noise_host (optional) = uniform numbers
 * [noise_rows][max_batch][top_k] consumed launch by launch (prefill = row 0) so a run is reproducible against
 * the oracle; beyond noise_rows, or with noise_host NULL, a counter-based generator keyed by `seed` is used. */
int b200asr_set_sampling(b200asr_engine* e, float temperature, int32_t top_k, float top_p, float repetition_penalty,
                         uint64_t seed, const float* noise_host, int32_t noise_rows);
int b200asr_prefill(b200asr_engine* e, const int32_t* prompt_ids, int32_t n_prompt, float* logits_out,
                    int32_t* first_token_out);
/* one decode launch.  token_in (optional) [batch] overrides the fed-back token
 * (teacher forcing); logits_out (optional) [batch][vocab]; token_out (optional) [batch]. */
int b200asr_decode_step(b200asr_engine* e, const int32_t* token_in, float* logits_out, int32_t* token_out);
/* device-resident greedy loop: up to max_steps decode launches without host
 * round trips; stops early when every utterance has latched a stop token or
 * hit generate_limit.  tokens_out [batch][tokens_ld], lens_out [batch]. */
int b200asr_decode(b200asr_engine* e, int32_t max_steps, int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out);
/* P(<|nospeech|>) from the last prefill's logits with the -128 suppress bias undone */
int b200asr_no_speech_prob(b200asr_engine* e, int32_t no_speech_token, float* prob_out);

/* whole path, one host call, one sync: encode + prefill + decode */
int b200asr_transcribe(b200asr_engine* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples,
                       const int32_t* prompt_ids, int32_t n_prompt, int32_t max_new, int32_t* tokens_out,
                       int32_t tokens_ld, int32_t* lens_out);
int b200asr_transcribe_ragged(b200asr_engine* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples,
                              const int32_t* lens, const int32_t* prompt_ids, int32_t n_prompt, int32_t max_new,
                              int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out);
/* same with PCM already uploaded (b200asr_upload_pcm / _ragged): device-resident timing */
int b200asr_transcribe_resident(b200asr_engine* e, const int32_t* prompt_ids, int32_t n_prompt, int32_t max_new,
                                int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out);

/* introspection (parity tests, profiling) ------------------------------------*/
/* copy an intermediate to host as fp32: "mel" [B][n_mels][T], "stem"/"enc_out" [B][T_enc][d],
 * "cross_k"/"cross_v" [B][L][H][T_enc][64], "self_k"/"self_v" [L][B][H][kv][64].
 * Returns the number of floats written through *numel_out. */
int b200asr_get_stage(b200asr_engine* e, const char* name, float* out, int64_t capacity, int64_t* numel_out);
/* options: "keep_stages" (0/1) keeps a copy of the conv-stem output for get_stage("stem");
 * decoder kernel selection (bf16 mode; all default 1): "stream" = the split-K tensor-core streaming kernel (prefill + greedy loop,
 * batch <= 8), "stream_multi" = prompt rows of all clips as one multi-row prefill iteration, "stream_l2_hint" = L2 evict-first
 * policy on the streamed boxes; with "stream" 0: "ring" = round 1's CUDA-core streaming kernel (batch <= 4), else "mega" = the
 * grid-barrier kernel, else the per-op CUDA graph.  "mega_timing" 1 selects the instrumented build (per-phase stamps, read back
 * with get_stage("mega_timing")).  "fp8" (default 0): decoder matrices + tied head as E4M3 with per-row scales in the streaming
 * kernel (quantised on device at the next launch; batch <= 4; any launch the FP8 kernel cannot take is refused with
 * B200ASR_E_INVALID).  "stream_lean" (default 1): plain greedy decode launches use the instantiation with the prompt / penalty /
 * logits / ragged branches compiled out.  "pdl" (default 1): encoder GEMMs and LayerNorm kernels launch as programmatic
 * dependents; "enc_graph" (default 1): the encoder replays as one CUDA graph per (batch, length). */
int b200asr_set_option(b200asr_engine* e, const char* key, int64_t value);
void* b200asr_stream(b200asr_engine* e);                 /* cudaStream_t of the engine */
int b200asr_synchronize(b200asr_engine* e);
int64_t b200asr_kernel_launches(const b200asr_engine* e);   /* kernels launched by this engine so far */
int b200asr_num_sms(const b200asr_engine* e);

/* standalone operator tests (the GEMM the encoder is built on) ----------------
 * C[M][N] = act(A[M][K] . B[N][K]^T + bias) + residual, host fp32 in/out, computed
 * on device in bf16 by the tcgen05 kernel (impl 1) or the CUDA-core kernel (impl 0). */
int b200asr_test_gemm(int32_t device, int32_t impl, int32_t M, int32_t N, int32_t K, const float* A, const float* B,
                      const float* bias, const float* residual, int32_t act, float* C, char* err, int32_t err_len);

/* ---------------------------------------------------------------------------------------------
 * Non-autoregressive models (SenseVoiceSmall): one call = front end + encoder + CTC head.
 * Replaces ort_session_A.run_with_iobinding of SenseVoice/Inference_SenseVoice_ONNX.py:303
 * (graph inputs `audio` [1,1,N] + `language_idx` [1], outputs `token_ids` [num_token] + `num_id` [1];
 * math SenseVoice/Export_SenseVoice.py:271-296).  The reference graph is batch 1; a batch here is that
 * many independent clips of equal length.
 * ------------------------------------------------------------------------------------------- */
#define B200ASR_NAR_SENSEVOICE 0
#define B200ASR_NAR_PARAFORMER 1   /* Paraformer/Non-Streaming/Inference_Paraformer_ONNX.py:293; math Export_Paraformer.py:474-563 */

typedef struct b200asr_nar b200asr_nar;

typedef struct b200asr_nar_config {
  int32_t kind;                    /* B200ASR_NAR_* */
  int32_t n_mels, nfft, win, hop;  /* 80 / 512 / 400 / 160 : Kaldi fbank (snip-edges framing) */
  int32_t lfr_m, lfr_n;            /* 7 / 6 : low-frame-rate stacking */
  int32_t d_model, n_heads, ffn;   /* 512 / 4 / 2048 */
  int32_t n_blocks0, n_blocks, n_tp_blocks;   /* 1 / 49 / 20 SANM blocks; after_norm sits before the tp blocks */
  int32_t vocab, blank_id;
  int32_t n_prompt;                /* prompt rows in front of the speech rows: 1 language + 3 system = 4 */
  int32_t n_lang;                  /* rows of the language prompt table (7) */
  int32_t fsmn_kernel;             /* 11 */
  int32_t max_batch, max_samples;
  int32_t precision, device, use_tensor_cores;
  float ln_eps;                    /* LayerNorm epsilon of the checkpoint's norm modules */
  /* Paraformer only (ignored for SenseVoice): n_tp_blocks = 0, n_prompt = 0, n_lang = 0 */
  int32_t dec_att_blocks, dec_ffn_blocks, dec_ffn;   /* 16 / 1 / 2048 */
  int32_t cif_kernel;              /* 3 */
  float tail_threshold;            /* 0.45 appended to the alpha sequence */
  float dec_ln_eps;
} b200asr_nar_config;

int b200asr_nar_create(const b200asr_nar_config* cfg, b200asr_nar** out);
void b200asr_nar_destroy(b200asr_nar* e);
const char* b200asr_nar_last_error(const b200asr_nar* e);
/* tensors (fp32, as the exported graph holds them): fbank_kernel [2F][win], mel_filters [F][n_mels], cmvn_means,
 * cmvn_vars [feat], speech_position [>= max T_lfr][feat], language_embed [n_lang][feat], system_embed [n_prompt-1][feat],
 * blk{i}.{norm1.g,norm1.b,qkv.w,qkv.b,fsmn.w,fsmn.b,out.w,norm2.g,norm2.b,w1.w,w1.b,w2.w,w2.b}, after_norm.{g,b},
 * tp_norm.{g,b}, ctc.{w,b}
 * Paraformer: fbank_kernel, mel_filters, cmvn_vars, encoder_input_bias [>= max T_lfr][feat], enc{i}.{qkv.w,qkv.b,fsmn.w,out.w,
 * out.b,w1.w,w1.b,w2.w,w2.b} (LayerNorm affines already folded), enc_after_norm.{g,b}, cif.conv.{w [D][D][k],b},
 * cif.out.{w,b}, dec{i}.{w1.w,w1.b,w2.w,w2.b[,norm2.g,norm2.b,fsmn.w,q.w,q.b,kv.w,kv.b,cout.w,cout.b]}, out.{w,b} */
int b200asr_nar_set_tensor(b200asr_nar* e, const char* name, const float* host_data, int64_t numel);
int b200asr_nar_finalize_weights(b200asr_nar* e);
/* pcm [batch][n_samples]: int16, or float32 carrying int16-range values (audio_pcm_scale = 1); language_idx [batch]
 * selects the language prompt row (ignored, may be NULL, for Paraformer); tokens_out [batch][tokens_ld], lens_out [batch] */
int b200asr_nar_run(b200asr_nar* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples,
                    const int32_t* language_idx, int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out);
/* same, split so a benchmark can time with the PCM already resident in HBM */
int b200asr_nar_upload(b200asr_nar* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples,
                       const int32_t* language_idx);
int b200asr_nar_run_resident(b200asr_nar* e, int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out);
/* ragged batch: clip b has lens[b] samples (win <= lens[b] <= n_samples), rows of pcm_host are n_samples (= the longest clip)
 * apart.  Each clip gets what the reference's dynamic-length graph (audio axis: SenseVoice/Export_SenseVoice.py:19,379,
 * Paraformer/Non-Streaming/Export_Paraformer.py:74,603) gives it when run alone: its own frame count, LFR tail, FSMN and
 * CIF-conv zero padding, attention key range, CTC roll and CIF tail threshold position. */
int b200asr_nar_run_ragged(b200asr_nar* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples,
                           const int32_t* lens, const int32_t* language_idx, int32_t* tokens_out, int32_t tokens_ld,
                           int32_t* lens_out);
int b200asr_nar_upload_ragged(b200asr_nar* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples,
                              const int32_t* lens, const int32_t* language_idx);
/* "mel" [B][frames][n_mels], "feats" [B][T][feat], "enc_out" [B][T][d], "logits" [B][T][vocab], "frame_ids" [B][T] */
int b200asr_nar_get_stage(b200asr_nar* e, const char* name, float* out, int64_t capacity, int64_t* numel_out);
int64_t b200asr_nar_kernel_launches(const b200asr_nar* e);
/* options: "graph" (0/1, default 1): replay the SenseVoice forward as one CUDA graph per (batch, n_samples) */
int b200asr_nar_set_option(b200asr_nar* e, const char* key, int64_t value);
void* b200asr_nar_stream(b200asr_nar* e);

/* ---------------------------------------------------------------------------------------------
 * Qwen3-ASR: log-mel -> chunked Conv2d stem -> windowed audio encoder -> prompt concat -> decoder with KV cache.
 * Replaces, in Qwen_ASR/Inference_Qwen_ASR_ONNX.py, prefill_session.run (:656: encoder + concat + rotary/mask +
 * decoder prefill + first-token arg-max) and the per-token embed_session.run + decode_session.run pair (:703-717).
 * Math: Qwen_ASR/Export_Qwen_ASR.py QWEN3_ASR_ENCODER.forward :850-927, QWEN3_ASR_ROTARY_MASK_* :933-1025,
 * QWEN3_ASR_DECODER_MAIN.forward :1265-1336, ARGMAX :1418-1420, CONCAT_EMBED :1428-1435.  The reference graph is
 * batch 1; a batch here is that many independent clips of equal length sharing one prompt.
 * ------------------------------------------------------------------------------------------- */
typedef struct b200asr_qwen b200asr_qwen;

typedef struct b200asr_qwen_config {
  int32_t device, max_batch, max_samples, precision, use_tensor_cores;
  int32_t n_mels, n_fft, hop;                                   /* 128 / 400 / 160 */
  int32_t enc_layers, enc_d, enc_heads, enc_ffn;                /* 0.6B: 18 / 896 / 14 / 3584 */
  int32_t conv_ch, out_dim, chunks_per_window;                  /* 480 / 1024 / 8 (n_window_infer / (2 n_window)) */
  float enc_ln_eps;                                             /* 1e-5 */
  int32_t vocab, hidden, inter, dec_layers, heads, kv_heads, head_dim;   /* 151936 / 1024 / 3072 / 28 / 16 / 8 / 128 */
  int32_t max_seq_len;                                          /* MAX_SEQ_LEN of the export (1024): prompt + audio + generated */
  float rms_eps;                                                /* 1e-6 */
} b200asr_qwen_config;

int b200asr_qwen_create(const b200asr_qwen_config* cfg, b200asr_qwen** out);
const char* b200asr_qwen_create_error(void);
void b200asr_qwen_destroy(b200asr_qwen* e);
const char* b200asr_qwen_last_error(const b200asr_qwen* e);
/* tensors (fp32, folded as the exported graphs hold them; Conv2d / Linear layouts as PyTorch stores them):
 * stft_kernel [2F][n_fft], mel_fbank [n_mels][F], conv{1,2,3}.{w,b}, conv_out.w, enc_pos [13][enc_d],
 * enc{i}.{qkv.w,qkv.b,out.w,out.b,fc1.w,fc1.b,fc2.w,fc2.b} (LayerNorm affines and sqrt(scaling) folded, :829-848),
 * proj1.{w,b} (ln_post folded), proj2.{w,b}, embed.w [vocab][hidden], lm_head.w (optional; absent = tied to embed.w),
 * final_norm.g, rope_cos / rope_sin [>= max_seq_len][head_dim/2],
 * dec{i}.{qkv.w (input_layernorm weight folded), qk_norm.g [2][head_dim] (q then k, head_dim^-0.25 folded), o.w,
 * gate_up.w (post_attention_layernorm weight folded; gate rows then up rows), down.w} (:1141-1190) */
int b200asr_qwen_set_tensor(b200asr_qwen* e, const char* name, const float* host_data, int64_t numel);
int b200asr_qwen_finalize_weights(b200asr_qwen* e);
/* token ids the exporter bakes around the audio (:1540-1586) and the stop set (:1503) */
int b200asr_qwen_set_prompt(b200asr_qwen* e, const int32_t* head_ids, int32_t n_head, const int32_t* suffix_ids, int32_t n_suffix,
                            const int32_t* tail_ids, int32_t n_tail, const int32_t* stop_ids, int32_t n_stop);
/* decode strategy (Inference_Qwen_ASR_ONNX.py:369-376): repeat_penalty == 1 -> greedy; otherwise penalty-greedy, the
 * script's default (REPEAT_PENALTY 0.8, PENALTY_RANGE 10, :90-91): from the first decode step on, the logits of the last
 * penalty_range selected ids are multiplied by repeat_penalty before the arg-max (APPLY_PENALTY, Export_Qwen_ASR.py:1403-1415).
 * The engine starts in greedy mode. */
int b200asr_qwen_set_decode_options(b200asr_qwen* e, float repeat_penalty, int32_t penalty_range);
/* sampling strategy (USE_SAMPLING, Inference_Qwen_ASR_ONNX.py:85-89): temperature > 0 selects TOPK_TOPP_SAMPLING
 * (Export_Qwen_ASR.py:1348-1400, the same head as Whisper's) on the prefill and every decode head; temperature <= 0 returns to
 * the arg-max heads.  noise_host [noise_rows][max_batch][top_k] uniform (0,1) makes a run reproducible (NULL = counter hash of seed). */
int b200asr_qwen_set_sampling(b200asr_qwen* e, float temperature, int32_t top_k, float top_p, float repetition_penalty, uint64_t seed,
                              const float* noise_host, int32_t noise_rows);
/* pcm [batch][n_samples]: int16 (scaled by 1/32768 on device) or float32 in [-1,1].  Leaves the prompt embedding
 * [head | query | suffix | audio | tail | language tail] in HBM; n_prompt_out = its length. */
int b200asr_qwen_encode(b200asr_qwen* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples,
                        const int32_t* query_ids, int32_t n_query, const int32_t* language_tail_ids, int32_t n_language_tail,
                        int32_t* n_prompt_out);
/* decoder over the whole prompt; logits_out [batch][vocab] (nullable), token_out [batch] = arg-max */
int b200asr_qwen_prefill(b200asr_qwen* e, float* logits_out, int32_t* token_out);
/* one token: token_in [batch] (NULL = the token selected by the previous call, as the script feeds max_logits_idx back) */
int b200asr_qwen_decode_step(b200asr_qwen* e, const int32_t* token_in, float* logits_out, int32_t* token_out);
/* greedy loop on the device until every clip hit a stop id or generation_limit = max_seq_len - 10 - n_prompt (:666) */
int b200asr_qwen_decode(b200asr_qwen* e, int32_t max_new, int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out);
/* encode + prefill + decode in one call; max_new < 0 = the script's generation_limit */
int b200asr_qwen_transcribe(b200asr_qwen* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples,
                            const int32_t* query_ids, int32_t n_query, const int32_t* language_tail_ids, int32_t n_language_tail,
                            int32_t max_new, int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out);
/* same, split so a benchmark can time with the PCM already resident in HBM */
int b200asr_qwen_upload(b200asr_qwen* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples);
int b200asr_qwen_transcribe_resident(b200asr_qwen* e, const int32_t* query_ids, int32_t n_query, const int32_t* language_tail_ids,
                                     int32_t n_language_tail, int32_t max_new, int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out);
/* Ragged batches: clips of different lengths in one batch, each with the tokens it gets when it runs alone.  pcm
 * [batch][n_samples] with n_samples = the longest clip; lens[b] (n_fft <= lens[b] <= n_samples) = samples of clip b, the rest of
 * its row is ignored.  Per clip: its own reflect padding and log-mel maximum, chunk / window key counts, audio rows, prompt length
 * (n_prompt_out [batch]), RoPE positions, cache length and generation_limit = max_seq_len - 10 - its prompt length (the reference
 * runs one clip per call, Inference_Qwen_ASR_ONNX.py:586-745).  After _upload_ragged / _encode_ragged the prefill / decode_step /
 * decode / transcribe_resident calls above work on the ragged batch; get_stage rows past a clip's own length are padding. */
int b200asr_qwen_upload_ragged(b200asr_qwen* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples, const int32_t* lens);
int b200asr_qwen_encode_ragged(b200asr_qwen* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples, const int32_t* lens,
                               const int32_t* query_ids, int32_t n_query, const int32_t* language_tail_ids, int32_t n_language_tail,
                               int32_t* n_prompt_out);
int b200asr_qwen_transcribe_ragged(b200asr_qwen* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples, const int32_t* lens,
                                   const int32_t* query_ids, int32_t n_query, const int32_t* language_tail_ids, int32_t n_language_tail,
                                   int32_t max_new, int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out);
/* "features" [B][frames][n_mels], "audio_hidden" [B][n_audio][out_dim], "prompt_embed" [B][n_prompt][hidden] (before
 * prefill), "logits" [B][vocab] */
int b200asr_qwen_get_stage(b200asr_qwen* e, const char* name, float* out, int64_t capacity, int64_t* numel_out);
int64_t b200asr_qwen_kernel_launches(const b200asr_qwen* e);
/* options: "graph" (0/1, default 1): replay a decode step as one CUDA graph; "attn_tc" (0/1): fused tcgen05 encoder attention;
 * "attn_split" (0/1, default 1, bf16 cache): key-split decode attention with register-resident cache rows;
 * "pdl" (0/1, default 1): launch decode-step kernels as programmatic dependents (batches of 1-2 clips);
 * "persist" (0/1, default 0; bf16, <= 4 clips, arg-max heads): the 5 x n_layers decoder-layer launches of a decode step as one
 * cooperative kernel with grid barriers (csrc/qwen_persist.cuh) -- same tokens, measured slower than the default on B200;
 * "persist_timing" (0/1): block 0 stamps %globaltimer at every phase of that kernel, read back with get_stage("persist_timing") */
int b200asr_qwen_set_option(b200asr_qwen* e, const char* key, int64_t value);
void* b200asr_qwen_stream(b200asr_qwen* e);

#ifdef __cplusplus
}
#endif
#endif /* B200ASR_H_ */
