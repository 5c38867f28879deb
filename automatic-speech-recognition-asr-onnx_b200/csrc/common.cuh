// Shared declarations for the b200asr CUDA engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <atomic>
#include <string>

namespace b200asr {

enum DType : int { kF32 = 0, kBF16 = 1 };
enum Act : int { kActNone = 0, kActGelu = 1, kActRelu = 2, kActGeluTanh = 3 };

typedef __nv_bfloat16 bf16;

__host__ __device__ inline size_t dtype_size(int dt) { return dt == kF32 ? 4 : 2; }

// C[z][m][n] = act(sum_k A[z][m][k] * B[z][n][k] + bias[n]) + residual[z][m][n]
// Batch index z is split as z = zo * batch_inner + zi so one launch can walk
// (utterance, head) with two independent strides per operand.
struct GemmArgs {
  const void* A = nullptr; int64_t lda = 0, sAo = 0, sAi = 0; int a_dtype = kF32;
  const void* B = nullptr; int64_t ldb = 0, sBo = 0, sBi = 0; int b_dtype = kF32;
  int transB = 0;                     // 0: B is [N][K] (K contiguous); 1: B is [K][N]
  void* C = nullptr; int64_t ldc = 0, sCo = 0, sCi = 0; int c_dtype = kF32;
  const float* bias = nullptr; int64_t sBias = 0;   // bias advances by sBias per batch index z
  const float* residual = nullptr; int64_t ldr = 0, sRo = 0, sRi = 0;
  int act = kActNone;
  int M = 0, N = 0, K = 0;
  int batch = 1, batch_inner = 1;
  int pdl = 0;                        // gemm_tc only: launch as a programmatic dependent (B must be a weight matrix no kernel writes)
};

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: a process-wide "done" flag would leave a second engine on
// another GPU without it.  One bit per (device, slot), set atomically; returns true when the caller must set the attribute.
struct AttrOnce {
  std::atomic<unsigned long long> bits[64];
  bool need(int slot = 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    const unsigned long long m = 1ull << (slot & 63);
    return (bits[dev].fetch_or(m) & m) == 0;
  }
};

// ---- device helpers ---------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// exact (erf) GELU: torch.nn.functional.gelu default, Whisper/Export_Whisper.py:428,437,662
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// tanh-approximated GELU: torch.nn.GELU(approximate="tanh"), Qwen_ASR/Export_Qwen_ASR.py:716-719,873-875
__device__ __forceinline__ float gelu_tanh(float x) {
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.0f + tanhf(u));
}

// Programmatic dependent launch (no-ops for a kernel launched without the attribute)
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// order-preserving float <-> int key (for atomicMax on floats of either sign)
__device__ __forceinline__ int float_to_key(float f) {
  int b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float key_to_float(int k) {
  return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff);
}

// ---- launch wrappers (one per .cu) -------------------------------------------
// frontend.cu
cudaError_t launch_logmel(const void* pcm, int pcm_is_f32, int batch, int n_samples, int64_t pcm_stride,
                          const float* basis_t /*[n_fft][2F]*/, const float* fbank /*[n_mels][F]*/,
                          const int* fb_start, const int* fb_len, int n_fft, int hop, int n_mels,
                          float* mel_raw /*[B][T][n_mels]*/, int* max_key /*[B]*/, cudaStream_t st,
                          const int* n_per_clip = nullptr /*ragged batch: samples per clip (device)*/);
cudaError_t launch_zero_tail_rows(void* buf, int dtype, const int* n_per_clip, int hop, int batch, int T, int d, cudaStream_t st);
cudaError_t launch_mel_finalize(const float* mel_raw, const int* max_key, int batch, int T, int n_mels,
                                void* mel_pad /*[B][T+2][n_mels]*/, int out_dtype, cudaStream_t st,
                                const int* n_per_clip = nullptr, int hop = 1);
cudaError_t launch_fill_i32(int* p, int v, int n, cudaStream_t st);

// gemm_simt.cu
cudaError_t launch_gemm_simt(const GemmArgs& g, cudaStream_t st);

// gemm_tc.cu  (tcgen05 + TMA; bf16 operands, fp32 accumulate)
bool gemm_tc_supported(const GemmArgs& g);
cudaError_t launch_gemm_tc(const GemmArgs& g, int num_sms, cudaStream_t st, std::string* err);

// layers.cu
cudaError_t launch_layernorm(const float* x, int64_t ldx, const float* gamma, const float* beta, void* out,
                             int out_dtype, int64_t ldo, int rows, int d, float eps, cudaStream_t st, int pdl = 0 /*1: launched as a programmatic dependent of the kernel that writes x*/);
cudaError_t launch_softmax_rows(const float* s, void* p, int p_dtype, int64_t rows, int cols, cudaStream_t st,
                                const int* valid = nullptr /*[entries]: key columns per entry*/, int64_t rows_per_entry = 1);

// decoder.cu
struct DecState {           // lives in device memory, one per engine
  int kv_len;               // tokens already in the self-KV cache
  int step;                 // decode launches since prefill (0 = prefill head)
  int all_done;             // 1 when every utterance has latched a stop token / hit the limit
  int pad;
};
struct DecLinearArgs {
  const float* x; int64_t ldx;          // fp32 rows [rows][K]
  int ln_mode;                          // 0 none, 1 affine-less LN, 2 affine LN (gamma/beta), 3 affine-less RMS norm
  const float* gamma; const float* beta; float eps;
  const void* W; int w_dtype;           // [N][K]
  const float* bias;                    // [N] or null
  int act;
  const float* residual; int64_t ldr;   // fp32 [rows][N] or null (may alias out)
  float* out; int64_t ldo;              // fp32 [rows][N]        (mode 0)
  // mode 1: fused-QKV scatter: cols [0,d) -> out (q, ld=ldo); [d,2d) -> K cache; [2d,3d) -> V cache
  int mode; void* kcache; void* vcache; int kv_dtype;
  int n_new, n_heads, head_dim, max_target, batch; const DecState* state;
  int rows, N, K;
};
cudaError_t launch_dec_linear(const DecLinearArgs& a, cudaStream_t st);
cudaError_t launch_dec_embed(const int* tokens /*[B][n_new]*/, const void* embed, int w_dtype, const void* pos,
                             int batch, int n_new, int d, const DecState* state, float* x, cudaStream_t st);
cudaError_t launch_dec_self_attn(const float* q /*[rows][d]*/, const void* kcache, const void* vcache, int kv_dtype,
                                 int batch, int n_new, int n_heads, int head_dim, int max_target,
                                 const DecState* state, float* ctx /*[rows][d]*/, cudaStream_t st);
cudaError_t launch_dec_cross_attn(const float* q, const void* cross_kv /*[2L][B][T][d]: K layers then V layers*/, int kv_dtype, int layer,
                                  int n_layers, int batch, int n_new, int n_heads, int head_dim, int T,
                                  float* ctx, cudaStream_t st, const int* t_valid = nullptr /*[batch]: keys per clip (ragged batch)*/);
struct SelectArgs {
  float* logits; int vocab; int batch;
  const int* cand_idx;       // optional [B][vocab]: `logits` holds per-slice maxima and the token id is cand_idx[b][arg-max] (plain arg-max heads only)
  const float* begin_bias;   // [vocab] added on the fly for the prefill head only (nullable)
  int* cur_token;            // [B] token fed to the next step
  int* tokens; int tokens_ld;// [B][tokens_ld] accepted (non-stop) tokens
  int* n_gen;                // [B]
  int* finished;             // [B]
  int* save_id; int save_ld; // [B][save_ld] every selected id (GREEDY_SEARCH history)
  int* n_save;               // [B]
  int* selected_hist; int sel_ld;   // [B][sel_ld] selected id per launch (prefill = 0)
  const int* stop_ids; int n_stop;
  int limit;
  float penalty_value; int penalty_range;   // penalty_value == 1 -> plain argmax
  DecState* state; int n_new;               // state->kv_len += n_new, step += 1
  // TOPK_TOPP_SAMPLING (Export_Whisper.py:263-307); temperature <= 0 selects the argmax heads above
  float temperature; int top_k; float top_p; float rep_penalty; unsigned long long seed;
  const float* noise; int noise_ld; int noise_rows;   // optional uniform noise [launch][max_batch][top_k] (reproducible runs)
  int noise_batch;                                    // max_batch: the row stride of `noise` in utterances
  const int* limit_v;                                 // optional [B]: per-utterance generation limit, min(limit, limit_v[b]) (ragged Qwen3-ASR batches)
};
cudaError_t launch_select_token(const SelectArgs& a, cudaStream_t st);
cudaError_t launch_softmax_pick(const float* logits, const float* add_bias, int vocab, int batch, int index,
                                float* out, cudaStream_t st);

// decoder_mega.cu: persistent prefill + greedy-loop kernel (one cooperative launch)
struct MegaLayer {
  const void *qkv_w, *out_w, *cq_w, *cout_w, *fc1_w, *fc2_w;
  const float *qkv_b, *out_b, *cq_b, *cout_b, *fc1_b, *fc2_b;
};
// one contiguous region of the per-step read stream; `start` is its offset in the padded stream
// (every block is padded to a multiple of the super-chunk so a super-chunk never straddles blocks)
struct PfBlock { const char* ptr; long long bytes; long long padded; long long start; };

struct MegaArgs {
  const MegaLayer* layers; int n_layers;
  const void* embed; const float* pos; const float* ln_g; const float* ln_b;
  const float* suppress_bias; const float* begin_bias;
  void* kcache; void* vcache; const void* cross_kv; int T;
  const int* t_valid;                    // ragged batch: encoder positions per clip [batch] (device); nullptr = T for every clip
  // streaming kernel, sub-batch launch (multi-row prefill of a batch in groups): the launch covers clips [clip0, clip0 + batch) of a
  // batch of batch_stride clips -- per-clip arrays arrive offset by clip0, the K/V caches and cross-KV keep the whole batch's strides
  int batch_stride, clip0;               // batch_stride 0 = batch
  int batch, d, ffn, n_heads, vocab, max_target;
  float* x; float* q; float* ctx; float* f; float* logits;     // logits may be null
  const int* first_tokens; int first_n_new;                      // iteration 0: [B][first_n_new]
  int* cur_token; int* tokens; int tokens_ld; int* n_gen; int* finished;
  int* save_id; int save_ld; int* n_save; int* selected_hist; int sel_ld;
  const int* stop_ids; int n_stop; int limit; float penalty_value; int penalty_range;
  DecState* state;
  unsigned int* bar;
  float* cand_val; int* cand_idx;        // [gridDim.x][batch]
  int n_iters; int first_is_prefill;     // begin-suppress bias applies to iteration 0 only when set
  const PfBlock* pf_blocks; int n_pf_blocks; long long pf_total; long long pf_ahead;
  float eps;
  unsigned long long* timing; int timing_cap;   // optional: globaltimer after every grid barrier (block 0)
};

size_t mega_smem_bytes(int d, int ffn, int T, int max_target);
bool mega_supported(int batch, int first_n_new, int d, int ffn);
int mega_pf_piece();
cudaError_t launch_decoder_mega(const MegaArgs& a, int w_dtype, int num_sms, cudaStream_t st);

// decoder_ring.cu: streaming greedy-decode kernel (TMA weight ring + flag-in-data exchanges); bf16, batch <= 4
struct RingArgs {
  MegaArgs m;
  unsigned long long* ll; long long ll_stride;   // 4 exchange buffers of ll_stride 8-byte words each (zeroed per launch)
  int ld_vec;                                    // words per utterance row in an exchange buffer
  int n_stages, stage_bytes;                     // shared-memory ring
  int box_rows;                                  // cross-attention K/V rows per ring stage (TMA box)
  int tc;                                        // 1: mma.sync dot products (weight rows skewed by 16 B in the ring)
  int task_inv;                                  // inverse (mod grid) of the attention-task -> CTA stride
  int part_cap, sc_cap;
  int debug;                                     // timing experiments: skip parts of a phase (see decoder_ring.cu)
  int fine_timing;                               // stamp inside linear phases too (after gather / LN / ring / publish)
};
constexpr int kRingTaskMul = 7;                  // task t of layer l -> CTA (t * 7 + offset(l)) % grid
bool ring_supported(int batch, int d, int ffn, int n_heads, int vocab, int num_sms);
bool ring_plan(const MegaArgs& a, int num_sms, bool tc, RingArgs* ra, size_t* smem_bytes);
size_t ring_exchange_words(int batch, int d, int ffn, int num_sms);
cudaError_t launch_decoder_ring(const RingArgs& ra, const CUtensorMap& cross_map, int num_sms, size_t smem_bytes,
                                cudaStream_t st);
// decoder_stream.cu: split-K tensor-core streaming decode kernel (tcgen05 on a TMA-fed weight ring, fixed-point
// accumulate-in-L2 exchanges); bf16, batch <= 8.  The product path for the Whisper greedy loop and prefill.
struct StreamLayer {        // fp32 vectors of one decoder layer: biases and the LayerNorm-fold row sums (sum_k W[n][k])
  const float *qkv_b, *qkv_ws, *out_b, *cq_b, *cq_ws, *cout_b, *fc1_b, *fc1_ws, *fc2_b;
  const float *qkv_s, *out_s, *cq_s, *cout_s, *fc1_s, *fc2_s;     // FP8 weight path: per-row scales (nullptr in bf16 mode)
};
struct StreamArgs {
  MegaArgs m;                          // model dims, device pointers, token bookkeeping (shared with the other decoder kernels)
  const StreamLayer* sl;               // [L]
  const CUtensorMap* wmaps;            // [6L + 1] SWIZZLE_128B maps, box [128 rows][64 k]: qkv, out, cq, cout, fc1, fc2 per layer, then the tied head
  const float* head_g; const float* head_b;   // [vocab] sum_k E[n][k] gamma[k], sum_k E[n][k] beta[k] (final LayerNorm folded around the tied head)
  const float* head_s; int fp8;               // FP8 weight path (E4M3 weights + per-row scale, atoms of 128 k): head scales; 1 = on
  unsigned long long* acc;             // [2][set_words] accumulator words (12-bit count | 52-bit fixed point), zero at launch
  long long set_words, layer_words;
  unsigned long long* cand;            // [2][grid][NRT][2] flag-in-data arg-max candidates
  const int4* sched;                   // [grid][6L + 1] {atom begin, atom end, first tile, first k-atom}
  const unsigned char* cnt;            // [L][3][cnt_ld] contributors per 128-row output tile of qkv / cq / fc1
  const unsigned short* xexp;          // [L][3][xt] cumulative contributors per residual-stream tile after out / cout / fc2
  int cnt_ld, xt;
  int n_stages, n_slots;               // ring stages (16 KB each); B-operand k-atom slots
  int task_inv;                        // inverse (mod grid) of the attention-task -> CTA stride
  int l2_hint;                         // 1: weight / KV boxes are loaded with an L2 evict-first policy
  int multi;                           // 1: the m.first_n_new prompt positions of every clip run as rows of ONE iteration (single-iteration launch)
  int keep_state;                      // 1: leave DecState (kv_len, step) as it is -- another sub-batch launch of the same positions follows
  int lean;                            // 1: a plain greedy decode launch may take the instantiation with the rarely used branches compiled out
  int debug;
};
constexpr int kStreamMaxBatch = 8;
bool stream_supported(int batch, int d, int ffn, int n_heads, int vocab, int T, int num_sms);
// host-side plan: schedule tables (filled into the vectors), shared-memory size; false when the shape does not fit
struct StreamPlan {
  int n_stages = 0, n_slots = 0, cnt_ld = 0, xt = 0, nrt = 0;
  size_t smem_bytes = 0; long long set_words = 0, layer_words = 0; size_t cand_words = 0;
};
bool stream_plan(int batch, int d, int ffn, int n_heads, int vocab, int n_layers, int T, int max_target, int num_sms,
                 StreamPlan* plan, void* sched_out /*std::vector<int4>*/, void* cnt_out /*std::vector<unsigned char>*/,
                 void* xexp_out /*std::vector<unsigned short>*/, int fp8 = 0);
cudaError_t launch_quant_rows_e4m3(const void* W_bf16, void* W8, float* scale, int N, int K, cudaStream_t st);
cudaError_t launch_rowdot_e4m3(const void* W8, const float* scale, const float* vec /*nullable: ones*/, float* out, int N, int K, cudaStream_t st);
cudaError_t launch_decoder_stream(const StreamArgs& sa, const CUtensorMap& cross_map, const CUtensorMap& kc_map,
                                  const CUtensorMap& vc_map, int nrt, int num_sms, size_t smem_bytes, cudaStream_t st);
cudaError_t launch_rowdot_bf16(const void* W, const float* vec /*nullable: ones*/, float* out, int N, int K, cudaStream_t st);
// gemm_tc.cu: SWIZZLE_128B bf16 tensor map over [rows][ld] with a [box_rows][64] box (3-D form, batch 1)
bool make_tmap_rows_sw128(CUtensorMap* tm, const void* base, int64_t cols, int64_t rows, int64_t ld, int box_rows,
                          std::string* err);
// same over bytes (fp8 weights): [rows][ld] bytes with a [box_rows][128] box
bool make_tmap_rows_sw128_u8(CUtensorMap* tm, const void* base, int64_t cols, int64_t rows, int64_t ld, int box_rows,
                             std::string* err);
// attention_tc.cu: fused softmax(Q K^T) V per (utterance, head) on tcgen05; qkv bf16 [batch*T][3d], ctx bf16 [batch*T][d]
bool attention_tc_supported(int T, int d, int n_heads);
cudaError_t launch_attention_tc(const void* qkv, void* ctx, int batch, int T, int d, int n_heads, cudaStream_t st,
                                std::string* err, const int* kv_valid = nullptr, float mask_add = 0.f);
// gemm_tc.cu: plain (unswizzled) 2-D bf16 tensor map over [rows][ld] with a [box_rows][box_cols] box
bool make_tmap_2d_plain(CUtensorMap* tm, const void* base, int64_t cols, int64_t rows, int64_t ld, int box_cols,
                        int box_rows, std::string* err);

}  // namespace b200asr
