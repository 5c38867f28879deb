// Engine object + C ABI (include/b200asr.h).  Host-side orchestration only: every
// arithmetic step is one of the kernels in frontend.cu / gemm_tc.cu / gemm_simt.cu /
// layers.cu / decoder.cu, enqueued on the engine's single stream.
#include "common.cuh"
#include "../../include/b200asr.h"

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace b200asr;

namespace {

std::string g_create_error;

struct DevTensor {
  void* ptr = nullptr;
  int64_t numel = 0;
  int dtype = kF32;
};

__global__ void f32_to_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void bf16_to_f32_kernel(const bf16* __restrict__ in, float* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __bfloat162float(in[i]);
}

}  // namespace

struct b200asr_engine {
  b200asr_config cfg{};
  cudaStream_t st = nullptr;
  std::string err;
  int num_sms = 148;
  int64_t launches = 0;
  bool finalized = false;
  bool keep_stages = false;
  bool warned_tmap = false; int64_t tmap_fallbacks = 0;
  int act_dtype = kF32;      // activations fed to GEMMs, weights, KV caches
  size_t es = 4;

  std::map<std::string, DevTensor> w;
  float* stage_buf = nullptr; int64_t stage_cap = 0;     // fp32 staging for uploads / stage dumps
  float* basis_t = nullptr; int* fb_start = nullptr; int* fb_len = nullptr;

  // encoder state
  int B = 0, n_samples = 0, T_mel = 0, T_enc = 0, pcm_dtype = B200ASR_PCM_I16;
  void* pcm = nullptr; void* pcm_pinned = nullptr;
  // ragged batch (b200asr_*_ragged): samples per clip and encoder positions per clip, device [2][max_batch]; ragged = in force
  int* d_lens = nullptr; bool ragged = false; std::vector<int> h_lens;
  const int* dev_nsamp() const { return ragged ? d_lens : nullptr; }
  const int* dev_tvalid() const { return ragged ? d_lens + cfg.max_batch : nullptr; }
  float* mel_raw = nullptr; int* max_key = nullptr;
  void* mel_pad = nullptr; void* h1_pad = nullptr;
  float* hidden = nullptr; float* stem = nullptr;
  void* xhat = nullptr; void* qkv = nullptr; void* ctx = nullptr; void* ffn = nullptr;
  float* S = nullptr; void* P = nullptr;
  void* cross_kv = nullptr;
  // decoder state
  void* kcache = nullptr; void* vcache = nullptr;
  float *dx = nullptr, *dq = nullptr, *dctx = nullptr, *dffn = nullptr, *logits = nullptr, *prob = nullptr;
  int *d_prompt = nullptr, *cur_token = nullptr, *tokens = nullptr, *n_gen = nullptr, *finished = nullptr;
  int *save_id = nullptr, *n_save = nullptr, *selected_hist = nullptr, *d_stop = nullptr;
  DecState* dstate = nullptr;
  int* h_pinned = nullptr;          // pinned scratch for small D2H results
  size_t h_pinned_bytes = 0;
  std::vector<int> stop_ids; int limit = 0; int limit_cfg = 0; float repeat_penalty = 1.0f; int penalty_range = 20;
  int n_prompt = 0; bool prefilled = false; bool encoded = false;
  cudaGraphExec_t step_graph = nullptr; int64_t step_graph_nodes = 0;
  cudaGraphExec_t enc_graph = nullptr; int64_t enc_graph_nodes = 0; std::string enc_graph_key; bool use_enc_graph = true;
  // persistent decoder kernel state
  bool use_mega = true; long long pf_ahead = 0;
  MegaLayer* mega_layers = nullptr; PfBlock* pf_blocks = nullptr; int n_pf_blocks = 0; long long pf_total = 0;
  int pf_B = -1, pf_T = -1;
  unsigned int* mega_bar = nullptr; float* cand_val = nullptr; int* cand_idx = nullptr;
  bool mega_timing = false; unsigned long long* timing = nullptr; static constexpr int kTimingCap = 16384;
  // streaming decode kernel (decoder_ring.cu)
  bool use_attn_tc = true;
  bool use_pdl = true;
  // sampling head (TOPK_TOPP_SAMPLING); temperature <= 0 = argmax heads
  float samp_temperature = 0.f; int samp_top_k = 10; float samp_top_p = 0.95f; float samp_rep = 1.0f;
  unsigned long long samp_seed = 0; float* samp_noise = nullptr; int samp_noise_rows = 0, samp_noise_ld = 0;
  bool use_ring = true; bool ring_tc = false;   // mma.sync dot products: correct but measured slower than the CUDA-core path
   bool ring_fine = false; int ring_debug = 0; unsigned long long* ring_ll = nullptr; size_t ring_ll_words = 0;
  CUtensorMap cross_map{}; int cmap_B = -1, cmap_T = -1, cmap_rows = -1; int ring_task_inv = 0;
  // split-K tensor-core streaming decode kernel (decoder_stream.cu): the product path for prefill + greedy loop
  bool use_stream = true; bool stream_l2_hint = true; int stream_debug = 0; bool stream_multi = true;
  // per weight format (0: bf16, 1: FP8 E4M3 + per-row scale, `set_option("fp8", 1)`): tensor maps, fold vectors, schedule
  struct StreamTables {
    StreamLayer* layers = nullptr; CUtensorMap* wmaps = nullptr; float* fold = nullptr;   // fold vectors (row sums, head g / b)
    float* head_g = nullptr; float* head_b = nullptr; float* head_s = nullptr;
    uint8_t* w8 = nullptr; float* scales = nullptr;                                        // FP8: quantised decoder matrices + tied head, their row scales
    int4* sched = nullptr; unsigned char* cnt = nullptr; unsigned short* xexp = nullptr;
    StreamPlan plans[4]; bool plan_ok[4] = {false, false, false, false}; bool tables = false;   // one plan per row-count class (1, 2, 4, 8)
  } stt[2];
  bool use_fp8 = false;
  bool stream_lean = true;
  unsigned long long* st_acc = nullptr; size_t st_acc_words = 0; unsigned long long* st_cand = nullptr; size_t st_cand_words = 0;
  CUtensorMap st_cross{}, st_kc{}, st_vc{}; int st_map_B = -1, st_map_T = -1;
  std::string graph_key;

  int fail(int code, const std::string& m) { err = m; return code; }
  int cuda_fail(cudaError_t e, const char* what) {
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return B200ASR_E_CUDA;
  }
};


// Stream-ordered copy + sync on the engine's (non-blocking) stream.  A plain cudaMemcpy from
// pageable memory may return before the DMA lands and is only ordered against the legacy
// stream, so kernels on e->st could read a half-written buffer.
static cudaError_t b200_copy_sync(b200asr_engine* e, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
  cudaError_t r = cudaMemcpyAsync(dst, src, bytes, kind, e->st);
  if (r != cudaSuccess) return r;
  return cudaStreamSynchronize(e->st);
}

#define CK(expr)                                                           \
  do {                                                                     \
    cudaError_t _e = (expr);                                               \
    if (_e != cudaSuccess) return e->cuda_fail(_e, #expr);                 \
  } while (0)
#define KL(expr)                                                           \
  do {                                                                     \
    cudaError_t _e = (expr);                                               \
    e->launches++;                                                         \
    if (_e != cudaSuccess) return e->cuda_fail(_e, #expr);                 \
  } while (0)
#define RET(expr)                                                          \
  do {                                                                     \
    int _r = (expr);                                                       \
    if (_r != B200ASR_OK) return _r;                                       \
  } while (0)

namespace {

bool is_weight_matrix(const std::string& n) {
  return (n.size() > 2 && n.compare(n.size() - 2, 2, ".w") == 0) || n == "dec.embed";
}

int ensure_stage(b200asr_engine* e, int64_t numel) {
  if (numel <= e->stage_cap) return B200ASR_OK;
  if (e->stage_buf) cudaFree(e->stage_buf);
  e->stage_buf = nullptr; e->stage_cap = 0;
  CK(cudaMalloc(&e->stage_buf, (size_t)numel * sizeof(float)));
  e->stage_cap = numel;
  return B200ASR_OK;
}

template <typename T>
int dmalloc(b200asr_engine* e, T** p, size_t bytes) {
  CK(cudaMalloc(reinterpret_cast<void**>(p), bytes ? bytes : 16));
  CK(cudaMemsetAsync(*p, 0, bytes ? bytes : 16, e->st));
  return B200ASR_OK;
}

const DevTensor* find(b200asr_engine* e, const std::string& n) {
  auto it = e->w.find(n);
  return it == e->w.end() ? nullptr : &it->second;
}

int need(b200asr_engine* e, const std::string& n, int64_t numel) {
  const DevTensor* t = find(e, n);
  if (!t) return e->fail(B200ASR_E_MISSING, "missing weight tensor '" + n + "'");
  if (t->numel != numel)
    return e->fail(B200ASR_E_INVALID, "tensor '" + n + "' has " + std::to_string(t->numel) + " elements, expected " +
                                          std::to_string(numel));
  return B200ASR_OK;
}

const void* W(b200asr_engine* e, const std::string& n) { return e->w[n].ptr; }
const float* WF(b200asr_engine* e, const std::string& n) { return reinterpret_cast<const float*>(e->w[n].ptr); }

// one Linear on the encoder side: tcgen05 when possible, CUDA cores otherwise
int gemm(b200asr_engine* e, const GemmArgs& g) {
  if (e->act_dtype == kBF16 && e->cfg.use_tensor_cores && gemm_tc_supported(g)) {
    std::string msg;
    cudaError_t r = launch_gemm_tc(g, e->num_sms, e->st, &msg);
    if (r == cudaErrorNotSupported) {
      // the driver refused the tensor map (e.g. an overlapping strided view): CUDA-core GEMM for this call
      cudaGetLastError();
      if (!e->warned_tmap) { fprintf(stderr, "b200asr: %s; using the CUDA-core GEMM for this operand view\n", msg.c_str()); e->warned_tmap = true; }
      e->tmap_fallbacks++;
    } else {
      e->launches++;
      if (r != cudaSuccess) return e->fail(B200ASR_E_CUDA, "gemm_tc: " + (msg.empty() ? std::string(cudaGetErrorString(r)) : msg));
      return B200ASR_OK;
    }
  }
  KL(launch_gemm_simt(g, e->st));
  return B200ASR_OK;
}

GemmArgs linear_args(b200asr_engine* e, const void* A, int64_t lda, const std::string& wname, const std::string& bname,
                     void* C, int64_t ldc, int c_dtype, int M, int N, int K) {
  GemmArgs g;
  g.A = A; g.lda = lda; g.a_dtype = e->act_dtype;
  g.B = W(e, wname); g.ldb = K; g.b_dtype = e->act_dtype;
  g.C = C; g.ldc = ldc; g.c_dtype = c_dtype;
  g.bias = bname.empty() ? nullptr : WF(e, bname);
  g.M = M; g.N = N; g.K = K;
  g.pdl = e->use_pdl ? 1 : 0;          // B is a weight matrix: the tcgen05 GEMM may start under its predecessor's tail
  return g;
}

int enqueue_encoder(b200asr_engine* e) {
  const b200asr_config& c = e->cfg;
  const int B = e->B, d = c.d_model, H = c.n_heads, Tm = e->T_mel, T = e->T_enc, L = c.dec_layers;
  const int M = B * T;
  const int ad = e->act_dtype;
  const size_t es = e->es;
  // ---- front end ----
  KL(launch_fill_i32(e->max_key, INT_MIN, B, e->st));
  KL(launch_logmel(e->pcm, e->pcm_dtype == B200ASR_PCM_F32, B, e->n_samples, e->n_samples, e->basis_t,
                   WF(e, "mel_fbank"), e->fb_start, e->fb_len, c.n_fft, c.hop, c.n_mels, e->mel_raw, e->max_key, e->st, e->dev_nsamp()));
  // zero the conv padding rows (row 0 and row Tm+1 of every utterance)
  CK(cudaMemset2DAsync(e->mel_pad, (size_t)(Tm + 2) * c.n_mels * es, 0, (size_t)c.n_mels * es, B, e->st));
  CK(cudaMemset2DAsync((char*)e->mel_pad + (size_t)(Tm + 1) * c.n_mels * es, (size_t)(Tm + 2) * c.n_mels * es, 0,
                       (size_t)c.n_mels * es, B, e->st));
  CK(cudaMemset2DAsync(e->h1_pad, (size_t)(Tm + 2) * d * es, 0, (size_t)d * es, B, e->st));
  CK(cudaMemset2DAsync((char*)e->h1_pad + (size_t)(Tm + 1) * d * es, (size_t)(Tm + 2) * d * es, 0, (size_t)d * es, B, e->st));
  KL(launch_mel_finalize(e->mel_raw, e->max_key, B, Tm, c.n_mels, e->mel_pad, ad, e->st, e->dev_nsamp(), c.hop));
  // ---- conv stem as strided-view GEMMs (Export_Whisper.py:428-429) ----
  {
    GemmArgs g = linear_args(e, e->mel_pad, c.n_mels, "enc.conv1.w", "enc.conv1.b", (char*)e->h1_pad + (size_t)d * es, d,
                             ad, Tm, d, 3 * c.n_mels);
    g.sAo = (int64_t)(Tm + 2) * c.n_mels; g.sCo = (int64_t)(Tm + 2) * d; g.batch = B; g.act = kActGelu;
    RET(gemm(e, g));
    // ragged batch: conv2 must see its `padding=1` zero row right after each clip's own last frame
    if (e->ragged) KL(launch_zero_tail_rows(e->h1_pad, ad, e->dev_nsamp(), c.hop, B, Tm, d, e->st));
    GemmArgs g2 = linear_args(e, e->h1_pad, 2 * d, "enc.conv2.w", "enc.conv2.b", e->hidden, d, kF32, T, d, 3 * d);
    g2.sAo = (int64_t)(Tm + 2) * d; g2.sCo = (int64_t)T * d; g2.batch = B; g2.act = kActGelu;
    g2.residual = WF(e, "enc.pos"); g2.ldr = d; g2.sRo = 0;
    RET(gemm(e, g2));
  }
  if (e->keep_stages) CK(cudaMemcpyAsync(e->stem, e->hidden, (size_t)M * d * 4, cudaMemcpyDeviceToDevice, e->st));
  // ---- encoder layers (Export_Whisper.py:430-437) ----
  for (int l = 0; l < c.enc_layers; ++l) {
    const std::string p = "enc.L" + std::to_string(l) + ".";
    KL(launch_layernorm(e->hidden, d, nullptr, nullptr, e->xhat, ad, d, M, d, 1e-5f, e->st, e->use_pdl ? 1 : 0));
    RET(gemm(e, linear_args(e, e->xhat, d, p + "qkv.w", p + "qkv.b", e->qkv, 3 * d, ad, M, 3 * d, d)));
    if (ad == kBF16 && c.use_tensor_cores && e->use_attn_tc && attention_tc_supported(T, d, H)) {
      // fused softmax(Q K^T) V on tcgen05: scores and probabilities never leave the SM
      std::string msg;
      cudaError_t r = launch_attention_tc(e->qkv, e->ctx, B, T, d, H, e->st, &msg, e->dev_tvalid(), -1e30f);
      e->launches++;
      if (r != cudaSuccess) return e->fail(B200ASR_E_CUDA, "attention_tc: " + (msg.empty() ? std::string(cudaGetErrorString(r)) : msg));
    } else {  // per-(utterance, head) softmax(Q K^T) V ; scale pre-folded into q and k
      GemmArgs s;
      s.A = e->qkv; s.lda = 3 * d; s.sAo = (int64_t)T * 3 * d; s.sAi = 64; s.a_dtype = ad;
      s.B = (char*)e->qkv + (size_t)d * es; s.ldb = 3 * d; s.sBo = (int64_t)T * 3 * d; s.sBi = 64; s.b_dtype = ad;
      s.C = e->S; s.ldc = T; s.sCo = (int64_t)H * T * T; s.sCi = (int64_t)T * T; s.c_dtype = kF32;
      s.M = T; s.N = T; s.K = 64; s.batch = B * H; s.batch_inner = H;
      KL(launch_gemm_simt(s, e->st));
      KL(launch_softmax_rows(e->S, e->P, ad, (int64_t)B * H * T, T, e->st, e->dev_tvalid(), (int64_t)H * T));
      GemmArgs o;
      o.A = e->P; o.lda = T; o.sAo = (int64_t)H * T * T; o.sAi = (int64_t)T * T; o.a_dtype = ad;
      o.B = (char*)e->qkv + (size_t)2 * d * es; o.ldb = 3 * d; o.sBo = (int64_t)T * 3 * d; o.sBi = 64; o.b_dtype = ad;
      o.transB = 1;
      o.C = e->ctx; o.ldc = d; o.sCo = (int64_t)T * d; o.sCi = 64; o.c_dtype = ad;
      o.M = T; o.N = 64; o.K = T; o.batch = B * H; o.batch_inner = H;
      KL(launch_gemm_simt(o, e->st));
    }
    {
      GemmArgs g = linear_args(e, e->ctx, d, p + "out.w", p + "out.b", e->hidden, d, kF32, M, d, d);
      g.residual = e->hidden; g.ldr = d;
      RET(gemm(e, g));
    }
    KL(launch_layernorm(e->hidden, d, nullptr, nullptr, e->xhat, ad, d, M, d, 1e-5f, e->st, e->use_pdl ? 1 : 0));
    {
      GemmArgs g = linear_args(e, e->xhat, d, p + "fc1.w", p + "fc1.b", e->ffn, c.ffn, ad, M, c.ffn, d);
      g.act = kActGelu;
      RET(gemm(e, g));
      GemmArgs g2 = linear_args(e, e->ffn, c.ffn, p + "fc2.w", p + "fc2.b", e->hidden, d, kF32, M, d, c.ffn);
      g2.residual = e->hidden; g2.ldr = d;
      RET(gemm(e, g2));
    }
  }
  // ---- final LN (affine) + fused cross-KV projection (Export_Whisper.py:438-447) ----
  KL(launch_layernorm(e->hidden, d, WF(e, "enc.ln_post.g"), WF(e, "enc.ln_post.b"), e->xhat, ad, d, M, d, 1e-5f, e->st, e->use_pdl ? 1 : 0));
  {  // 2L projections of the same rows: batched over z = (kind, layer) so each layer's K / V lands contiguous:
     // cross_kv[z][b*T + t][d], z = l for K (pre-scaled), z = L + l for V
    GemmArgs g = linear_args(e, e->xhat, d, "enc.cross_kv.w", "enc.cross_kv.b", e->cross_kv, d, ad, M, d, d);
    g.batch = 2 * L; g.sAo = 0; g.sBo = (int64_t)d * d; g.sCo = (int64_t)M * d; g.sBias = d;
    RET(gemm(e, g));
  }
  return B200ASR_OK;
}

// The encoder as one CUDA graph per (batch, clip length, options): ~230 launches, each GEMM's two tensor maps encoded on the host,
// are built once instead of per call (the maps are kernel parameters, so they live in the graph's nodes).
int run_encoder(b200asr_engine* e) {
  char key[160];
  snprintf(key, sizeof key, "%d/%d/%d/%d/%d/%d/%d/%d", e->B, e->n_samples, e->pcm_dtype, e->keep_stages ? 1 : 0, e->use_attn_tc ? 1 : 0,
           e->use_pdl ? 1 : 0, e->cfg.use_tensor_cores, e->ragged ? 1 : 0);   // (per-clip lengths are read from device memory: one graph serves every ragged batch of this shape)
  if (e->use_enc_graph && e->act_dtype == kBF16) {
    if (!e->enc_graph || e->enc_graph_key != key) {
      if (e->enc_graph) { cudaGraphExecDestroy(e->enc_graph); e->enc_graph = nullptr; }
      cudaGraph_t graph = nullptr;
      const int64_t before = e->launches;
      CK(cudaStreamBeginCapture(e->st, cudaStreamCaptureModeThreadLocal));
      const int r = enqueue_encoder(e);
      const cudaError_t ce = cudaStreamEndCapture(e->st, &graph);
      e->enc_graph_nodes = e->launches - before;
      e->launches = before;
      if (r != B200ASR_OK) { if (graph) cudaGraphDestroy(graph); return r; }
      if (ce != cudaSuccess) return e->cuda_fail(ce, "cudaStreamEndCapture (encoder)");
      const cudaError_t ci = cudaGraphInstantiate(&e->enc_graph, graph, 0);
      cudaGraphDestroy(graph);
      if (ci != cudaSuccess) { e->enc_graph = nullptr; return e->cuda_fail(ci, "cudaGraphInstantiate (encoder)"); }
      e->enc_graph_key = key;
    }
    CK(cudaGraphLaunch(e->enc_graph, e->st));
    e->launches += e->enc_graph_nodes;
  } else {
    RET(enqueue_encoder(e));
  }
  e->encoded = true;
  e->prefilled = false;
  return B200ASR_OK;
}

// enqueue one decoder launch over n_new tokens per utterance (tokens on device, [B][n_new])
int enqueue_decoder(b200asr_engine* e, const int* tokens_dev, int n_new, bool first) {
  const b200asr_config& c = e->cfg;
  const int B = e->B, d = c.d_model, H = c.n_heads, L = c.dec_layers, T = e->T_enc;
  const int rows = B * n_new;
  const int wd = e->act_dtype;
  KL(launch_dec_embed(tokens_dev, W(e, "dec.embed"), wd, W(e, "dec.pos"), B, n_new, d, e->dstate, e->dx, e->st));
  const size_t layer_kv = (size_t)B * H * c.max_target * 64 * e->es;
  for (int l = 0; l < L; ++l) {
    const std::string p = "dec.L" + std::to_string(l) + ".";
    DecLinearArgs a{};
    a.eps = 1e-5f; a.w_dtype = wd; a.n_new = n_new; a.n_heads = H; a.head_dim = 64; a.max_target = c.max_target;
    a.batch = B; a.state = e->dstate; a.rows = rows; a.kv_dtype = wd;
    // LN + fused QKV, K/V appended straight into the resident cache
    DecLinearArgs q = a;
    q.x = e->dx; q.ldx = d; q.ln_mode = 1; q.W = W(e, p + "qkv.w"); q.bias = WF(e, p + "qkv.b");
    q.out = e->dq; q.ldo = d; q.mode = 1; q.kcache = (char*)e->kcache + l * layer_kv; q.vcache = (char*)e->vcache + l * layer_kv;
    q.N = 3 * d; q.K = d;
    KL(launch_dec_linear(q, e->st));
    KL(launch_dec_self_attn(e->dq, q.kcache, q.vcache, wd, B, n_new, H, 64, c.max_target, e->dstate, e->dctx, e->st));
    DecLinearArgs o = a;
    o.x = e->dctx; o.ldx = d; o.W = W(e, p + "out.w"); o.bias = WF(e, p + "out.b"); o.residual = e->dx; o.ldr = d;
    o.out = e->dx; o.ldo = d; o.N = d; o.K = d;
    KL(launch_dec_linear(o, e->st));
    DecLinearArgs cq = a;
    cq.x = e->dx; cq.ldx = d; cq.ln_mode = 1; cq.W = W(e, p + "cq.w"); cq.bias = WF(e, p + "cq.b");
    cq.out = e->dq; cq.ldo = d; cq.N = d; cq.K = d;
    KL(launch_dec_linear(cq, e->st));
    KL(launch_dec_cross_attn(e->dq, e->cross_kv, wd, l, L, B, n_new, H, 64, T, e->dctx, e->st, e->dev_tvalid()));
    DecLinearArgs co = a;
    co.x = e->dctx; co.ldx = d; co.W = W(e, p + "cout.w"); co.bias = WF(e, p + "cout.b"); co.residual = e->dx; co.ldr = d;
    co.out = e->dx; co.ldo = d; co.N = d; co.K = d;
    KL(launch_dec_linear(co, e->st));
    DecLinearArgs f1 = a;
    f1.x = e->dx; f1.ldx = d; f1.ln_mode = 1; f1.W = W(e, p + "fc1.w"); f1.bias = WF(e, p + "fc1.b"); f1.act = kActGelu;
    f1.out = e->dffn; f1.ldo = c.ffn; f1.N = c.ffn; f1.K = d;
    KL(launch_dec_linear(f1, e->st));
    DecLinearArgs f2 = a;
    f2.x = e->dffn; f2.ldx = c.ffn; f2.W = W(e, p + "fc2.w"); f2.bias = WF(e, p + "fc2.b"); f2.residual = e->dx; f2.ldr = d;
    f2.out = e->dx; f2.ldo = d; f2.N = d; f2.K = c.ffn;
    KL(launch_dec_linear(f2, e->st));
  }
  {  // last-token LN (affine) + tied lm head + permanent suppress bias (Export_Whisper.py:663-666)
    DecLinearArgs h{};
    h.eps = 1e-5f; h.w_dtype = wd; h.rows = B; h.state = e->dstate;
    h.x = e->dx + (size_t)(n_new - 1) * d; h.ldx = (int64_t)n_new * d; h.ln_mode = 2;
    h.gamma = WF(e, "dec.ln.g"); h.beta = WF(e, "dec.ln.b");
    h.W = W(e, "dec.embed"); h.bias = WF(e, "dec.suppress_bias"); h.out = e->logits; h.ldo = c.vocab;
    h.N = c.vocab; h.K = d;
    KL(launch_dec_linear(h, e->st));
  }
  SelectArgs s{};
  s.logits = e->logits; s.vocab = c.vocab; s.batch = B;
  s.begin_bias = first ? WF(e, "dec.begin_suppress_bias") : nullptr;
  s.cur_token = e->cur_token; s.tokens = e->tokens; s.tokens_ld = c.max_target; s.n_gen = e->n_gen;
  s.finished = e->finished; s.save_id = e->save_id; s.save_ld = c.max_target; s.n_save = e->n_save;
  s.selected_hist = e->selected_hist; s.sel_ld = c.max_target;
  s.stop_ids = e->d_stop; s.n_stop = (int)e->stop_ids.size(); s.limit = e->limit;
  s.penalty_value = e->repeat_penalty; s.penalty_range = e->penalty_range;
  s.state = e->dstate; s.n_new = n_new;
  s.temperature = e->samp_temperature; s.top_k = e->samp_top_k; s.top_p = e->samp_top_p; s.rep_penalty = e->samp_rep;
  s.seed = e->samp_seed; s.noise = e->samp_noise; s.noise_ld = e->samp_noise_ld; s.noise_rows = e->samp_noise_rows;
  s.noise_batch = e->cfg.max_batch;
  KL(launch_select_token(s, e->st));
  e->launches++;   // select = 2 kernels
  return B200ASR_OK;
}

int ensure_step_graph(b200asr_engine* e) {
  char key[384];
  snprintf(key, sizeof key, "%d/%d/%d/%g/%d/%zu/%g/%d/%g/%g/%llu/%p/%d", e->B, e->T_enc, e->limit, e->repeat_penalty, e->penalty_range,
           e->stop_ids.size(), e->samp_temperature, e->samp_top_k, e->samp_top_p, e->samp_rep, e->samp_seed, (void*)e->samp_noise,
           e->ragged ? 1 : 0);          // (the per-clip key-count pointer of the cross-attention launches is baked into the nodes)
  if (e->step_graph && e->graph_key == key) return B200ASR_OK;
  if (e->step_graph) { cudaGraphExecDestroy(e->step_graph); e->step_graph = nullptr; }
  cudaGraph_t graph = nullptr;
  const int64_t before = e->launches;
  CK(cudaStreamBeginCapture(e->st, cudaStreamCaptureModeThreadLocal));
  int r = enqueue_decoder(e, e->cur_token, 1, false);
  cudaError_t ce = cudaStreamEndCapture(e->st, &graph);
  e->step_graph_nodes = e->launches - before;
  e->launches = before;
  if (r != B200ASR_OK) { if (graph) cudaGraphDestroy(graph); return r; }
  if (ce != cudaSuccess) return e->cuda_fail(ce, "cudaStreamEndCapture");
  ce = cudaGraphInstantiate(&e->step_graph, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) return e->cuda_fail(ce, "cudaGraphInstantiate");
  e->graph_key = key;
  return B200ASR_OK;
}

int launch_step(b200asr_engine* e) {
  RET(ensure_step_graph(e));
  CK(cudaGraphLaunch(e->step_graph, e->st));
  e->launches += e->step_graph_nodes;
  return B200ASR_OK;
}


// ---- persistent decoder kernel plumbing ------------------------------------------------------
bool mega_ok(b200asr_engine* e, int first_n_new) {
  return e->use_mega && e->samp_temperature <= 0.f && mega_supported(e->B, first_n_new, e->cfg.d_model, e->cfg.ffn) && e->penalty_range <= 32 &&
         e->B * (first_n_new > 1 ? first_n_new : 1) <= 64 &&
         mega_smem_bytes(e->cfg.d_model, e->cfg.ffn, e->T_enc, e->cfg.max_target) <= 220 * 1024;
}

int build_mega_tables(b200asr_engine* e) {
  const b200asr_config& c = e->cfg;
  const int L = c.dec_layers;
  const long long d = c.d_model, f = c.ffn;
  const long long es = (long long)e->es;
  if (!e->mega_layers) {
    std::vector<MegaLayer> hl(L);
    for (int l = 0; l < L; ++l) {
      const std::string p = "dec.L" + std::to_string(l) + ".";
      hl[l].qkv_w = W(e, p + "qkv.w"); hl[l].out_w = W(e, p + "out.w"); hl[l].cq_w = W(e, p + "cq.w");
      hl[l].cout_w = W(e, p + "cout.w"); hl[l].fc1_w = W(e, p + "fc1.w"); hl[l].fc2_w = W(e, p + "fc2.w");
      hl[l].qkv_b = WF(e, p + "qkv.b"); hl[l].out_b = WF(e, p + "out.b"); hl[l].cq_b = WF(e, p + "cq.b");
      hl[l].cout_b = WF(e, p + "cout.b"); hl[l].fc1_b = WF(e, p + "fc1.b"); hl[l].fc2_b = WF(e, p + "fc2.b");
    }
    CK(cudaMalloc(&e->mega_layers, sizeof(MegaLayer) * L));
    CK(b200_copy_sync(e, e->mega_layers, hl.data(), sizeof(MegaLayer) * L, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&e->pf_blocks, sizeof(PfBlock) * (8 * L + 1)));
    CK(cudaMalloc(&e->mega_bar, 64));
    CK(cudaMalloc(&e->cand_val, sizeof(float) * e->num_sms * c.max_batch));
    CK(cudaMalloc(&e->cand_idx, sizeof(int) * e->num_sms * c.max_batch));
  }
  if (e->pf_B == e->B && e->pf_T == e->T_enc) return B200ASR_OK;
  // per-step read stream in phase order: qkv, out, cq, crossK, crossV, cout, fc1, fc2 per layer, then the lm head
  std::vector<PfBlock> hb;
  const long long super = (long long)e->num_sms * mega_pf_piece();
  long long pos = 0;
  auto add = [&](const void* ptr, long long bytes) {
    PfBlock b; b.ptr = (const char*)ptr; b.bytes = bytes; b.padded = (bytes + super - 1) / super * super; b.start = pos;
    pos += b.padded; hb.push_back(b);
  };
  const long long kvb = (long long)e->B * e->T_enc * d * es;
  for (int l = 0; l < L; ++l) {
    const std::string p = "dec.L" + std::to_string(l) + ".";
    add(W(e, p + "qkv.w"), 3 * d * d * es); add(W(e, p + "out.w"), d * d * es); add(W(e, p + "cq.w"), d * d * es);
    add((const char*)e->cross_kv + (long long)l * kvb, kvb); add((const char*)e->cross_kv + (long long)(L + l) * kvb, kvb);
    add(W(e, p + "cout.w"), d * d * es); add(W(e, p + "fc1.w"), f * d * es); add(W(e, p + "fc2.w"), d * f * es);
  }
  add(W(e, "dec.embed"), (long long)c.vocab * d * es);
  e->n_pf_blocks = (int)hb.size(); e->pf_total = pos;
  CK(b200_copy_sync(e, e->pf_blocks, hb.data(), sizeof(PfBlock) * hb.size(), cudaMemcpyHostToDevice));
  e->pf_B = e->B; e->pf_T = e->T_enc;
  return B200ASR_OK;
}

void fill_mega_args(b200asr_engine* e, MegaArgs& a, int n_iters, const int* first_tokens, int first_n_new,
                    bool first_is_prefill, bool want_logits) {
  const b200asr_config& c = e->cfg;
  a.layers = e->mega_layers; a.n_layers = c.dec_layers;
  a.embed = W(e, "dec.embed"); a.pos = WF(e, "dec.pos"); a.ln_g = WF(e, "dec.ln.g"); a.ln_b = WF(e, "dec.ln.b");
  a.suppress_bias = WF(e, "dec.suppress_bias"); a.begin_bias = WF(e, "dec.begin_suppress_bias");
  a.kcache = e->kcache; a.vcache = e->vcache; a.cross_kv = e->cross_kv; a.T = e->T_enc; a.t_valid = e->dev_tvalid();
  a.batch = e->B; a.d = c.d_model; a.ffn = c.ffn; a.n_heads = c.n_heads; a.vocab = c.vocab; a.max_target = c.max_target;
  a.x = e->dx; a.q = e->dq; a.ctx = e->dctx; a.f = e->dffn; a.logits = want_logits ? e->logits : nullptr;
  a.first_tokens = first_tokens; a.first_n_new = first_n_new;
  a.cur_token = e->cur_token; a.tokens = e->tokens; a.tokens_ld = c.max_target; a.n_gen = e->n_gen; a.finished = e->finished;
  a.save_id = e->save_id; a.save_ld = c.max_target; a.n_save = e->n_save; a.selected_hist = e->selected_hist; a.sel_ld = c.max_target;
  a.stop_ids = e->d_stop; a.n_stop = (int)e->stop_ids.size(); a.limit = e->limit;
  a.penalty_value = e->repeat_penalty; a.penalty_range = e->penalty_range;
  a.state = e->dstate; a.bar = e->mega_bar; a.cand_val = e->cand_val; a.cand_idx = e->cand_idx;
  a.n_iters = n_iters; a.first_is_prefill = first_is_prefill ? 1 : 0;
  a.pf_blocks = e->pf_blocks; a.n_pf_blocks = e->n_pf_blocks;
  a.pf_total = e->pf_ahead > 0 ? e->pf_total : 0; a.pf_ahead = e->pf_ahead;
  a.eps = 1e-5f;
  a.timing = nullptr; a.timing_cap = 0;
}

int arm_timing(b200asr_engine* e, MegaArgs& a) {
  if (!e->mega_timing) return B200ASR_OK;
  if (!e->timing) CK(cudaMalloc(&e->timing, sizeof(unsigned long long) * b200asr_engine::kTimingCap));
  CK(cudaMemsetAsync(e->timing, 0, sizeof(unsigned long long) * b200asr_engine::kTimingCap, e->st));
  a.timing = e->timing; a.timing_cap = b200asr_engine::kTimingCap;
  return B200ASR_OK;
}

// one cooperative launch: iteration 0 consumes first_tokens [B][first_n_new], later iterations feed back the argmax
int run_mega(b200asr_engine* e, int n_iters, const int* first_tokens, int first_n_new, bool first_is_prefill,
             bool want_logits) {
  RET(build_mega_tables(e));
  MegaArgs a{};
  fill_mega_args(e, a, n_iters, first_tokens, first_n_new, first_is_prefill, want_logits);
  RET(arm_timing(e, a));
  CK(cudaMemsetAsync(e->mega_bar, 0, 64, e->st));
  KL(launch_decoder_mega(a, e->act_dtype, e->num_sms, e->st));
  return B200ASR_OK;
}

// ---- streaming decode kernel plumbing (decoder_ring.cu) ----------------------------------------
bool ring_ok(b200asr_engine* e) {
  const b200asr_config& c = e->cfg;
  if (!e->use_ring || !e->use_mega || e->samp_temperature > 0.f || e->act_dtype != kBF16 || e->penalty_range > 32) return false;
  if (e->ragged) return false;                   // round-1 kernel: no per-clip key count (the tensor-core kernel and the barrier kernel have one)
  if (e->num_sms % kRingTaskMul == 0) return false;
  if (!ring_supported(e->B, c.d_model, c.ffn, c.n_heads, c.vocab, e->num_sms)) return false;
  if (c.d_model / 8 > 256) return false;         // TMA box rows
  MegaArgs a{}; a.batch = e->B; a.d = c.d_model; a.ffn = c.ffn; a.vocab = c.vocab; a.T = e->T_enc; a.max_target = c.max_target;
  RingArgs ra{}; size_t smem = 0;
  const bool tc = e->ring_tc && !e->ring_fine && (!e->ring_debug || e->ring_debug == 32) && !e->mega_timing;
  return ring_plan(a, e->num_sms, tc, &ra, &smem);
}

// n_iters single-token iterations starting from cur_token (the prefill has run)
int run_ring(b200asr_engine* e, int n_iters, bool want_logits) {
  const b200asr_config& c = e->cfg;
  RET(build_mega_tables(e));
  RingArgs ra{};
  fill_mega_args(e, ra.m, n_iters, e->cur_token, 1, false, want_logits);
  RET(arm_timing(e, ra.m));
  size_t smem = 0;
  const bool tc = e->ring_tc && !e->ring_fine && (!e->ring_debug || e->ring_debug == 32) && !e->mega_timing;
  if (!ring_plan(ra.m, e->num_sms, tc, &ra, &smem)) return e->fail(B200ASR_E_INVALID, "decoder_ring: shared-memory plan does not fit");
  const size_t words = ring_exchange_words(c.max_batch < 4 ? c.max_batch : 4, c.d_model, c.ffn, e->num_sms);
  if (!e->ring_ll) {
    CK(cudaMalloc(&e->ring_ll, words * 4 * sizeof(unsigned long long)));
    e->ring_ll_words = words;
    int inv = 0;
    for (int i = 1; i < e->num_sms; ++i) if ((i * kRingTaskMul) % e->num_sms == 1) inv = i;
    e->ring_task_inv = inv;
  }
  if (e->cmap_B != e->B || e->cmap_T != e->T_enc || e->cmap_rows != ra.box_rows) {
    std::string msg;
    const int64_t rows = (int64_t)2 * c.dec_layers * e->B * e->T_enc;
    if (!make_tmap_2d_plain(&e->cross_map, e->cross_kv, c.d_model, rows, c.d_model, 64, ra.box_rows, &msg))
      return e->fail(B200ASR_E_CUDA, "decoder_ring: " + msg);
    e->cmap_B = e->B; e->cmap_T = e->T_enc; e->cmap_rows = ra.box_rows;
  }
  ra.ll = e->ring_ll; ra.ll_stride = (long long)e->ring_ll_words;
  ra.ld_vec = c.d_model * 3 > c.ffn ? c.d_model * 3 : c.ffn;
  ra.task_inv = e->ring_task_inv;
  ra.fine_timing = e->ring_fine ? 1 : 0;
  ra.debug = e->ring_debug;
  CK(cudaMemsetAsync(e->ring_ll, 0, e->ring_ll_words * 4 * sizeof(unsigned long long), e->st));
  KL(launch_decoder_ring(ra, e->cross_map, e->num_sms, smem, e->st));
  return B200ASR_OK;
}

// ---- split-K tensor-core streaming decode kernel plumbing (decoder_stream.cu) -------------------
bool stream_ok(b200asr_engine* e) {
  const b200asr_config& c = e->cfg;
  if (!e->use_stream || e->samp_temperature > 0.f || e->act_dtype != kBF16 || e->penalty_range > 32) return false;
  if (!stream_supported(e->B, c.d_model, c.ffn, c.n_heads, c.vocab, e->T_enc, e->num_sms)) return false;
  const int slot = e->B <= 1 ? 0 : (e->B <= 2 ? 1 : (e->B <= 4 ? 2 : 3));
  if (e->stt[e->use_fp8 ? 1 : 0].plan_ok[slot]) return true;
  StreamPlan pl;
  return stream_plan(e->B, c.d_model, c.ffn, c.n_heads, c.vocab, c.dec_layers, e->T_enc, c.max_target, e->num_sms, &pl,
                     nullptr, nullptr, nullptr, e->use_fp8 ? 1 : 0);
}

int build_stream_tables(b200asr_engine* e, int rows) {
  const b200asr_config& c = e->cfg;
  const int L = c.dec_layers, d = c.d_model, f = c.ffn;
  const bool fp8 = e->use_fp8;
  b200asr_engine::StreamTables& t = e->stt[fp8 ? 1 : 0];
  if (!t.layers) {
    // LayerNorm-fold operands: row sums of the LN-consuming matrices, and the final LayerNorm folded around the tied head
    const size_t per_layer = (size_t)3 * d + d + f;
    CK(cudaMalloc(&t.fold, (per_layer * L + 2 * (size_t)c.vocab) * sizeof(float)));
    std::vector<StreamLayer> hl(L);
    std::vector<CUtensorMap> maps((size_t)6 * L + 1);
    std::string msg;
    const char* names[6] = {"qkv.w", "out.w", "cq.w", "cout.w", "fc1.w", "fc2.w"};
    const int rowsN[6] = {3 * d, d, d, d, f, d}, cols[6] = {d, d, d, d, d, f};
    const size_t layer_rows = (size_t)3 * d + 3 * (size_t)d + f + d;           // weight rows of one layer = scale entries
    const size_t layer_bytes = (size_t)3 * d * d + 3 * (size_t)d * d + 2 * (size_t)f * d;
    if (fp8) {
      // FP8 weight path (the reference's q8 plans: Optimize_ONNX_Common.py:3860-4100): E4M3 with one scale per
      // weight row, quantised once from the bf16 matrices; halves the bytes a greedy step streams
      CK(cudaMalloc(&t.w8, layer_bytes * L + (size_t)c.vocab * d));
      CK(cudaMalloc(&t.scales, (layer_rows * L + (size_t)c.vocab) * sizeof(float)));
    }
    for (int l = 0; l < L; ++l) {
      const std::string p = "dec.L" + std::to_string(l) + ".";
      float* base = t.fold + per_layer * l;
      hl[l] = StreamLayer{};
      hl[l].qkv_b = WF(e, p + "qkv.b"); hl[l].qkv_ws = base; hl[l].out_b = WF(e, p + "out.b");
      hl[l].cq_b = WF(e, p + "cq.b"); hl[l].cq_ws = base + 3 * d; hl[l].cout_b = WF(e, p + "cout.b");
      hl[l].fc1_b = WF(e, p + "fc1.b"); hl[l].fc1_ws = base + 4 * d; hl[l].fc2_b = WF(e, p + "fc2.b");
      if (!fp8) {
        KL(launch_rowdot_bf16(W(e, p + "qkv.w"), nullptr, base, 3 * d, d, e->st));
        KL(launch_rowdot_bf16(W(e, p + "cq.w"), nullptr, base + 3 * d, d, d, e->st));
        KL(launch_rowdot_bf16(W(e, p + "fc1.w"), nullptr, base + 4 * d, f, d, e->st));
        for (int i = 0; i < 6; ++i)
          if (!make_tmap_rows_sw128(&maps[(size_t)l * 6 + i], W(e, p + names[i]), cols[i], rowsN[i], cols[i], 128, &msg))
            return e->fail(B200ASR_E_CUDA, "decoder_stream: " + msg);
      } else {
        uint8_t* wq = t.w8 + layer_bytes * l;
        float* sc = t.scales + layer_rows * l;
        const float* scs[6];
        for (int i = 0; i < 6; ++i) {
          KL(launch_quant_rows_e4m3(W(e, p + names[i]), wq, sc, rowsN[i], cols[i], e->st));
          if (!make_tmap_rows_sw128_u8(&maps[(size_t)l * 6 + i], wq, cols[i], rowsN[i], cols[i], 128, &msg))
            return e->fail(B200ASR_E_CUDA, "decoder_stream: " + msg);
          if (i == 0) KL(launch_rowdot_e4m3(wq, sc, nullptr, base, 3 * d, d, e->st));
          if (i == 2) KL(launch_rowdot_e4m3(wq, sc, nullptr, base + 3 * d, d, d, e->st));
          if (i == 4) KL(launch_rowdot_e4m3(wq, sc, nullptr, base + 4 * d, f, d, e->st));
          scs[i] = sc;
          wq += (size_t)rowsN[i] * cols[i]; sc += rowsN[i];
        }
        hl[l].qkv_s = scs[0]; hl[l].out_s = scs[1]; hl[l].cq_s = scs[2]; hl[l].cout_s = scs[3]; hl[l].fc1_s = scs[4]; hl[l].fc2_s = scs[5];
      }
    }
    t.head_g = t.fold + per_layer * L; t.head_b = t.head_g + c.vocab;
    if (!fp8) {
      if (!make_tmap_rows_sw128(&maps[(size_t)6 * L], W(e, "dec.embed"), d, c.vocab, d, 128, &msg))
        return e->fail(B200ASR_E_CUDA, "decoder_stream: " + msg);
      KL(launch_rowdot_bf16(W(e, "dec.embed"), WF(e, "dec.ln.g"), t.head_g, c.vocab, d, e->st));
      KL(launch_rowdot_bf16(W(e, "dec.embed"), WF(e, "dec.ln.b"), t.head_b, c.vocab, d, e->st));
    } else {
      uint8_t* wq = t.w8 + layer_bytes * L;
      t.head_s = t.scales + layer_rows * L;
      KL(launch_quant_rows_e4m3(W(e, "dec.embed"), wq, t.head_s, c.vocab, d, e->st));
      if (!make_tmap_rows_sw128_u8(&maps[(size_t)6 * L], wq, d, c.vocab, d, 128, &msg))
        return e->fail(B200ASR_E_CUDA, "decoder_stream: " + msg);
      KL(launch_rowdot_e4m3(wq, t.head_s, WF(e, "dec.ln.g"), t.head_g, c.vocab, d, e->st));
      KL(launch_rowdot_e4m3(wq, t.head_s, WF(e, "dec.ln.b"), t.head_b, c.vocab, d, e->st));
    }
    CK(cudaMalloc(&t.layers, sizeof(StreamLayer) * L));
    CK(b200_copy_sync(e, t.layers, hl.data(), sizeof(StreamLayer) * L, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&t.wmaps, sizeof(CUtensorMap) * maps.size()));
    CK(b200_copy_sync(e, t.wmaps, maps.data(), sizeof(CUtensorMap) * maps.size(), cudaMemcpyHostToDevice));
    int inv = 0;
    for (int i = 1; i < e->num_sms; ++i) if ((i * kRingTaskMul) % e->num_sms == 1) inv = i;
    e->ring_task_inv = inv;
  }
  const int slot = rows <= 1 ? 0 : (rows <= 2 ? 1 : (rows <= 4 ? 2 : 3));
  if (!t.plan_ok[slot]) {
    // the schedule tables depend on the model dimensions only; the plan (shared-memory split, accumulator sizes) on the row class
    std::vector<int4> sched; std::vector<unsigned char> cnt; std::vector<unsigned short> xexp;
    StreamPlan pl;
    if (!stream_plan(rows, d, f, c.n_heads, c.vocab, L, e->T_enc, c.max_target, e->num_sms, &pl, &sched, &cnt, &xexp, fp8 ? 1 : 0))
      return e->fail(B200ASR_E_INVALID, "decoder_stream: plan does not fit");
    CK(cudaStreamSynchronize(e->st));
    if (!t.tables) {
      CK(cudaMalloc(&t.sched, sched.size() * sizeof(int4)));
      CK(cudaMalloc(&t.cnt, cnt.size() + 16));
      CK(cudaMalloc(&t.xexp, xexp.size() * 2 + 16));
      CK(b200_copy_sync(e, t.sched, sched.data(), sched.size() * sizeof(int4), cudaMemcpyHostToDevice));
      CK(b200_copy_sync(e, t.cnt, cnt.data(), cnt.size(), cudaMemcpyHostToDevice));
      CK(b200_copy_sync(e, t.xexp, xexp.data(), xexp.size() * 2, cudaMemcpyHostToDevice));
      t.tables = true;
    }
    const size_t words = (size_t)pl.set_words * 2;
    if (words > e->st_acc_words) {
      if (e->st_acc) cudaFree(e->st_acc);
      CK(cudaMalloc(&e->st_acc, words * 8));
      e->st_acc_words = words;
    }
    if (pl.cand_words > e->st_cand_words) {
      if (e->st_cand) cudaFree(e->st_cand);
      CK(cudaMalloc(&e->st_cand, pl.cand_words * 8));
      e->st_cand_words = pl.cand_words;
    }
    t.plans[slot] = pl; t.plan_ok[slot] = true;
  }
  if (e->st_map_B != e->B || e->st_map_T != e->T_enc) {
    std::string msg;
    const int64_t crows = (int64_t)2 * L * e->B * e->T_enc;
    const int64_t krows = (int64_t)L * e->B * c.n_heads * c.max_target;
    if (!make_tmap_2d_plain(&e->st_cross, e->cross_kv, d, crows, d, 64, 128, &msg) ||
        !make_tmap_2d_plain(&e->st_kc, e->kcache, 64, krows, 64, 64, 128, &msg) ||
        !make_tmap_2d_plain(&e->st_vc, e->vcache, 64, krows, 64, 64, 128, &msg))
      return e->fail(B200ASR_E_CUDA, "decoder_stream: " + msg);
    e->st_map_B = e->B; e->st_map_T = e->T_enc;
  }
  return B200ASR_OK;
}

// one cooperative launch: n_first forced tokens per utterance ([B][n_first] on device; the last one feeds the first head),
// then n_heads - 1 further greedy iterations.  first_is_prefill: the first head applies the begin-suppress bias.
// clip0 / nclips: a sub-batch launch over clips [clip0, clip0 + nclips) of the engine's batch (nclips 0 = the whole batch)
static int launch_stream(b200asr_engine* e, int rows, int n_heads_iters, const int* first_tokens, int n_first, bool first_is_prefill,
                         bool want_logits, bool multi, int clip0 = 0, int nclips = 0, bool keep_state = false) {
  RET(build_stream_tables(e, rows));
  const int slot = rows <= 1 ? 0 : (rows <= 2 ? 1 : (rows <= 4 ? 2 : 3));
  const b200asr_engine::StreamTables& t = e->stt[e->use_fp8 ? 1 : 0];
  const StreamPlan& pl = t.plans[slot];
  StreamArgs sa{};
  fill_mega_args(e, sa.m, n_heads_iters, first_tokens, n_first, first_is_prefill, want_logits);
  RET(arm_timing(e, sa.m));
  sa.sl = t.layers; sa.wmaps = t.wmaps; sa.head_g = t.head_g; sa.head_b = t.head_b; sa.head_s = t.head_s; sa.fp8 = e->use_fp8 ? 1 : 0;
  sa.acc = e->st_acc; sa.set_words = pl.set_words; sa.layer_words = pl.layer_words;
  sa.cand = e->st_cand; sa.sched = t.sched; sa.cnt = t.cnt; sa.xexp = t.xexp;
  sa.cnt_ld = pl.cnt_ld; sa.xt = pl.xt; sa.n_stages = pl.n_stages; sa.n_slots = pl.n_slots;
  sa.task_inv = e->ring_task_inv; sa.l2_hint = e->stream_l2_hint ? 1 : 0; sa.debug = e->stream_debug; sa.multi = multi ? 1 : 0;
  sa.keep_state = keep_state ? 1 : 0;
  sa.lean = e->stream_lean ? 1 : 0;
  if (nclips > 0 && nclips < e->B) {
    MegaArgs& m = sa.m;
    const b200asr_config& c = e->cfg;
    m.batch_stride = e->B; m.clip0 = clip0; m.batch = nclips;
    m.cur_token += clip0; m.n_gen += clip0; m.finished += clip0; m.n_save += clip0;
    m.tokens += (size_t)clip0 * m.tokens_ld; m.save_id += (size_t)clip0 * m.save_ld; m.selected_hist += (size_t)clip0 * m.sel_ld;
    if (m.logits) m.logits += (size_t)clip0 * c.vocab;
    if (m.t_valid) m.t_valid += clip0;
  }
  CK(cudaMemsetAsync(e->st_acc, 0, (size_t)pl.set_words * 2 * 8, e->st));
  CK(cudaMemsetAsync(e->st_cand, 0, pl.cand_words * 8, e->st));
  KL(launch_decoder_stream(sa, e->st_cross, e->st_kc, e->st_vc, pl.nrt, e->num_sms, pl.smem_bytes, e->st));
  return B200ASR_OK;
}

// n_first forced tokens per utterance ([B][n_first] on device; the last one feeds the first head), then n_heads - 1 further
// greedy iterations.  first_is_prefill: the first head applies the begin-suppress bias.  When the prompt rows of all clips fit
// the kernel's 8 activation rows, they run together as one multi-row iteration (its own launch), followed by a decode launch;
// otherwise the prompt is fed token by token inside a single launch.
int run_stream(b200asr_engine* e, int n_heads_iters, const int* first_tokens, int n_first, bool first_is_prefill, bool want_logits) {
  const int B = e->B;
  const int max_rows = e->use_fp8 ? 4 : kStreamMaxBatch;      // the FP8 kernel stages four E5M2 rows per activation row: 16 / 4
  if (n_first > 1 && n_first <= max_rows && e->stream_multi) {
    // multi-row prefill: the prompt rows of as many clips as fit the kernel's rows per launch (all of them when B * n_first fits;
    // e.g. 4 clips x 4 prompt tokens = two launches of 8 rows: 2.2 ms instead of 3.9 ms token by token), then the decode launch
    const int per = max_rows / n_first;
    for (int b0 = 0; b0 < B; b0 += per) {
      const int nb = std::min(per, B - b0);
      RET(launch_stream(e, nb * n_first, 1, first_tokens + (size_t)b0 * n_first, n_first, first_is_prefill, want_logits, true,
                        b0, nb, b0 + nb < B));
    }
    if (n_heads_iters > 1) RET(launch_stream(e, B, n_heads_iters - 1, e->cur_token, 1, false, want_logits, false));
    return B200ASR_OK;
  }
  return launch_stream(e, B, n_heads_iters, first_tokens, n_first, first_is_prefill, want_logits, false);
}

// lens: optional samples per clip (ragged batch); n_samples is then the row stride of pcm_host and the batch maximum
int do_upload(b200asr_engine* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples,
              const int32_t* lens = nullptr) {
  const b200asr_config& c = e->cfg;
  if (!e->finalized) return e->fail(B200ASR_E_INVALID, "weights not finalized");
  if (!pcm_host || batch <= 0 || batch > c.max_batch) return e->fail(B200ASR_E_INVALID, "batch out of range");
  if (n_samples < c.n_fft || n_samples > c.max_samples) return e->fail(B200ASR_E_INVALID, "n_samples out of range");
  if (pcm_dtype != B200ASR_PCM_I16 && pcm_dtype != B200ASR_PCM_F32) return e->fail(B200ASR_E_INVALID, "bad pcm dtype");
  const int Tm = n_samples / c.hop;
  const int T = (Tm + 1) / 2;
  if (T > c.max_source) return e->fail(B200ASR_E_INVALID, "audio longer than max_source positions");
  bool ragged = false;
  if (lens) {
    int longest = 0;
    for (int b = 0; b < batch; ++b) {
      if (lens[b] < c.n_fft || lens[b] > n_samples) return e->fail(B200ASR_E_INVALID, "per-clip length out of [n_fft, n_samples]");
      longest = std::max(longest, (int)lens[b]);
      ragged |= lens[b] != n_samples;
    }
    if (ragged) {
      // the encoder grid is sized by n_samples; a shorter maximum would only run padding rows, so ask for the tight stride
      if (longest / c.hop != Tm) return e->fail(B200ASR_E_INVALID, "n_samples must be the longest clip's length (rounded up within one hop)");
      e->h_lens.assign((size_t)2 * c.max_batch, 0);
      for (int b = 0; b < batch; ++b) { e->h_lens[b] = lens[b]; e->h_lens[c.max_batch + b] = (lens[b] / c.hop + 1) / 2; }
      CK(cudaMemcpyAsync(e->d_lens, e->h_lens.data(), e->h_lens.size() * sizeof(int), cudaMemcpyHostToDevice, e->st));
    }
  }
  e->ragged = ragged;
  e->B = batch; e->n_samples = n_samples; e->T_mel = Tm; e->T_enc = T; e->pcm_dtype = pcm_dtype;
  const size_t bytes = (size_t)batch * n_samples * (pcm_dtype == B200ASR_PCM_F32 ? 4 : 2);
  CK(cudaMemcpyAsync(e->pcm, pcm_host, bytes, cudaMemcpyHostToDevice, e->st));
  e->encoded = false; e->prefilled = false;
  return B200ASR_OK;
}

// set_option("fp8", 1) is a request for the FP8 streaming kernel: no silent fall-back to a bf16 decoder
int fp8_guard(b200asr_engine* e) {
  if (e->use_fp8 && !stream_ok(e))
    return e->fail(B200ASR_E_INVALID, "fp8 weights run in the streaming decode kernel only: bf16 engine, batch <= 4, d_model and ffn multiples of 128, arg-max heads");
  return B200ASR_OK;
}

int do_prefill(b200asr_engine* e, const int32_t* prompt_ids, int32_t n_prompt, int extra_iters = 0) {
  const b200asr_config& c = e->cfg;
  if (!e->encoded) return e->fail(B200ASR_E_INVALID, "prefill before encode");
  RET(fp8_guard(e));
  if (!prompt_ids || n_prompt <= 0 || n_prompt >= c.max_target) return e->fail(B200ASR_E_INVALID, "bad prompt");
  const int B = e->B;
  for (int i = 0; i < B * n_prompt; ++i)
    if (prompt_ids[i] < 0 || prompt_ids[i] >= c.vocab) return e->fail(B200ASR_E_INVALID, "prompt id out of range [0, vocab)");
  CK(cudaMemcpyAsync(e->d_prompt, prompt_ids, (size_t)B * n_prompt * 4, cudaMemcpyHostToDevice, e->st));
  CK(cudaMemsetAsync(e->dstate, 0, sizeof(DecState), e->st));
  CK(cudaMemsetAsync(e->n_gen, 0, (size_t)B * 4, e->st));
  CK(cudaMemsetAsync(e->finished, 0, (size_t)B * 4, e->st));
  CK(cudaMemsetAsync(e->n_save, 0, (size_t)B * 4, e->st));
  e->n_prompt = n_prompt;
  // generate_limit = MAX_SEQ_LEN - prompt length (Inference_Whisper_ONNX.py:821), optionally tightened
  e->limit = c.max_target - n_prompt;
  if (e->limit_cfg > 0 && e->limit_cfg < e->limit) e->limit = e->limit_cfg;
  if (extra_iters < 0) {          // caller only wanted the reset (it launches the fused prefill+decode itself)
    e->prefilled = true;
    return B200ASR_OK;
  }
  if (stream_ok(e)) RET(run_stream(e, 1 + extra_iters, e->d_prompt, n_prompt, true, true));
  else if (mega_ok(e, n_prompt)) RET(run_mega(e, 1 + extra_iters, e->d_prompt, n_prompt, true, true));
  else RET(enqueue_decoder(e, e->d_prompt, n_prompt, true));
  e->prefilled = true;
  return B200ASR_OK;
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

const char* b200asr_last_error(const b200asr_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int b200asr_create(const b200asr_config* cfg, b200asr_engine** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return B200ASR_E_INVALID; }
  *out = nullptr;
  if (cfg->d_model <= 0 || cfg->n_heads <= 0 || cfg->d_model != cfg->n_heads * 64) {
    g_create_error = "d_model must equal n_heads * 64 (Whisper head_dim)"; return B200ASR_E_INVALID;
  }
  if (cfg->d_model % 8 || cfg->ffn % 8 || cfg->n_mels % 8 || cfg->n_fft % 2 || cfg->hop <= 0 || cfg->max_batch <= 0 ||
      cfg->max_samples < cfg->n_fft || cfg->vocab <= 0 || cfg->max_target <= 1 || cfg->enc_layers <= 0 ||
      cfg->dec_layers <= 0) {
    g_create_error = "invalid model dimensions"; return B200ASR_E_INVALID;
  }
  if (cfg->precision != B200ASR_PRECISION_F32 && cfg->precision != B200ASR_PRECISION_BF16) {
    g_create_error = "invalid precision"; return B200ASR_E_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) {
    g_create_error = "no CUDA device: the b200asr engine has no CPU fallback"; return B200ASR_E_NOGPU;
  }
  cudaDeviceProp prop;
  if (cudaSetDevice(cfg->device) != cudaSuccess || cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) {
    g_create_error = "cudaSetDevice failed"; return B200ASR_E_CUDA;
  }
  if (prop.major != 10) {
    g_create_error = "device is sm_" + std::to_string(prop.major * 10 + prop.minor) + "; this build targets sm_100a only";
    return B200ASR_E_NOGPU;
  }
  b200asr_engine* e = new b200asr_engine();
  e->cfg = *cfg;
  e->num_sms = prop.multiProcessorCount;
  e->act_dtype = cfg->precision == B200ASR_PRECISION_BF16 ? kBF16 : kF32;
  e->es = dtype_size(e->act_dtype);
  if (cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking) != cudaSuccess) {
    g_create_error = "cudaStreamCreate failed"; delete e; return B200ASR_E_CUDA;
  }
  *out = e;
  return B200ASR_OK;
}

void b200asr_destroy(b200asr_engine* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  cudaStreamSynchronize(e->st);
  if (e->step_graph) cudaGraphExecDestroy(e->step_graph);
  if (e->enc_graph) cudaGraphExecDestroy(e->enc_graph);
  for (auto& kv : e->w) cudaFree(kv.second.ptr);
  void* bufs[] = {e->stage_buf, e->basis_t, e->fb_start, e->fb_len, e->pcm, e->d_lens, e->mel_raw, e->max_key, e->mel_pad, e->h1_pad,
                  e->hidden, e->stem, e->xhat, e->qkv, e->ctx, e->ffn, e->S, e->P, e->cross_kv, e->kcache, e->vcache,
                  e->dx, e->dq, e->dctx, e->dffn, e->logits, e->prob, e->d_prompt, e->cur_token, e->tokens, e->n_gen,
                  e->finished, e->save_id, e->n_save, e->selected_hist, e->d_stop, e->dstate, e->mega_layers, e->pf_blocks,
                  e->mega_bar, e->cand_val, e->cand_idx, e->timing, e->ring_ll, e->samp_noise,
                  e->st_acc, e->st_cand,
                  e->stt[0].layers, e->stt[0].wmaps, e->stt[0].fold, e->stt[0].sched, e->stt[0].cnt, e->stt[0].xexp,
                  e->stt[1].layers, e->stt[1].wmaps, e->stt[1].fold, e->stt[1].sched, e->stt[1].cnt, e->stt[1].xexp,
                  e->stt[1].w8, e->stt[1].scales};
  for (void* p : bufs) if (p) cudaFree(p);
  if (e->h_pinned) cudaFreeHost(e->h_pinned);
  cudaStreamDestroy(e->st);
  delete e;
}

int b200asr_set_option(b200asr_engine* e, const char* key, int64_t value) {
  if (!e || !key) return B200ASR_E_INVALID;
  if (!strcmp(key, "keep_stages")) { e->keep_stages = value != 0; return B200ASR_OK; }
  if (!strcmp(key, "mega_timing")) { e->mega_timing = value != 0; return B200ASR_OK; }
  if (!strcmp(key, "mega")) { e->use_mega = value != 0; return B200ASR_OK; }
  if (!strcmp(key, "attn_tc")) { e->use_attn_tc = value != 0; return B200ASR_OK; }
  if (!strcmp(key, "pdl")) { e->use_pdl = value != 0; return B200ASR_OK; }
  if (!strcmp(key, "enc_graph")) { e->use_enc_graph = value != 0; return B200ASR_OK; }
  if (!strcmp(key, "stream")) { e->use_stream = value != 0; return B200ASR_OK; }
  if (!strcmp(key, "stream_l2_hint")) { e->stream_l2_hint = value != 0; return B200ASR_OK; }
  if (!strcmp(key, "stream_debug")) { e->stream_debug = (int)value; return B200ASR_OK; }
  if (!strcmp(key, "stream_multi")) { e->stream_multi = value != 0; return B200ASR_OK; }
  if (!strcmp(key, "fp8")) { e->use_fp8 = value != 0; return B200ASR_OK; }
  if (!strcmp(key, "stream_lean")) { e->stream_lean = value != 0; return B200ASR_OK; }
  if (!strcmp(key, "ring")) { e->use_ring = value != 0; return B200ASR_OK; }
  if (!strcmp(key, "ring_tc")) { e->ring_tc = value != 0; return B200ASR_OK; }
  if (!strcmp(key, "ring_debug")) { e->ring_debug = (int)value; return B200ASR_OK; }
  if (!strcmp(key, "ring_fine_timing")) { e->ring_fine = value != 0; return B200ASR_OK; }
  if (!strcmp(key, "pf_ahead_mb")) { e->pf_ahead = (long long)value << 20; return B200ASR_OK; }
  return e->fail(B200ASR_E_INVALID, std::string("unknown option ") + key);
}

int b200asr_set_tensor(b200asr_engine* e, const char* name_c, const float* host, int64_t numel) {
  if (!e || !name_c || !host || numel <= 0) return B200ASR_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  const std::string name(name_c);
  const b200asr_config& c = e->cfg;
  const int F = c.n_fft / 2 + 1;
  if (name == "stft_kernel") {
    // [2F][n_fft] (Conv1d weight of STFT_Process.py:136-150) -> basis_t [n_fft][2F] for coalesced reads
    if (numel != (int64_t)2 * F * c.n_fft) return e->fail(B200ASR_E_INVALID, "stft_kernel size mismatch");
    std::vector<float> t((size_t)numel);
    for (int r = 0; r < 2 * F; ++r)
      for (int k = 0; k < c.n_fft; ++k) t[(size_t)k * 2 * F + r] = host[(size_t)r * c.n_fft + k];
    if (!e->basis_t) CK(cudaMalloc(&e->basis_t, (size_t)numel * 4));
    CK(b200_copy_sync(e, e->basis_t, t.data(), (size_t)numel * 4, cudaMemcpyHostToDevice));
  }
  if (name == "mel_fbank") {
    if (numel != (int64_t)c.n_mels * F) return e->fail(B200ASR_E_INVALID, "mel_fbank size mismatch");
    std::vector<int> s0(c.n_mels), ln(c.n_mels);
    for (int m = 0; m < c.n_mels; ++m) {
      int lo = F, hi = -1;
      for (int f = 0; f < F; ++f) if (host[(size_t)m * F + f] != 0.f) { if (f < lo) lo = f; hi = f; }
      s0[m] = hi < 0 ? 0 : lo; ln[m] = hi < 0 ? 0 : hi - lo + 1;
    }
    if (!e->fb_start) { CK(cudaMalloc(&e->fb_start, c.n_mels * 4)); CK(cudaMalloc(&e->fb_len, c.n_mels * 4)); }
    CK(b200_copy_sync(e, e->fb_start, s0.data(), c.n_mels * 4, cudaMemcpyHostToDevice));
    CK(b200_copy_sync(e, e->fb_len, ln.data(), c.n_mels * 4, cudaMemcpyHostToDevice));
  }
  DevTensor t;
  t.numel = numel;
  t.dtype = is_weight_matrix(name) ? e->act_dtype : kF32;
  auto it = e->w.find(name);
  if (it != e->w.end()) { cudaFree(it->second.ptr); e->w.erase(it); }
  CK(cudaMalloc(&t.ptr, (size_t)numel * dtype_size(t.dtype)));
  if (t.dtype == kF32) {
    CK(b200_copy_sync(e, t.ptr, host, (size_t)numel * 4, cudaMemcpyHostToDevice));
  } else {
    RET(ensure_stage(e, numel));
    CK(b200_copy_sync(e, e->stage_buf, host, (size_t)numel * 4, cudaMemcpyHostToDevice));
    f32_to_bf16_kernel<<<1024, 256, 0, e->st>>>(e->stage_buf, (bf16*)t.ptr, numel);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->st));
  }
  e->w[name] = t;
  e->finalized = false;
  if (e->enc_graph) { cudaGraphExecDestroy(e->enc_graph); e->enc_graph = nullptr; }      // weight pointers are baked into its nodes
  // tables that cache weight pointers (tensor maps, fold vectors, layer tables) are rebuilt on the next decoder launch
  for (auto& t : e->stt) {
    if (!t.layers) continue;
    cudaFree(t.layers); cudaFree(t.wmaps); cudaFree(t.fold);
    t.layers = nullptr; t.wmaps = nullptr; t.fold = nullptr;
    if (t.w8) { cudaFree(t.w8); cudaFree(t.scales); t.w8 = nullptr; t.scales = nullptr; t.head_s = nullptr; }
  }
  if (e->mega_layers) { cudaFree(e->mega_layers); e->mega_layers = nullptr; cudaFree(e->pf_blocks); e->pf_blocks = nullptr;
                        cudaFree(e->mega_bar); e->mega_bar = nullptr; cudaFree(e->cand_val); e->cand_val = nullptr;
                        cudaFree(e->cand_idx); e->cand_idx = nullptr; e->pf_B = -1; e->pf_T = -1; }
  return B200ASR_OK;
}

int b200asr_finalize_weights(b200asr_engine* e) {
  if (!e) return B200ASR_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  const b200asr_config& c = e->cfg;
  const int64_t d = c.d_model, f = c.ffn, L = c.dec_layers, F = c.n_fft / 2 + 1;
  RET(need(e, "stft_kernel", 2 * F * c.n_fft));
  RET(need(e, "mel_fbank", c.n_mels * F));
  RET(need(e, "enc.conv1.w", d * 3 * c.n_mels)); RET(need(e, "enc.conv1.b", d));
  RET(need(e, "enc.conv2.w", d * 3 * d)); RET(need(e, "enc.conv2.b", d));
  RET(need(e, "enc.pos", (int64_t)c.max_source * d));
  for (int l = 0; l < c.enc_layers; ++l) {
    const std::string p = "enc.L" + std::to_string(l) + ".";
    RET(need(e, p + "qkv.w", 3 * d * d)); RET(need(e, p + "qkv.b", 3 * d));
    RET(need(e, p + "out.w", d * d)); RET(need(e, p + "out.b", d));
    RET(need(e, p + "fc1.w", f * d)); RET(need(e, p + "fc1.b", f));
    RET(need(e, p + "fc2.w", d * f)); RET(need(e, p + "fc2.b", d));
  }
  RET(need(e, "enc.ln_post.g", d)); RET(need(e, "enc.ln_post.b", d));
  RET(need(e, "enc.cross_kv.w", 2 * L * d * d)); RET(need(e, "enc.cross_kv.b", 2 * L * d));
  RET(need(e, "dec.embed", (int64_t)c.vocab * d)); RET(need(e, "dec.pos", (int64_t)c.max_target * d));
  for (int l = 0; l < c.dec_layers; ++l) {
    const std::string p = "dec.L" + std::to_string(l) + ".";
    RET(need(e, p + "qkv.w", 3 * d * d)); RET(need(e, p + "qkv.b", 3 * d));
    RET(need(e, p + "out.w", d * d)); RET(need(e, p + "out.b", d));
    RET(need(e, p + "cq.w", d * d)); RET(need(e, p + "cq.b", d));
    RET(need(e, p + "cout.w", d * d)); RET(need(e, p + "cout.b", d));
    RET(need(e, p + "fc1.w", f * d)); RET(need(e, p + "fc1.b", f));
    RET(need(e, p + "fc2.w", d * f)); RET(need(e, p + "fc2.b", d));
  }
  RET(need(e, "dec.ln.g", d)); RET(need(e, "dec.ln.b", d));
  RET(need(e, "dec.suppress_bias", c.vocab)); RET(need(e, "dec.begin_suppress_bias", c.vocab));
  if (e->finalized) return B200ASR_OK;
  if (e->stage_buf) { cudaFree(e->stage_buf); e->stage_buf = nullptr; e->stage_cap = 0; }
  if (!e->pcm) {
    const size_t es = e->es;
    const int64_t B = c.max_batch, Tm = c.max_samples / c.hop, T = (Tm + 1) / 2, M = B * T, H = c.n_heads;
    const int64_t rows = B * 8;    // decoder rows: up to 8 prompt tokens per utterance
    RET(dmalloc(e, &e->pcm, (size_t)B * c.max_samples * 4));
    RET(dmalloc(e, &e->d_lens, (size_t)2 * B * 4));
    RET(dmalloc(e, &e->mel_raw, (size_t)B * Tm * c.n_mels * 4));
    RET(dmalloc(e, &e->max_key, (size_t)B * 4));
    RET(dmalloc(e, &e->mel_pad, (size_t)B * (Tm + 2) * c.n_mels * es));
    RET(dmalloc(e, &e->h1_pad, (size_t)B * (Tm + 2) * d * es));
    RET(dmalloc(e, &e->hidden, (size_t)M * d * 4));
    RET(dmalloc(e, &e->stem, (size_t)M * d * 4));
    RET(dmalloc(e, &e->xhat, (size_t)M * d * es));
    RET(dmalloc(e, &e->qkv, (size_t)M * 3 * d * es));
    RET(dmalloc(e, &e->ctx, (size_t)M * d * es));
    RET(dmalloc(e, &e->ffn, (size_t)M * f * es));
    RET(dmalloc(e, &e->S, (size_t)B * H * T * T * 4));
    RET(dmalloc(e, &e->P, (size_t)B * H * T * T * es));
    RET(dmalloc(e, &e->cross_kv, (size_t)M * 2 * L * d * es));
    RET(dmalloc(e, &e->kcache, (size_t)L * B * H * c.max_target * 64 * es));
    RET(dmalloc(e, &e->vcache, (size_t)L * B * H * c.max_target * 64 * es));
    RET(dmalloc(e, &e->dx, (size_t)rows * d * 4));
    RET(dmalloc(e, &e->dq, (size_t)rows * d * 4));
    RET(dmalloc(e, &e->dctx, (size_t)rows * d * 4));
    RET(dmalloc(e, &e->dffn, (size_t)rows * f * 4));
    RET(dmalloc(e, &e->logits, (size_t)B * c.vocab * 4));
    RET(dmalloc(e, &e->prob, (size_t)B * 4));
    RET(dmalloc(e, &e->d_prompt, (size_t)rows * 4));
    RET(dmalloc(e, &e->cur_token, (size_t)B * 4));
    RET(dmalloc(e, &e->tokens, (size_t)B * c.max_target * 4));
    RET(dmalloc(e, &e->n_gen, (size_t)B * 4));
    RET(dmalloc(e, &e->finished, (size_t)B * 4));
    RET(dmalloc(e, &e->save_id, (size_t)B * c.max_target * 4));
    RET(dmalloc(e, &e->n_save, (size_t)B * 4));
    RET(dmalloc(e, &e->selected_hist, (size_t)B * c.max_target * 4));
    RET(dmalloc(e, &e->d_stop, 64 * 4));
    RET(dmalloc(e, &e->dstate, sizeof(DecState)));
    e->h_pinned_bytes = (size_t)B * (c.max_target + 4) * 4 + 256;
    CK(cudaMallocHost(&e->h_pinned, e->h_pinned_bytes));
  }
  CK(cudaStreamSynchronize(e->st));
  e->finalized = true;
  return B200ASR_OK;
}

int b200asr_upload_pcm(b200asr_engine* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples) {
  if (!e) return B200ASR_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  RET(do_upload(e, pcm_host, pcm_dtype, batch, n_samples));
  CK(cudaStreamSynchronize(e->st));
  return B200ASR_OK;
}

int b200asr_encode_resident(b200asr_engine* e) {
  if (!e) return B200ASR_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  if (e->B <= 0) return e->fail(B200ASR_E_INVALID, "no PCM uploaded");
  RET(run_encoder(e));
  CK(cudaStreamSynchronize(e->st));
  return B200ASR_OK;
}

int b200asr_encode(b200asr_engine* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples) {
  if (!e) return B200ASR_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  RET(do_upload(e, pcm_host, pcm_dtype, batch, n_samples));
  RET(run_encoder(e));
  CK(cudaStreamSynchronize(e->st));
  return B200ASR_OK;
}

// Ragged batch: clip b has lens[b] samples, rows of pcm_host are n_samples apart (n_samples = the longest clip).  Every clip
// is processed exactly as if it were encoded alone (its own reflect pad, mel maximum, conv zero pad, attention keys), which is
// what the reference's dynamic audio axis (Export_Whisper.py:743) gives a caller who runs the clips one by one.
int b200asr_upload_pcm_ragged(b200asr_engine* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples,
                              const int32_t* lens) {
  if (!e) return B200ASR_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  if (!lens) return e->fail(B200ASR_E_INVALID, "null lens");
  RET(do_upload(e, pcm_host, pcm_dtype, batch, n_samples, lens));
  CK(cudaStreamSynchronize(e->st));
  return B200ASR_OK;
}

int b200asr_encode_ragged(b200asr_engine* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples,
                          const int32_t* lens) {
  if (!e) return B200ASR_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  if (!lens) return e->fail(B200ASR_E_INVALID, "null lens");
  RET(do_upload(e, pcm_host, pcm_dtype, batch, n_samples, lens));
  RET(run_encoder(e));
  CK(cudaStreamSynchronize(e->st));
  return B200ASR_OK;
}

int b200asr_set_decode_options(b200asr_engine* e, const int32_t* stop_ids, int32_t n_stop, int32_t generate_limit,
                               float repeat_penalty, int32_t penalty_range) {
  if (!e) return B200ASR_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  if (n_stop < 0 || n_stop > 64 || (n_stop > 0 && !stop_ids)) return e->fail(B200ASR_E_INVALID, "bad stop id list");
  if (!(repeat_penalty > 0.f) || penalty_range < 0) return e->fail(B200ASR_E_INVALID, "bad penalty options");
  if (!e->finalized) return e->fail(B200ASR_E_INVALID, "weights not finalized");
  e->stop_ids.assign(stop_ids, stop_ids + n_stop);
  if (n_stop) CK(b200_copy_sync(e, e->d_stop, stop_ids, (size_t)n_stop * 4, cudaMemcpyHostToDevice));
  e->limit_cfg = generate_limit > 0 ? generate_limit : 0;
  e->repeat_penalty = repeat_penalty;
  e->penalty_range = penalty_range;
  if (e->step_graph) { cudaGraphExecDestroy(e->step_graph); e->step_graph = nullptr; }
  return B200ASR_OK;
}

int b200asr_set_sampling(b200asr_engine* e, float temperature, int32_t top_k, float top_p, float repetition_penalty,
                         uint64_t seed, const float* noise_host, int32_t noise_rows) {
  if (!e) return B200ASR_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  if (temperature > 0.f && (top_k < 1 || top_k > 64 || top_k > e->cfg.vocab || !(top_p > 0.f) || !(repetition_penalty > 0.f)))
    return e->fail(B200ASR_E_INVALID, "bad sampling options (1 <= top_k <= 64, top_p > 0, repetition_penalty > 0)");
  e->samp_temperature = temperature; e->samp_top_k = top_k; e->samp_top_p = top_p; e->samp_rep = repetition_penalty;
  e->samp_seed = seed;
  if (e->samp_noise) { cudaFree(e->samp_noise); e->samp_noise = nullptr; }
  e->samp_noise_rows = 0; e->samp_noise_ld = top_k;
  if (temperature > 0.f && noise_host && noise_rows > 0) {
    const size_t n = (size_t)noise_rows * e->cfg.max_batch * top_k;
    CK(cudaMalloc(&e->samp_noise, n * 4));
    CK(b200_copy_sync(e, e->samp_noise, noise_host, n * 4, cudaMemcpyHostToDevice));
    e->samp_noise_rows = noise_rows;
  }
  if (e->step_graph) { cudaGraphExecDestroy(e->step_graph); e->step_graph = nullptr; }
  return B200ASR_OK;
}

int b200asr_prefill(b200asr_engine* e, const int32_t* prompt_ids, int32_t n_prompt, float* logits_out,
                    int32_t* first_token_out) {
  if (!e) return B200ASR_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  if (n_prompt > 8) return e->fail(B200ASR_E_INVALID, "prompt longer than 8 tokens");
  RET(do_prefill(e, prompt_ids, n_prompt));
  if (logits_out) CK(cudaMemcpyAsync(logits_out, e->logits, (size_t)e->B * e->cfg.vocab * 4, cudaMemcpyDeviceToHost, e->st));
  if (first_token_out) CK(cudaMemcpyAsync(first_token_out, e->cur_token, (size_t)e->B * 4, cudaMemcpyDeviceToHost, e->st));
  CK(cudaStreamSynchronize(e->st));
  return B200ASR_OK;
}

int b200asr_decode_step(b200asr_engine* e, const int32_t* token_in, float* logits_out, int32_t* token_out) {
  if (!e) return B200ASR_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  if (!e->prefilled) return e->fail(B200ASR_E_INVALID, "decode before prefill");
  DecState hs;
  CK(cudaMemcpyAsync(&hs, e->dstate, sizeof hs, cudaMemcpyDeviceToHost, e->st));
  CK(cudaStreamSynchronize(e->st));
  if (hs.kv_len + 1 > e->cfg.max_target) return e->fail(B200ASR_E_INVALID, "KV cache full");
  if (token_in) {
    for (int b = 0; b < e->B; ++b)
      if (token_in[b] < 0 || token_in[b] >= e->cfg.vocab) return e->fail(B200ASR_E_INVALID, "token id out of range [0, vocab)");
    CK(cudaMemcpyAsync(e->cur_token, token_in, (size_t)e->B * 4, cudaMemcpyHostToDevice, e->st));
  }
  if (stream_ok(e)) RET(run_stream(e, 1, e->cur_token, 1, false, true));
  else if (ring_ok(e)) RET(run_ring(e, 1, true));
  else if (mega_ok(e, 1)) RET(run_mega(e, 1, e->cur_token, 1, false, true));
  else RET(launch_step(e));
  if (logits_out) CK(cudaMemcpyAsync(logits_out, e->logits, (size_t)e->B * e->cfg.vocab * 4, cudaMemcpyDeviceToHost, e->st));
  if (token_out) CK(cudaMemcpyAsync(token_out, e->cur_token, (size_t)e->B * 4, cudaMemcpyDeviceToHost, e->st));
  CK(cudaStreamSynchronize(e->st));
  return B200ASR_OK;
}

static int fetch_tokens(b200asr_engine* e, int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out);
static int decode_loop(b200asr_engine* e, int max_steps, int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out) {
  const b200asr_config& c = e->cfg;
  if (!e->prefilled) return e->fail(B200ASR_E_INVALID, "decode before prefill");
  if (!tokens_out || !lens_out || tokens_ld <= 0) return e->fail(B200ASR_E_INVALID, "null output");
  // the loop needs at most limit-1 launches after the prefill head produced token #1
  int steps = e->limit - 1;
  if (max_steps >= 0 && max_steps < steps) steps = max_steps;
  const int room = c.max_target - e->n_prompt;    // cache positions left after the prompt
  if (steps > room) steps = room;
  int* h_done = e->h_pinned;
  if (steps > 0 && stream_ok(e)) {
    RET(run_stream(e, steps, e->cur_token, 1, false, false));   // the kernel leaves the loop itself when all latched
  } else if (steps > 0 && ring_ok(e)) {
    RET(run_ring(e, steps, false));
  } else if (steps > 0 && mega_ok(e, 1)) {
    RET(run_mega(e, steps, e->cur_token, 1, false, false));
  } else {
    for (int s = 0; s < steps; ++s) {
      RET(launch_step(e));
      if (!e->stop_ids.empty() && (s % 8) == 7 && s + 1 < steps) {
        CK(cudaMemcpyAsync(h_done, &e->dstate->all_done, 4, cudaMemcpyDeviceToHost, e->st));
        CK(cudaStreamSynchronize(e->st));
        if (*h_done) break;
      }
    }
  }
  return fetch_tokens(e, tokens_out, tokens_ld, lens_out);
}

static int fetch_tokens(b200asr_engine* e, int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out) {
  const b200asr_config& c = e->cfg;
  const int B = e->B;
  int* h_len = e->h_pinned + 64;
  int* h_tok = h_len + B;
  CK(cudaMemcpyAsync(h_len, e->n_gen, (size_t)B * 4, cudaMemcpyDeviceToHost, e->st));
  CK(cudaMemcpyAsync(h_tok, e->tokens, (size_t)B * c.max_target * 4, cudaMemcpyDeviceToHost, e->st));
  CK(cudaStreamSynchronize(e->st));
  for (int b = 0; b < B; ++b) {
    const int n = h_len[b] < tokens_ld ? h_len[b] : tokens_ld;
    lens_out[b] = n;
    memcpy(tokens_out + (size_t)b * tokens_ld, h_tok + (size_t)b * c.max_target, (size_t)n * 4);
  }
  return B200ASR_OK;
}

int b200asr_decode(b200asr_engine* e, int32_t max_steps, int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out) {
  if (!e) return B200ASR_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  return decode_loop(e, max_steps, tokens_out, tokens_ld, lens_out);
}

int b200asr_no_speech_prob(b200asr_engine* e, int32_t no_speech_token, float* prob_out) {
  if (!e || !prob_out) return B200ASR_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  if (!e->prefilled) return e->fail(B200ASR_E_INVALID, "no_speech_prob before prefill");
  if (no_speech_token < 0 || no_speech_token >= e->cfg.vocab) return e->fail(B200ASR_E_INVALID, "bad token id");
  // unsuppress bias = -suppress_bias (+128 on suppressed ids): computed once
  if (!find(e, "dec.unsuppress_bias")) {
    std::vector<float> h((size_t)e->cfg.vocab);
    CK(b200_copy_sync(e, h.data(), W(e, "dec.suppress_bias"), h.size() * 4, cudaMemcpyDeviceToHost));
    for (auto& v : h) v = (v != 0.f) ? 128.0f : 0.f;
    DevTensor t; t.numel = e->cfg.vocab; t.dtype = kF32;
    CK(cudaMalloc(&t.ptr, h.size() * 4));
    CK(b200_copy_sync(e, t.ptr, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    e->w["dec.unsuppress_bias"] = t;
  }
  KL(launch_softmax_pick(e->logits, WF(e, "dec.unsuppress_bias"), e->cfg.vocab, e->B, no_speech_token, e->prob, e->st));
  CK(cudaMemcpyAsync(prob_out, e->prob, (size_t)e->B * 4, cudaMemcpyDeviceToHost, e->st));
  CK(cudaStreamSynchronize(e->st));
  return B200ASR_OK;
}

int b200asr_transcribe_resident(b200asr_engine* e, const int32_t* prompt_ids, int32_t n_prompt, int32_t max_new,
                                int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out) {
  if (!e) return B200ASR_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  if (e->B <= 0) return e->fail(B200ASR_E_INVALID, "no PCM uploaded");
  if (n_prompt > 8) return e->fail(B200ASR_E_INVALID, "prompt longer than 8 tokens");
  RET(run_encoder(e));
  if (!tokens_out || !lens_out || tokens_ld <= 0) return e->fail(B200ASR_E_INVALID, "null output");
  const int saved = e->limit_cfg;
  if (max_new > 0 && (saved == 0 || max_new < saved)) e->limit_cfg = max_new;
  int r;
  if (stream_ok(e)) {
    // prompt prefill + the whole greedy loop in one launch of the streaming kernel
    r = do_prefill(e, prompt_ids, n_prompt, -1);                 // reset + upload the prompt only
    if (r == B200ASR_OK) {
      int heads = e->limit;                                      // first head = token #1, then limit - 1 further launches
      const int room = e->cfg.max_target - n_prompt + 1;
      if (heads > room) heads = room;
      if (heads < 1) heads = 1;
      r = run_stream(e, heads, e->d_prompt, n_prompt, true, false);
    }
    e->limit_cfg = saved;
    RET(r);
    return fetch_tokens(e, tokens_out, tokens_ld, lens_out);
  }
  if (mega_ok(e, n_prompt)) {
    r = do_prefill(e, prompt_ids, n_prompt, -1);                 // reset + upload the prompt only
    if (r == B200ASR_OK && ring_ok(e)) {
      // prefill launch (multi-token rows) in the barrier kernel, then the streaming kernel for the greedy loop
      r = run_mega(e, 1, e->d_prompt, n_prompt, true, false);
      int steps = e->limit - 1;
      const int room = e->cfg.max_target - n_prompt;
      if (steps > room) steps = room;
      if (r == B200ASR_OK && steps > 0) r = run_ring(e, steps, false);
    } else if (r == B200ASR_OK) {
      r = run_mega(e, e->limit, e->d_prompt, n_prompt, true, false);   // prefill + (limit-1) decode launches
    }
    e->limit_cfg = saved;
    RET(r);
    return fetch_tokens(e, tokens_out, tokens_ld, lens_out);
  }
  r = do_prefill(e, prompt_ids, n_prompt);
  e->limit_cfg = saved;
  RET(r);
  return decode_loop(e, -1, tokens_out, tokens_ld, lens_out);
}

int b200asr_transcribe(b200asr_engine* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples,
                       const int32_t* prompt_ids, int32_t n_prompt, int32_t max_new, int32_t* tokens_out,
                       int32_t tokens_ld, int32_t* lens_out) {
  if (!e) return B200ASR_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  RET(do_upload(e, pcm_host, pcm_dtype, batch, n_samples));
  return b200asr_transcribe_resident(e, prompt_ids, n_prompt, max_new, tokens_out, tokens_ld, lens_out);
}

int b200asr_transcribe_ragged(b200asr_engine* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples,
                              const int32_t* lens, const int32_t* prompt_ids, int32_t n_prompt, int32_t max_new,
                              int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out) {
  if (!e) return B200ASR_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  if (!lens) return e->fail(B200ASR_E_INVALID, "null lens");
  RET(do_upload(e, pcm_host, pcm_dtype, batch, n_samples, lens));
  return b200asr_transcribe_resident(e, prompt_ids, n_prompt, max_new, tokens_out, tokens_ld, lens_out);
}

int b200asr_get_stage(b200asr_engine* e, const char* name_c, float* out, int64_t capacity, int64_t* numel_out) {
  if (!e || !name_c || !out) return B200ASR_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  const b200asr_config& c = e->cfg;
  const std::string name(name_c);
  const int64_t B = e->B, d = c.d_model, H = c.n_heads, L = c.dec_layers, T = e->T_enc, Tm = e->T_mel;
  if (!e->encoded) return e->fail(B200ASR_E_INVALID, "get_stage before encode");
  CK(cudaStreamSynchronize(e->st));
  auto fetch = [&](const void* src, int64_t n, int dtype, std::vector<float>& h) -> int {
    h.resize((size_t)n);
    if (dtype == kF32) { CK(b200_copy_sync(e, h.data(), src, (size_t)n * 4, cudaMemcpyDeviceToHost)); return B200ASR_OK; }
    RET(ensure_stage(e, n));
    bf16_to_f32_kernel<<<1024, 256, 0, e->st>>>((const bf16*)src, e->stage_buf, n);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->st));
    CK(b200_copy_sync(e, h.data(), e->stage_buf, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return B200ASR_OK;
  };
  std::vector<float> h;
  int64_t n = 0;
  if (name == "mel") {                       // -> [B][n_mels][Tm]
    RET(fetch(e->mel_pad, B * (Tm + 2) * c.n_mels, e->act_dtype, h));
    n = B * c.n_mels * Tm;
    if (n > capacity) return e->fail(B200ASR_E_INVALID, "stage buffer too small");
    for (int64_t b = 0; b < B; ++b)
      for (int64_t t = 0; t < Tm; ++t)
        for (int64_t m = 0; m < c.n_mels; ++m)
          out[(b * c.n_mels + m) * Tm + t] = h[(size_t)((b * (Tm + 2) + 1 + t) * c.n_mels + m)];
  } else if (name == "stem" || name == "hidden") {
    if (name == "stem" && !e->keep_stages) return e->fail(B200ASR_E_INVALID, "set option keep_stages=1 before encode");
    RET(fetch(name == "stem" ? e->stem : e->hidden, B * T * d, kF32, h));
    n = B * T * d;
    if (n > capacity) return e->fail(B200ASR_E_INVALID, "stage buffer too small");
    memcpy(out, h.data(), (size_t)n * 4);
  } else if (name == "enc_out") {
    RET(fetch(e->xhat, B * T * d, e->act_dtype, h));
    n = B * T * d;
    if (n > capacity) return e->fail(B200ASR_E_INVALID, "stage buffer too small");
    memcpy(out, h.data(), (size_t)n * 4);
  } else if (name == "cross_k" || name == "cross_v") {      // -> [B][L][H][T][64]
    RET(fetch(e->cross_kv, B * T * 2 * L * d, e->act_dtype, h));
    n = B * L * H * T * 64;
    if (n > capacity) return e->fail(B200ASR_E_INVALID, "stage buffer too small");
    const int64_t zoff = name == "cross_v" ? L : 0;
    for (int64_t b = 0; b < B; ++b)
      for (int64_t l = 0; l < L; ++l)
        for (int64_t hh = 0; hh < H; ++hh)
          for (int64_t t = 0; t < T; ++t)
            memcpy(out + ((((b * L + l) * H + hh) * T + t) * 64),
                   h.data() + (size_t)((((zoff + l) * B + b) * T + t) * d + hh * 64), 64 * 4);
  } else if (name == "self_k" || name == "self_v") {        // -> [L][B][H][kv][64]
    DecState hs;
    CK(b200_copy_sync(e, &hs, e->dstate, sizeof hs, cudaMemcpyDeviceToHost));
    const int64_t kv = hs.kv_len, Bm = B, mt = c.max_target;
    n = L * Bm * H * kv * 64;
    if (n > capacity) return e->fail(B200ASR_E_INVALID, "stage buffer too small");
    for (int64_t l = 0; l < L; ++l) {
      const char* base = (const char*)(name == "self_k" ? e->kcache : e->vcache) + (size_t)l * Bm * H * mt * 64 * e->es;
      RET(fetch(base, Bm * H * mt * 64, e->act_dtype, h));
      for (int64_t bh = 0; bh < Bm * H; ++bh)
        memcpy(out + (size_t)((l * Bm * H + bh) * kv * 64), h.data() + (size_t)(bh * mt * 64), (size_t)kv * 64 * 4);
    }
  } else if (name == "mega_timing_raw") {                   // the raw 64-bit stamps, two floats' worth of bits each
    if (!e->timing) return e->fail(B200ASR_E_INVALID, "set option mega_timing=1 first");
    n = capacity < 2 * (int64_t)b200asr_engine::kTimingCap ? capacity : 2 * (int64_t)b200asr_engine::kTimingCap;
    CK(b200_copy_sync(e, out, e->timing, (size_t)n * 4, cudaMemcpyDeviceToHost));
  } else if (name == "mega_timing") {                       // us between consecutive grid barriers of the last launch
    if (!e->timing) return e->fail(B200ASR_E_INVALID, "set option mega_timing=1 first");
    std::vector<unsigned long long> ht(b200asr_engine::kTimingCap);
    CK(b200_copy_sync(e, ht.data(), e->timing, ht.size() * 8, cudaMemcpyDeviceToHost));
    n = 0;
    for (size_t i = 1; i < ht.size() && ht[i] != 0 && n < capacity; ++i) out[n++] = (float)((double)(ht[i] - ht[i - 1]) * 1e-3);
  } else if (name == "stream_acc_val" || name == "stream_acc_cnt") {   // debug: decoded accumulator words of both step sets
    if (!e->st_acc) return e->fail(B200ASR_E_INVALID, "the streaming decode kernel has not run");
    n = (int64_t)e->st_acc_words;
    if (n > capacity) return e->fail(B200ASR_E_INVALID, "stage buffer too small");
    std::vector<unsigned long long> hw((size_t)n);
    CK(b200_copy_sync(e, hw.data(), e->st_acc, (size_t)n * 8, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n; ++i) {
      const unsigned long long w = hw[(size_t)i];
      const unsigned long long cnt = (w + (1ull << 51)) >> 52;
      out[i] = name == "stream_acc_cnt" ? (float)cnt : (float)((double)(long long)(w - (cnt << 52)) / 16777216.0);
    }
  } else if (name == "selected") {                          // [B][step] as float
    DecState hs;
    CK(b200_copy_sync(e, &hs, e->dstate, sizeof hs, cudaMemcpyDeviceToHost));
    std::vector<int> hi((size_t)B * c.max_target);
    CK(b200_copy_sync(e, hi.data(), e->selected_hist, hi.size() * 4, cudaMemcpyDeviceToHost));
    n = B * hs.step;
    if (n > capacity) return e->fail(B200ASR_E_INVALID, "stage buffer too small");
    for (int64_t b = 0; b < B; ++b)
      for (int64_t s = 0; s < hs.step; ++s) out[b * hs.step + s] = (float)hi[(size_t)(b * c.max_target + s)];
  } else {
    return e->fail(B200ASR_E_INVALID, "unknown stage '" + name + "'");
  }
  if (numel_out) *numel_out = n;
  return B200ASR_OK;
}

void* b200asr_stream(b200asr_engine* e) { return e ? (void*)e->st : nullptr; }
int b200asr_synchronize(b200asr_engine* e) {
  if (!e) return B200ASR_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaStreamSynchronize(e->st));
  return B200ASR_OK;
}
int64_t b200asr_kernel_launches(const b200asr_engine* e) { return e ? e->launches : 0; }
int b200asr_num_sms(const b200asr_engine* e) { return e ? e->num_sms : 0; }

int b200asr_test_gemm(int32_t device, int32_t impl, int32_t M, int32_t N, int32_t K, const float* A, const float* B,
                      const float* bias, const float* residual, int32_t act, float* C, char* err, int32_t err_len) {
  auto fail = [&](const std::string& m, int code) {
    if (err && err_len > 0) { strncpy(err, m.c_str(), err_len - 1); err[err_len - 1] = 0; }
    return code;
  };
  if (cudaSetDevice(device) != cudaSuccess) return fail("cudaSetDevice failed", B200ASR_E_NOGPU);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  float *dA32 = nullptr, *dB32 = nullptr, *dC = nullptr, *dbias = nullptr, *dres = nullptr;
  bf16 *dA = nullptr, *dB = nullptr;
  const int64_t Kp = (K + 7) / 8 * 8;       // row pitch padded to 16 bytes for TMA
  cudaError_t ce = cudaSuccess;
  auto ok = [&](cudaError_t x) { if (ce == cudaSuccess) ce = x; return x == cudaSuccess; };
  ok(cudaMalloc(&dA32, (size_t)M * Kp * 4)); ok(cudaMalloc(&dB32, (size_t)N * Kp * 4));
  ok(cudaMalloc(&dA, (size_t)M * Kp * 2)); ok(cudaMalloc(&dB, (size_t)N * Kp * 2));
  ok(cudaMalloc(&dC, (size_t)M * N * 4));
  if (ce == cudaSuccess) {
    ok(cudaMemset(dA32, 0, (size_t)M * Kp * 4)); ok(cudaMemset(dB32, 0, (size_t)N * Kp * 4));
    ok(cudaMemcpy2D(dA32, Kp * 4, A, (size_t)K * 4, (size_t)K * 4, M, cudaMemcpyHostToDevice));
    ok(cudaMemcpy2D(dB32, Kp * 4, B, (size_t)K * 4, (size_t)K * 4, N, cudaMemcpyHostToDevice));
    f32_to_bf16_kernel<<<256, 256>>>(dA32, dA, (int64_t)M * Kp);
    f32_to_bf16_kernel<<<256, 256>>>(dB32, dB, (int64_t)N * Kp);
    if (bias) { ok(cudaMalloc(&dbias, (size_t)N * 4)); ok(cudaMemcpy(dbias, bias, (size_t)N * 4, cudaMemcpyHostToDevice)); }
    if (residual) { ok(cudaMalloc(&dres, (size_t)M * N * 4)); ok(cudaMemcpy(dres, residual, (size_t)M * N * 4, cudaMemcpyHostToDevice)); }
  }
  std::string msg;
  if (ce == cudaSuccess) {
    GemmArgs g;
    g.A = dA; g.lda = Kp; g.a_dtype = kBF16; g.B = dB; g.ldb = Kp; g.b_dtype = kBF16;
    g.C = dC; g.ldc = N; g.c_dtype = kF32; g.bias = dbias; g.residual = dres; g.ldr = N; g.act = act;
    g.M = M; g.N = N; g.K = K;
    if (impl == 1) ok(launch_gemm_tc(g, prop.multiProcessorCount, 0, &msg));
    else ok(launch_gemm_simt(g, 0));
    ok(cudaDeviceSynchronize());
    if (ce == cudaSuccess) ok(cudaMemcpy(C, dC, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
  }
  cudaFree(dA32); cudaFree(dB32); cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dbias); cudaFree(dres);
  if (ce != cudaSuccess) return fail(msg.empty() ? std::string(cudaGetErrorString(ce)) : msg, B200ASR_E_CUDA);
  return B200ASR_OK;
}

}  // extern "C"
