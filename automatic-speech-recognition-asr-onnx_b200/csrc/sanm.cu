// Non-autoregressive SANM path (SenseVoiceSmall): Kaldi fbank front end -> LFR + CMVN + prompts -> SANM encoder
// blocks (self-attention || FSMN memory, position-wise FFN) -> CTC head with greedy collapse, one call per batch
// of equal-length clips.  Replaces `ort_session_A.run_with_iobinding` of
// /root/reference/SenseVoice/Inference_SenseVoice_ONNX.py:303; the math follows SENSE_VOICE.forward,
// /root/reference/SenseVoice/Export_SenseVoice.py:271-296 (block :227-258, folds :208-220, front end :139-169).
//
// Linear layers run on the engine's GEMMs (tcgen05 in bf16 mode, CUDA cores in the fp32 parity mode); attention uses
// the fused tcgen05 kernel (64- and 128-wide heads; SenseVoiceSmall / Paraformer have 128) up to 256 positions and the
// unfused batched products beyond or in fp32 mode; the kernels in this file are the HBM-bound pieces around them.  The
// Paraformer decoder runs once over the stacked token rows of all clips of a batch (segment-aware FSMN / cross-attention).
#include "common.cuh"
#include "../../include/b200asr.h"

#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace b200asr;

namespace {

constexpr int kFbFrames = 8;          // frames per CTA of the fbank kernel
constexpr int kFbBins = 288;          // >= nfft/2 + 1 = 257 frequency bins
constexpr int kFbThreads = 2 * kFbBins;   // two thread groups split the window taps (the loop is bound by the basis reads)

// ---- Kaldi fbank: conv1d with the folded [2F][win] basis (stride hop, snip-edges) -> power -> mel -> log ----
__global__ void __launch_bounds__(kFbThreads)
kaldi_fbank_kernel(const void* __restrict__ pcm, int pcm_is_f32, int n_samples, const float* __restrict__ basis_t /*[win][2F]*/,
                   const float* __restrict__ melw /*[F][n_mels]*/, int win, int hop, int F, int n_mels, int frames,
                   float log_floor, float* __restrict__ mel /*[B][frames][n_mels]*/) {
  extern __shared__ float fsm[];
  const int span = (kFbFrames - 1) * hop + win;
  float* xs = fsm;                       // [span]
  float* pw = fsm + span;                // [kFbFrames][F]
  const int b = blockIdx.y, f0 = blockIdx.x * kFbFrames;
  const int64_t base = (int64_t)b * n_samples + (int64_t)f0 * hop;
  for (int i = threadIdx.x; i < span; i += blockDim.x) {
    const int64_t s = (int64_t)f0 * hop + i;
    float v = 0.f;
    if (s < n_samples) v = pcm_is_f32 ? reinterpret_cast<const float*>(pcm)[base + i]
                                      : (float)reinterpret_cast<const short*>(pcm)[base + i];
    xs[i] = v;
  }
  __syncthreads();
  const int grp = threadIdx.x / kFbBins, f = threadIdx.x - grp * kFbBins;
  float re[kFbFrames], im[kFbFrames];
#pragma unroll
  for (int j = 0; j < kFbFrames; ++j) { re[j] = 0.f; im[j] = 0.f; }
  if (f < F) {
    const int k_lo = grp ? win / 2 : 0, k_hi = grp ? win : win / 2;
#pragma unroll 4
    for (int k = k_lo; k < k_hi; ++k) {
      const float c = basis_t[(int64_t)k * 2 * F + f];
      const float s = basis_t[(int64_t)k * 2 * F + F + f];
#pragma unroll
      for (int j = 0; j < kFbFrames; ++j) {
        const float x = xs[j * hop + k];
        re[j] = fmaf(c, x, re[j]);
        im[j] = fmaf(s, x, im[j]);
      }
    }
  }
  float* part = pw + kFbFrames * F;      // [kFbFrames][F][2] partial sums of the second tap group
  if (grp == 1 && f < F) {
#pragma unroll
    for (int j = 0; j < kFbFrames; ++j) { part[(j * F + f) * 2] = re[j]; part[(j * F + f) * 2 + 1] = im[j]; }
  }
  __syncthreads();
  if (grp == 0 && f < F) {
#pragma unroll
    for (int j = 0; j < kFbFrames; ++j) {
      const float r = re[j] + part[(j * F + f) * 2], i = im[j] + part[(j * F + f) * 2 + 1];
      pw[j * F + f] = r * r + i * i;
    }
  }
  __syncthreads();
  for (int o = threadIdx.x; o < kFbFrames * n_mels; o += blockDim.x) {
    const int j = o / n_mels, m = o - j * n_mels;
    if (f0 + j >= frames) continue;
    float acc = 0.f;
    for (int q = 0; q < F; ++q) acc = fmaf(pw[j * F + q], melw[(int64_t)q * n_mels + m], acc);
    mel[((int64_t)b * frames + f0 + j) * n_mels + m] = logf(fmaxf(acc, log_floor));
  }
}

// ---- LFR stack (m frames every n, clamped gather) + CMVN + sinusoid position; prompt rows in front ----
__global__ void lfr_cmvn_kernel(const float* __restrict__ mel, int frames, int n_mels, int lfr_m, int lfr_n, int T_lfr, int n_prompt,
                                const float* __restrict__ means, const float* __restrict__ vars, const float* __restrict__ pos,
                                const float* __restrict__ lang_embed, const float* __restrict__ sys_embed,
                                const int* __restrict__ lang_idx, float* __restrict__ x /*[B][n_prompt + T_lfr][feat]*/,
                                const int* __restrict__ frames_per_clip /*ragged batch: fbank frames of each clip, or null*/) {
  const int feat = n_mels * lfr_m;
  const int t = blockIdx.x, b = blockIdx.y;
  const int frames_b = frames_per_clip ? frames_per_clip[b] : frames;     // the clamped gather repeats the clip's OWN last frame
  float* xr = x + ((int64_t)b * (n_prompt + T_lfr) + t) * feat;
  if (t < n_prompt) {
    const float* src = t == 0 ? lang_embed + (int64_t)lang_idx[b] * feat : sys_embed + (int64_t)(t - 1) * feat;
    for (int c = threadIdx.x; c < feat; c += blockDim.x) xr[c] = src[c];
    return;
  }
  const int tl = t - n_prompt;
  const int half = (lfr_m - 1) / 2;
  for (int c = threadIdx.x; c < feat; c += blockDim.x) {
    const int j = c / n_mels, m = c - j * n_mels;
    int fr = tl * lfr_n + j - half;
    fr = max(0, min(fr, frames_b - 1));
    const float v = mel[((int64_t)b * frames + fr) * n_mels + m];
    xr[c] = (means ? (v + means[c]) * vars[c] : v * vars[c]) + pos[(int64_t)tl * feat + c];
  }
}

// ---- FSMN memory: depth-wise conv over time on the value rows (+ bias = linear_out's bias, + residual) ----
template <typename T>
__global__ void fsmn_kernel(const T* __restrict__ qkv, int64_t ld_qkv, int v_off, const float* __restrict__ w /*[D][k]*/,
                            const float* __restrict__ bias, const float* __restrict__ resid /*[M][D] or null*/, int Tn, int D, int ksz,
                            float* __restrict__ out /*[M][D]*/, const int* __restrict__ t_valid = nullptr /*ragged batch: rows per clip*/) {
  const int t = blockIdx.x, b = blockIdx.y;
  const int half = (ksz - 1) / 2;
  const int64_t row0 = (int64_t)b * Tn;
  const int Tv = t_valid ? t_valid[b] : Tn;             // the conv's zero padding starts at the clip's own end
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float acc = bias[c];
    for (int j = 0; j < ksz; ++j) {
      const int tt = t + j - half;
      if (tt >= 0 && tt < Tv) acc = fmaf(w[c * ksz + j], to_f<T>(qkv[(row0 + tt) * ld_qkv + v_off + c]), acc);
    }
    if (resid) acc += resid[(row0 + t) * D + c];
    out[(row0 + t) * D + c] = acc;
  }
}

// ---- CTC head: per-frame argmax (warp per row), then greedy collapse per utterance ----
__global__ void __launch_bounds__(256)
row_argmax_kernel(const float* __restrict__ logits, int rows, int vocab, int* __restrict__ ids) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* lr = logits + (int64_t)row * vocab;
  float best = -INFINITY; int bi = 0x7fffffff;
  for (int i = lane; i < vocab; i += 32) { const float v = lr[i]; if (v > best) { best = v; bi = i; } }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) ids[row] = bi == 0x7fffffff ? 0 : bi;
}
// keep frame t when id[t] != id[(t+1) % T] and id[t] != blank (Export_SenseVoice.py:289-294)
__global__ void ctc_collapse_kernel(const int* __restrict__ ids, int Tn, int blank, int* __restrict__ tokens, int tokens_ld,
                                    int* __restrict__ lens, const int* __restrict__ t_valid) {
  const int b = blockIdx.x;
  if (threadIdx.x != 0) return;
  const int* r = ids + (int64_t)b * Tn;
  const int Tv = t_valid ? t_valid[b] : Tn;             // ragged batch: the clip's own frames (the roll wraps at ITS end)
  int n = 0;
  for (int t = 0; t < Tv; ++t) {
    const int id = r[t], nx = r[t + 1 == Tv ? 0 : t + 1];
    if (id != nx && id != blank && n < tokens_ld) tokens[(int64_t)b * tokens_ld + n++] = id;
  }
  lens[b] = n;
}

// ---- Paraformer CIF predictor (Export_Paraformer.py:499-521) ----
// alpha[t] = sigmoid(w . relu_conv[t] + b); one warp per encoder frame
template <typename T>
__global__ void __launch_bounds__(256)
cif_alpha_kernel(const T* __restrict__ conv, const float* __restrict__ w, const float* __restrict__ b, int rows, int D,
                 float* __restrict__ alphas /*[B][Tn + 1]*/, int Tn) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int c = lane; c < D; c += 32) acc = fmaf(to_f<T>(conv[(int64_t)row * D + c]), w[c], acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    const int bb = row / Tn, t = row - bb * Tn;
    alphas[(int64_t)bb * (Tn + 1) + t] = 1.0f / (1.0f + expf(-(acc + b[0])));
  }
}
// integrate-and-fire: float64 prefix sum of alpha (+ tail), fire where its floor advances, acoustic embedding n =
// difference of the weighted hidden prefix sums completed at consecutive fires.  One CTA per utterance.
__global__ void __launch_bounds__(256)
cif_scan_kernel(float* __restrict__ alphas /*[B][Tn+1]: tail written here*/, float tail, const float* __restrict__ enc /*[B][Tn][D]*/,
                int Tn_ld, int D, float* __restrict__ acoustic /*[B][Tn+1][D]*/, int* __restrict__ n_tok,
                const int* __restrict__ t_valid /*ragged batch: encoder frames per clip, or null*/) {
  extern __shared__ float cs[];            // prefix[Tn+1] | fire flags as float [Tn+1]
  float* prefix = cs;
  float* fire = cs + Tn_ld + 1;
  const int b = blockIdx.x;
  const int Tn = t_valid ? t_valid[b] : Tn_ld;          // the tail threshold is appended after the clip's own last frame
  float* a = alphas + (int64_t)b * (Tn_ld + 1);
  if (threadIdx.x == 0) {
    a[Tn] = tail;
    double acc = 0.0;
    float prev_floor = 0.f;
    for (int t = 0; t <= Tn; ++t) {
      acc += (double)a[t];
      const float p = (float)acc;
      const float fl = floorf(p);
      prefix[t] = p;
      fire[t] = fl > prev_floor ? 1.f : 0.f;
      prev_floor = fl;
    }
    n_tok[b] = (int)floorf(prefix[Tn]);
  }
  __syncthreads();
  const float* h = enc + (int64_t)b * Tn_ld * D;
  float* out = acoustic + (int64_t)b * (Tn_ld + 1) * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    double acc = 0.0;                       // torch's CPU cumsum accumulates float32 inputs in double
    float prev = 0.f;
    int n = 0;
    for (int t = 0; t <= Tn; ++t) {
      const float hv = t < Tn ? h[(int64_t)t * D + c] : 0.f;
      acc += (double)(a[t] * hv);
      if (fire[t] != 0.f) {
        const float remains = prefix[t] - floorf(prefix[t]);
        const float completed = (float)acc - remains * hv;
        out[(int64_t)n * D + c] = completed - prev;
        prev = completed;
        ++n;
      }
    }
  }
}
// rows [0, n) <- src rows, rows beyond zero-filled up to `rows` (the zero-fire guard's dummy frame)
__global__ void copy_rows_kernel(const float* __restrict__ src, int n, int rows, int D, float* __restrict__ dst) {
  const int r = blockIdx.x;
  for (int c = threadIdx.x; c < D; c += blockDim.x) dst[(int64_t)r * D + c] = r < n ? src[(int64_t)r * D + c] : 0.f;
}
template <typename T>
__global__ void pad_rows_kernel(const float* __restrict__ src /*[B][Tn][D]*/, int Tn, int D, int pad, T* __restrict__ dst /*[B][Tn+2*pad][D]*/,
                                const int* __restrict__ t_valid) {
  const int t = blockIdx.x, b = blockIdx.y;      // t in [0, Tn + 2*pad)
  const int ts = t - pad;
  const int Tv = t_valid ? t_valid[b] : Tn;      // ragged batch: zeros from the clip's own end on
  for (int c = threadIdx.x; c < D; c += blockDim.x)
    dst[((int64_t)b * (Tn + 2 * pad) + t) * D + c] = from_f<T>((ts >= 0 && ts < Tv) ? src[((int64_t)b * Tn + ts) * D + c] : 0.f);
}

// ---- Paraformer decoder over all clips of a batch at once: the decoder rows of the clips are stacked (clip b owns rows
//      seg_off[b] .. seg_off[b+1]), so every row-wise layer is one launch; the three kernels below are the pieces that
//      need the clip boundaries ----
__device__ __forceinline__ int seg_find(const int* __restrict__ seg_off, int B, int r) {
  int b = 0;
  while (b + 1 < B && r >= seg_off[b + 1]) ++b;
  return b;
}

__global__ void para_gather_rows_kernel(const float* __restrict__ acoustic /*[B][T+1][D]*/, const int* __restrict__ n_tok,
                                        const int* __restrict__ seg_off, int B, int T, int D, float* __restrict__ dst) {
  const int r = blockIdx.x;
  const int b = seg_find(seg_off, B, r), i = r - seg_off[b];
  const float* src = acoustic + ((int64_t)b * (T + 1) + i) * D;
  const bool live = i < n_tok[b];                     // zero-fire guard: a clip without fires decodes one zero row (:523-528)
  for (int c = threadIdx.x; c < D; c += blockDim.x) dst[(int64_t)r * D + c] = live ? src[c] : 0.f;
}

__global__ void fsmn_seg_kernel(const float* __restrict__ in /*[rows][D]*/, const float* __restrict__ w /*[D][k]*/,
                                const float* __restrict__ resid, const int* __restrict__ seg_off, int B, int D, int ksz,
                                float* __restrict__ out) {
  const int r = blockIdx.x;
  const int b = seg_find(seg_off, B, r);
  const int lo = seg_off[b], hi = seg_off[b + 1];
  const int half = (ksz - 1) / 2;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < ksz; ++j) {
      const int rr = r + j - half;
      if (rr >= lo && rr < hi) acc = fmaf(w[c * ksz + j], in[(int64_t)rr * D + c], acc);
    }
    out[(int64_t)r * D + c] = acc + resid[(int64_t)r * D + c];
  }
}

template <typename T> struct NVec;
template <> struct NVec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float* o) { const float4 u = *reinterpret_cast<const float4*>(p); o[0] = u.x; o[1] = u.y; o[2] = u.z; o[3] = u.w; }
};
template <> struct NVec<bf16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const bf16* p, float* o) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
  }
};

// cross-attention of stacked decoder rows on their own clip's encoder memory: one warp per (row, head), keys on lanes
// for the scores, head dims on lanes for P V; probabilities stay fp32
template <typename T, int DH>
__global__ void __launch_bounds__(128)
para_cross_attn_kernel(const T* __restrict__ q /*[rows][D]*/, const T* __restrict__ kv /*[B*Tn][2D]: k | v*/, const int* __restrict__ seg_off,
                       int B, int Tn_ld, int H, int total, T* __restrict__ ctx /*[rows][D]*/, const int* __restrict__ t_valid) {
  extern __shared__ float csm[];                 // per warp: q[DH] + scores[Tn]
  constexpr int M = DH / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * (blockDim.x >> 5) + warp;
  if (item >= total) return;
  const int row = item / H, h = item - row * H;
  const int D = H * DH;
  const int b = seg_find(seg_off, B, row);
  const int Tn = t_valid ? t_valid[b] : Tn_ld;          // ragged batch: the clip's own encoder memory
  float* qs = csm + warp * (DH + Tn_ld);
  float* sc = qs + DH;
#pragma unroll
  for (int m = 0; m < M; ++m) qs[lane + 32 * m] = to_f<T>(q[(int64_t)row * D + h * DH + lane + 32 * m]);
  __syncwarp();
  const T* K = kv + (int64_t)b * Tn_ld * 2 * D + h * DH;
  const T* V = K + D;
  float mx = -INFINITY;
  for (int j = lane; j < Tn; j += 32) {
    const T* kr = K + (int64_t)j * 2 * D;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < DH; c += NVec<T>::N) {
      float kk[NVec<T>::N];
      NVec<T>::load(kr + c, kk);
#pragma unroll
      for (int e = 0; e < NVec<T>::N; ++e) s = fmaf(kk[e], qs[c + e], s);
    }
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < Tn; j += 32) { const float p = expf(sc[j] - mx); sc[j] = p; sum += p; }
  sum = warp_sum(sum);
  __syncwarp();
  float acc[M];
#pragma unroll
  for (int m = 0; m < M; ++m) acc[m] = 0.f;
#pragma unroll 4
  for (int j = 0; j < Tn; ++j) {
    const float p = sc[j];
    const T* vr = V + (int64_t)j * 2 * D;
#pragma unroll
    for (int m = 0; m < M; ++m) acc[m] = fmaf(p, to_f<T>(vr[lane + 32 * m]), acc[m]);
  }
  const float inv = 1.0f / sum;
#pragma unroll
  for (int m = 0; m < M; ++m) ctx[(int64_t)row * D + h * DH + lane + 32 * m] = from_f<T>(acc[m] * inv);
}

__global__ void para_scatter_tokens_kernel(const int* __restrict__ ids, const int* __restrict__ seg_off, int B, int max_T, int rows,
                                           int* __restrict__ tokens) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int b = seg_find(seg_off, B, r);
  tokens[(int64_t)b * max_T + (r - seg_off[b])] = ids[r];
}

__global__ void nar_f32_to_bf16(const float* __restrict__ in, bf16* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void nar_bf16_to_f32(const bf16* __restrict__ in, float* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __bfloat162float(in[i]);
}

struct NarTensor { void* ptr = nullptr; int64_t numel = 0; int dtype = kF32; };
std::string g_nar_create_error;

}  // namespace

struct b200asr_nar {
  b200asr_nar_config cfg{};
  cudaStream_t st = nullptr;
  std::string err;
  int num_sms = 148;
  int64_t launches = 0;
  bool finalized = false;
  int act = kF32; size_t es = 4;
  std::map<std::string, NarTensor> w;
  float* basis_t = nullptr;
  float* stage_buf = nullptr; int64_t stage_cap = 0;
  // per-call state
  int B = 0, n_samples = 0, frames = 0, T_lfr = 0, T = 0, pcm_dtype = B200ASR_PCM_I16;
  void* pcm = nullptr; float* mel = nullptr; float* feats = nullptr; float* hidden = nullptr; float* resid = nullptr;
  void *xhat = nullptr, *qkv = nullptr, *ctx = nullptr, *ffn = nullptr, *P = nullptr;
  float* S = nullptr; float* logits = nullptr; float* enc_out = nullptr;
  int *frame_ids = nullptr, *tokens = nullptr, *lens = nullptr, *lang = nullptr;
  // Paraformer
  void *enc_pad = nullptr, *conv_out = nullptr, *kvbuf = nullptr, *dq = nullptr;
  float *alphas = nullptr, *acoustic = nullptr, *dec = nullptr, *dx = nullptr, *f32buf = nullptr, *sa_in = nullptr, *dec_logits = nullptr;
  int* n_tok = nullptr; int last_rows = 0; int* seg_off = nullptr; bool batched_decoder = true;
  // SenseVoice: the whole forward is one CUDA graph per (batch, n_samples) -- ~700 small launches are host-bound otherwise
  bool use_attn_tc = true, use_pdl = true;
  bool use_graph = true; cudaGraphExec_t graph = nullptr; int graph_B = -1, graph_N = -1, graph_dtype = -1; int64_t graph_nodes = 0;
  int* h_pinned = nullptr;
  int max_frames = 0, max_T = 0;
  // ragged batch (b200asr_nar_*_ragged): device [2][max_batch] = fbank frames per clip | encoder rows per clip (prompt rows included)
  int* d_meta = nullptr; bool ragged = false; std::vector<int> h_meta;
  const int* dev_frames() const { return ragged ? d_meta : nullptr; }
  const int* dev_tvalid() const { return ragged ? d_meta + cfg.max_batch : nullptr; }

  int fail(int code, const std::string& m) { err = m; return code; }
  int cuda_fail(cudaError_t e, const char* what) { err = std::string(what) + ": " + cudaGetErrorString(e); return B200ASR_E_CUDA; }
};

#define NCK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return e->cuda_fail(_e, #expr); } while (0)
#define NKL(expr) do { cudaError_t _e = (expr); e->launches++; if (_e != cudaSuccess) return e->cuda_fail(_e, #expr); } while (0)
#define NRET(expr) do { int _r = (expr); if (_r != B200ASR_OK) return _r; } while (0)

namespace {

bool nar_is_matrix(const std::string& n) {
  return n.size() > 2 && n.compare(n.size() - 2, 2, ".w") == 0 && n.find("fsmn") == std::string::npos && n != "cif.out.w";
}
const void* NW(b200asr_nar* e, const std::string& n) { return e->w[n].ptr; }
const float* NWF(b200asr_nar* e, const std::string& n) { return reinterpret_cast<const float*>(e->w[n].ptr); }

int nar_need(b200asr_nar* e, const std::string& n, int64_t numel) {
  auto it = e->w.find(n);
  if (it == e->w.end()) return e->fail(B200ASR_E_MISSING, "missing weight tensor '" + n + "'");
  if (it->second.numel != numel)
    return e->fail(B200ASR_E_INVALID, "tensor '" + n + "' has " + std::to_string(it->second.numel) + " elements, expected " + std::to_string(numel));
  return B200ASR_OK;
}

int nar_gemm(b200asr_nar* e, const GemmArgs& g) {
  if (e->act == kBF16 && e->cfg.use_tensor_cores && gemm_tc_supported(g)) {
    std::string msg;
    cudaError_t r = launch_gemm_tc(g, e->num_sms, e->st, &msg);
    if (r != cudaErrorNotSupported) {
      e->launches++;
      if (r != cudaSuccess) return e->fail(B200ASR_E_CUDA, "gemm_tc: " + (msg.empty() ? std::string(cudaGetErrorString(r)) : msg));
      return B200ASR_OK;
    }
    cudaGetLastError();
  }
  NKL(launch_gemm_simt(g, e->st));
  return B200ASR_OK;
}

GemmArgs nar_linear(b200asr_nar* e, const void* A, int64_t lda, const std::string& wn, const std::string& bn, void* C, int64_t ldc,
                    int c_dtype, int M, int N, int K) {
  GemmArgs g;
  g.A = A; g.lda = lda; g.a_dtype = e->act;
  g.B = NW(e, wn); g.ldb = K; g.b_dtype = e->act;
  g.C = C; g.ldc = ldc; g.c_dtype = c_dtype;
  g.bias = bn.empty() ? nullptr : NWF(e, bn);
  g.M = M; g.N = N; g.K = K;
  g.pdl = e->use_pdl ? 1 : 0;          // B is a weight matrix: the tcgen05 GEMM may start under its predecessor's tail
  return g;
}

template <typename T>
int nar_alloc(b200asr_nar* e, T** p, size_t bytes) {
  NCK(cudaMalloc(reinterpret_cast<void**>(p), bytes ? bytes : 16));
  NCK(cudaMemsetAsync(*p, 0, bytes ? bytes : 16, e->st));
  return B200ASR_OK;
}

const float* NWF_opt(b200asr_nar* e, const std::string& n) {
  auto it = e->w.find(n);
  return it == e->w.end() ? nullptr : reinterpret_cast<const float*>(it->second.ptr);
}

// one SANM block; `p` = tensor-name prefix.  LayerNorm affines are optional (Paraformer folds them into the Linears);
// the FSMN bias operand is `fsmn_bias` (SenseVoice: the conv's own bias = linear_out's; Paraformer: linear_out's bias)
int nar_block(b200asr_nar* e, const std::string& p, const std::string& fsmn_bias, const float* x_in, int din) {
  const b200asr_nar_config& c = e->cfg;
  const int D = c.d_model, H = c.n_heads, dh = D / H, T = e->T, B = e->B, M = B * T, ad = e->act;
  const size_t es = e->es;
  NKL(launch_layernorm(x_in, din, NWF_opt(e, p + "norm1.g"), NWF_opt(e, p + "norm1.b"), e->xhat, ad, din, M, din, c.ln_eps, e->st, e->use_pdl ? 1 : 0));
  NRET(nar_gemm(e, nar_linear(e, e->xhat, din, p + "qkv.w", p + "qkv.b", e->qkv, 3 * D, ad, M, 3 * D, din)));
  // FSMN memory (+ x when the block keeps its width) -> the fp32 residual the out-projection adds
  const float* res_in = (din == D) ? x_in : nullptr;
  if (ad == kBF16)
    fsmn_kernel<bf16><<<dim3(T, B), 256, 0, e->st>>>((const bf16*)e->qkv, 3 * D, 2 * D, NWF(e, p + "fsmn.w"), NWF(e, fsmn_bias), res_in, T, D, c.fsmn_kernel, e->resid, e->dev_tvalid());
  else
    fsmn_kernel<float><<<dim3(T, B), 256, 0, e->st>>>((const float*)e->qkv, 3 * D, 2 * D, NWF(e, p + "fsmn.w"), NWF(e, fsmn_bias), res_in, T, D, c.fsmn_kernel, e->resid, e->dev_tvalid());
  NKL(cudaGetLastError());
  if (ad == kBF16 && c.use_tensor_cores && e->use_attn_tc && attention_tc_supported(T, D, H)) {
    std::string msg;
    cudaError_t r = launch_attention_tc(e->qkv, e->ctx, B, T, D, H, e->st, &msg, e->dev_tvalid(), -1e30f);
    e->launches++;
    if (r != cudaSuccess) return e->fail(B200ASR_E_CUDA, "attention_tc: " + (msg.empty() ? std::string(cudaGetErrorString(r)) : msg));
  } else {
    GemmArgs s;
    s.A = e->qkv; s.lda = 3 * D; s.sAo = (int64_t)T * 3 * D; s.sAi = dh; s.a_dtype = ad;
    s.B = (char*)e->qkv + (size_t)D * es; s.ldb = 3 * D; s.sBo = (int64_t)T * 3 * D; s.sBi = dh; s.b_dtype = ad;
    s.C = e->S; s.ldc = T; s.sCo = (int64_t)H * T * T; s.sCi = (int64_t)T * T; s.c_dtype = kF32;
    s.M = T; s.N = T; s.K = dh; s.batch = B * H; s.batch_inner = H;
    NKL(launch_gemm_simt(s, e->st));
    NKL(launch_softmax_rows(e->S, e->P, ad, (int64_t)B * H * T, T, e->st, e->dev_tvalid(), (int64_t)H * T));
    GemmArgs o;
    o.A = e->P; o.lda = T; o.sAo = (int64_t)H * T * T; o.sAi = (int64_t)T * T; o.a_dtype = ad;
    o.B = (char*)e->qkv + (size_t)2 * D * es; o.ldb = 3 * D; o.sBo = (int64_t)T * 3 * D; o.sBi = dh; o.b_dtype = ad;
    o.transB = 1;
    o.C = e->ctx; o.ldc = D; o.sCo = (int64_t)T * D; o.sCi = dh; o.c_dtype = ad;
    o.M = T; o.N = dh; o.K = T; o.batch = B * H; o.batch_inner = H;
    NKL(launch_gemm_simt(o, e->st));
  }
  {
    GemmArgs g = nar_linear(e, e->ctx, D, p + "out.w", "", e->hidden, D, kF32, M, D, D);
    g.residual = e->resid; g.ldr = D;
    NRET(nar_gemm(e, g));
  }
  NKL(launch_layernorm(e->hidden, D, NWF_opt(e, p + "norm2.g"), NWF_opt(e, p + "norm2.b"), e->xhat, ad, D, M, D, c.ln_eps, e->st, e->use_pdl ? 1 : 0));
  {
    GemmArgs g = nar_linear(e, e->xhat, D, p + "w1.w", p + "w1.b", e->ffn, c.ffn, ad, M, c.ffn, D);
    g.act = kActRelu;
    NRET(nar_gemm(e, g));
    GemmArgs g2 = nar_linear(e, e->ffn, c.ffn, p + "w2.w", p + "w2.b", e->hidden, D, kF32, M, D, c.ffn);
    g2.residual = e->hidden; g2.ldr = D;
    NRET(nar_gemm(e, g2));
  }
  return B200ASR_OK;
}

int nar_fbank(b200asr_nar* e) {
  const b200asr_nar_config& c = e->cfg;
  const int F = c.nfft / 2 + 1, B = e->B;
  const int span = (kFbFrames - 1) * c.hop + c.win;
  const size_t smem = (size_t)(span + 3 * kFbFrames * F) * sizeof(float);
  dim3 grid((e->frames + kFbFrames - 1) / kFbFrames, B);
  kaldi_fbank_kernel<<<grid, kFbThreads, smem, e->st>>>(e->pcm, e->pcm_dtype == B200ASR_PCM_F32, e->n_samples, e->basis_t,
                                                        NWF(e, "mel_filters"), c.win, c.hop, F, c.n_mels, e->frames,
                                                        1.1920928955078125e-07f, e->mel);
  NKL(cudaGetLastError());
  return B200ASR_OK;
}

int sensevoice_forward(b200asr_nar* e) {
  const b200asr_nar_config& c = e->cfg;
  const int feat = c.n_mels * c.lfr_m, D = c.d_model, B = e->B, T = e->T, M = B * T;
  NRET(nar_fbank(e));
  lfr_cmvn_kernel<<<dim3(T, B), 256, 0, e->st>>>(e->mel, e->frames, c.n_mels, c.lfr_m, c.lfr_n, e->T_lfr, c.n_prompt,
                                                 NWF(e, "cmvn_means"), NWF(e, "cmvn_vars"), NWF(e, "speech_position"),
                                                 NWF(e, "language_embed"), NWF(e, "system_embed"), e->lang, e->feats, e->dev_frames());
  NKL(cudaGetLastError());
  const int n_main = c.n_blocks0 + c.n_blocks, n_all = n_main + c.n_tp_blocks;
  for (int i = 0; i < n_all; ++i) {
    const std::string p = "blk" + std::to_string(i) + ".";
    NRET(nar_block(e, p, p + "fsmn.b", i == 0 ? e->feats : e->hidden, i == 0 ? feat : D));
    if (i == n_main - 1)      // after_norm between the encoder blocks and the transformer-postnet blocks (:266)
      NKL(launch_layernorm(e->hidden, D, NWF(e, "after_norm.g"), NWF(e, "after_norm.b"), e->hidden, kF32, D, M, D, c.ln_eps, e->st, e->use_pdl ? 1 : 0));
  }
  NKL(launch_layernorm(e->hidden, D, NWF(e, "tp_norm.g"), NWF(e, "tp_norm.b"), e->enc_out, kF32, D, M, D, c.ln_eps, e->st, e->use_pdl ? 1 : 0));
  NKL(launch_layernorm(e->hidden, D, NWF(e, "tp_norm.g"), NWF(e, "tp_norm.b"), e->xhat, e->act, D, M, D, c.ln_eps, e->st, e->use_pdl ? 1 : 0));
  NRET(nar_gemm(e, nar_linear(e, e->xhat, D, "ctc.w", "ctc.b", e->logits, c.vocab, kF32, M, c.vocab, D)));
  row_argmax_kernel<<<(M + 7) / 8, 256, 0, e->st>>>(e->logits, M, c.vocab, e->frame_ids);
  NKL(cudaGetLastError());
  ctc_collapse_kernel<<<B, 32, 0, e->st>>>(e->frame_ids, T, c.blank_id, e->tokens, e->max_T, e->lens, e->dev_tvalid());
  NKL(cudaGetLastError());
  return B200ASR_OK;
}

// ---- Paraformer (Export_Paraformer.py:474-563) ----
int paraformer_decode_one(b200asr_nar* e, int b, int n_tok) {
  const b200asr_nar_config& c = e->cfg;
  const int D = c.d_model, H = c.n_heads, dh = D / H, T = e->T, ad = e->act, Fd = c.dec_ffn;
  const size_t es = e->es;
  const int rows = n_tok > 0 ? n_tok : 1;                      // zero-fire guard: one zero frame, dropped again below (:523-528)
  copy_rows_kernel<<<rows, 128, 0, e->st>>>(e->acoustic + (int64_t)b * (T + 1) * D, n_tok, rows, D, e->dec);
  NKL(cudaGetLastError());
  const char* memory = (const char*)e->xhat + (size_t)b * T * D * es;           // encoder_out of utterance b in the activation dtype
  auto ffn = [&](const std::string& p, float* out, const float* resid) -> int {
    NKL(launch_layernorm(e->dec, D, nullptr, nullptr, e->dq, ad, D, rows, D, c.dec_ln_eps, e->st, e->use_pdl ? 1 : 0));
    {
      GemmArgs g = nar_linear(e, e->dq, D, p + "w1.w", p + "w1.b", e->f32buf, Fd, kF32, rows, Fd, D);
      g.act = kActRelu;
      NRET(nar_gemm(e, g));
    }
    NKL(launch_layernorm(e->f32buf, Fd, nullptr, nullptr, e->ffn, ad, Fd, rows, Fd, c.dec_ln_eps, e->st, e->use_pdl ? 1 : 0));
    GemmArgs g2 = nar_linear(e, e->ffn, Fd, p + "w2.w", p + "w2.b", out, D, kF32, rows, D, Fd);
    if (resid) { g2.residual = resid; g2.ldr = D; }
    return nar_gemm(e, g2);
  };
  for (int i = 0; i < c.dec_att_blocks; ++i) {
    const std::string p = "dec" + std::to_string(i) + ".";
    NRET(ffn(p, e->dx, nullptr));                                                               // x = FFN(dec)
    NKL(launch_layernorm(e->dx, D, NWF(e, p + "norm2.g"), NWF(e, p + "norm2.b"), e->sa_in, kF32, D, rows, D, c.dec_ln_eps, e->st, e->use_pdl ? 1 : 0));
    fsmn_kernel<float><<<dim3(rows, 1), 256, 0, e->st>>>(e->sa_in, D, 0, NWF(e, p + "fsmn.w"), NWF(e, "zero_bias"), e->dec, rows, D,
                                                       c.fsmn_kernel, e->dx);                  // x = dec + fsmn(norm2(x))
    NKL(cudaGetLastError());
    NKL(launch_layernorm(e->dx, D, nullptr, nullptr, e->dq, ad, D, rows, D, c.dec_ln_eps, e->st, e->use_pdl ? 1 : 0));
    NRET(nar_gemm(e, nar_linear(e, e->dq, D, p + "q.w", p + "q.b", e->qkv, D, ad, rows, D, D)));
    NRET(nar_gemm(e, nar_linear(e, memory, D, p + "kv.w", p + "kv.b", e->kvbuf, 2 * D, ad, T, 2 * D, D)));
    GemmArgs sgm;
    sgm.A = e->qkv; sgm.lda = D; sgm.sAi = dh; sgm.a_dtype = ad;
    sgm.B = e->kvbuf; sgm.ldb = 2 * D; sgm.sBi = dh; sgm.b_dtype = ad;
    sgm.C = e->S; sgm.ldc = T; sgm.sCi = (int64_t)rows * T; sgm.c_dtype = kF32;
    sgm.M = rows; sgm.N = T; sgm.K = dh; sgm.batch = H; sgm.batch_inner = H;
    NKL(launch_gemm_simt(sgm, e->st));
    NKL(launch_softmax_rows(e->S, e->P, ad, (int64_t)H * rows, T, e->st, e->ragged ? e->dev_tvalid() + b : nullptr, (int64_t)H * rows));
    GemmArgs o;
    o.A = e->P; o.lda = T; o.sAi = (int64_t)rows * T; o.a_dtype = ad;
    o.B = (char*)e->kvbuf + (size_t)D * es; o.ldb = 2 * D; o.sBi = dh; o.b_dtype = ad; o.transB = 1;
    o.C = e->ctx; o.ldc = D; o.sCi = dh; o.c_dtype = ad;
    o.M = rows; o.N = dh; o.K = T; o.batch = H; o.batch_inner = H;
    NKL(launch_gemm_simt(o, e->st));
    GemmArgs g = nar_linear(e, e->ctx, D, p + "cout.w", p + "cout.b", e->dec, D, kF32, rows, D, D);
    g.residual = e->dx; g.ldr = D;
    NRET(nar_gemm(e, g));                                                                       // dec = x + cross_out
  }
  for (int i = c.dec_att_blocks; i < c.dec_att_blocks + c.dec_ffn_blocks; ++i) {
    NRET(ffn("dec" + std::to_string(i) + ".", e->dx, nullptr));
    NCK(cudaMemcpyAsync(e->dec, e->dx, (size_t)rows * D * 4, cudaMemcpyDeviceToDevice, e->st));
  }
  NKL(launch_layernorm(e->dec, D, nullptr, nullptr, e->dq, ad, D, rows, D, c.dec_ln_eps, e->st, e->use_pdl ? 1 : 0));
  NRET(nar_gemm(e, nar_linear(e, e->dq, D, "out.w", "out.b", e->dec_logits, c.vocab, kF32, rows, c.vocab, D)));
  row_argmax_kernel<<<(rows + 7) / 8, 256, 0, e->st>>>(e->dec_logits, rows, c.vocab, e->tokens + (int64_t)b * e->max_T);
  NKL(cudaGetLastError());
  e->last_rows = rows;
  return B200ASR_OK;
}

// decoder of every clip in one pass (stacked rows; clip boundaries in e->seg_off)
template <int DH>
int paraformer_decode_all(b200asr_nar* e, int rows) {
  const b200asr_nar_config& c = e->cfg;
  const int D = c.d_model, H = c.n_heads, T = e->T, B = e->B, ad = e->act, Fd = c.dec_ffn;
  para_gather_rows_kernel<<<rows, 128, 0, e->st>>>(e->acoustic, e->n_tok, e->seg_off, B, T, D, e->dec);
  NKL(cudaGetLastError());
  auto ffn = [&](const std::string& p, float* out) -> int {
    NKL(launch_layernorm(e->dec, D, nullptr, nullptr, e->dq, ad, D, rows, D, c.dec_ln_eps, e->st, e->use_pdl ? 1 : 0));
    {
      GemmArgs g = nar_linear(e, e->dq, D, p + "w1.w", p + "w1.b", e->f32buf, Fd, kF32, rows, Fd, D);
      g.act = kActRelu;
      NRET(nar_gemm(e, g));
    }
    NKL(launch_layernorm(e->f32buf, Fd, nullptr, nullptr, e->ffn, ad, Fd, rows, Fd, c.dec_ln_eps, e->st, e->use_pdl ? 1 : 0));
    return nar_gemm(e, nar_linear(e, e->ffn, Fd, p + "w2.w", p + "w2.b", out, D, kF32, rows, D, Fd));
  };
  const size_t smem = (size_t)4 * (DH + T) * sizeof(float);
  for (int i = 0; i < c.dec_att_blocks; ++i) {
    const std::string p = "dec" + std::to_string(i) + ".";
    NRET(ffn(p, e->dx));                                                                        // x = FFN(dec)
    NKL(launch_layernorm(e->dx, D, NWF(e, p + "norm2.g"), NWF(e, p + "norm2.b"), e->sa_in, kF32, D, rows, D, c.dec_ln_eps, e->st, e->use_pdl ? 1 : 0));
    fsmn_seg_kernel<<<rows, 256, 0, e->st>>>(e->sa_in, NWF(e, p + "fsmn.w"), e->dec, e->seg_off, B, D, c.fsmn_kernel, e->dx);   // x = dec + fsmn(norm2(x))
    NKL(cudaGetLastError());
    NKL(launch_layernorm(e->dx, D, nullptr, nullptr, e->dq, ad, D, rows, D, c.dec_ln_eps, e->st, e->use_pdl ? 1 : 0));
    NRET(nar_gemm(e, nar_linear(e, e->dq, D, p + "q.w", p + "q.b", e->qkv, D, ad, rows, D, D)));
    NRET(nar_gemm(e, nar_linear(e, e->xhat, D, p + "kv.w", p + "kv.b", e->kvbuf, 2 * D, ad, B * T, 2 * D, D)));
    const int total = rows * H;
    if (ad == kBF16) para_cross_attn_kernel<bf16, DH><<<(total + 3) / 4, 128, smem, e->st>>>((const bf16*)e->qkv, (const bf16*)e->kvbuf, e->seg_off, B, T, H, total, (bf16*)e->ctx, e->dev_tvalid());
    else para_cross_attn_kernel<float, DH><<<(total + 3) / 4, 128, smem, e->st>>>((const float*)e->qkv, (const float*)e->kvbuf, e->seg_off, B, T, H, total, (float*)e->ctx, e->dev_tvalid());
    NKL(cudaGetLastError());
    GemmArgs g = nar_linear(e, e->ctx, D, p + "cout.w", p + "cout.b", e->dec, D, kF32, rows, D, D);
    g.residual = e->dx; g.ldr = D;
    NRET(nar_gemm(e, g));                                                                       // dec = x + cross_out
  }
  for (int i = c.dec_att_blocks; i < c.dec_att_blocks + c.dec_ffn_blocks; ++i) {
    NRET(ffn("dec" + std::to_string(i) + ".", e->dx));
    NCK(cudaMemcpyAsync(e->dec, e->dx, (size_t)rows * D * 4, cudaMemcpyDeviceToDevice, e->st));
  }
  NKL(launch_layernorm(e->dec, D, nullptr, nullptr, e->dq, ad, D, rows, D, c.dec_ln_eps, e->st, e->use_pdl ? 1 : 0));
  NRET(nar_gemm(e, nar_linear(e, e->dq, D, "out.w", "out.b", e->dec_logits, c.vocab, kF32, rows, c.vocab, D)));
  row_argmax_kernel<<<(rows + 7) / 8, 256, 0, e->st>>>(e->dec_logits, rows, c.vocab, e->frame_ids);
  NKL(cudaGetLastError());
  para_scatter_tokens_kernel<<<(rows + 127) / 128, 128, 0, e->st>>>(e->frame_ids, e->seg_off, B, e->max_T, rows, e->tokens);
  NKL(cudaGetLastError());
  e->last_rows = rows;
  return B200ASR_OK;
}

// front end + encoder + CIF up to the fire scan: fixed shapes per (batch, clip length), so it replays as one CUDA graph
int paraformer_encoder(b200asr_nar* e) {
  const b200asr_nar_config& c = e->cfg;
  const int feat = c.n_mels * c.lfr_m, D = c.d_model, B = e->B, T = e->T, M = B * T, ad = e->act;
  NRET(nar_fbank(e));
  lfr_cmvn_kernel<<<dim3(T, B), 256, 0, e->st>>>(e->mel, e->frames, c.n_mels, c.lfr_m, c.lfr_n, e->T_lfr, 0, nullptr,
                                                 NWF(e, "cmvn_vars"), NWF(e, "encoder_input_bias"), nullptr, nullptr, e->lang, e->feats, e->dev_frames());
  NKL(cudaGetLastError());
  const int n_enc = c.n_blocks0 + c.n_blocks;
  for (int i = 0; i < n_enc; ++i) {
    const std::string p = "enc" + std::to_string(i) + ".";
    NRET(nar_block(e, p, p + "out.b", i == 0 ? e->feats : e->hidden, i == 0 ? feat : D));
  }
  NKL(launch_layernorm(e->hidden, D, NWF(e, "enc_after_norm.g"), NWF(e, "enc_after_norm.b"), e->enc_out, kF32, D, M, D, c.ln_eps, e->st, e->use_pdl ? 1 : 0));
  NKL(launch_layernorm(e->hidden, D, NWF(e, "enc_after_norm.g"), NWF(e, "enc_after_norm.b"), e->xhat, ad, D, M, D, c.ln_eps, e->st, e->use_pdl ? 1 : 0));
  // CIF: conv k over time as a GEMM on the zero-padded, overlapping-row view (row t = k*D contiguous values from padded row t)
  const int pad = (c.cif_kernel - 1) / 2;
  if (ad == kBF16) pad_rows_kernel<bf16><<<dim3(T + 2 * pad, B), 128, 0, e->st>>>(e->enc_out, T, D, pad, (bf16*)e->enc_pad, e->dev_tvalid());
  else pad_rows_kernel<float><<<dim3(T + 2 * pad, B), 128, 0, e->st>>>(e->enc_out, T, D, pad, (float*)e->enc_pad, e->dev_tvalid());
  NKL(cudaGetLastError());
  {
    GemmArgs g = nar_linear(e, e->enc_pad, D, "cif.conv.w", "cif.conv.b", e->conv_out, D, ad, T, D, c.cif_kernel * D);
    g.sAo = (int64_t)(T + 2 * pad) * D; g.sCo = (int64_t)T * D; g.batch = B; g.act = kActRelu;
    NRET(nar_gemm(e, g));
  }
  if (ad == kBF16) cif_alpha_kernel<bf16><<<(M + 7) / 8, 256, 0, e->st>>>((const bf16*)e->conv_out, NWF(e, "cif.out.w"), NWF(e, "cif.out.b"), M, D, e->alphas, T);
  else cif_alpha_kernel<float><<<(M + 7) / 8, 256, 0, e->st>>>((const float*)e->conv_out, NWF(e, "cif.out.w"), NWF(e, "cif.out.b"), M, D, e->alphas, T);
  NKL(cudaGetLastError());
  cif_scan_kernel<<<B, 256, (size_t)2 * (T + 1) * sizeof(float), e->st>>>(e->alphas, c.tail_threshold, e->enc_out, T, D, e->acoustic, e->n_tok, e->dev_tvalid());
  NKL(cudaGetLastError());
  return B200ASR_OK;
}

int nar_graph_run(b200asr_nar* e, int (*fn)(b200asr_nar*));

int paraformer_forward(b200asr_nar* e) {
  const int B = e->B, T = e->T;
  NRET(e->use_graph ? nar_graph_run(e, paraformer_encoder) : paraformer_encoder(e));
  // the token count sizes the decoder: one small device->host read (the reference graph has the same data-dependent shape)
  int* h_n = e->h_pinned;
  NCK(cudaMemcpyAsync(h_n, e->n_tok, (size_t)B * 4, cudaMemcpyDeviceToHost, e->st));
  NCK(cudaStreamSynchronize(e->st));
  std::vector<int> counts(h_n, h_n + B);
  for (int b = 0; b < B; ++b)
    if (counts[b] > T + 1) return e->fail(B200ASR_E_CUDA, "CIF fired more tokens than frames");
  const int dh = e->cfg.d_model / e->cfg.n_heads;
  if (e->batched_decoder && (dh == 64 || dh == 128)) {
    int* h_off = e->h_pinned + B;
    h_off[0] = 0;
    for (int b = 0; b < B; ++b) h_off[b + 1] = h_off[b] + (counts[b] > 0 ? counts[b] : 1);
    NCK(cudaMemcpyAsync(e->seg_off, h_off, (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, e->st));
    NRET(dh == 128 ? paraformer_decode_all<128>(e, h_off[B]) : paraformer_decode_all<64>(e, h_off[B]));
  } else {
    for (int b = 0; b < B; ++b) NRET(paraformer_decode_one(e, b, counts[b]));
  }
  NCK(cudaMemcpyAsync(e->lens, e->n_tok, (size_t)B * 4, cudaMemcpyDeviceToDevice, e->st));
  return B200ASR_OK;
}

int nar_graph_run(b200asr_nar* e, int (*fn)(b200asr_nar*)) {
  const int dtype_key = e->pcm_dtype + (e->ragged ? 16 : 0);          // (per-clip lengths are read from device memory: one graph per shape)
  if (!e->graph || e->graph_B != e->B || e->graph_N != e->n_samples || e->graph_dtype != dtype_key) {
    if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }
    cudaGraph_t g = nullptr;
    const int64_t before = e->launches;
    NCK(cudaStreamBeginCapture(e->st, cudaStreamCaptureModeThreadLocal));
    const int r = fn(e);
    const cudaError_t ce = cudaStreamEndCapture(e->st, &g);
    e->graph_nodes = e->launches - before;
    e->launches = before;
    if (r != B200ASR_OK) { if (g) cudaGraphDestroy(g); return r; }
    if (ce != cudaSuccess) return e->cuda_fail(ce, "cudaStreamEndCapture");
    const cudaError_t ci = cudaGraphInstantiate(&e->graph, g, 0);
    cudaGraphDestroy(g);
    if (ci != cudaSuccess) return e->cuda_fail(ci, "cudaGraphInstantiate");
    e->graph_B = e->B; e->graph_N = e->n_samples; e->graph_dtype = dtype_key;
  }
  NCK(cudaGraphLaunch(e->graph, e->st));
  e->launches += e->graph_nodes;
  return B200ASR_OK;
}

int nar_forward(b200asr_nar* e) {
  if (e->cfg.kind == B200ASR_NAR_PARAFORMER) return paraformer_forward(e);      // data-dependent decoder size: host read inside
  return e->use_graph ? nar_graph_run(e, sensevoice_forward) : sensevoice_forward(e);
}

}  // namespace

extern "C" {

const char* b200asr_nar_last_error(const b200asr_nar* e) { return e ? e->err.c_str() : g_nar_create_error.c_str(); }

int b200asr_nar_create(const b200asr_nar_config* cfg, b200asr_nar** out) {
  if (!cfg || !out) { g_nar_create_error = "null argument"; return B200ASR_E_INVALID; }
  *out = nullptr;
  if (cfg->kind != B200ASR_NAR_SENSEVOICE && cfg->kind != B200ASR_NAR_PARAFORMER) { g_nar_create_error = "unknown model kind"; return B200ASR_E_INVALID; }
  if (cfg->kind == B200ASR_NAR_PARAFORMER && (cfg->dec_att_blocks < 0 || cfg->dec_ffn_blocks < 0 || cfg->dec_ffn <= 0 || cfg->dec_ffn % 8 || cfg->dec_ffn > 2048 ||
                                              cfg->cif_kernel < 1 || (cfg->cif_kernel & 1) == 0)) { g_nar_create_error = "invalid Paraformer decoder dimensions"; return B200ASR_E_INVALID; }
  if (cfg->d_model <= 0 || cfg->n_heads <= 0 || cfg->d_model % cfg->n_heads || cfg->d_model % 8 || cfg->ffn % 8 ||
      (cfg->n_mels * cfg->lfr_m) % 8 || cfg->nfft / 2 + 1 > kFbBins || cfg->win <= 0 || cfg->hop <= 0 || cfg->vocab <= 0 ||
      cfg->max_batch <= 0 || cfg->max_samples < cfg->win || cfg->n_prompt < 0 || cfg->fsmn_kernel < 1 || (cfg->fsmn_kernel & 1) == 0) {
    g_nar_create_error = "invalid model dimensions"; return B200ASR_E_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) {
    g_nar_create_error = "no CUDA device: the b200asr engine has no CPU fallback"; return B200ASR_E_NOGPU;
  }
  cudaDeviceProp prop;
  if (cudaSetDevice(cfg->device) != cudaSuccess || cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) {
    g_nar_create_error = "cudaSetDevice failed"; return B200ASR_E_CUDA;
  }
  if (prop.major != 10) { g_nar_create_error = "this build targets sm_100a only"; return B200ASR_E_NOGPU; }
  b200asr_nar* e = new b200asr_nar();
  e->cfg = *cfg;
  e->num_sms = prop.multiProcessorCount;
  e->act = cfg->precision == B200ASR_PRECISION_BF16 ? kBF16 : kF32;
  e->es = dtype_size(e->act);
  if (cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking) != cudaSuccess) { g_nar_create_error = "cudaStreamCreate failed"; delete e; return B200ASR_E_CUDA; }
  *out = e;
  return B200ASR_OK;
}

void b200asr_nar_destroy(b200asr_nar* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  cudaStreamSynchronize(e->st);
  if (e->graph) cudaGraphExecDestroy(e->graph);
  for (auto& kv : e->w) cudaFree(kv.second.ptr);
  void* bufs[] = {e->d_meta, e->basis_t, e->stage_buf, e->pcm, e->mel, e->feats, e->hidden, e->resid, e->xhat, e->qkv, e->ctx, e->ffn, e->P, e->S,
                  e->logits, e->enc_out, e->frame_ids, e->tokens, e->lens, e->lang, e->enc_pad, e->conv_out, e->kvbuf, e->dq, e->alphas,
                  e->acoustic, e->dec, e->dx, e->f32buf, e->sa_in, e->dec_logits, e->n_tok, e->seg_off};
  for (void* p : bufs) if (p) cudaFree(p);
  if (e->h_pinned) cudaFreeHost(e->h_pinned);
  cudaStreamDestroy(e->st);
  delete e;
}

int b200asr_nar_set_tensor(b200asr_nar* e, const char* name_c, const float* host, int64_t numel) {
  if (!e || !name_c || !host || numel <= 0) return B200ASR_E_INVALID;
  NCK(cudaSetDevice(e->cfg.device));
  const std::string name(name_c);
  const b200asr_nar_config& c = e->cfg;
  const int F = c.nfft / 2 + 1;
  if (name == "fbank_kernel") {        // [2F][win] Conv1d weight -> [win][2F] so a thread's frequency bin is the fast axis
    if (numel != (int64_t)2 * F * c.win) return e->fail(B200ASR_E_INVALID, "fbank_kernel size mismatch");
    std::vector<float> t((size_t)numel);
    for (int r = 0; r < 2 * F; ++r)
      for (int k = 0; k < c.win; ++k) t[(size_t)k * 2 * F + r] = host[(size_t)r * c.win + k];
    if (!e->basis_t) NCK(cudaMalloc(&e->basis_t, (size_t)numel * 4));
    NCK(cudaMemcpyAsync(e->basis_t, t.data(), (size_t)numel * 4, cudaMemcpyHostToDevice, e->st));
    NCK(cudaStreamSynchronize(e->st));
  }
  std::vector<float> relaid;
  if (name == "cif.conv.w") {         // Conv1d weight [co][ci][k] -> [co][k*D + ci]: the GEMM's K axis walks k then ci
    const int D = c.d_model, K = c.cif_kernel;
    if (numel != (int64_t)D * D * K) return e->fail(B200ASR_E_INVALID, "cif.conv.w size mismatch");
    relaid.resize((size_t)numel);
    for (int co = 0; co < D; ++co)
      for (int ci = 0; ci < D; ++ci)
        for (int k = 0; k < K; ++k) relaid[((size_t)co * K + k) * D + ci] = host[((size_t)co * D + ci) * K + k];
    host = relaid.data();
  }
  NarTensor t;
  t.numel = numel;
  t.dtype = nar_is_matrix(name) ? e->act : kF32;
  auto it = e->w.find(name);
  if (it != e->w.end()) { cudaFree(it->second.ptr); e->w.erase(it); }
  NCK(cudaMalloc(&t.ptr, (size_t)numel * dtype_size(t.dtype)));
  if (t.dtype == kF32) {
    NCK(cudaMemcpyAsync(t.ptr, host, (size_t)numel * 4, cudaMemcpyHostToDevice, e->st));
    NCK(cudaStreamSynchronize(e->st));
  } else {
    if (numel > e->stage_cap) {
      if (e->stage_buf) cudaFree(e->stage_buf);
      e->stage_buf = nullptr; e->stage_cap = 0;
      NCK(cudaMalloc(&e->stage_buf, (size_t)numel * 4));
      e->stage_cap = numel;
    }
    NCK(cudaMemcpyAsync(e->stage_buf, host, (size_t)numel * 4, cudaMemcpyHostToDevice, e->st));
    nar_f32_to_bf16<<<1024, 256, 0, e->st>>>(e->stage_buf, (bf16*)t.ptr, numel);
    NCK(cudaGetLastError());
    NCK(cudaStreamSynchronize(e->st));
  }
  e->w[name] = t;
  e->finalized = false;
  return B200ASR_OK;
}

int b200asr_nar_finalize_weights(b200asr_nar* e) {
  if (!e) return B200ASR_E_INVALID;
  NCK(cudaSetDevice(e->cfg.device));
  const b200asr_nar_config& c = e->cfg;
  const int64_t F = c.nfft / 2 + 1, feat = c.n_mels * c.lfr_m, D = c.d_model, f = c.ffn;
  e->max_frames = (c.max_samples - c.win) / c.hop + 1;
  const int max_lfr = (e->max_frames + c.lfr_n - 1) / c.lfr_n;
  const bool para = c.kind == B200ASR_NAR_PARAFORMER;
  e->max_T = max_lfr + c.n_prompt + (para ? 1 : 0);            // Paraformer: up to T + 1 decoder rows (tail fire)
  NRET(nar_need(e, "fbank_kernel", 2 * F * c.win)); NRET(nar_need(e, "mel_filters", F * c.n_mels));
  NRET(nar_need(e, "cmvn_vars", feat));
  if (!para) {
    NRET(nar_need(e, "cmvn_means", feat));
    NRET(nar_need(e, "language_embed", (int64_t)c.n_lang * feat)); NRET(nar_need(e, "system_embed", (int64_t)(c.n_prompt - 1) * feat));
    auto it = e->w.find("speech_position");
    if (it == e->w.end()) return e->fail(B200ASR_E_MISSING, "missing weight tensor 'speech_position'");
    if (it->second.numel < (int64_t)max_lfr * feat) return e->fail(B200ASR_E_INVALID, "speech_position shorter than max_samples needs");
    const int n_all = c.n_blocks0 + c.n_blocks + c.n_tp_blocks;
    for (int i = 0; i < n_all; ++i) {
      const std::string p = "blk" + std::to_string(i) + ".";
      const int64_t din = i == 0 ? feat : D;
      NRET(nar_need(e, p + "norm1.g", din)); NRET(nar_need(e, p + "norm1.b", din));
      NRET(nar_need(e, p + "qkv.w", 3 * D * din)); NRET(nar_need(e, p + "qkv.b", 3 * D));
      NRET(nar_need(e, p + "fsmn.w", D * c.fsmn_kernel)); NRET(nar_need(e, p + "fsmn.b", D));
      NRET(nar_need(e, p + "out.w", D * D));
      NRET(nar_need(e, p + "norm2.g", D)); NRET(nar_need(e, p + "norm2.b", D));
      NRET(nar_need(e, p + "w1.w", f * D)); NRET(nar_need(e, p + "w1.b", f));
      NRET(nar_need(e, p + "w2.w", D * f)); NRET(nar_need(e, p + "w2.b", D));
    }
    NRET(nar_need(e, "after_norm.g", D)); NRET(nar_need(e, "after_norm.b", D));
    NRET(nar_need(e, "tp_norm.g", D)); NRET(nar_need(e, "tp_norm.b", D));
    NRET(nar_need(e, "ctc.w", (int64_t)c.vocab * D)); NRET(nar_need(e, "ctc.b", c.vocab));
  } else {
    const int64_t Fd = c.dec_ffn;
    {
      auto it = e->w.find("encoder_input_bias");
      if (it == e->w.end()) return e->fail(B200ASR_E_MISSING, "missing weight tensor 'encoder_input_bias'");
      if (it->second.numel < (int64_t)max_lfr * feat) return e->fail(B200ASR_E_INVALID, "encoder_input_bias shorter than max_samples needs");
    }
    for (int i = 0; i < c.n_blocks0 + c.n_blocks; ++i) {
      const std::string p = "enc" + std::to_string(i) + ".";
      const int64_t din = i == 0 ? feat : D;
      NRET(nar_need(e, p + "qkv.w", 3 * D * din)); NRET(nar_need(e, p + "qkv.b", 3 * D));
      NRET(nar_need(e, p + "fsmn.w", D * c.fsmn_kernel));
      NRET(nar_need(e, p + "out.w", D * D)); NRET(nar_need(e, p + "out.b", D));
      NRET(nar_need(e, p + "w1.w", f * D)); NRET(nar_need(e, p + "w1.b", f));
      NRET(nar_need(e, p + "w2.w", D * f)); NRET(nar_need(e, p + "w2.b", D));
    }
    NRET(nar_need(e, "enc_after_norm.g", D)); NRET(nar_need(e, "enc_after_norm.b", D));
    NRET(nar_need(e, "cif.conv.w", D * D * c.cif_kernel)); NRET(nar_need(e, "cif.conv.b", D));
    NRET(nar_need(e, "cif.out.w", D)); NRET(nar_need(e, "cif.out.b", 1));
    for (int i = 0; i < c.dec_att_blocks + c.dec_ffn_blocks; ++i) {
      const std::string p = "dec" + std::to_string(i) + ".";
      NRET(nar_need(e, p + "w1.w", Fd * D)); NRET(nar_need(e, p + "w1.b", Fd));
      NRET(nar_need(e, p + "w2.w", D * Fd)); NRET(nar_need(e, p + "w2.b", D));
      if (i < c.dec_att_blocks) {
        NRET(nar_need(e, p + "norm2.g", D)); NRET(nar_need(e, p + "norm2.b", D));
        NRET(nar_need(e, p + "fsmn.w", D * c.fsmn_kernel));
        NRET(nar_need(e, p + "q.w", D * D)); NRET(nar_need(e, p + "q.b", D));
        NRET(nar_need(e, p + "kv.w", 2 * D * D)); NRET(nar_need(e, p + "kv.b", 2 * D));
        NRET(nar_need(e, p + "cout.w", D * D)); NRET(nar_need(e, p + "cout.b", D));
      }
    }
    NRET(nar_need(e, "out.w", (int64_t)c.vocab * D)); NRET(nar_need(e, "out.b", c.vocab));
    if (e->w.find("zero_bias") == e->w.end()) {
      NarTensor z; z.numel = D; z.dtype = kF32;
      NCK(cudaMalloc(&z.ptr, (size_t)D * 4));
      NCK(cudaMemsetAsync(z.ptr, 0, (size_t)D * 4, e->st));
      e->w["zero_bias"] = z;
    }
  }
  if (e->finalized) return B200ASR_OK;
  if (e->stage_buf) { cudaFree(e->stage_buf); e->stage_buf = nullptr; e->stage_cap = 0; }
  if (!e->pcm) {
    const size_t es = e->es;
    const int64_t B = c.max_batch, M = B * e->max_T, H = c.n_heads, Tm = e->max_T;
    const int64_t wide = feat > D ? feat : D;
    NRET(nar_alloc(e, &e->pcm, (size_t)B * c.max_samples * 4));
    NRET(nar_alloc(e, &e->mel, (size_t)B * e->max_frames * c.n_mels * 4));
    NRET(nar_alloc(e, &e->feats, (size_t)M * feat * 4));
    NRET(nar_alloc(e, &e->hidden, (size_t)M * D * 4));
    NRET(nar_alloc(e, &e->resid, (size_t)M * D * 4));
    NRET(nar_alloc(e, &e->enc_out, (size_t)M * D * 4));
    NRET(nar_alloc(e, &e->xhat, (size_t)M * wide * es));
    NRET(nar_alloc(e, &e->qkv, (size_t)M * 3 * D * es));
    NRET(nar_alloc(e, &e->ctx, (size_t)M * D * es));
    NRET(nar_alloc(e, &e->ffn, (size_t)M * (f > c.dec_ffn ? f : c.dec_ffn) * es));
    NRET(nar_alloc(e, &e->S, (size_t)B * H * Tm * Tm * 4));
    NRET(nar_alloc(e, &e->P, (size_t)B * H * Tm * Tm * es));
    NRET(nar_alloc(e, &e->logits, (size_t)M * c.vocab * 4));
    NRET(nar_alloc(e, &e->frame_ids, (size_t)M * 4));
    NRET(nar_alloc(e, &e->tokens, (size_t)B * Tm * 4));
    NRET(nar_alloc(e, &e->lens, (size_t)B * 4));
    NRET(nar_alloc(e, &e->lang, (size_t)B * 4));
    NRET(nar_alloc(e, &e->d_meta, (size_t)2 * B * 4));
    if (c.kind == B200ASR_NAR_PARAFORMER) {
      const int64_t pad = (c.cif_kernel - 1) / 2;
      NRET(nar_alloc(e, &e->enc_pad, (size_t)B * (Tm + 2 * pad) * D * es));
      NRET(nar_alloc(e, &e->conv_out, (size_t)M * D * es));
      NRET(nar_alloc(e, &e->kvbuf, (size_t)M * 2 * D * es));            // decoder buffers hold the stacked rows of the whole batch
      NRET(nar_alloc(e, &e->dq, (size_t)M * D * es));
      NRET(nar_alloc(e, &e->alphas, (size_t)B * (Tm + 1) * 4));
      NRET(nar_alloc(e, &e->acoustic, (size_t)B * (Tm + 1) * D * 4));
      NRET(nar_alloc(e, &e->dec, (size_t)M * D * 4));
      NRET(nar_alloc(e, &e->dx, (size_t)M * D * 4));
      NRET(nar_alloc(e, &e->sa_in, (size_t)M * D * 4));
      NRET(nar_alloc(e, &e->f32buf, (size_t)M * c.dec_ffn * 4));
      NRET(nar_alloc(e, &e->dec_logits, (size_t)M * c.vocab * 4));
      NRET(nar_alloc(e, &e->seg_off, (size_t)(B + 1) * 4));
      NRET(nar_alloc(e, &e->n_tok, (size_t)B * 4));
    }
    NCK(cudaMallocHost(&e->h_pinned, (size_t)B * (Tm + 2) * 4 + 64));
    const int span = (kFbFrames - 1) * c.hop + c.win;
    NCK(cudaFuncSetAttribute(kaldi_fbank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((span + 3 * kFbFrames * F) * sizeof(float))));
  }
  NCK(cudaStreamSynchronize(e->st));
  e->finalized = true;
  return B200ASR_OK;
}

static int nar_upload(b200asr_nar* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples,
                      const int32_t* language_idx, const int32_t* lens = nullptr) {
  const b200asr_nar_config& c = e->cfg;
  if (!e->finalized) return e->fail(B200ASR_E_INVALID, "weights not finalized");
  const bool para = c.kind == B200ASR_NAR_PARAFORMER;
  if (!pcm_host || (!language_idx && !para)) return e->fail(B200ASR_E_INVALID, "null argument");
  if (batch <= 0 || batch > c.max_batch) return e->fail(B200ASR_E_INVALID, "batch out of range");
  if (n_samples < c.win || n_samples > c.max_samples) return e->fail(B200ASR_E_INVALID, "n_samples out of range");
  if (pcm_dtype != B200ASR_PCM_I16 && pcm_dtype != B200ASR_PCM_F32) return e->fail(B200ASR_E_INVALID, "bad pcm dtype");
  if (!para)
    for (int b = 0; b < batch; ++b)
      if (language_idx[b] < 0 || language_idx[b] >= c.n_lang) return e->fail(B200ASR_E_INVALID, "language_idx out of range");
  e->B = batch; e->n_samples = n_samples; e->pcm_dtype = pcm_dtype;
  e->frames = (n_samples - c.win) / c.hop + 1;
  e->T_lfr = (e->frames + c.lfr_n - 1) / c.lfr_n;
  e->T = e->T_lfr + c.n_prompt;
  bool ragged = false;
  if (lens) {
    int longest = 0;
    for (int b = 0; b < batch; ++b) {
      if (lens[b] < c.win || lens[b] > n_samples) return e->fail(B200ASR_E_INVALID, "per-clip length out of [win, n_samples]");
      longest = longest > lens[b] ? longest : lens[b];
      ragged |= lens[b] != n_samples;
    }
    if (ragged) {
      if ((longest - c.win) / c.hop + 1 != e->frames) return e->fail(B200ASR_E_INVALID, "n_samples must be the longest clip's length (within one hop)");
      e->h_meta.assign((size_t)2 * c.max_batch, 0);
      for (int b = 0; b < batch; ++b) {
        const int fr = (lens[b] - c.win) / c.hop + 1;
        e->h_meta[b] = fr;
        e->h_meta[c.max_batch + b] = (fr + c.lfr_n - 1) / c.lfr_n + c.n_prompt;
      }
      NCK(cudaMemcpyAsync(e->d_meta, e->h_meta.data(), e->h_meta.size() * sizeof(int), cudaMemcpyHostToDevice, e->st));
    }
  }
  e->ragged = ragged;
  NCK(cudaMemcpyAsync(e->pcm, pcm_host, (size_t)batch * n_samples * (pcm_dtype == B200ASR_PCM_F32 ? 4 : 2), cudaMemcpyHostToDevice, e->st));
  if (!para) NCK(cudaMemcpyAsync(e->lang, language_idx, (size_t)batch * 4, cudaMemcpyHostToDevice, e->st));
  return B200ASR_OK;
}

static int nar_fetch(b200asr_nar* e, int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out) {
  if (!tokens_out || !lens_out || tokens_ld <= 0) return e->fail(B200ASR_E_INVALID, "null output");
  const int batch = e->B;
  int* h_len = e->h_pinned;
  int* h_tok = e->h_pinned + batch;
  NCK(cudaMemcpyAsync(h_len, e->lens, (size_t)batch * 4, cudaMemcpyDeviceToHost, e->st));
  NCK(cudaMemcpyAsync(h_tok, e->tokens, (size_t)batch * e->max_T * 4, cudaMemcpyDeviceToHost, e->st));
  NCK(cudaStreamSynchronize(e->st));
  for (int b = 0; b < batch; ++b) {
    const int n = h_len[b] < tokens_ld ? h_len[b] : tokens_ld;
    lens_out[b] = n;
    memcpy(tokens_out + (size_t)b * tokens_ld, h_tok + (size_t)b * e->max_T, (size_t)n * 4);
  }
  return B200ASR_OK;
}

int b200asr_nar_run(b200asr_nar* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples,
                    const int32_t* language_idx, int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out) {
  if (!e) return B200ASR_E_INVALID;
  NCK(cudaSetDevice(e->cfg.device));
  NRET(nar_upload(e, pcm_host, pcm_dtype, batch, n_samples, language_idx));
  NRET(nar_forward(e));
  return nar_fetch(e, tokens_out, tokens_ld, lens_out);
}

// Ragged batch: clip b has lens[b] samples; rows of pcm_host are n_samples (= the longest clip) apart.  Every clip gets the
// result of running the reference's dynamic-length graph on it alone: its own frame count, LFR tail, FSMN / CIF-conv zero
// padding, attention keys, CTC roll and CIF tail.
int b200asr_nar_run_ragged(b200asr_nar* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples,
                           const int32_t* lens, const int32_t* language_idx, int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out) {
  if (!e) return B200ASR_E_INVALID;
  NCK(cudaSetDevice(e->cfg.device));
  if (!lens) return e->fail(B200ASR_E_INVALID, "null lens");
  NRET(nar_upload(e, pcm_host, pcm_dtype, batch, n_samples, language_idx, lens));
  NRET(nar_forward(e));
  return nar_fetch(e, tokens_out, tokens_ld, lens_out);
}

int b200asr_nar_upload_ragged(b200asr_nar* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples,
                              const int32_t* lens, const int32_t* language_idx) {
  if (!e) return B200ASR_E_INVALID;
  NCK(cudaSetDevice(e->cfg.device));
  if (!lens) return e->fail(B200ASR_E_INVALID, "null lens");
  NRET(nar_upload(e, pcm_host, pcm_dtype, batch, n_samples, language_idx, lens));
  NCK(cudaStreamSynchronize(e->st));
  return B200ASR_OK;
}

int b200asr_nar_upload(b200asr_nar* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples,
                       const int32_t* language_idx) {
  if (!e) return B200ASR_E_INVALID;
  NCK(cudaSetDevice(e->cfg.device));
  NRET(nar_upload(e, pcm_host, pcm_dtype, batch, n_samples, language_idx));
  NCK(cudaStreamSynchronize(e->st));
  return B200ASR_OK;
}

int b200asr_nar_run_resident(b200asr_nar* e, int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out) {
  if (!e) return B200ASR_E_INVALID;
  NCK(cudaSetDevice(e->cfg.device));
  if (e->B <= 0) return e->fail(B200ASR_E_INVALID, "no PCM uploaded");
  NRET(nar_forward(e));
  return nar_fetch(e, tokens_out, tokens_ld, lens_out);
}

int b200asr_nar_get_stage(b200asr_nar* e, const char* name_c, float* out, int64_t capacity, int64_t* numel_out) {
  if (!e || !name_c || !out) return B200ASR_E_INVALID;
  NCK(cudaSetDevice(e->cfg.device));
  const b200asr_nar_config& c = e->cfg;
  const std::string name(name_c);
  if (e->B <= 0) return e->fail(B200ASR_E_INVALID, "get_stage before run");
  const int64_t B = e->B, T = e->T, feat = c.n_mels * c.lfr_m;
  const void* src = nullptr; int64_t n = 0; bool is_int = false;
  if (name == "mel") { src = e->mel; n = B * e->frames * c.n_mels; }
  else if (name == "feats") { src = e->feats; n = B * T * feat; }
  else if (name == "enc_out") { src = e->enc_out; n = B * T * c.d_model; }
  else if (name == "logits") { src = e->logits; n = B * T * c.vocab; }
  else if (name == "frame_ids") { src = e->frame_ids; n = B * T; is_int = true; }
  else if (name == "alphas" && e->alphas) { src = e->alphas; n = B * (T + 1); }
  else if (name == "acoustic" && e->acoustic) { src = e->acoustic; n = B * (T + 1) * c.d_model; }
  else if (name == "dec_logits" && e->dec_logits) { src = e->dec_logits; n = (int64_t)e->last_rows * c.vocab; }
  else if (name == "n_tok" && e->n_tok) { src = e->n_tok; n = B; is_int = true; }
  else return e->fail(B200ASR_E_INVALID, "unknown stage '" + name + "'");
  if (n > capacity) return e->fail(B200ASR_E_INVALID, "stage buffer too small");
  NCK(cudaStreamSynchronize(e->st));
  if (!is_int) {
    NCK(cudaMemcpy(out, src, (size_t)n * 4, cudaMemcpyDeviceToHost));
  } else {
    std::vector<int> h((size_t)n);
    NCK(cudaMemcpy(h.data(), src, (size_t)n * 4, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n; ++i) out[i] = (float)h[(size_t)i];
  }
  if (numel_out) *numel_out = n;
  return B200ASR_OK;
}

int64_t b200asr_nar_kernel_launches(const b200asr_nar* e) { return e ? e->launches : 0; }
int b200asr_nar_set_option(b200asr_nar* e, const char* key, int64_t value) {
  if (!e || !key) return B200ASR_E_INVALID;
  if (!strcmp(key, "graph")) { e->use_graph = value != 0; return B200ASR_OK; }
  if (!strcmp(key, "batched_decoder")) { e->batched_decoder = value != 0; return B200ASR_OK; }
  if (!strcmp(key, "pdl")) { e->use_pdl = value != 0; if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; } return B200ASR_OK; }
  if (!strcmp(key, "attn_tc")) { e->use_attn_tc = value != 0; if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; } return B200ASR_OK; }
  return e->fail(B200ASR_E_INVALID, std::string("unknown option ") + key);
}
void* b200asr_nar_stream(b200asr_nar* e) { return e ? (void*)e->st : nullptr; }

}  // extern "C"
