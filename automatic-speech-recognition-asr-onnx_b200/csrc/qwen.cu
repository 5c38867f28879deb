// Qwen3-ASR path: Whisper-style log-mel -> 100-frame chunks -> 3 x (Conv2d stride 2 + tanh-GELU) -> Linear -> windowed
// encoder (8 chunks = 104 tokens per window, additive -128 key mask) -> ln_post/proj1/proj2 -> prompt concat ->
// decoder (RMS norm, fused QKV, per-head QK-norm, RoPE, grouped-query causal attention over a resident KV cache,
// SwiGLU) -> tied/untied lm_head -> arg-max with stop latch.
// Replaces prefill_session.run / embed_session.run + decode_session.run of
// /root/reference/Qwen_ASR/Inference_Qwen_ASR_ONNX.py:656,703-717; the math follows QWEN3_ASR_ENCODER.forward
// (/root/reference/Qwen_ASR/Export_Qwen_ASR.py:850-927), QWEN3_ASR_ROTARY_MASK_* (:933-1025) and
// QWEN3_ASR_DECODER_MAIN.forward (:1265-1336) with the folds of :829-848 and :1141-1190 done once on the host.
//
// Data layout in HBM: conv activations are channel-last [chunk][time][mel][channel] so the last conv's output IS the
// [chunk*13][16*C] operand of conv_out (its weight columns are permuted once on the host); conv2/conv3 are im2col +
// tcgen05 GEMM with the tanh-GELU in the epilogue; the residual streams stay fp32, GEMM operands are bf16 in bf16 mode.
// KV cache: [layer][utterance][kv_head][position][head_dim] in the activation dtype.  The prefill runs through the
// engine's GEMMs and a tiled causal attention; a decode step (M = batch rows) is 5 launches per layer -- four
// qwen_gemv_kernel launches that keep a whole weight matrix in flight (raw-register prefetch, one resident wave) and one
// attention launch fused with QK-norm / RoPE / cache append -- replayed as one CUDA graph with programmatic dependent
// launches; positions come from the device-side DecState, not launch arguments.
#include "common.cuh"
#include "../../include/b200asr.h"

#include <climits>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace b200asr;

namespace {

constexpr int kChunk = 100;        // mel frames per conv chunk (2 * n_window, Export_Qwen_ASR.py:744)
constexpr int kChunkTok = 13;      // tokens a full chunk yields after three stride-2 convs (:519-527)

// Programmatic dependent launch: a decode-step kernel is allowed to start while its predecessor drains; everything that
// does not depend on the predecessor (weight / cache-row requests) is issued first, then pdl_wait() orders the rest.
// Both are no-ops for a kernel launched without the attribute.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

template <typename T> __device__ __forceinline__ void store_as(void* p, int64_t i, float v) { reinterpret_cast<T*>(p)[i] = from_f<T>(v); }

// ---- features: max(x, amax - 8) -> (x + 4) / 4, zero rows up to the chunk multiple (:857-865) ----
//      `frames` = rows of the batch grid (the longest clip); a clip's own rows end at n_per_clip[b] / hop (ragged batches)
__global__ void qwen_feat_kernel(const float* __restrict__ mel_raw, const int* __restrict__ max_key, const int* __restrict__ n_per_clip,
                                 int hop, int frames, int frames_pad, int n_mels, float* __restrict__ feat) {
  const int b = blockIdx.y;
  const float floor_v = key_to_float(max_key[b]) - 8.0f;
  const int64_t n = (int64_t)frames_pad * n_mels, ld = (int64_t)frames * n_mels, nv = (int64_t)(n_per_clip[b] / hop) * n_mels;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i < nv) v = (fmaxf(mel_raw[b * ld + i], floor_v) + 4.0f) * 0.25f;
    feat[b * n + i] = v;
  }
}

// ---- conv2d1 (1 -> C, 3x3, stride 2, pad 1) + tanh-GELU, channel-last output [chunk][50][64][C]: weights and bias live in
//      shared memory, a thread keeps the 9 input taps of one output position in registers and emits 8 channels per store ----
template <typename OutT>
__global__ void __launch_bounds__(256)
qwen_conv1_kernel(const float* __restrict__ feat /*[chunks][100][n_mels]*/, const float* __restrict__ w /*[C][3 mel][3 time]*/,
                  const float* __restrict__ bias, int n_mels, int C, int To, int Fo, int64_t total /*positions x C/8*/, OutT* __restrict__ out) {
  extern __shared__ float cw[];                  // [C][9] weights | [C] bias
  for (int i = threadIdx.x; i < C * 9; i += blockDim.x) cw[i] = w[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) cw[C * 9 + i] = bias[i];
  __syncthreads();
  // work item -> (channel group, position): gw neighbouring lanes take neighbouring 8-channel groups of one position (one
  // contiguous store run), the next lanes the next positions; a group block [gw] is the slowest index so the weight reads of a
  // warp touch few distinct rows
  const int cg = C / 8;
  const int gw = (cg % 4 == 0) ? 4 : ((cg % 2 == 0) ? 2 : 1);
  const int64_t n_pos = total / cg;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g_lo = (int)(i % gw);
    int64_t r = (i / gw) % n_pos;
    const int c0 = ((int)(i / (gw * n_pos)) * gw + g_lo) * 8;
    const int fo = (int)(r % Fo); r /= Fo;
    const int to = (int)(r % To);
    const int64_t chunk = r / To;
    const float* in = feat + chunk * kChunk * n_mels;
    float tap[9];
#pragma unroll
    for (int kf = 0; kf < 3; ++kf) {
      const int f = 2 * fo - 1 + kf;
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) {
        const int t = 2 * to - 1 + kt;
        tap[kf * 3 + kt] = (f >= 0 && f < n_mels && t >= 0 && t < kChunk) ? in[t * n_mels + f] : 0.f;
      }
    }
    float o[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float* wc = cw + (c0 + c) * 9;
      float acc = cw[C * 9 + c0 + c];
#pragma unroll
      for (int k = 0; k < 9; ++k) acc = fmaf(wc[k], tap[k], acc);
      o[c] = gelu_tanh(acc);
    }
    OutT* dst = out + ((chunk * To + to) * Fo + fo) * (int64_t)C + c0;
#pragma unroll
    for (int c = 0; c < 8; ++c) dst[c] = from_f<OutT>(o[c]);
  }
}

// ---- im2col for a 3x3 stride-2 pad-1 conv on channel-last activations: column = (kt*3 + kf)*C + ci ----
template <typename T>
__global__ void __launch_bounds__(256)
qwen_im2col_kernel(const T* __restrict__ in /*[chunks][Ti][Fi][C]*/, int Ti, int Fi, int To, int Fo, int C, int64_t total_vec,
                   T* __restrict__ col /*[chunks*To*Fo][9*C]*/) {
  constexpr int V = 16 / sizeof(T);
  const int cv = C / V;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    int64_t r = i / cv;
    const int tap = (int)(r % 9); r /= 9;
    const int fo = (int)(r % Fo); r /= Fo;
    const int to = (int)(r % To);
    const int64_t chunk = r / To;
    const int kt = tap / 3, kf = tap - kt * 3;
    const int t = 2 * to - 1 + kt, f = 2 * fo - 1 + kf;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (t >= 0 && t < Ti && f >= 0 && f < Fi)
      v = *reinterpret_cast<const uint4*>(in + ((chunk * Ti + t) * Fi + f) * C + c8 * V);
    *reinterpret_cast<uint4*>(col + i * V) = v;
  }
}

// ---- stem rows [B][chunks*13][d] + pos[row % 13] -> window-padded hidden [B][n_win*tpw][d] (:880-897) ----
//      rows past the clip's own last chunk are zero (ragged batches: exactly what the clip sees when it runs alone)
__global__ void qwen_window_kernel(const float* __restrict__ stem, const float* __restrict__ pos, const int* __restrict__ n_per_clip, int hop,
                                   int n_chunks, int rows_win, int d, float* __restrict__ h) {
  const int r = blockIdx.x, b = blockIdx.y;
  const int valid_rows = n_chunks * kChunkTok;
  const int own_rows = ((n_per_clip[b] / hop + kChunk - 1) / kChunk) * kChunkTok;
  float* dst = h + ((int64_t)b * rows_win + r) * d;
  if (r < own_rows) {
    const float* src = stem + ((int64_t)b * valid_rows + r) * d;
    const float* p = pos + (int64_t)(r % kChunkTok) * d;
    for (int i = threadIdx.x; i < d; i += blockDim.x) dst[i] = src[i] + p[i];
  } else {
    for (int i = threadIdx.x; i < d; i += blockDim.x) dst[i] = 0.f;
  }
}

// ---- softmax(s + key mask) for the CUDA-core attention path: keys >= valid[window] get -128 added (:768-775,904) ----
template <typename OutT>
__global__ void __launch_bounds__(256)
qwen_mask_softmax_kernel(const float* __restrict__ s, OutT* __restrict__ p, int64_t rows, int cols, int rows_per_win,
                         const int* __restrict__ valid, float mask_add) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nv = valid[row / rows_per_win];
  const float* sr = s + row * cols;
  float m = -INFINITY;
  for (int c = lane; c < cols; c += 32) m = fmaxf(m, sr[c] + (c >= nv ? mask_add : 0.f));
  m = warp_max(m);
  float sum = 0.f;
  for (int c = lane; c < cols; c += 32) sum += expf(sr[c] + (c >= nv ? mask_add : 0.f) - m);
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  OutT* pr = p + row * cols;
  for (int c = lane; c < cols; c += 32) pr[c] = from_f<OutT>(expf(sr[c] + (c >= nv ? mask_add : 0.f) - m) * inv);
}

// ---- prompt rows: token embedding or audio row (:925, CONCAT_EMBED :1428-1435) ----
template <typename WT>
__global__ void qwen_prompt_kernel(const int* __restrict__ src /*[B][n_prompt]: id >= 0, or -(audio row) - 1*/, const WT* __restrict__ embed,
                                   const float* __restrict__ enc_out, int64_t enc_stride, int n_prompt, int d, float* __restrict__ x) {
  const int pos = blockIdx.x, b = blockIdx.y;
  const int s = src[(int64_t)b * n_prompt + pos];
  float* dst = x + ((int64_t)b * n_prompt + pos) * d;
  if (s >= 0) {
    const WT* e = embed + (int64_t)s * d;
    for (int i = threadIdx.x; i < d; i += blockDim.x) dst[i] = to_f<WT>(e[i]);
  } else {
    const float* e = enc_out + (int64_t)b * enc_stride + (int64_t)(-s - 1) * d;
    for (int i = threadIdx.x; i < d; i += blockDim.x) dst[i] = e[i];
  }
}

template <typename WT>
__global__ void qwen_embed_kernel(const int* __restrict__ tokens, const WT* __restrict__ embed, int d, int vocab, float* __restrict__ x) {
  const int b = blockIdx.x;
  int t = tokens[b];
  t = t < 0 ? 0 : (t >= vocab ? vocab - 1 : t);
  const WT* e = embed + (int64_t)t * d;
  for (int i = threadIdx.x; i < d; i += blockDim.x) x[(int64_t)b * d + i] = to_f<WT>(e[i]);
}

// ---- RMS norm: y = x * rsqrt(mean(x^2) + eps) [* gamma]; source row = r * row_mul + row_off (:1043-1077,1331) ----
template <typename OutT>
__global__ void __launch_bounds__(256)
qwen_rmsnorm_kernel(const float* __restrict__ x, int64_t ldx, int row_mul, int row_off, const float* __restrict__ gamma, float eps,
                    OutT* __restrict__ out, int64_t ldo, int rows, int d, const int* __restrict__ row_off_v = nullptr) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + ((int64_t)r * row_mul + row_off + (row_off_v ? row_off_v[r] : 0)) * ldx;      // (ragged prefill: a clip's own last prompt row)
  float q = 0.f;
  for (int i = lane; i < d; i += 32) q += xr[i] * xr[i];
  const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
  for (int i = lane; i < d; i += 32) {
    float v = xr[i] * rstd;
    if (gamma) v *= gamma[i];
    out[(int64_t)r * ldo + i] = from_f<OutT>(v);
  }
}

// ---- per-head QK RMS norm (d^-0.25 folded into g) + RoPE; q -> fp32 rows, k / v -> the resident cache (:1281-1312) ----
template <typename KT, int DH>
__global__ void __launch_bounds__(128)
qwen_qk_rope_kernel(const float* __restrict__ qkv /*[rows][(H+2KH)*DH]*/, const float* __restrict__ g /*[2][DH]*/,
                    const float* __restrict__ cosT, const float* __restrict__ sinT /*[max_seq][DH/2]*/, float eps, int n_new, int H,
                    int KH, int max_seq, int64_t cache_layer_off, int total, const DecState* __restrict__ state,
                    float* __restrict__ q /*[rows][H*DH]*/, KT* __restrict__ kc, KT* __restrict__ vc) {
  constexpr int M = DH / 32, half = DH / 2;
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= total) return;
  const int lane = threadIdx.x & 31;
  const int NHD = H + 2 * KH;
  const int row = item / NHD, head = item - row * NHD;
  const int b = row / n_new, i = row - b * n_new;
  const int pos = state->kv_len + i;
  const float* src = qkv + ((int64_t)row * NHD + head) * DH;
  float v[M];
#pragma unroll
  for (int m = 0; m < M; ++m) v[m] = src[lane + 32 * m];
  if (head < H + KH) {
    float ss = 0.f;
#pragma unroll
    for (int m = 0; m < M; ++m) ss += v[m] * v[m];
    const float r = rsqrtf(warp_sum(ss) / (float)DH + eps);
    const float* gg = g + (head < H ? 0 : DH);
#pragma unroll
    for (int m = 0; m < M; ++m) v[m] *= r * gg[lane + 32 * m];
#pragma unroll
    for (int m = 0; m < M / 2; ++m) {
      const int j = lane + 32 * m;
      const float c = cosT[(int64_t)pos * half + j], s = sinT[(int64_t)pos * half + j];
      const float a = v[m], bb = v[m + M / 2];
      v[m] = a * c - bb * s;
      v[m + M / 2] = bb * c + a * s;
    }
  }
  if (head < H) {
    float* dst = q + ((int64_t)row * H + head) * DH;
#pragma unroll
    for (int m = 0; m < M; ++m) dst[lane + 32 * m] = v[m];
  } else {
    const int kh = head < H + KH ? head - H : head - H - KH;
    KT* dst = (head < H + KH ? kc : vc) + cache_layer_off + (((int64_t)b * KH + kh) * max_seq + pos) * DH;
#pragma unroll
    for (int m = 0; m < M; ++m) dst[lane + 32 * m] = from_f<KT>(v[m]);
  }
}

template <typename KT> struct KVec;
template <> struct KVec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float* o) { const float4 u = *reinterpret_cast<const float4*>(p); o[0] = u.x; o[1] = u.y; o[2] = u.z; o[3] = u.w; }
};
template <> struct KVec<bf16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const bf16* p, float* o) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
  }
};

// ---- causal grouped-query attention over the cache: one warp per (row, query head).  Keys after the row's own
//      position are left out, which equals the reference's additive -128 (:959-963) whenever exp(-128 + delta)
//      underflows in fp32, i.e. unless a masked score exceeds the row maximum by more than ~25 ----
template <typename KT, typename OutT, int DH>
__global__ void __launch_bounds__(128)
qwen_attn_kernel(const float* __restrict__ q, const KT* __restrict__ kc, const KT* __restrict__ vc, int64_t cache_layer_off, int n_new,
                 int H, int KH, int max_seq, int total, const DecState* __restrict__ state, OutT* __restrict__ ctx) {
  extern __shared__ float qsm[];                 // per warp: q[DH] + scores[max_seq]
  constexpr int M = DH / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * (blockDim.x >> 5) + warp;
  if (item >= total) return;
  const int row = item / H, h = item - row * H;
  const int b = row / n_new, i = row - b * n_new;
  const int n_keys = state->kv_len + i + 1;
  const int kh = h / (H / KH);
  float* qs = qsm + warp * (DH + max_seq);
  float* sc = qs + DH;
  const float* qr = q + ((int64_t)row * H + h) * DH;
#pragma unroll
  for (int m = 0; m < M; ++m) qs[lane + 32 * m] = qr[lane + 32 * m];
  __syncwarp();
  const KT* K = kc + cache_layer_off + ((int64_t)b * KH + kh) * max_seq * DH;
  const KT* V = vc + cache_layer_off + ((int64_t)b * KH + kh) * max_seq * DH;
  float mx = -INFINITY;
  for (int j = lane; j < n_keys; j += 32) {
    const KT* kr = K + (int64_t)j * DH;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < DH; c += KVec<KT>::N) {
      float kk[KVec<KT>::N];
      KVec<KT>::load(kr + c, kk);
#pragma unroll
      for (int e = 0; e < KVec<KT>::N; ++e) s = fmaf(kk[e], qs[c + e], s);
    }
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < n_keys; j += 32) { const float p = expf(sc[j] - mx); sc[j] = p; sum += p; }
  sum = warp_sum(sum);
  __syncwarp();
  float acc[M];
#pragma unroll
  for (int m = 0; m < M; ++m) acc[m] = 0.f;
  for (int j = 0; j < n_keys; ++j) {
    const float p = sc[j];
    const KT* vr = V + (int64_t)j * DH;
#pragma unroll
    for (int m = 0; m < M; ++m) acc[m] = fmaf(p, to_f<KT>(vr[lane + 32 * m]), acc[m]);
  }
  const float inv = 1.0f / sum;
  OutT* dst = ctx + ((int64_t)row * H + h) * DH;
#pragma unroll
  for (int m = 0; m < M; ++m) dst[lane + 32 * m] = from_f<OutT>(acc[m] * inv);
}

// ---- prefill attention, tiled: one CTA per (utterance, query head, 8 consecutive rows).  K / V tiles of 32 positions are
//      staged once in shared memory (rows padded by 16 bytes: a lane reads its own key row conflict-free) and shared by the
//      8 warps, each of which owns one query row with an online soft-max; lanes hold the keys for the scores and the head
//      dims for P V.  Causal: a row stops at its own position (see qwen_attn_kernel on the -128 mask) ----
constexpr int kPfRows = 8, kPfKeys = 32;
template <typename KT, typename OutT, int DH>
__global__ void __launch_bounds__(kPfRows * 32)
qwen_attn_prefill_kernel(const float* __restrict__ q, const KT* __restrict__ kc, const KT* __restrict__ vc, int64_t cache_layer_off,
                         int n_new, int H, int KH, int max_seq, const DecState* __restrict__ state, OutT* __restrict__ ctx) {
  constexpr int EPL = DH / 32, LDK = DH + 16 / (int)sizeof(KT);
  __shared__ __align__(16) KT sk[kPfKeys * LDK];
  __shared__ __align__(16) KT sv[kPfKeys * DH];
  __shared__ float sq[kPfRows][DH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int i = tile * kPfRows + warp;                 // this warp's row within the utterance's new positions
  const bool live = i < n_new;
  const int base = state->kv_len;
  const int row = b * n_new + i;
  const int kh = h / (H / KH);
  const int my_keys = base + i + 1;                    // causal extent of this row
  const int last_row = min(n_new, (tile + 1) * kPfRows) - 1;
  const int cta_keys = base + last_row + 1;
  if (live) {
    const float* qr = q + ((int64_t)row * H + h) * DH;
#pragma unroll
    for (int m = 0; m < EPL; ++m) sq[warp][lane + 32 * m] = qr[lane + 32 * m];
  }
  const KT* K = kc + cache_layer_off + ((int64_t)b * KH + kh) * max_seq * DH;
  const KT* V = vc + cache_layer_off + ((int64_t)b * KH + kh) * max_seq * DH;
  float mx = -INFINITY, sum = 0.f, acc[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) acc[e] = 0.f;
  constexpr int VPR = DH * (int)sizeof(KT) / 16;        // 16-byte vectors per cache row
  for (int j0 = 0; j0 < cta_keys; j0 += kPfKeys) {
    __syncthreads();
    for (int v = threadIdx.x; v < kPfKeys * VPR; v += kPfRows * 32) {
      const int r = v / VPR, c = v - r * VPR;
      uint4 kk = make_uint4(0u, 0u, 0u, 0u), vv = kk;
      if (j0 + r < cta_keys) {
        kk = reinterpret_cast<const uint4*>(K + (int64_t)(j0 + r) * DH)[c];
        vv = reinterpret_cast<const uint4*>(V + (int64_t)(j0 + r) * DH)[c];
      }
      *reinterpret_cast<uint4*>(reinterpret_cast<char*>(sk) + (size_t)r * LDK * sizeof(KT) + c * 16) = kk;
      *reinterpret_cast<uint4*>(reinterpret_cast<char*>(sv) + (size_t)r * DH * sizeof(KT) + c * 16) = vv;
    }
    __syncthreads();
    if (!live || j0 >= my_keys) continue;
    const int j = j0 + lane;
    float s = -INFINITY;
    if (j < my_keys) {
      const KT* kr = sk + lane * LDK;
      s = 0.f;
#pragma unroll
      for (int c = 0; c < DH; c += KVec<KT>::N) {
        float kk[KVec<KT>::N];
        KVec<KT>::load(kr + c, kk);
#pragma unroll
        for (int e = 0; e < KVec<KT>::N; ++e) s = fmaf(kk[e], sq[warp][c + e], s);
      }
    }
    const float m_new = fmaxf(mx, warp_max(s));
    const float scale = expf(mx - m_new);              // 0 on the first tile (mx = -inf)
    const float p = j < my_keys ? expf(s - m_new) : 0.f;
    sum = sum * scale + warp_sum(p);
#pragma unroll
    for (int e = 0; e < EPL; ++e) acc[e] *= scale;
    mx = m_new;
    const int n_here = min(kPfKeys, my_keys - j0);
    for (int t = 0; t < n_here; ++t) {
      const float pt = __shfl_sync(0xffffffffu, p, t);
      const KT* vr = sv + t * DH + lane * EPL;
#pragma unroll
      for (int e = 0; e < EPL; ++e) acc[e] = fmaf(pt, to_f<KT>(vr[e]), acc[e]);
    }
  }
  if (live) {
    const float inv = 1.0f / sum;
    OutT* dst = ctx + ((int64_t)row * H + h) * DH + lane * EPL;
#pragma unroll
    for (int e = 0; e < EPL; ++e) dst[e] = from_f<OutT>(acc[e] * inv);
  }
}

// ---- decode-step attention, fused with the QK-norm / RoPE / cache append of the new position: one CTA per (utterance,
//      query head).  Warps 0-2 normalise and rotate this head's q and its kv head's new k / v (every query head of a
//      group writes the same cache row, a benign duplicate); then the keys are spread over all 512 threads for the scores
//      and over the 16 warps for P V (lanes own head dims, unrolled so several cache rows are in flight at once) ----
constexpr int kAttDecThreads = 512;
template <typename KT, int DH>
__global__ void __launch_bounds__(kAttDecThreads)
qwen_attn_decode_kernel(const float* __restrict__ qkv /*[B][(H+2KH)*DH]*/, const float* __restrict__ g, const float* __restrict__ cosT,
                        const float* __restrict__ sinT, float eps, KT* __restrict__ kc, KT* __restrict__ vc, int64_t cache_layer_off,
                        int H, int KH, int max_seq, const DecState* __restrict__ state, const int* __restrict__ kv_off, float* __restrict__ ctx) {
  extern __shared__ float dsm[];                 // q[DH] | k_new[DH] | v_new[DH] | scores[max_seq] | part[16][DH]
  constexpr int NW = kAttDecThreads / 32;
  __shared__ float red[NW];
  constexpr int M = DH / 32, half = DH / 2, EPL = DH / 32;
  float* qs = dsm;
  float* kn = dsm + DH;
  float* vn = dsm + 2 * DH;
  float* sc = dsm + 3 * DH;
  float* part = sc + max_seq;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  // kv_off[b] <= 0: this clip's prompt is that much shorter than the longest.  A clip that reached its own generation_limit keeps
  // stepping until the batch's last clip has (its tokens are no longer accepted): its position stays inside its own cache rows
  const int pos = min(state->kv_len + kv_off[b], max_seq - 1), n_keys = pos + 1;
  const int kh = h / (H / KH);
  const int NHD = H + 2 * KH;
  KT* K = kc + cache_layer_off + ((int64_t)b * KH + kh) * max_seq * DH;
  KT* V = vc + cache_layer_off + ((int64_t)b * KH + kh) * max_seq * DH;
  pdl_launch_dependents();
  pdl_wait();                                    // qkv rows come from the previous kernel
  if (warp < 3) {
    const int head = warp == 0 ? h : (warp == 1 ? H + kh : H + KH + kh);
    const float* src = qkv + ((int64_t)b * NHD + head) * DH;
    float v[M];
#pragma unroll
    for (int m = 0; m < M; ++m) v[m] = src[lane + 32 * m];
    if (warp < 2) {
      float ss = 0.f;
#pragma unroll
      for (int m = 0; m < M; ++m) ss += v[m] * v[m];
      const float r = rsqrtf(warp_sum(ss) / (float)DH + eps);
      const float* gg = g + (warp == 0 ? 0 : DH);
#pragma unroll
      for (int m = 0; m < M; ++m) v[m] *= r * gg[lane + 32 * m];
#pragma unroll
      for (int m = 0; m < M / 2; ++m) {
        const int j = lane + 32 * m;
        const float c = cosT[(int64_t)pos * half + j], sn = sinT[(int64_t)pos * half + j];
        const float a = v[m], bb = v[m + M / 2];
        v[m] = a * c - bb * sn;
        v[m + M / 2] = bb * c + a * sn;
      }
    }
    if (warp == 0) {
#pragma unroll
      for (int m = 0; m < M; ++m) qs[lane + 32 * m] = v[m];
    } else {
      KT* dst = (warp == 1 ? K : V) + (int64_t)pos * DH;
      float* keep = warp == 1 ? kn : vn;
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const KT rv = from_f<KT>(v[m]);
        dst[lane + 32 * m] = rv;
        keep[lane + 32 * m] = to_f<KT>(rv);          // what later steps will read back from the cache
      }
    }
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < n_keys; j += kAttDecThreads) {
    float s = 0.f;
    if (j == pos) {
#pragma unroll 8
      for (int c = 0; c < DH; ++c) s = fmaf(kn[c], qs[c], s);
    } else {
      const KT* kr = K + (int64_t)j * DH;
#pragma unroll
      for (int c = 0; c < DH; c += KVec<KT>::N) {
        float kk[KVec<KT>::N];
        KVec<KT>::load(kr + c, kk);
#pragma unroll
        for (int e = 0; e < KVec<KT>::N; ++e) s = fmaf(kk[e], qs[c + e], s);
      }
    }
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int j = threadIdx.x; j < n_keys; j += kAttDecThreads) { const float p = expf(sc[j] - mx); sc[j] = p; sum += p; }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < NW; ++w) sum += red[w];
  float acc[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) acc[e] = 0.f;
#pragma unroll 4
  for (int j = warp; j < pos; j += NW) {
    const float p = sc[j];
    const KT* vr = V + (int64_t)j * DH + lane * EPL;
#pragma unroll
    for (int e = 0; e < EPL; ++e) acc[e] = fmaf(p, to_f<KT>(vr[e]), acc[e]);
  }
  if (warp == pos % NW) {
    const float p = sc[pos];
#pragma unroll
    for (int e = 0; e < EPL; ++e) acc[e] = fmaf(p, vn[lane * EPL + e], acc[e]);
  }
#pragma unroll
  for (int e = 0; e < EPL; ++e) part[warp * DH + lane * EPL + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < DH) {
    float o = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) o += part[w * DH + threadIdx.x];
    ctx[((int64_t)b * H + h) * DH + threadIdx.x] = o / sum;
  }
}

// ---- decode-step attention for the bf16 cache, split over the keys: grid (utterance x query head, S key ranges of at
//      most 128 keys).  Every cache row a CTA needs is requested into registers at kernel entry -- K: one key per thread,
//      V: keys strided over the 8 warps with lanes on head dims -- while warps 0-2 normalise / rotate q and the new k, v,
//      so the whole step costs one memory round trip; the CTA that finishes last (ticket counter) merges the S partial
//      (max, sum, unnormalised output) triples into the context row ----
constexpr int kSplitKeys = 128;
// SK = keys per range (<= 256 = one per thread); VPF = V rows requested into registers up front (SK 128) or walked in an
// unrolled loop (SK 256: batches of 3-4 clips, where 4 ranges per head keep the grid inside one resident wave)
template <int DH, int SK, bool VPF>
__global__ void __launch_bounds__(256, VPF ? 1 : 2)
qwen_attn_split_kernel(const float* __restrict__ qkv, const float* __restrict__ g, const float* __restrict__ cosT,
                       const float* __restrict__ sinT, float eps, bf16* __restrict__ kc, bf16* __restrict__ vc, int64_t cache_layer_off,
                       int H, int KH, int max_seq, const DecState* __restrict__ state, const int* __restrict__ kv_off,
                       float* __restrict__ part /*[B*H][S][DH+2]*/,
                       int* __restrict__ counter /*[B*H]*/, float* __restrict__ ctx) {
  constexpr int M = DH / 32, half = DH / 2, EPL = DH / 32, VK = SK / 8;
  __shared__ float qs[DH], kn[DH], vn[DH], sc[SK], red[8], pw[8][DH];
  __shared__ int s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x, b = bh / H, h = bh - b * H;
  const int S = gridDim.y, sp = blockIdx.y;
  const int pos = min(state->kv_len + kv_off[b], max_seq - 1), n_keys = pos + 1;      // (clamp: see qwen_attn_decode_kernel)
  const int per = (n_keys + S - 1) / S;
  const int lo = sp * per, hi = min(n_keys, lo + per);
  const int kh = h / (H / KH);
  const int NHD = H + 2 * KH;
  bf16* K = kc + cache_layer_off + ((int64_t)b * KH + kh) * max_seq * DH;
  bf16* V = vc + cache_layer_off + ((int64_t)b * KH + kh) * max_seq * DH;
  // ---- every cache load of this CTA, up front ----
  uint4 kreg[DH / 8];
  const int jk = lo + threadIdx.x;
  const bool k_cached = threadIdx.x < SK && jk < hi && jk != pos;
  if (k_cached) {
    const uint4* kr = reinterpret_cast<const uint4*>(K + (int64_t)jk * DH);
#pragma unroll
    for (int c = 0; c < DH / 8; ++c) kreg[c] = kr[c];
  }
  uint32_t vreg[VPF ? VK : 1][EPL / 2];
  if (VPF) {
#pragma unroll
    for (int i = 0; i < VK; ++i) {
      const int j = lo + warp + 8 * i;
      if (j < hi && j != pos) {
        const uint32_t* vr = reinterpret_cast<const uint32_t*>(V + (int64_t)j * DH + lane * EPL);
#pragma unroll
        for (int e = 0; e < EPL / 2; ++e) vreg[VPF ? i : 0][e] = vr[e];
      }
    }
  }
  pdl_launch_dependents();
  pdl_wait();                                    // qkv rows come from the previous kernel
  if (warp < 3) {
    const int head = warp == 0 ? h : (warp == 1 ? H + kh : H + KH + kh);
    const float* src = qkv + ((int64_t)b * NHD + head) * DH;
    float v[M];
#pragma unroll
    for (int m = 0; m < M; ++m) v[m] = src[lane + 32 * m];
    if (warp < 2) {
      float ss = 0.f;
#pragma unroll
      for (int m = 0; m < M; ++m) ss += v[m] * v[m];
      const float r = rsqrtf(warp_sum(ss) / (float)DH + eps);
      const float* gg = g + (warp == 0 ? 0 : DH);
#pragma unroll
      for (int m = 0; m < M; ++m) v[m] *= r * gg[lane + 32 * m];
#pragma unroll
      for (int m = 0; m < M / 2; ++m) {
        const int j = lane + 32 * m;
        const float c = cosT[(int64_t)pos * half + j], sn = sinT[(int64_t)pos * half + j];
        const float a = v[m], bb = v[m + M / 2];
        v[m] = a * c - bb * sn;
        v[m + M / 2] = bb * c + a * sn;
      }
    }
    if (warp == 0) {
#pragma unroll
      for (int m = 0; m < M; ++m) qs[lane + 32 * m] = v[m];
    } else {
      bf16* dst = (warp == 1 ? K : V) + (int64_t)pos * DH;
      float* keep = warp == 1 ? kn : vn;
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const bf16 rv = __float2bfloat16_rn(v[m]);
        if (sp == 0) dst[lane + 32 * m] = rv;          // one writer per (kv head, query head); duplicates across the group are identical
        keep[lane + 32 * m] = __bfloat162float(rv);     // what later steps will read back from the cache
      }
    }
  }
  __syncthreads();
  float s = -INFINITY;
  if (threadIdx.x < SK && jk < hi) {
    s = 0.f;
    if (jk == pos) {
#pragma unroll 8
      for (int c = 0; c < DH; ++c) s = fmaf(kn[c], qs[c], s);
    } else {
#pragma unroll
      for (int c = 0; c < DH / 8; ++c) {
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&kreg[c]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(h2[e]);
          s = fmaf(f.x, qs[c * 8 + 2 * e], s);
          s = fmaf(f.y, qs[c * 8 + 2 * e + 1], s);
        }
      }
    }
  }
  float mx = warp_max(s);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float p = 0.f;
  if (threadIdx.x < SK && jk < hi) { p = expf(s - mx); sc[threadIdx.x] = p; }
  float sum = warp_sum(p);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += red[w];
  float acc[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) acc[e] = 0.f;
  auto pv_step = [&](int j, const uint32_t* vv) {
    const float pj = sc[j - lo];
    if (j == pos) {
#pragma unroll
      for (int e = 0; e < EPL; ++e) acc[e] = fmaf(pj, vn[lane * EPL + e], acc[e]);
    } else {
#pragma unroll
      for (int e = 0; e < EPL / 2; ++e) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&vv[e]));
        acc[2 * e] = fmaf(pj, f.x, acc[2 * e]);
        acc[2 * e + 1] = fmaf(pj, f.y, acc[2 * e + 1]);
      }
    }
  };
  if constexpr (VPF) {
#pragma unroll
    for (int i = 0; i < VK; ++i) {
      const int j = lo + warp + 8 * i;
      if (j < hi) pv_step(j, vreg[i]);
    }
  } else {
#pragma unroll 8
    for (int i = 0; i < VK; ++i) {
      const int j = lo + warp + 8 * i;
      if (j < hi) {
        uint32_t vv[EPL / 2];
        if (j != pos) {
          const uint32_t* vr = reinterpret_cast<const uint32_t*>(V + (int64_t)j * DH + lane * EPL);
#pragma unroll
          for (int e = 0; e < EPL / 2; ++e) vv[e] = vr[e];
        }
        pv_step(j, vv);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < EPL; ++e) pw[warp][lane * EPL + e] = acc[e];
  __syncthreads();
  float* mine = part + ((int64_t)bh * S + sp) * (DH + 2);
  if (threadIdx.x < DH) {
    float o = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) o += pw[w][threadIdx.x];
    mine[threadIdx.x] = o;
  }
  if (threadIdx.x == 0) { mine[DH] = mx; mine[DH + 1] = sum; }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&counter[bh], 1) == S - 1) ? 1 : 0;
  __syncthreads();
  if (s_last) {
    __threadfence();
    if (threadIdx.x < DH) {
      const float* pb = part + (int64_t)bh * S * (DH + 2);
      float Mx = -INFINITY;
      for (int q = 0; q < S; ++q) Mx = fmaxf(Mx, __ldcg(pb + q * (DH + 2) + DH));
      float num = 0.f, den = 0.f;
      for (int q = 0; q < S; ++q) {
        const float wq = expf(__ldcg(pb + q * (DH + 2) + DH) - Mx);
        num = fmaf(wq, __ldcg(pb + q * (DH + 2) + threadIdx.x), num);
        den = fmaf(wq, __ldcg(pb + q * (DH + 2) + DH + 1), den);
      }
      ctx[(int64_t)bh * DH + threadIdx.x] = num / den;
    }
    if (threadIdx.x == 0) counter[bh] = 0;             // ready for the next layer / graph replay
  }
}

// ---- weight-streaming GEMV for a decode step: out[r][n] = (rms? rstd[r] : 1) * sum_k W[n][k] * x'[r][k] (+ residual), with
//      x' = x or silu(x[:K]) * x[K:2K] (SwiGLU of the fused gate_up rows).  A layer is a few MB, so the kernel is built to
//      have the whole matrix in flight at once: a warp owns NC output columns and 1/KS of their K range (at most four
//      256-wide chunks), i.e. up to 8 outstanding 16-byte loads per thread, requested as raw registers BEFORE the
//      activations are staged; the RMS statistics ride along with the staging pass and scale the finished dot products.
//      Wide layers (the vocabulary head) loop over column passes with the next pass's loads issued ahead of the math ----
struct QGemvArgs {
  const float* x; int64_t ldx; int rms; float eps; int swiglu;
  const void* W; const float* residual; int64_t ldr; float* out; int64_t ldo;
  int rows, N, K;
};
constexpr int kGemvRows = 4;
constexpr int kGemvCH = 4;                       // 256-wide k chunks per warp (K <= KS * 1024)
template <typename WT> struct WRaw;
template <> struct WRaw<bf16> {
  uint4 v;
  __device__ __forceinline__ void load(const bf16* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void get(float* o) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
  }
};
template <> struct WRaw<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) { a = *reinterpret_cast<const float4*>(p); b = *reinterpret_cast<const float4*>(p + 4); }
  __device__ __forceinline__ void get(float* o) const { o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w; }
};

// LOOP = false: the grid covers every column pass (one resident wave at 3 CTAs per SM), no second register buffer
template <typename WT, int KS, int NC, bool LOOP>
__global__ void __launch_bounds__(256, LOOP ? 2 : 3)
qwen_gemv_kernel(QGemvArgs a) {
  extern __shared__ float gx[];                  // [kGemvRows][K]
  __shared__ float red[8][kGemvRows];
  __shared__ float psum[8][NC][kGemvRows];
  constexpr int CPB = (8 / KS) * NC;             // output columns per CTA pass
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ks = warp % KS, cw = warp / KS;
  const int K = a.K, N = a.N;
  const WT* W = reinterpret_cast<const WT*>(a.W);
  const int k0 = (ks * 32 + lane) * 8, kstep = KS * 256;
  const int n_pass = (N + CPB - 1) / CPB;
  WRaw<WT> cur[NC][kGemvCH], nxt[LOOP ? NC : 1][LOOP ? kGemvCH : 1];
  auto fetch = [&](auto& buf, int pass) {
    const int c0 = pass * CPB + cw * NC;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      if (pass < n_pass && c0 + j < N) {
        const WT* wr = W + (int64_t)(c0 + j) * K;
#pragma unroll
        for (int i = 0; i < kGemvCH; ++i) {
          const int k = k0 + i * kstep;
          if (k < K) buf[j][i].load(wr + k);
        }
      }
    }
  };
  fetch(cur, blockIdx.x);
  pdl_launch_dependents();
  pdl_wait();                                    // activations / residual come from the previous kernel
  for (int r0 = 0; r0 < a.rows; r0 += kGemvRows) {
    const int nr = min(kGemvRows, a.rows - r0);
    __syncthreads();
    float ss[kGemvRows];
#pragma unroll
    for (int r = 0; r < kGemvRows; ++r) {
      ss[r] = 0.f;
      if (r < nr) {
        const float* xr = a.x + (int64_t)(r0 + r) * a.ldx;
        for (int k = threadIdx.x * 4; k < K; k += 1024) {          // K % 8 == 0 and 16-byte aligned rows (checked by the launcher)
          float4 v = *reinterpret_cast<const float4*>(xr + k);
          if (a.swiglu) {
            const float4 u = *reinterpret_cast<const float4*>(xr + K + k);
            v.x = v.x / (1.0f + expf(-v.x)) * u.x; v.y = v.y / (1.0f + expf(-v.y)) * u.y;
            v.z = v.z / (1.0f + expf(-v.z)) * u.z; v.w = v.w / (1.0f + expf(-v.w)) * u.w;
          }
          *reinterpret_cast<float4*>(gx + r * K + k) = v;
          ss[r] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
      }
    }
    if (a.rms) {
#pragma unroll
      for (int r = 0; r < kGemvRows; ++r) { const float t = warp_sum(ss[r]); if (lane == 0) red[warp][r] = t; }
    }
    __syncthreads();
    float rstd[kGemvRows];
#pragma unroll
    for (int r = 0; r < kGemvRows; ++r) {
      rstd[r] = 1.f;
      if (a.rms) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][r];
        rstd[r] = rsqrtf(t / (float)K + a.eps);
      }
    }
    if (r0 > 0) fetch(cur, blockIdx.x);                              // batches beyond four rows walk the matrix again (L2)
    for (int pass = blockIdx.x; pass < n_pass; pass += gridDim.x) {
      if (LOOP) fetch(nxt, pass + gridDim.x);                         // next pass's loads are in flight during this pass's math
      float acc[NC][kGemvRows];
#pragma unroll
      for (int j = 0; j < NC; ++j)
#pragma unroll
        for (int r = 0; r < kGemvRows; ++r) acc[j][r] = 0.f;
      const int c0 = pass * CPB + cw * NC;
#pragma unroll
      for (int i = 0; i < kGemvCH; ++i) {
        const int k = k0 + i * kstep;
        if (k < K) {
          float w[NC][8];
#pragma unroll
          for (int j = 0; j < NC; ++j) cur[j][i].get(w[j]);
#pragma unroll
          for (int r = 0; r < kGemvRows; ++r) {
            if (r < nr) {
              const float4 x0 = *reinterpret_cast<const float4*>(gx + r * K + k);
              const float4 x1 = *reinterpret_cast<const float4*>(gx + r * K + k + 4);
#pragma unroll
              for (int j = 0; j < NC; ++j) {
                float t = acc[j][r];
                t = fmaf(w[j][0], x0.x, t); t = fmaf(w[j][1], x0.y, t); t = fmaf(w[j][2], x0.z, t); t = fmaf(w[j][3], x0.w, t);
                t = fmaf(w[j][4], x1.x, t); t = fmaf(w[j][5], x1.y, t); t = fmaf(w[j][6], x1.z, t); t = fmaf(w[j][7], x1.w, t);
                acc[j][r] = t;
              }
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < NC; ++j)
#pragma unroll
        for (int r = 0; r < kGemvRows; ++r) acc[j][r] = warp_sum(acc[j][r]);
      if (KS > 1) {
        if (lane == 0) {
#pragma unroll
          for (int j = 0; j < NC; ++j)
#pragma unroll
            for (int r = 0; r < kGemvRows; ++r) psum[warp][j][r] = acc[j][r];
        }
        __syncthreads();
      }
      if (ks == 0 && lane < NC * kGemvRows) {
        const int j = lane / kGemvRows, r = lane - j * kGemvRows;
        const int c = c0 + j;
        if (r < nr && c < N) {
          float v = 0.f;
          if (KS > 1) {
#pragma unroll
            for (int q = 0; q < KS; ++q) v += psum[cw * KS + q][j][r];
          } else {
#pragma unroll
            for (int jj = 0; jj < NC; ++jj)
#pragma unroll
              for (int rr = 0; rr < kGemvRows; ++rr) if (jj == j && rr == r) v = acc[jj][rr];
          }
          float rs = 1.f;
#pragma unroll
          for (int rr = 0; rr < kGemvRows; ++rr) if (rr == r) rs = rstd[rr];
          v *= rs;
          const int row = r0 + r;
          if (a.residual) v += a.residual[(int64_t)row * a.ldr + c];
          a.out[(int64_t)row * a.ldo + c] = v;
        }
      }
      if (KS > 1) __syncthreads();
      if (LOOP) {
#pragma unroll
        for (int j = 0; j < NC; ++j)
#pragma unroll
          for (int i = 0; i < kGemvCH; ++i) cur[j][i] = nxt[LOOP ? j : 0][LOOP ? i : 0];
      }
    }
  }
}

#include "qwen_persist.cuh"

// ---- first stage of the vocabulary arg-max: one CTA per (slice, utterance) keeps its slice's best (value, lowest id) ----
constexpr int kArgSlices = 128;
__global__ void __launch_bounds__(256)
qwen_argmax_slices_kernel(const float* __restrict__ logits, int vocab, float* __restrict__ cand_val, int* __restrict__ cand_idx) {
  __shared__ float sv[8];
  __shared__ int si[8];
  const int b = blockIdx.y, sl = blockIdx.x;
  const int per = (vocab + kArgSlices - 1) / kArgSlices;
  const int lo = sl * per, hi = min(vocab, lo + per);
  const float* lg = logits + (int64_t)b * vocab;
  float best = -INFINITY; int besti = 0x7fffffff;
  for (int i = lo + threadIdx.x; i < hi; i += 256) {
    const float v = lg[i];
    if (v > best || (v == best && i < besti)) { best = v; besti = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sv[warp] = best; si[warp] = besti; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) if (sv[w] > best || (sv[w] == best && si[w] < besti)) { best = sv[w]; besti = si[w]; }
    cand_val[b * kArgSlices + sl] = best;
    cand_idx[b * kArgSlices + sl] = besti == 0x7fffffff ? 0 : besti;
  }
}

// ---- SwiGLU: silu(gate) * up on the fused gate_up rows (:1327-1329) ----
template <typename OutT>
__global__ void qwen_swiglu_kernel(const float* __restrict__ gu, int inter, int64_t total, OutT* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / inter; const int c = (int)(i - r * inter);
    const float g = gu[r * 2 * inter + c], u = gu[r * 2 * inter + inter + c];
    out[i] = from_f<OutT>(g / (1.0f + expf(-g)) * u);
  }
}

__global__ void qwen_f32_to_bf16(const float* __restrict__ in, bf16* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}

// ---- penalty-greedy (the script's default strategy): logits of the last `range` selected ids *= value before the
//      arg-max; gather-then-scatter semantics, a repeated id is scaled once (APPLY_PENALTY, Export_Qwen_ASR.py:669-694,1403-1415) ----
__global__ void qwen_penalty_kernel(float* __restrict__ logits, int vocab, const int* __restrict__ save_id, int save_ld,
                                    const int* __restrict__ n_save, int range, float value) {
  const int b = blockIdx.x;
  if (threadIdx.x != 0) return;
  float* lg = logits + (int64_t)b * vocab;
  const int ns = n_save[b];
  const int first = ns - range > 0 ? ns - range : 0;
  for (int j = first; j < ns; ++j) {
    const int id = save_id[(int64_t)b * save_ld + j];
    bool seen = false;
    for (int k = first; k < j; ++k) seen |= (save_id[(int64_t)b * save_ld + k] == id);
    if (!seen && id >= 0 && id < vocab) lg[id] *= value;
  }
}

__global__ void qwen_reset_kernel(DecState* st, int* n_gen, int* finished, int* n_save, int batch) {
  if (threadIdx.x == 0) { st->kv_len = 0; st->step = 0; st->all_done = 0; st->pad = 0; }
  if (threadIdx.x < batch) { n_gen[threadIdx.x] = 0; finished[threadIdx.x] = 0; n_save[threadIdx.x] = 0; }
}

struct QTensor { void* ptr = nullptr; int64_t numel = 0; int dtype = kF32; };
std::string g_qwen_create_error;

int aftercnn_len(int n) {
  if (n >= kChunk) return kChunkTok;
  if (n <= 0) return 0;
  const int a = (n - 1) / 2 + 1, b = (a - 1) / 2 + 1;
  return (b - 1) / 2 + 1;
}

}  // namespace

struct b200asr_qwen {
  b200asr_qwen_config cfg{};
  cudaStream_t st = nullptr;
  std::string err;
  int num_sms = 148;
  int64_t launches = 0;
  bool finalized = false;
  int act = kF32; size_t es = 4;
  std::map<std::string, QTensor> w;
  float* basis_t = nullptr; int* fb_start = nullptr; int* fb_len = nullptr;
  float* stage_buf = nullptr; int64_t stage_cap = 0;
  std::vector<int> head_ids, suffix_ids, tail_ids, stop_ids;
  int* d_stop = nullptr;
  int max_frames = 0, max_chunks = 0, max_win = 0;
  // per call
  //   n_samples / frames / n_chunks / n_win / n_audio / n_prompt describe the LONGEST clip of the batch (= every clip of a uniform
  //   batch); clip_ns / clip_prompt hold each clip's own values, and the device reads them from d_ns / kv_off / d_limit
  int B = 0, n_samples = 0, frames = 0, n_chunks = 0, n_win = 0, n_audio = 0, n_prompt = 0, pcm_dtype = B200ASR_PCM_I16;
  std::vector<int> clip_ns, clip_prompt;
  bool prefilled = false;
  int limit = 0;
  // encoder buffers
  void* pcm = nullptr; float* mel_raw = nullptr; int* max_key = nullptr; float* feat = nullptr;
  void *c1 = nullptr, *col = nullptr, *c2 = nullptr, *c3 = nullptr;
  float *stem = nullptr, *h = nullptr, *S = nullptr, *enc_out = nullptr;
  void *xhat = nullptr, *qkv = nullptr, *ctx = nullptr, *ffn = nullptr, *P = nullptr;
  int *win_valid = nullptr, *prompt_src = nullptr;
  int *d_ns = nullptr;        // [max_batch] samples per clip (front end: reflect pad and arg-max at the clip's own end)
  int *kv_off = nullptr;      // [max_batch] clip prompt length - longest prompt length (<= 0): cache position = DecState.kv_len + kv_off[b]
  int *d_limit = nullptr;     // [max_batch] generation limit of each clip (max_seq_len - 10 - its prompt length)
  // decoder buffers
  float *x = nullptr, *qkvf = nullptr, *q = nullptr, *gu = nullptr, *xl = nullptr, *logits = nullptr, *cand_val = nullptr;
  int* cand_idx = nullptr;
  void *xn = nullptr, *actx = nullptr, *mlp = nullptr, *kc = nullptr, *vc = nullptr;
  DecState* dstate = nullptr;
  int *cur_token = nullptr, *tokens = nullptr, *n_gen = nullptr, *finished = nullptr, *save_id = nullptr, *n_save = nullptr;
  cudaGraphExec_t step_graph = nullptr; int graph_B = -1, graph_limit = -1; int64_t graph_nodes = 0; bool use_graph = true;
  bool use_attn_tc = true, use_attn_split = true, use_pdl = true, use_attn_tiled = true;
  float repeat_penalty = 1.0f; int penalty_range = 10;
  float samp_temperature = 0.f, samp_top_p = 1.f, samp_rep = 1.f; int samp_top_k = 0; unsigned long long samp_seed = 0;
  float* samp_noise = nullptr; int samp_noise_rows = 0;
  bool pdl_all = false;
  float* att_part = nullptr; int* att_counter = nullptr;
  int* h_pinned = nullptr;
  // persistent decode-layer kernel (qwen_persist.cuh): layer table, barrier counter and its host-side running total
  QPersistLayer* p_layers = nullptr; unsigned* p_bar = nullptr; unsigned p_bar_count = 0;
  bool use_persist = false;                        // off by default: measured slower than the PDL graph
  int persist_grid = 0, persist_per_sm = 2, persist_dbg = 0;
  unsigned long long* p_timing = nullptr;          // 512 stamps of the last persistent launch ("persist_timing" option)
  //   persist_grid: 0 = not probed yet, -1 = unsupported here, else CTAs of the cooperative launch

  int fail(int code, const std::string& m) { err = m; return code; }
  int cuda_fail(cudaError_t e, const char* what) { err = std::string(what) + ": " + cudaGetErrorString(e); return B200ASR_E_CUDA; }
};

#define QCK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return e->cuda_fail(_e, #expr); } while (0)
#define QKL(expr) do { cudaError_t _e = (expr); e->launches++; if (_e != cudaSuccess) return e->cuda_fail(_e, #expr); } while (0)
#define QRET(expr) do { int _r = (expr); if (_r != B200ASR_OK) return _r; } while (0)

namespace {

bool qwen_is_matrix(const std::string& n) {
  if (n == "conv1.w") return false;                      // 9-tap CUDA-core conv reads fp32 weights
  return n.size() > 2 && n.compare(n.size() - 2, 2, ".w") == 0;
}
const void* QW(b200asr_qwen* e, const std::string& n) { return e->w[n].ptr; }
const float* QWF(b200asr_qwen* e, const std::string& n) { return reinterpret_cast<const float*>(e->w[n].ptr); }

int qwen_need(b200asr_qwen* e, const std::string& n, int64_t numel, bool at_least = false) {
  auto it = e->w.find(n);
  if (it == e->w.end()) return e->fail(B200ASR_E_MISSING, "missing weight tensor '" + n + "'");
  if (at_least ? it->second.numel < numel : it->second.numel != numel)
    return e->fail(B200ASR_E_INVALID, "tensor '" + n + "' has " + std::to_string(it->second.numel) + " elements, expected " + std::to_string(numel));
  return B200ASR_OK;
}

int qwen_gemm(b200asr_qwen* e, const GemmArgs& g) {
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
  QKL(launch_gemm_simt(g, e->st));
  return B200ASR_OK;
}

GemmArgs qwen_linear(b200asr_qwen* e, const void* A, int64_t lda, const std::string& wn, const std::string& bn, void* C, int64_t ldc,
                     int c_dtype, int M, int N, int K) {
  GemmArgs g;
  g.A = A; g.lda = lda; g.a_dtype = e->act;
  g.B = QW(e, wn); g.ldb = K; g.b_dtype = e->act;
  g.C = C; g.ldc = ldc; g.c_dtype = c_dtype;
  g.bias = bn.empty() ? nullptr : QWF(e, bn);
  g.M = M; g.N = N; g.K = K;
  g.pdl = e->use_pdl ? 1 : 0;          // B is a weight matrix: the tcgen05 GEMM may start under its predecessor's tail
  return g;
}

template <typename T>
int qwen_alloc(b200asr_qwen* e, T** p, size_t bytes) {
  QCK(cudaMalloc(reinterpret_cast<void**>(p), bytes ? bytes : 16));
  QCK(cudaMemsetAsync(*p, 0, bytes ? bytes : 16, e->st));
  return B200ASR_OK;
}

inline unsigned grid_for(int64_t n, int block = 256, int64_t cap = 148 * 16) {
  int64_t g = (n + block - 1) / block;
  if (g > cap) g = cap;
  return (unsigned)(g < 1 ? 1 : g);
}

// ---- audio encoder: PCM resident in e->pcm -> enc_out [B][n_win*tpw][out_dim] (rows >= n_audio are padding) ----
int qwen_encoder(b200asr_qwen* e) {
  const b200asr_qwen_config& c = e->cfg;
  const int B = e->B, C = c.conv_ch, D = c.enc_d, H = c.enc_heads, dh = D / H, ad = e->act;
  const size_t es = e->es;
  const int frames_pad = e->n_chunks * kChunk;
  const int64_t chunks = (int64_t)B * e->n_chunks;
  QKL(launch_fill_i32(e->max_key, INT_MIN, B, e->st));
  QKL(launch_logmel(e->pcm, e->pcm_dtype == B200ASR_PCM_F32, B, e->n_samples, e->n_samples, e->basis_t, QWF(e, "mel_fbank"),
                    e->fb_start, e->fb_len, c.n_fft, c.hop, c.n_mels, e->mel_raw, e->max_key, e->st, e->d_ns));
  qwen_feat_kernel<<<dim3(grid_for((int64_t)frames_pad * c.n_mels, 256, 64), B), 256, 0, e->st>>>(e->mel_raw, e->max_key, e->d_ns, c.hop, e->frames,
                                                                                                 frames_pad, c.n_mels, e->feat);
  QKL(cudaGetLastError());
  // conv stem: time 100 -> 50 -> 25 -> 13, mel n_mels -> /2 -> /4 -> /8
  const int T1 = 50, T2 = 25, T3 = 13;
  const int F1 = (c.n_mels + 1) / 2, F2 = (F1 + 1) / 2, F3 = (F2 + 1) / 2;
  {
    const int64_t total = chunks * T1 * F1 * (C / 8);
    const size_t csm = (size_t)C * 10 * sizeof(float);
    if (ad == kBF16) qwen_conv1_kernel<bf16><<<grid_for(total, 256, 148 * 8), 256, csm, e->st>>>(e->feat, QWF(e, "conv1.w"), QWF(e, "conv1.b"), c.n_mels, C, T1, F1, total, (bf16*)e->c1);
    else qwen_conv1_kernel<float><<<grid_for(total, 256, 148 * 8), 256, csm, e->st>>>(e->feat, QWF(e, "conv1.w"), QWF(e, "conv1.b"), c.n_mels, C, T1, F1, total, (float*)e->c1);
    QKL(cudaGetLastError());
  }
  auto conv = [&](const void* in, int Ti, int Fi, int To, int Fo, const char* wn, const char* bn, void* out) -> int {
    const int64_t rows = chunks * To * Fo;
    const int V = ad == kBF16 ? 8 : 4;
    const int64_t total_vec = rows * 9 * (C / V);
    if (ad == kBF16) qwen_im2col_kernel<bf16><<<grid_for(total_vec, 256, 148 * 32), 256, 0, e->st>>>((const bf16*)in, Ti, Fi, To, Fo, C, total_vec, (bf16*)e->col);
    else qwen_im2col_kernel<float><<<grid_for(total_vec, 256, 148 * 32), 256, 0, e->st>>>((const float*)in, Ti, Fi, To, Fo, C, total_vec, (float*)e->col);
    QKL(cudaGetLastError());
    GemmArgs g = qwen_linear(e, e->col, 9 * C, wn, bn, out, C, ad, (int)rows, C, 9 * C);
    g.act = kActGeluTanh;
    return qwen_gemm(e, g);
  };
  QRET(conv(e->c1, T1, F1, T2, F2, "conv2.w", "conv2.b", e->c2));
  QRET(conv(e->c2, T2, F2, T3, F3, "conv3.w", "conv3.b", e->c3));
  const int rows_stem = (int)(chunks * kChunkTok);
  QRET(qwen_gemm(e, qwen_linear(e, e->c3, (int64_t)F3 * C, "conv_out.w", "", e->stem, D, kF32, rows_stem, D, F3 * C)));
  const int tpw = c.chunks_per_window * kChunkTok;
  const int rows_win = e->n_win * tpw;
  qwen_window_kernel<<<dim3(rows_win, B), 128, 0, e->st>>>(e->stem, QWF(e, "enc_pos"), e->d_ns, c.hop, e->n_chunks, rows_win, D, e->h);
  QKL(cudaGetLastError());
  const int M = B * rows_win, NW = B * e->n_win;
  for (int i = 0; i < c.enc_layers; ++i) {
    const std::string p = "enc" + std::to_string(i) + ".";
    QKL(launch_layernorm(e->h, D, nullptr, nullptr, e->xhat, ad, D, M, D, c.enc_ln_eps, e->st));
    QRET(qwen_gemm(e, qwen_linear(e, e->xhat, D, p + "qkv.w", p + "qkv.b", e->qkv, 3 * D, ad, M, 3 * D, D)));
    if (ad == kBF16 && c.use_tensor_cores && e->use_attn_tc && attention_tc_supported(tpw, D, H)) {
      std::string msg;
      cudaError_t r = launch_attention_tc(e->qkv, e->ctx, NW, tpw, D, H, e->st, &msg, e->win_valid, -128.0f);
      e->launches++;
      if (r != cudaSuccess) return e->fail(B200ASR_E_CUDA, "attention_tc: " + (msg.empty() ? std::string(cudaGetErrorString(r)) : msg));
    } else {
      const int T = tpw;
      GemmArgs s;
      s.A = e->qkv; s.lda = 3 * D; s.sAo = (int64_t)T * 3 * D; s.sAi = dh; s.a_dtype = ad;
      s.B = (char*)e->qkv + (size_t)D * es; s.ldb = 3 * D; s.sBo = (int64_t)T * 3 * D; s.sBi = dh; s.b_dtype = ad;
      s.C = e->S; s.ldc = T; s.sCo = (int64_t)H * T * T; s.sCi = (int64_t)T * T; s.c_dtype = kF32;
      s.M = T; s.N = T; s.K = dh; s.batch = NW * H; s.batch_inner = H;
      QKL(launch_gemm_simt(s, e->st));
      const int64_t rows = (int64_t)NW * H * T;
      if (ad == kBF16) qwen_mask_softmax_kernel<bf16><<<(unsigned)((rows + 7) / 8), 256, 0, e->st>>>(e->S, (bf16*)e->P, rows, T, H * T, e->win_valid, -128.0f);
      else qwen_mask_softmax_kernel<float><<<(unsigned)((rows + 7) / 8), 256, 0, e->st>>>(e->S, (float*)e->P, rows, T, H * T, e->win_valid, -128.0f);
      QKL(cudaGetLastError());
      GemmArgs o;
      o.A = e->P; o.lda = T; o.sAo = (int64_t)H * T * T; o.sAi = (int64_t)T * T; o.a_dtype = ad;
      o.B = (char*)e->qkv + (size_t)2 * D * es; o.ldb = 3 * D; o.sBo = (int64_t)T * 3 * D; o.sBi = dh; o.b_dtype = ad;
      o.transB = 1;
      o.C = e->ctx; o.ldc = D; o.sCo = (int64_t)T * D; o.sCi = dh; o.c_dtype = ad;
      o.M = T; o.N = dh; o.K = T; o.batch = NW * H; o.batch_inner = H;
      QKL(launch_gemm_simt(o, e->st));
    }
    {
      GemmArgs g = qwen_linear(e, e->ctx, D, p + "out.w", p + "out.b", e->h, D, kF32, M, D, D);
      g.residual = e->h; g.ldr = D;
      QRET(qwen_gemm(e, g));
    }
    QKL(launch_layernorm(e->h, D, nullptr, nullptr, e->xhat, ad, D, M, D, c.enc_ln_eps, e->st));
    {
      GemmArgs g = qwen_linear(e, e->xhat, D, p + "fc1.w", p + "fc1.b", e->ffn, c.enc_ffn, ad, M, c.enc_ffn, D);
      g.act = kActGeluTanh;
      QRET(qwen_gemm(e, g));
      GemmArgs g2 = qwen_linear(e, e->ffn, c.enc_ffn, p + "fc2.w", p + "fc2.b", e->h, D, kF32, M, D, c.enc_ffn);
      g2.residual = e->h; g2.ldr = D;
      QRET(qwen_gemm(e, g2));
    }
  }
  QKL(launch_layernorm(e->h, D, nullptr, nullptr, e->xhat, ad, D, M, D, c.enc_ln_eps, e->st));
  {
    GemmArgs g = qwen_linear(e, e->xhat, D, "proj1.w", "proj1.b", e->ffn, D, ad, M, D, D);
    g.act = kActGeluTanh;
    QRET(qwen_gemm(e, g));
    QRET(qwen_gemm(e, qwen_linear(e, e->ffn, D, "proj2.w", "proj2.b", e->enc_out, c.out_dim, kF32, M, c.out_dim, D)));
  }
  return B200ASR_OK;
}

template <int DH>
int qwen_attention_launch(b200asr_qwen* e, int layer, int rows, int n_new, bool gemv) {
  const b200asr_qwen_config& c = e->cfg;
  const int H = c.heads, KH = c.kv_heads, ad = e->act;
  const int64_t layer_off = (int64_t)layer * c.max_batch * KH * c.max_seq_len * DH;
  const float* g = QWF(e, "dec" + std::to_string(layer) + ".qk_norm.g");
  const float *cosT = QWF(e, "rope_cos"), *sinT = QWF(e, "rope_sin");
  if (gemv) {        // one fused launch: QK-norm + RoPE + cache append + attention
    const size_t dsmem = (size_t)(3 * DH + c.max_seq_len + (kAttDecThreads / 32) * DH) * sizeof(float);
    const int S = (c.max_seq_len + kSplitKeys - 1) / kSplitKeys, S2 = (c.max_seq_len + 2 * kSplitKeys - 1) / (2 * kSplitKeys);
    const bool pdl_att = e->use_pdl && (rows <= 2 || e->pdl_all);
    if (ad == kBF16 && e->use_attn_split && rows * H * S <= 2 * e->num_sms) {      // 1-2 clips: (head, 128-key range) CTAs with every cache row in registers (measured better than 256-key ranges at 2 clips: 1.10 vs 1.16 ms/step)
      QKL(launch_pdl(qwen_attn_split_kernel<DH, kSplitKeys, true>, dim3(rows * H, S), dim3(256), 0, e->st, pdl_att, (const float*)e->qkvf, g, cosT, sinT, c.rms_eps, (bf16*)e->kc, (bf16*)e->vc, layer_off, H, KH, c.max_seq_len, (const DecState*)e->dstate, (const int*)e->kv_off, e->att_part, e->att_counter, (float*)e->actx));
      return B200ASR_OK;
    }
    if (ad == kBF16 && e->use_attn_split && rows * H * S2 <= 2 * e->num_sms) {     // 3-4 clips: 256-key ranges keep the grid inside one wave
      QKL(launch_pdl(qwen_attn_split_kernel<DH, 2 * kSplitKeys, false>, dim3(rows * H, S2), dim3(256), 0, e->st, pdl_att, (const float*)e->qkvf, g, cosT, sinT, c.rms_eps, (bf16*)e->kc, (bf16*)e->vc, layer_off, H, KH, c.max_seq_len, (const DecState*)e->dstate, (const int*)e->kv_off, e->att_part, e->att_counter, (float*)e->actx));
      return B200ASR_OK;
    }
    if (ad == kBF16) QKL(launch_pdl(qwen_attn_decode_kernel<bf16, DH>, dim3(rows * H), dim3(kAttDecThreads), dsmem, e->st, e->use_pdl && (rows <= 2 || e->pdl_all), (const float*)e->qkvf, g, cosT, sinT, c.rms_eps, (bf16*)e->kc, (bf16*)e->vc, layer_off, H, KH, c.max_seq_len, (const DecState*)e->dstate, (const int*)e->kv_off, (float*)e->actx));
    else QKL(launch_pdl(qwen_attn_decode_kernel<float, DH>, dim3(rows * H), dim3(kAttDecThreads), dsmem, e->st, e->use_pdl && (rows <= 2 || e->pdl_all), (const float*)e->qkvf, g, cosT, sinT, c.rms_eps, (float*)e->kc, (float*)e->vc, layer_off, H, KH, c.max_seq_len, (const DecState*)e->dstate, (const int*)e->kv_off, (float*)e->actx));
    return B200ASR_OK;
  }
  const int total_qk = rows * (H + 2 * KH), total_at = rows * H;
  const size_t smem = (size_t)4 * (DH + c.max_seq_len) * sizeof(float);
  if (ad == kBF16) {
    qwen_qk_rope_kernel<bf16, DH><<<(total_qk + 3) / 4, 128, 0, e->st>>>(e->qkvf, g, cosT, sinT, c.rms_eps, n_new, H, KH, c.max_seq_len, layer_off, total_qk, e->dstate, e->q, (bf16*)e->kc, (bf16*)e->vc);
    QKL(cudaGetLastError());
    if (e->use_attn_tiled) qwen_attn_prefill_kernel<bf16, bf16, DH><<<dim3((n_new + kPfRows - 1) / kPfRows, H, rows / n_new), kPfRows * 32, 0, e->st>>>(e->q, (const bf16*)e->kc, (const bf16*)e->vc, layer_off, n_new, H, KH, c.max_seq_len, e->dstate, (bf16*)e->actx);
    else qwen_attn_kernel<bf16, bf16, DH><<<(total_at + 3) / 4, 128, smem, e->st>>>(e->q, (const bf16*)e->kc, (const bf16*)e->vc, layer_off, n_new, H, KH, c.max_seq_len, total_at, e->dstate, (bf16*)e->actx);
  } else {
    qwen_qk_rope_kernel<float, DH><<<(total_qk + 3) / 4, 128, 0, e->st>>>(e->qkvf, g, cosT, sinT, c.rms_eps, n_new, H, KH, c.max_seq_len, layer_off, total_qk, e->dstate, e->q, (float*)e->kc, (float*)e->vc);
    QKL(cudaGetLastError());
    if (e->use_attn_tiled) qwen_attn_prefill_kernel<float, float, DH><<<dim3((n_new + kPfRows - 1) / kPfRows, H, rows / n_new), kPfRows * 32, 0, e->st>>>(e->q, (const float*)e->kc, (const float*)e->vc, layer_off, n_new, H, KH, c.max_seq_len, e->dstate, (float*)e->actx);
    else qwen_attn_kernel<float, float, DH><<<(total_at + 3) / 4, 128, smem, e->st>>>(e->q, (const float*)e->kc, (const float*)e->vc, layer_off, n_new, H, KH, c.max_seq_len, total_at, e->dstate, (float*)e->actx);
  }
  QKL(cudaGetLastError());
  return B200ASR_OK;
}

template <typename WT, int KS>
cudaError_t qwen_gemv_launch_ks(const QGemvArgs& a, int num_sms, size_t smem, cudaStream_t st, bool pdl) {
  static AttrOnce attr;
  if (attr.need()) {
    cudaFuncSetAttribute(qwen_gemv_kernel<WT, KS, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(qwen_gemv_kernel<WT, KS, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  const int cpb = (8 / KS) * 2;
  const int n_pass = (a.N + cpb - 1) / cpb;
  // layers that fit one resident wave (3 CTAs per SM without the second register buffer) run without a column loop;
  // wider ones (the vocabulary head) loop inside 2 CTAs per SM with the next pass's loads issued ahead of the math
  if (n_pass <= 3 * num_sms) return launch_pdl(qwen_gemv_kernel<WT, KS, 2, false>, dim3(n_pass), dim3(256), smem, st, pdl, a);
  return launch_pdl(qwen_gemv_kernel<WT, KS, 2, true>, dim3(2 * num_sms), dim3(256), smem, st, pdl, a);
}

template <typename WT>
cudaError_t qwen_gemv_launch(const QGemvArgs& a, int num_sms, cudaStream_t st, bool pdl) {
  if (a.K % 8 != 0 || a.ldx % 4 != 0 || (reinterpret_cast<uintptr_t>(a.x) & 15)) return cudaErrorInvalidValue;
  const size_t smem = (size_t)kGemvRows * a.K * sizeof(float);
  // warps per column: a warp covers at most four 256-wide chunks of its columns' K range
  int ks = 1;
  while (ks < 8 && a.K > ks * 256 * kGemvCH) ks *= 2;
  if (a.K > 8 * 256 * kGemvCH) return cudaErrorInvalidValue;
  switch (ks) {
    case 1: return qwen_gemv_launch_ks<WT, 1>(a, num_sms, smem, st, pdl);
    case 2: return qwen_gemv_launch_ks<WT, 2>(a, num_sms, smem, st, pdl);
    case 4: return qwen_gemv_launch_ks<WT, 4>(a, num_sms, smem, st, pdl);
    default: return qwen_gemv_launch_ks<WT, 8>(a, num_sms, smem, st, pdl);
  }
}

int qwen_gemv(b200asr_qwen* e, const float* x, int64_t ldx, bool rms, bool swiglu, const std::string& wn, const float* residual, int64_t ldr,
              float* out, int64_t ldo, int rows, int N, int K) {
  QGemvArgs a{};
  a.x = x; a.ldx = ldx; a.rms = rms ? 1 : 0; a.eps = e->cfg.rms_eps; a.swiglu = swiglu ? 1 : 0;
  a.W = QW(e, wn); a.residual = residual; a.ldr = ldr; a.out = out; a.ldo = ldo;
  a.rows = rows; a.N = N; a.K = K;
  const bool pdl = e->use_pdl && (rows <= 2 || e->pdl_all);      // measured: helps at 1-2 rows (1.22 -> 1.09 ms/step), hurts at 4 (1.86 -> 2.10)
  QKL(e->act == kBF16 ? qwen_gemv_launch<bf16>(a, e->num_sms, e->st, pdl) : qwen_gemv_launch<float>(a, e->num_sms, e->st, pdl));
  return B200ASR_OK;
}

// ---- the decoder layers of one decode step as one cooperative launch (bf16, <= 4 clips, known reduction-length classes) ----
struct QPersistPlan { const void* fn; size_t smem; };
template <int DH, int KSH, int KSQ, int KSI>
QPersistPlan qwen_persist_plan_t(size_t smem) { return {(const void*)qwen_persist_kernel<DH, KSH, KSQ, KSI>, smem}; }

QPersistPlan qwen_persist_plan(const b200asr_qwen_config& c) {
  const int Hd = c.hidden, QD = c.heads * c.head_dim, I = c.inter;
  const int cap = 8 * 256 * kGemvCH;
  if (Hd % 8 || QD % 8 || I % 8 || Hd > cap || QD > cap || I > cap) return {nullptr, 0};
  if ((c.max_seq_len + kPersistSK - 1) / kPersistSK > (c.max_seq_len + kSplitKeys - 1) / kSplitKeys) return {nullptr, 0};      // att_part capacity
  const int mk = Hd > QD ? (Hd > I ? Hd : I) : (QD > I ? QD : I);
  const size_t smem = (size_t)kGemvRows * mk * sizeof(float);
  const int ksh = qp_ks_for(Hd), ksq = qp_ks_for(QD), ksi = qp_ks_for(I);
  if (c.head_dim == 128 && ksh == 1 && ksq == 2 && ksi == 4) return qwen_persist_plan_t<128, 1, 2, 4>(smem);      // Qwen3-ASR-0.6B
  if (c.head_dim == 128 && ksh == 2 && ksq == 2 && ksi == 8) return qwen_persist_plan_t<128, 2, 2, 8>(smem);      // Qwen3-ASR-1.7B
  if (c.head_dim == 64 && ksh == 1 && ksq == 1 && ksi == 1) return qwen_persist_plan_t<64, 1, 1, 1>(smem);        // test dims
  return {nullptr, 0};
}

// once, at finalize: can 2 CTAs per SM of the kernel for these dims be co-resident?  persist_grid = CTAs, or -1
void qwen_persist_probe(b200asr_qwen* e) {
  e->persist_grid = -1;
  if (e->act != kBF16) return;
  const QPersistPlan pl = qwen_persist_plan(e->cfg);
  if (!pl.fn) return;
  int per_sm = 0, coop = 0;
  // (always: the 48 KB default covers static + dynamic shared memory together)
  if (cudaFuncSetAttribute(pl.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem) != cudaSuccess) { cudaGetLastError(); return; }
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pl.fn, kPersistThreads, pl.smem) != cudaSuccess) { cudaGetLastError(); return; }
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, e->cfg.device);
  if (per_sm >= 2 && coop) e->persist_grid = 2 * e->num_sms;
}
bool qwen_persist_active(const b200asr_qwen* e, int rows) {
  return e->use_persist && e->persist_grid > 0 && e->p_layers && rows <= kGemvRows && e->samp_temperature <= 0.f ? true : false;
}

int qwen_persist_layers(b200asr_qwen* e, int rows) {
  const b200asr_qwen_config& c = e->cfg;
  const QPersistPlan pl = qwen_persist_plan(c);
  QPersistArgs a{};
  a.layers = e->p_layers; a.n_layers = c.dec_layers;
  a.x = e->x; a.qkvf = e->qkvf; a.actx = (float*)e->actx; a.gu = e->gu;
  a.rows = rows; a.hidden = c.hidden; a.heads = c.heads; a.kv_heads = c.kv_heads; a.inter = c.inter; a.max_seq = c.max_seq_len; a.eps = c.rms_eps;
  a.cosT = QWF(e, "rope_cos"); a.sinT = QWF(e, "rope_sin");
  a.kc = (bf16*)e->kc; a.vc = (bf16*)e->vc; a.cache_layer_stride = (long long)c.max_batch * c.kv_heads * c.max_seq_len * c.head_dim;
  a.state = e->dstate; a.kv_off = e->kv_off; a.att_part = e->att_part; a.att_counter = e->att_counter;
  a.S = (c.max_seq_len + kPersistSK - 1) / kPersistSK;
  a.bar = e->p_bar; a.bar_base = e->p_bar_count; a.dbg = e->persist_dbg;
  a.timing = e->p_timing; a.timing_cap = 512;
  const int grid = e->persist_per_sm == 1 ? e->persist_grid / 2 : e->persist_grid;
  void* params[] = {(void*)&a};
  QKL(cudaLaunchCooperativeKernel(pl.fn, dim3(grid), dim3(kPersistThreads), params, pl.smem, e->st));
  e->p_bar_count += 5u * (unsigned)c.dec_layers * (unsigned)grid;      // arrivals of this launch (unsigned wrap is fine: signed compare)
  return B200ASR_OK;
}

// ---- decoder over `n_new` new positions per utterance (x rows = [B][n_new][hidden], fp32), then head + selection ----
int qwen_decoder(b200asr_qwen* e, int n_new) {
  const b200asr_qwen_config& c = e->cfg;
  const int B = e->B, Hd = c.hidden, H = c.heads, KH = c.kv_heads, dh = c.head_dim, I = c.inter, ad = e->act;
  const int rows = B * n_new, NQ = (H + 2 * KH) * dh;
  const bool gemv = n_new == 1;
  const bool persist = gemv && qwen_persist_active(e, rows);
  if (persist) QRET(qwen_persist_layers(e, rows));
  for (int i = 0; i < c.dec_layers && !persist; ++i) {
    const std::string p = "dec" + std::to_string(i) + ".";
    if (gemv) {
      QRET(qwen_gemv(e, e->x, Hd, true, false, p + "qkv.w", nullptr, 0, e->qkvf, NQ, rows, NQ, Hd));
    } else {
      if (ad == kBF16) qwen_rmsnorm_kernel<bf16><<<(rows + 7) / 8, 256, 0, e->st>>>(e->x, Hd, 1, 0, nullptr, c.rms_eps, (bf16*)e->xn, Hd, rows, Hd);
      else qwen_rmsnorm_kernel<float><<<(rows + 7) / 8, 256, 0, e->st>>>(e->x, Hd, 1, 0, nullptr, c.rms_eps, (float*)e->xn, Hd, rows, Hd);
      QKL(cudaGetLastError());
      QRET(qwen_gemm(e, qwen_linear(e, e->xn, Hd, p + "qkv.w", "", e->qkvf, NQ, kF32, rows, NQ, Hd)));
    }
    if (dh == 128) QRET((qwen_attention_launch<128>(e, i, rows, n_new, gemv)));
    else QRET((qwen_attention_launch<64>(e, i, rows, n_new, gemv)));
    if (gemv) {
      QRET(qwen_gemv(e, (const float*)e->actx, H * dh, false, false, p + "o.w", e->x, Hd, e->x, Hd, rows, Hd, H * dh));
      QRET(qwen_gemv(e, e->x, Hd, true, false, p + "gate_up.w", nullptr, 0, e->gu, 2 * I, rows, 2 * I, Hd));
      QRET(qwen_gemv(e, e->gu, 2 * I, false, true, p + "down.w", e->x, Hd, e->x, Hd, rows, Hd, I));
    } else {
      GemmArgs g = qwen_linear(e, e->actx, H * dh, p + "o.w", "", e->x, Hd, kF32, rows, Hd, H * dh);
      g.residual = e->x; g.ldr = Hd;
      QRET(qwen_gemm(e, g));
      if (ad == kBF16) qwen_rmsnorm_kernel<bf16><<<(rows + 7) / 8, 256, 0, e->st>>>(e->x, Hd, 1, 0, nullptr, c.rms_eps, (bf16*)e->xn, Hd, rows, Hd);
      else qwen_rmsnorm_kernel<float><<<(rows + 7) / 8, 256, 0, e->st>>>(e->x, Hd, 1, 0, nullptr, c.rms_eps, (float*)e->xn, Hd, rows, Hd);
      QKL(cudaGetLastError());
      QRET(qwen_gemm(e, qwen_linear(e, e->xn, Hd, p + "gate_up.w", "", e->gu, 2 * I, kF32, rows, 2 * I, Hd)));
      if (ad == kBF16) qwen_swiglu_kernel<bf16><<<grid_for((int64_t)rows * I), 256, 0, e->st>>>(e->gu, I, (int64_t)rows * I, (bf16*)e->mlp);
      else qwen_swiglu_kernel<float><<<grid_for((int64_t)rows * I), 256, 0, e->st>>>(e->gu, I, (int64_t)rows * I, (float*)e->mlp);
      QKL(cudaGetLastError());
      GemmArgs g2 = qwen_linear(e, e->mlp, I, p + "down.w", "", e->x, Hd, kF32, rows, Hd, I);
      g2.residual = e->x; g2.ldr = Hd;
      QRET(qwen_gemm(e, g2));
    }
  }
  // final RMS norm with its learned weight on the last row of every utterance, then the vocabulary projection (:1331-1333)
  //   (prefill of a ragged batch: the shorter clips' last prompt rows sit kv_off[b] rows before the batch's last row)
  qwen_rmsnorm_kernel<float><<<(B + 7) / 8, 256, 0, e->st>>>(e->x, Hd, n_new, n_new - 1, QWF(e, "final_norm.g"), c.rms_eps, e->xl, Hd, B, Hd,
                                                             n_new > 1 ? e->kv_off : nullptr);
  QKL(cudaGetLastError());
  const std::string head = e->w.count("lm_head.w") ? "lm_head.w" : "embed.w";
  QRET(qwen_gemv(e, e->xl, Hd, false, false, head, nullptr, 0, e->logits, c.vocab, B, c.vocab, Hd));
  const bool sampling = e->samp_temperature > 0.f;
  if (!sampling && n_new == 1 && e->repeat_penalty != 1.0f && e->penalty_range > 0) {      // decode steps only: the prefill head is a plain arg-max
    qwen_penalty_kernel<<<B, 32, 0, e->st>>>(e->logits, c.vocab, e->save_id, c.max_seq_len, e->n_save, e->penalty_range, e->repeat_penalty);
    QKL(cudaGetLastError());
  }
  SelectArgs s{};
  if (!sampling) {
    qwen_argmax_slices_kernel<<<dim3(kArgSlices, B), 256, 0, e->st>>>(e->logits, c.vocab, e->cand_val, e->cand_idx);
    QKL(cudaGetLastError());
    s.logits = e->cand_val; s.vocab = kArgSlices; s.cand_idx = e->cand_idx;
  } else {                       // TOPK_TOPP_SAMPLING (:1348-1400) works on the full row, on the prefill head as well (:640-644)
    s.logits = e->logits; s.vocab = c.vocab; s.cand_idx = nullptr;
  }
  s.batch = B; s.begin_bias = nullptr;
  s.cur_token = e->cur_token; s.tokens = e->tokens; s.tokens_ld = c.max_seq_len; s.n_gen = e->n_gen;
  s.finished = e->finished; s.save_id = e->save_id; s.save_ld = c.max_seq_len; s.n_save = e->n_save;
  s.selected_hist = nullptr; s.sel_ld = 0;
  s.stop_ids = e->d_stop; s.n_stop = (int)e->stop_ids.size(); s.limit = e->limit;
  s.penalty_value = 1.0f; s.penalty_range = 0; s.state = e->dstate; s.n_new = n_new;
  s.temperature = e->samp_temperature; s.top_k = e->samp_top_k; s.top_p = e->samp_top_p; s.rep_penalty = e->samp_rep;
  s.seed = e->samp_seed; s.noise = e->samp_noise; s.noise_ld = e->samp_top_k; s.noise_rows = e->samp_noise_rows;
  s.noise_batch = e->cfg.max_batch;          // noise rows are laid out [launch][max_batch][top_k]
  s.limit_v = e->d_limit;
  QKL(launch_select_token(s, e->st));
  e->launches++;
  return B200ASR_OK;
}

int qwen_step(b200asr_qwen* e, const int* token_src) {
  const b200asr_qwen_config& c = e->cfg;
  if (e->act == kBF16) qwen_embed_kernel<bf16><<<e->B, 128, 0, e->st>>>(token_src, (const bf16*)QW(e, "embed.w"), c.hidden, c.vocab, e->x);
  else qwen_embed_kernel<float><<<e->B, 128, 0, e->st>>>(token_src, (const float*)QW(e, "embed.w"), c.hidden, c.vocab, e->x);
  QKL(cudaGetLastError());
  return qwen_decoder(e, 1);
}

int qwen_ensure_graph(b200asr_qwen* e) {
  if (e->step_graph && e->graph_B == e->B && e->graph_limit == e->limit) return B200ASR_OK;
  if (e->step_graph) { cudaGraphExecDestroy(e->step_graph); e->step_graph = nullptr; }
  cudaGraph_t graph = nullptr;
  const int64_t before = e->launches;
  QCK(cudaStreamBeginCapture(e->st, cudaStreamCaptureModeThreadLocal));
  const int r = qwen_step(e, e->cur_token);
  const cudaError_t ce = cudaStreamEndCapture(e->st, &graph);
  e->graph_nodes = e->launches - before;
  e->launches = before;
  if (r != B200ASR_OK) { if (graph) cudaGraphDestroy(graph); return r; }
  if (ce != cudaSuccess) return e->cuda_fail(ce, "cudaStreamEndCapture");
  const cudaError_t ie = cudaGraphInstantiate(&e->step_graph, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) return e->cuda_fail(ie, "cudaGraphInstantiate");
  e->graph_B = e->B; e->graph_limit = e->limit;
  return B200ASR_OK;
}

int qwen_fetch_logits_token(b200asr_qwen* e, float* logits_out, int32_t* token_out) {
  QCK(cudaStreamSynchronize(e->st));
  if (logits_out) QCK(cudaMemcpy(logits_out, e->logits, (size_t)e->B * e->cfg.vocab * 4, cudaMemcpyDeviceToHost));
  if (token_out) QCK(cudaMemcpy(token_out, e->cur_token, (size_t)e->B * 4, cudaMemcpyDeviceToHost));
  return B200ASR_OK;
}

// pinned host staging: [max_batch][max_seq_len] prompt layout | 16 (all-done flag) | [max_batch][max_win] window key counts | 4 x [max_batch]
int* hp_flag(b200asr_qwen* e) { return e->h_pinned + (size_t)e->cfg.max_batch * e->cfg.max_seq_len; }
int* hp_win(b200asr_qwen* e) { return hp_flag(e) + 16; }
int* hp_misc(b200asr_qwen* e, int k) { return hp_win(e) + (size_t)e->cfg.max_batch * e->max_win + (size_t)k * e->cfg.max_batch; }
int audio_rows(int frames) { return (frames / kChunk) * kChunkTok + aftercnn_len(frames % kChunk); }

// `n_samples` = row stride of `pcm_host` = the longest clip; `lens` (nullable) = samples of each clip, the rest of its row is padding
int qwen_upload(b200asr_qwen* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples, const int32_t* lens = nullptr) {
  const b200asr_qwen_config& c = e->cfg;
  if (!e->finalized) return e->fail(B200ASR_E_INVALID, "weights not finalized");
  if (!pcm_host) return e->fail(B200ASR_E_INVALID, "null argument");
  if (batch <= 0 || batch > c.max_batch) return e->fail(B200ASR_E_INVALID, "batch out of range");
  if (n_samples < c.n_fft || n_samples > c.max_samples) return e->fail(B200ASR_E_INVALID, "n_samples out of range");
  if (pcm_dtype != B200ASR_PCM_I16 && pcm_dtype != B200ASR_PCM_F32) return e->fail(B200ASR_E_INVALID, "bad pcm dtype");
  if (lens) {
    int longest = 0;
    for (int b = 0; b < batch; ++b) {
      if (lens[b] < c.n_fft || lens[b] > n_samples) return e->fail(B200ASR_E_INVALID, "clip length out of range (n_fft <= lens[b] <= n_samples)");
      longest = lens[b] > longest ? lens[b] : longest;
    }
    if (longest != n_samples) return e->fail(B200ASR_E_INVALID, "n_samples must equal the longest clip of a ragged batch");
  }
  e->B = batch; e->n_samples = n_samples; e->pcm_dtype = pcm_dtype;
  e->frames = n_samples / c.hop;
  e->n_chunks = (e->frames + kChunk - 1) / kChunk;
  e->n_win = (e->n_chunks + c.chunks_per_window - 1) / c.chunks_per_window;
  e->n_audio = audio_rows(e->frames);
  e->clip_ns.assign((size_t)batch, n_samples);
  if (lens) e->clip_ns.assign(lens, lens + batch);
  e->prefilled = false;
  int* hn = hp_misc(e, 0);
  memcpy(hn, e->clip_ns.data(), (size_t)batch * 4);
  QCK(cudaMemcpyAsync(e->d_ns, hn, (size_t)batch * 4, cudaMemcpyHostToDevice, e->st));
  QCK(cudaMemcpyAsync(e->pcm, pcm_host, (size_t)batch * n_samples * (pcm_dtype == B200ASR_PCM_F32 ? 4 : 2), cudaMemcpyHostToDevice, e->st));
  return B200ASR_OK;
}

// prompt layout + window key counts (host-side bookkeeping of :866-897,925 and CONCAT_EMBED), then the encoder
int qwen_encode_resident(b200asr_qwen* e, const int32_t* query_ids, int32_t n_query, const int32_t* lang_ids, int32_t n_lang) {
  const b200asr_qwen_config& c = e->cfg;
  if (e->B <= 0) return e->fail(B200ASR_E_INVALID, "no PCM uploaded");
  if (n_query < 0 || n_lang < 0 || (n_query && !query_ids) || (n_lang && !lang_ids)) return e->fail(B200ASR_E_INVALID, "bad prompt ids");
  // every clip's prompt = head | query | suffix | its own audio rows | tail | language tail; shorter clips are padded at the END with
  // token 0 up to the longest prompt: causal attention keeps those rows out of every real row, the head reads each clip's own last
  // row, and the decode steps overwrite their cache rows before reading them (cache position = kv_len + kv_off[b])
  std::vector<int> src;                                  // the longest clip's layout (range checks below), then every clip's
  auto push_ids = [&](const int* ids, size_t n) { for (size_t i = 0; i < n; ++i) src.push_back(ids[i]); };
  auto layout = [&](int n_audio) {
    src.clear();
    push_ids(e->head_ids.data(), e->head_ids.size());
    push_ids(query_ids, (size_t)n_query);
    push_ids(e->suffix_ids.data(), e->suffix_ids.size());
    for (int i = 0; i < n_audio; ++i) src.push_back(-i - 1);
    push_ids(e->tail_ids.data(), e->tail_ids.size());
    push_ids(lang_ids, (size_t)n_lang);
  };
  layout(e->n_audio);
  for (int i = 0; i < n_query; ++i) if (query_ids[i] < 0) return e->fail(B200ASR_E_INVALID, "prompt token id out of range");
  for (int i = 0; i < n_lang; ++i) if (lang_ids[i] < 0) return e->fail(B200ASR_E_INVALID, "prompt token id out of range");
  for (int v : e->head_ids) if (v < 0) return e->fail(B200ASR_E_INVALID, "prompt token id out of range");
  for (int v : e->suffix_ids) if (v < 0) return e->fail(B200ASR_E_INVALID, "prompt token id out of range");
  for (int v : e->tail_ids) if (v < 0) return e->fail(B200ASR_E_INVALID, "prompt token id out of range");
  for (int v : src) if (v >= c.vocab) return e->fail(B200ASR_E_INVALID, "prompt token id out of range");
  e->n_prompt = (int)src.size();
  if (e->n_prompt > c.max_seq_len) return e->fail(B200ASR_E_INVALID, "prompt longer than max_seq_len");
  int* hp = e->h_pinned;
  int* hv = hp_win(e);
  int *h_off = hp_misc(e, 1), *h_lim = hp_misc(e, 2);
  const int NP = e->n_prompt;
  e->clip_prompt.assign((size_t)e->B, NP);
  e->limit = 0;
  for (int b = 0; b < e->B; ++b) {
    const int frames_b = e->clip_ns[b] / c.hop;
    layout(audio_rows(frames_b));
    const int np = (int)src.size();
    e->clip_prompt[b] = np;
    memcpy(hp + (size_t)b * NP, src.data(), (size_t)np * 4);
    for (int i = np; i < NP; ++i) hp[(size_t)b * NP + i] = 0;
    h_off[b] = np - NP;
    int lim = c.max_seq_len - 10 - np;                       // Inference_Qwen_ASR_ONNX.py:666
    h_lim[b] = lim < 0 ? 0 : lim;
    e->limit = h_lim[b] > e->limit ? h_lim[b] : e->limit;    // the loop runs to the largest limit; the selection kernel holds each clip to its own
    for (int w = 0; w < e->n_win; ++w) {
      int n = 0;
      for (int k = 0; k < c.chunks_per_window; ++k) {
        const int ch = w * c.chunks_per_window + k;
        int len = frames_b - ch * kChunk;
        len = len < 0 ? 0 : (len > kChunk ? kChunk : len);
        if (ch < e->n_chunks) n += aftercnn_len(len);
      }
      hv[b * e->n_win + w] = n;
    }
  }
  QCK(cudaMemcpyAsync(e->prompt_src, hp, (size_t)e->B * NP * 4, cudaMemcpyHostToDevice, e->st));
  QCK(cudaMemcpyAsync(e->win_valid, hv, (size_t)e->B * e->n_win * 4, cudaMemcpyHostToDevice, e->st));
  QCK(cudaMemcpyAsync(e->kv_off, h_off, (size_t)e->B * 4, cudaMemcpyHostToDevice, e->st));
  QCK(cudaMemcpyAsync(e->d_limit, h_lim, (size_t)e->B * 4, cudaMemcpyHostToDevice, e->st));
  QRET(qwen_encoder(e));
  const int64_t enc_stride = (int64_t)e->n_win * c.chunks_per_window * kChunkTok * c.out_dim;
  if (e->act == kBF16) qwen_prompt_kernel<bf16><<<dim3(e->n_prompt, e->B), 128, 0, e->st>>>(e->prompt_src, (const bf16*)QW(e, "embed.w"), e->enc_out, enc_stride, e->n_prompt, c.hidden, e->x);
  else qwen_prompt_kernel<float><<<dim3(e->n_prompt, e->B), 128, 0, e->st>>>(e->prompt_src, (const float*)QW(e, "embed.w"), e->enc_out, enc_stride, e->n_prompt, c.hidden, e->x);
  QKL(cudaGetLastError());
  qwen_reset_kernel<<<1, 32, 0, e->st>>>(e->dstate, e->n_gen, e->finished, e->n_save, e->B);
  QKL(cudaGetLastError());
  e->prefilled = false;
  return B200ASR_OK;
}

int qwen_decode_loop(b200asr_qwen* e, int max_new) {
  // the prefill already selected (and possibly accepted) the first token; every replay adds at most one more
  int steps = e->limit - 1;
  if (max_new >= 0 && max_new - 1 < steps) steps = max_new - 1;
  const bool graph = e->use_graph && !qwen_persist_active(e, e->B);      // (a cooperative launch carries a per-launch barrier base: not replayed)
  if (graph) QRET(qwen_ensure_graph(e));
  for (int s = 0; s < steps; ++s) {
    if ((s & 15) == 0) {
      int* flag = hp_flag(e);
      QCK(cudaMemcpyAsync(flag, &e->dstate->all_done, 4, cudaMemcpyDeviceToHost, e->st));
      QCK(cudaStreamSynchronize(e->st));
      if (*flag) break;
    }
    if (graph) { QCK(cudaGraphLaunch(e->step_graph, e->st)); e->launches += e->graph_nodes; }
    else QRET(qwen_step(e, e->cur_token));
  }
  return B200ASR_OK;
}

int qwen_fetch_tokens(b200asr_qwen* e, int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out) {
  if (!tokens_out || !lens_out || tokens_ld <= 0) return e->fail(B200ASR_E_INVALID, "null output");
  const int B = e->B, ld = e->cfg.max_seq_len;
  std::vector<int> hl((size_t)B), ht((size_t)B * ld);
  QCK(cudaStreamSynchronize(e->st));
  QCK(cudaMemcpy(hl.data(), e->n_gen, (size_t)B * 4, cudaMemcpyDeviceToHost));
  QCK(cudaMemcpy(ht.data(), e->tokens, (size_t)B * ld * 4, cudaMemcpyDeviceToHost));
  for (int b = 0; b < B; ++b) {
    const int n = hl[b] < tokens_ld ? hl[b] : tokens_ld;
    lens_out[b] = n;
    memcpy(tokens_out + (size_t)b * tokens_ld, ht.data() + (size_t)b * ld, (size_t)n * 4);
  }
  return B200ASR_OK;
}

}  // namespace

extern "C" {

const char* b200asr_qwen_create_error(void) { return g_qwen_create_error.c_str(); }
const char* b200asr_qwen_last_error(const b200asr_qwen* e) { return e ? e->err.c_str() : g_qwen_create_error.c_str(); }

int b200asr_qwen_create(const b200asr_qwen_config* cfg, b200asr_qwen** out) {
  if (!cfg || !out) { g_qwen_create_error = "null argument"; return B200ASR_E_INVALID; }
  *out = nullptr;
  const b200asr_qwen_config& c = *cfg;
  auto bad = [&](const char* m) { g_qwen_create_error = m; return B200ASR_E_INVALID; };
  if (c.n_mels <= 0 || c.n_fft <= 0 || c.hop <= 0) return bad("bad front-end dims");
  if (c.enc_d <= 0 || c.enc_heads <= 0 || c.enc_d % c.enc_heads || c.enc_layers < 0 || c.enc_ffn <= 0) return bad("bad encoder dims");
  if (c.conv_ch <= 0 || c.conv_ch % 8) return bad("conv_ch must be a positive multiple of 8");
  if (c.chunks_per_window <= 0) return bad("bad chunks_per_window");
  if (c.out_dim != c.hidden) return bad("encoder output_dim must equal the decoder hidden size");
  if (c.hidden <= 0 || c.hidden % 8 || c.inter <= 0 || c.inter % 8 || c.vocab <= 0 || c.dec_layers <= 0) return bad("bad decoder dims");
  if (c.heads <= 0 || c.kv_heads <= 0 || c.heads % c.kv_heads) return bad("heads must be a multiple of kv_heads");
  if (c.head_dim != 64 && c.head_dim != 128) return bad("head_dim must be 64 or 128");
  if (c.max_batch <= 0 || c.max_batch > 32 || c.max_samples < c.n_fft || c.max_seq_len < 32) return bad("bad capacity");
  if (c.precision != B200ASR_PRECISION_F32 && c.precision != B200ASR_PRECISION_BF16) return bad("bad precision");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0 || c.device < 0 || c.device >= n_dev) {
    cudaGetLastError();
    g_qwen_create_error = "no usable CUDA device (the B200 engine has no CPU fallback)";
    return B200ASR_E_NOGPU;
  }
  cudaDeviceProp prop{};
  cudaGetDeviceProperties(&prop, c.device);
  if (prop.major != 10) { g_qwen_create_error = "device is not sm_100 (Blackwell B200)"; return B200ASR_E_NOGPU; }
  if (cudaSetDevice(c.device) != cudaSuccess) { g_qwen_create_error = "cudaSetDevice failed"; return B200ASR_E_CUDA; }
  b200asr_qwen* e = new b200asr_qwen();
  e->cfg = c;
  e->num_sms = prop.multiProcessorCount;
  e->act = c.precision == B200ASR_PRECISION_BF16 ? kBF16 : kF32;
  e->es = dtype_size(e->act);
  if (cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking) != cudaSuccess) { delete e; g_qwen_create_error = "stream creation failed"; return B200ASR_E_CUDA; }
  *out = e;
  return B200ASR_OK;
}

void b200asr_qwen_destroy(b200asr_qwen* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  cudaStreamSynchronize(e->st);
  if (e->step_graph) cudaGraphExecDestroy(e->step_graph);
  for (auto& kv : e->w) cudaFree(kv.second.ptr);
  void* bufs[] = {e->basis_t, e->fb_start, e->fb_len, e->stage_buf, e->d_stop, e->pcm, e->mel_raw, e->max_key, e->feat, e->c1, e->col, e->c2, e->c3,
                  e->stem, e->h, e->S, e->enc_out, e->xhat, e->qkv, e->ctx, e->ffn, e->P, e->win_valid, e->prompt_src, e->d_ns, e->kv_off, e->d_limit, e->x, e->qkvf, e->q, e->gu,
                  e->xl, e->logits, e->cand_val, e->cand_idx, e->att_part, e->att_counter, e->p_layers, e->p_bar, e->p_timing, e->samp_noise, e->xn, e->actx, e->mlp, e->kc, e->vc, e->dstate, e->cur_token, e->tokens, e->n_gen, e->finished, e->save_id, e->n_save};
  for (void* p : bufs) if (p) cudaFree(p);
  if (e->h_pinned) cudaFreeHost(e->h_pinned);
  cudaStreamDestroy(e->st);
  delete e;
}

int b200asr_qwen_set_tensor(b200asr_qwen* e, const char* name_c, const float* host, int64_t numel) {
  if (!e || !name_c || !host || numel <= 0) return B200ASR_E_INVALID;
  QCK(cudaSetDevice(e->cfg.device));
  const std::string name(name_c);
  const b200asr_qwen_config& c = e->cfg;
  const int F = c.n_fft / 2 + 1;
  if (name == "stft_kernel") {          // [2F][n_fft] Conv1d weight -> [n_fft][2F] (frequency bin = fast axis of a warp)
    if (numel != (int64_t)2 * F * c.n_fft) return e->fail(B200ASR_E_INVALID, "stft_kernel size mismatch");
    std::vector<float> t((size_t)numel);
    for (int r = 0; r < 2 * F; ++r)
      for (int k = 0; k < c.n_fft; ++k) t[(size_t)k * 2 * F + r] = host[(size_t)r * c.n_fft + k];
    if (!e->basis_t) QCK(cudaMalloc(&e->basis_t, (size_t)numel * 4));
    QCK(cudaMemcpy(e->basis_t, t.data(), (size_t)numel * 4, cudaMemcpyHostToDevice));
  }
  if (name == "mel_fbank") {            // [n_mels][F]: non-zero span of every triangular filter
    if (numel != (int64_t)c.n_mels * F) return e->fail(B200ASR_E_INVALID, "mel_fbank size mismatch");
    std::vector<int> s0(c.n_mels), ln(c.n_mels);
    for (int m = 0; m < c.n_mels; ++m) {
      int lo = F, hi = -1;
      for (int f = 0; f < F; ++f) if (host[(size_t)m * F + f] != 0.f) { if (f < lo) lo = f; hi = f; }
      s0[m] = hi < 0 ? 0 : lo; ln[m] = hi < 0 ? 0 : hi - lo + 1;
    }
    if (!e->fb_start) { QCK(cudaMalloc(&e->fb_start, c.n_mels * 4)); QCK(cudaMalloc(&e->fb_len, c.n_mels * 4)); }
    QCK(cudaMemcpy(e->fb_start, s0.data(), c.n_mels * 4, cudaMemcpyHostToDevice));
    QCK(cudaMemcpy(e->fb_len, ln.data(), c.n_mels * 4, cudaMemcpyHostToDevice));
  }
  std::vector<float> relaid;
  const int C = c.conv_ch;
  if (name == "conv2.w" || name == "conv3.w") {      // Conv2d [co][ci][k_mel][k_time] -> [co][(k_time*3 + k_mel)*C + ci] = the im2col column order
    if (numel != (int64_t)C * C * 9) return e->fail(B200ASR_E_INVALID, name + " size mismatch");
    relaid.resize((size_t)numel);
    for (int co = 0; co < C; ++co)
      for (int ci = 0; ci < C; ++ci)
        for (int kf = 0; kf < 3; ++kf)
          for (int kt = 0; kt < 3; ++kt)
            relaid[((size_t)co * 9 + kt * 3 + kf) * C + ci] = host[(((size_t)co * C + ci) * 3 + kf) * 3 + kt];
    host = relaid.data();
  }
  if (name == "conv_out.w") {           // Linear over (channel, mel) = c*F3 + f -> our channel-last rows (mel, channel) = f*C + c (:872-876)
    const int F3 = ((((c.n_mels + 1) / 2 + 1) / 2 + 1) / 2);
    if (numel != (int64_t)c.enc_d * C * F3) return e->fail(B200ASR_E_INVALID, "conv_out.w size mismatch");
    relaid.resize((size_t)numel);
    for (int o = 0; o < c.enc_d; ++o)
      for (int ch = 0; ch < C; ++ch)
        for (int f = 0; f < F3; ++f) relaid[((size_t)o * F3 + f) * C + ch] = host[((size_t)o * C + ch) * F3 + f];
    host = relaid.data();
  }
  QTensor t;
  t.numel = numel;
  t.dtype = qwen_is_matrix(name) ? e->act : kF32;
  auto it = e->w.find(name);
  if (it != e->w.end()) { cudaFree(it->second.ptr); e->w.erase(it); }
  QCK(cudaMalloc(&t.ptr, (size_t)numel * dtype_size(t.dtype)));
  if (t.dtype == kF32) {
    QCK(cudaMemcpyAsync(t.ptr, host, (size_t)numel * 4, cudaMemcpyHostToDevice, e->st));
    QCK(cudaStreamSynchronize(e->st));
  } else {
    if (numel > e->stage_cap) {
      if (e->stage_buf) cudaFree(e->stage_buf);
      e->stage_buf = nullptr; e->stage_cap = 0;
      QCK(cudaMalloc(&e->stage_buf, (size_t)numel * 4));
      e->stage_cap = numel;
    }
    QCK(cudaMemcpyAsync(e->stage_buf, host, (size_t)numel * 4, cudaMemcpyHostToDevice, e->st));
    qwen_f32_to_bf16<<<1024, 256, 0, e->st>>>(e->stage_buf, (bf16*)t.ptr, numel);
    QCK(cudaGetLastError());
    QCK(cudaStreamSynchronize(e->st));
  }
  e->w[name] = t;
  e->finalized = false;
  return B200ASR_OK;
}

int b200asr_qwen_set_prompt(b200asr_qwen* e, const int32_t* head_ids, int32_t n_head, const int32_t* suffix_ids, int32_t n_suffix,
                            const int32_t* tail_ids, int32_t n_tail, const int32_t* stop_ids, int32_t n_stop) {
  if (!e) return B200ASR_E_INVALID;
  QCK(cudaSetDevice(e->cfg.device));
  if (n_head < 0 || n_suffix < 0 || n_tail < 0 || n_stop < 0 || n_stop > 64) return e->fail(B200ASR_E_INVALID, "bad prompt sizes");
  if ((n_head && !head_ids) || (n_suffix && !suffix_ids) || (n_tail && !tail_ids) || (n_stop && !stop_ids)) return e->fail(B200ASR_E_INVALID, "null prompt ids");
  e->head_ids.assign(head_ids, head_ids + n_head);
  e->suffix_ids.assign(suffix_ids, suffix_ids + n_suffix);
  e->tail_ids.assign(tail_ids, tail_ids + n_tail);
  e->stop_ids.assign(stop_ids, stop_ids + n_stop);
  if (!e->d_stop) QCK(cudaMalloc(&e->d_stop, 64 * 4));
  if (n_stop) QCK(cudaMemcpy(e->d_stop, stop_ids, (size_t)n_stop * 4, cudaMemcpyHostToDevice));
  if (e->step_graph) { cudaGraphExecDestroy(e->step_graph); e->step_graph = nullptr; }
  return B200ASR_OK;
}

int b200asr_qwen_set_decode_options(b200asr_qwen* e, float repeat_penalty, int32_t penalty_range) {
  if (!e) return B200ASR_E_INVALID;
  if (penalty_range < 0 || !(repeat_penalty > 0.f)) return e->fail(B200ASR_E_INVALID, "bad penalty options");
  e->repeat_penalty = repeat_penalty; e->penalty_range = penalty_range;
  if (e->step_graph) { cudaSetDevice(e->cfg.device); cudaGraphExecDestroy(e->step_graph); e->step_graph = nullptr; }
  return B200ASR_OK;
}

int b200asr_qwen_set_sampling(b200asr_qwen* e, float temperature, int32_t top_k, float top_p, float repetition_penalty, uint64_t seed,
                              const float* noise_host, int32_t noise_rows) {
  if (!e) return B200ASR_E_INVALID;
  QCK(cudaSetDevice(e->cfg.device));
  if (temperature > 0.f && (top_k < 1 || top_k > 64 || top_k > e->cfg.vocab || !(top_p > 0.f) || !(repetition_penalty > 0.f)))
    return e->fail(B200ASR_E_INVALID, "bad sampling options (1 <= top_k <= 64, top_p > 0, repetition_penalty > 0)");
  e->samp_temperature = temperature; e->samp_top_k = top_k; e->samp_top_p = top_p; e->samp_rep = repetition_penalty; e->samp_seed = seed;
  if (e->samp_noise) { cudaFree(e->samp_noise); e->samp_noise = nullptr; }
  e->samp_noise_rows = 0;
  if (temperature > 0.f && noise_host && noise_rows > 0) {
    const size_t n = (size_t)noise_rows * e->cfg.max_batch * top_k;
    QCK(cudaMalloc(&e->samp_noise, n * 4));
    QCK(cudaMemcpy(e->samp_noise, noise_host, n * 4, cudaMemcpyHostToDevice));
    e->samp_noise_rows = noise_rows;
  }
  if (e->step_graph) { cudaGraphExecDestroy(e->step_graph); e->step_graph = nullptr; }
  return B200ASR_OK;
}

int b200asr_qwen_finalize_weights(b200asr_qwen* e) {
  if (!e) return B200ASR_E_INVALID;
  QCK(cudaSetDevice(e->cfg.device));
  const b200asr_qwen_config& c = e->cfg;
  const int64_t F = c.n_fft / 2 + 1, C = c.conv_ch, D = c.enc_d, Hd = c.hidden, I = c.inter, dh = c.head_dim;
  const int64_t F3 = ((((c.n_mels + 1) / 2 + 1) / 2 + 1) / 2);
  QRET(qwen_need(e, "stft_kernel", 2 * F * c.n_fft)); QRET(qwen_need(e, "mel_fbank", c.n_mels * F));
  QRET(qwen_need(e, "conv1.w", C * 9)); QRET(qwen_need(e, "conv1.b", C));
  QRET(qwen_need(e, "conv2.w", C * C * 9)); QRET(qwen_need(e, "conv2.b", C));
  QRET(qwen_need(e, "conv3.w", C * C * 9)); QRET(qwen_need(e, "conv3.b", C));
  QRET(qwen_need(e, "conv_out.w", D * C * F3)); QRET(qwen_need(e, "enc_pos", kChunkTok * D));
  for (int i = 0; i < c.enc_layers; ++i) {
    const std::string p = "enc" + std::to_string(i) + ".";
    QRET(qwen_need(e, p + "qkv.w", 3 * D * D)); QRET(qwen_need(e, p + "qkv.b", 3 * D));
    QRET(qwen_need(e, p + "out.w", D * D)); QRET(qwen_need(e, p + "out.b", D));
    QRET(qwen_need(e, p + "fc1.w", (int64_t)c.enc_ffn * D)); QRET(qwen_need(e, p + "fc1.b", c.enc_ffn));
    QRET(qwen_need(e, p + "fc2.w", (int64_t)c.enc_ffn * D)); QRET(qwen_need(e, p + "fc2.b", D));
  }
  QRET(qwen_need(e, "proj1.w", D * D)); QRET(qwen_need(e, "proj1.b", D));
  QRET(qwen_need(e, "proj2.w", (int64_t)c.out_dim * D)); QRET(qwen_need(e, "proj2.b", c.out_dim));
  QRET(qwen_need(e, "embed.w", (int64_t)c.vocab * Hd));
  if (e->w.count("lm_head.w")) QRET(qwen_need(e, "lm_head.w", (int64_t)c.vocab * Hd));     // absent = tied to embed.w
  QRET(qwen_need(e, "final_norm.g", Hd));
  QRET(qwen_need(e, "rope_cos", (int64_t)c.max_seq_len * dh / 2, true)); QRET(qwen_need(e, "rope_sin", (int64_t)c.max_seq_len * dh / 2, true));
  const int64_t NQ = (int64_t)(c.heads + 2 * c.kv_heads) * dh;
  for (int i = 0; i < c.dec_layers; ++i) {
    const std::string p = "dec" + std::to_string(i) + ".";
    QRET(qwen_need(e, p + "qkv.w", NQ * Hd)); QRET(qwen_need(e, p + "qk_norm.g", 2 * dh));
    QRET(qwen_need(e, p + "o.w", Hd * c.heads * dh));
    QRET(qwen_need(e, p + "gate_up.w", 2 * I * Hd)); QRET(qwen_need(e, p + "down.w", Hd * I));
  }
  if (e->finalized) return B200ASR_OK;
  if (e->stage_buf) { cudaFree(e->stage_buf); e->stage_buf = nullptr; e->stage_cap = 0; }
  if (!e->pcm) {
    const size_t es = e->es;
    const int64_t B = c.max_batch;
    e->max_frames = c.max_samples / c.hop;
    e->max_chunks = (e->max_frames + kChunk - 1) / kChunk;
    e->max_win = (e->max_chunks + c.chunks_per_window - 1) / c.chunks_per_window;
    const int64_t chunks = B * e->max_chunks, tpw = (int64_t)c.chunks_per_window * kChunkTok;
    const int64_t F1 = (c.n_mels + 1) / 2, F2 = (F1 + 1) / 2;
    const int64_t Me = B * e->max_win * tpw, rows = B * c.max_seq_len;
    QRET(qwen_alloc(e, &e->pcm, (size_t)B * c.max_samples * 4));
    QRET(qwen_alloc(e, &e->mel_raw, (size_t)B * e->max_frames * c.n_mels * 4));
    QRET(qwen_alloc(e, &e->max_key, (size_t)B * 4));
    QRET(qwen_alloc(e, &e->feat, (size_t)chunks * kChunk * c.n_mels * 4));
    QRET(qwen_alloc(e, &e->c1, (size_t)chunks * 50 * F1 * C * es));
    QRET(qwen_alloc(e, &e->col, (size_t)chunks * 25 * F2 * 9 * C * es));
    QRET(qwen_alloc(e, &e->c2, (size_t)chunks * 25 * F2 * C * es));
    QRET(qwen_alloc(e, &e->c3, (size_t)chunks * kChunkTok * F3 * C * es));
    QRET(qwen_alloc(e, &e->stem, (size_t)chunks * kChunkTok * D * 4));
    QRET(qwen_alloc(e, &e->h, (size_t)Me * D * 4));
    QRET(qwen_alloc(e, &e->enc_out, (size_t)Me * c.out_dim * 4));
    QRET(qwen_alloc(e, &e->xhat, (size_t)Me * D * es));
    QRET(qwen_alloc(e, &e->qkv, (size_t)Me * 3 * D * es));
    QRET(qwen_alloc(e, &e->ctx, (size_t)Me * D * es));
    QRET(qwen_alloc(e, &e->ffn, (size_t)Me * (c.enc_ffn > D ? c.enc_ffn : D) * es));
    QRET(qwen_alloc(e, &e->S, (size_t)B * e->max_win * c.enc_heads * tpw * tpw * 4));
    QRET(qwen_alloc(e, &e->P, (size_t)B * e->max_win * c.enc_heads * tpw * tpw * es));
    QRET(qwen_alloc(e, &e->win_valid, (size_t)B * e->max_win * 4));
    QRET(qwen_alloc(e, &e->prompt_src, (size_t)B * c.max_seq_len * 4));
    QRET(qwen_alloc(e, &e->d_ns, (size_t)B * 4));
    QRET(qwen_alloc(e, &e->kv_off, (size_t)B * 4));
    QRET(qwen_alloc(e, &e->d_limit, (size_t)B * 4));
    QRET(qwen_alloc(e, &e->x, (size_t)rows * Hd * 4));
    QRET(qwen_alloc(e, &e->xn, (size_t)rows * Hd * 4));
    QRET(qwen_alloc(e, &e->qkvf, (size_t)rows * NQ * 4));
    QRET(qwen_alloc(e, &e->q, (size_t)rows * c.heads * dh * 4));
    QRET(qwen_alloc(e, &e->actx, (size_t)rows * c.heads * dh * 4));
    QRET(qwen_alloc(e, &e->gu, (size_t)rows * 2 * I * 4));
    QRET(qwen_alloc(e, &e->mlp, (size_t)rows * I * 4));
    QRET(qwen_alloc(e, &e->xl, (size_t)B * Hd * 4));
    QRET(qwen_alloc(e, &e->logits, (size_t)B * c.vocab * 4));
    QRET(qwen_alloc(e, &e->cand_val, (size_t)B * kArgSlices * 4));
    QRET(qwen_alloc(e, &e->cand_idx, (size_t)B * kArgSlices * 4));
    QRET(qwen_alloc(e, &e->att_part, (size_t)B * c.heads * ((c.max_seq_len + kSplitKeys - 1) / kSplitKeys) * (dh + 2) * 4));
    QRET(qwen_alloc(e, &e->att_counter, (size_t)B * c.heads * 4));
    const size_t kv_bytes = (size_t)c.dec_layers * B * c.kv_heads * c.max_seq_len * dh * es;
    QRET(qwen_alloc(e, &e->kc, kv_bytes));
    QRET(qwen_alloc(e, &e->vc, kv_bytes));
    QRET(qwen_alloc(e, &e->dstate, sizeof(DecState)));
    QRET(qwen_alloc(e, &e->cur_token, (size_t)B * 4));
    QRET(qwen_alloc(e, &e->tokens, (size_t)B * c.max_seq_len * 4));
    QRET(qwen_alloc(e, &e->n_gen, (size_t)B * 4));
    QRET(qwen_alloc(e, &e->finished, (size_t)B * 4));
    QRET(qwen_alloc(e, &e->save_id, (size_t)B * c.max_seq_len * 4));
    QRET(qwen_alloc(e, &e->n_save, (size_t)B * 4));
    if (!e->d_stop) QCK(cudaMalloc(&e->d_stop, 64 * 4));
    QCK(cudaMallocHost(&e->h_pinned, ((size_t)B * c.max_seq_len + 16 + B * e->max_win + 4 * B + 64) * 4));   // layout: hp_flag / hp_win / hp_misc
    const size_t smem = (size_t)4 * (dh + c.max_seq_len) * sizeof(float);
    if (smem > 48 * 1024) {
      QCK(cudaFuncSetAttribute(qwen_attn_kernel<bf16, float, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      QCK(cudaFuncSetAttribute(qwen_attn_kernel<bf16, bf16, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      QCK(cudaFuncSetAttribute(qwen_attn_kernel<float, float, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      QCK(cudaFuncSetAttribute(qwen_attn_kernel<bf16, float, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      QCK(cudaFuncSetAttribute(qwen_attn_kernel<bf16, bf16, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      QCK(cudaFuncSetAttribute(qwen_attn_kernel<float, float, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
  }
  if (e->act == kBF16) {               // (rebuilt at every finalize: set_tensor replaces device buffers)
    std::vector<QPersistLayer> tab((size_t)c.dec_layers);
    for (int i = 0; i < c.dec_layers; ++i) {
      const std::string p = "dec" + std::to_string(i) + ".";
      tab[i].qkv_w = (const bf16*)QW(e, p + "qkv.w"); tab[i].o_w = (const bf16*)QW(e, p + "o.w");
      tab[i].gu_w = (const bf16*)QW(e, p + "gate_up.w"); tab[i].down_w = (const bf16*)QW(e, p + "down.w");
      tab[i].qk_g = QWF(e, p + "qk_norm.g");
    }
    QCK(cudaStreamSynchronize(e->st));
    if (!e->p_layers) QCK(cudaMalloc(&e->p_layers, tab.size() * sizeof(QPersistLayer)));
    QCK(cudaMemcpy(e->p_layers, tab.data(), tab.size() * sizeof(QPersistLayer), cudaMemcpyHostToDevice));
    if (!e->p_bar) { QRET(qwen_alloc(e, &e->p_bar, 16)); e->p_bar_count = 0; }
    if (e->persist_grid == 0) qwen_persist_probe(e);
  }
  QCK(cudaStreamSynchronize(e->st));
  e->finalized = true;
  return B200ASR_OK;
}

int b200asr_qwen_encode(b200asr_qwen* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples,
                        const int32_t* query_ids, int32_t n_query, const int32_t* language_tail_ids, int32_t n_language_tail,
                        int32_t* n_prompt_out) {
  if (!e) return B200ASR_E_INVALID;
  QCK(cudaSetDevice(e->cfg.device));
  QRET(qwen_upload(e, pcm_host, pcm_dtype, batch, n_samples));
  QRET(qwen_encode_resident(e, query_ids, n_query, language_tail_ids, n_language_tail));
  QCK(cudaStreamSynchronize(e->st));
  if (n_prompt_out) *n_prompt_out = e->n_prompt;
  return B200ASR_OK;
}

int b200asr_qwen_prefill(b200asr_qwen* e, float* logits_out, int32_t* token_out) {
  if (!e) return B200ASR_E_INVALID;
  QCK(cudaSetDevice(e->cfg.device));
  if (e->B <= 0 || e->n_prompt <= 0) return e->fail(B200ASR_E_INVALID, "prefill before encode");
  if (e->prefilled) return e->fail(B200ASR_E_INVALID, "prefill already done for this clip");
  QRET(qwen_decoder(e, e->n_prompt));
  e->prefilled = true;
  return qwen_fetch_logits_token(e, logits_out, token_out);
}

int b200asr_qwen_decode_step(b200asr_qwen* e, const int32_t* token_in, float* logits_out, int32_t* token_out) {
  if (!e) return B200ASR_E_INVALID;
  QCK(cudaSetDevice(e->cfg.device));
  if (!e->prefilled) return e->fail(B200ASR_E_INVALID, "decode_step before prefill");
  int kv_len = 0;
  QCK(cudaStreamSynchronize(e->st));
  QCK(cudaMemcpy(&kv_len, &e->dstate->kv_len, 4, cudaMemcpyDeviceToHost));
  if (kv_len + 1 > e->cfg.max_seq_len) return e->fail(B200ASR_E_INVALID, "KV cache full");
  if (token_in) {
    QCK(cudaMemcpyAsync(e->cur_token, token_in, (size_t)e->B * 4, cudaMemcpyHostToDevice, e->st));
    QRET(qwen_step(e, e->cur_token));
  } else if (e->use_graph && !qwen_persist_active(e, e->B)) {
    QRET(qwen_ensure_graph(e));
    QCK(cudaGraphLaunch(e->step_graph, e->st));
    e->launches += e->graph_nodes;
  } else {
    QRET(qwen_step(e, e->cur_token));
  }
  return qwen_fetch_logits_token(e, logits_out, token_out);
}

int b200asr_qwen_decode(b200asr_qwen* e, int32_t max_new, int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out) {
  if (!e) return B200ASR_E_INVALID;
  QCK(cudaSetDevice(e->cfg.device));
  if (!e->prefilled) return e->fail(B200ASR_E_INVALID, "decode before prefill");
  QRET(qwen_decode_loop(e, max_new));
  return qwen_fetch_tokens(e, tokens_out, tokens_ld, lens_out);
}

int b200asr_qwen_upload(b200asr_qwen* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples) {
  if (!e) return B200ASR_E_INVALID;
  QCK(cudaSetDevice(e->cfg.device));
  QRET(qwen_upload(e, pcm_host, pcm_dtype, batch, n_samples));
  QCK(cudaStreamSynchronize(e->st));
  return B200ASR_OK;
}

static int qwen_transcribe_resident(b200asr_qwen* e, const int32_t* query_ids, int32_t n_query, const int32_t* lang_ids, int32_t n_lang,
                                    int32_t max_new, int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out) {
  QRET(qwen_encode_resident(e, query_ids, n_query, lang_ids, n_lang));
  if (max_new >= 0 && max_new < e->limit) e->limit = max_new;      // caps the accepted-token count inside the selection kernel
  QRET(qwen_decoder(e, e->n_prompt));
  e->prefilled = true;
  QRET(qwen_decode_loop(e, max_new));
  return qwen_fetch_tokens(e, tokens_out, tokens_ld, lens_out);
}

int b200asr_qwen_transcribe(b200asr_qwen* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples,
                            const int32_t* query_ids, int32_t n_query, const int32_t* language_tail_ids, int32_t n_language_tail,
                            int32_t max_new, int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out) {
  if (!e) return B200ASR_E_INVALID;
  QCK(cudaSetDevice(e->cfg.device));
  QRET(qwen_upload(e, pcm_host, pcm_dtype, batch, n_samples));
  return qwen_transcribe_resident(e, query_ids, n_query, language_tail_ids, n_language_tail, max_new, tokens_out, tokens_ld, lens_out);
}

int b200asr_qwen_transcribe_resident(b200asr_qwen* e, const int32_t* query_ids, int32_t n_query, const int32_t* language_tail_ids,
                                     int32_t n_language_tail, int32_t max_new, int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out) {
  if (!e) return B200ASR_E_INVALID;
  QCK(cudaSetDevice(e->cfg.device));
  return qwen_transcribe_resident(e, query_ids, n_query, language_tail_ids, n_language_tail, max_new, tokens_out, tokens_ld, lens_out);
}

// ---- ragged batches: clips of different lengths in one batch, each with the result it has when it runs alone ----
int b200asr_qwen_upload_ragged(b200asr_qwen* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples, const int32_t* lens) {
  if (!e) return B200ASR_E_INVALID;
  QCK(cudaSetDevice(e->cfg.device));
  if (!lens) return e->fail(B200ASR_E_INVALID, "null argument");
  QRET(qwen_upload(e, pcm_host, pcm_dtype, batch, n_samples, lens));
  QCK(cudaStreamSynchronize(e->st));
  return B200ASR_OK;
}

int b200asr_qwen_encode_ragged(b200asr_qwen* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples, const int32_t* lens,
                               const int32_t* query_ids, int32_t n_query, const int32_t* language_tail_ids, int32_t n_language_tail,
                               int32_t* n_prompt_out) {
  if (!e) return B200ASR_E_INVALID;
  QCK(cudaSetDevice(e->cfg.device));
  if (!lens) return e->fail(B200ASR_E_INVALID, "null argument");
  QRET(qwen_upload(e, pcm_host, pcm_dtype, batch, n_samples, lens));
  QRET(qwen_encode_resident(e, query_ids, n_query, language_tail_ids, n_language_tail));
  QCK(cudaStreamSynchronize(e->st));
  if (n_prompt_out) for (int b = 0; b < e->B; ++b) n_prompt_out[b] = e->clip_prompt[b];
  return B200ASR_OK;
}

int b200asr_qwen_transcribe_ragged(b200asr_qwen* e, const void* pcm_host, int32_t pcm_dtype, int32_t batch, int32_t n_samples, const int32_t* lens,
                                   const int32_t* query_ids, int32_t n_query, const int32_t* language_tail_ids, int32_t n_language_tail,
                                   int32_t max_new, int32_t* tokens_out, int32_t tokens_ld, int32_t* lens_out) {
  if (!e) return B200ASR_E_INVALID;
  QCK(cudaSetDevice(e->cfg.device));
  if (!lens) return e->fail(B200ASR_E_INVALID, "null argument");
  QRET(qwen_upload(e, pcm_host, pcm_dtype, batch, n_samples, lens));
  return qwen_transcribe_resident(e, query_ids, n_query, language_tail_ids, n_language_tail, max_new, tokens_out, tokens_ld, lens_out);
}

int b200asr_qwen_get_stage(b200asr_qwen* e, const char* name_c, float* out, int64_t capacity, int64_t* numel_out) {
  if (!e || !name_c || !out) return B200ASR_E_INVALID;
  QCK(cudaSetDevice(e->cfg.device));
  const b200asr_qwen_config& c = e->cfg;
  const std::string name(name_c);
  if (e->B <= 0) return e->fail(B200ASR_E_INVALID, "get_stage before encode");
  QCK(cudaStreamSynchronize(e->st));
  const int64_t B = e->B;
  if (name == "features") {            // [B][frames][n_mels]
    const int64_t n = B * e->frames * c.n_mels;
    if (n > capacity) return e->fail(B200ASR_E_INVALID, "stage buffer too small");
    QCK(cudaMemcpy2D(out, (size_t)e->frames * c.n_mels * 4, e->feat, (size_t)e->n_chunks * kChunk * c.n_mels * 4, (size_t)e->frames * c.n_mels * 4, B, cudaMemcpyDeviceToHost));
    if (numel_out) *numel_out = n;
    return B200ASR_OK;
  }
  if (name == "audio_hidden") {        // [B][n_audio][out_dim]
    const int64_t n = B * e->n_audio * c.out_dim;
    if (n > capacity) return e->fail(B200ASR_E_INVALID, "stage buffer too small");
    const size_t stride = (size_t)e->n_win * c.chunks_per_window * kChunkTok * c.out_dim * 4;
    QCK(cudaMemcpy2D(out, (size_t)e->n_audio * c.out_dim * 4, e->enc_out, stride, (size_t)e->n_audio * c.out_dim * 4, B, cudaMemcpyDeviceToHost));
    if (numel_out) *numel_out = n;
    return B200ASR_OK;
  }
  if (name == "persist_timing") {      // 512 x u64 nanosecond stamps as 1024 floats' worth of raw bytes
    if (!e->p_timing || capacity < 1024) return e->fail(B200ASR_E_INVALID, "persist_timing is off or the buffer is too small");
    QCK(cudaMemcpy(out, e->p_timing, 512 * 8, cudaMemcpyDeviceToHost));
    if (numel_out) *numel_out = 1024;
    return B200ASR_OK;
  }
  const void* src = nullptr; int64_t n = 0;
  if (name == "logits") { src = e->logits; n = B * c.vocab; }
  else if (name == "prompt_embed" && !e->prefilled) { src = e->x; n = B * e->n_prompt * c.hidden; }
  else return e->fail(B200ASR_E_INVALID, "unknown stage '" + name + "'");
  if (n > capacity) return e->fail(B200ASR_E_INVALID, "stage buffer too small");
  QCK(cudaMemcpy(out, src, (size_t)n * 4, cudaMemcpyDeviceToHost));
  if (numel_out) *numel_out = n;
  return B200ASR_OK;
}

int64_t b200asr_qwen_kernel_launches(const b200asr_qwen* e) { return e ? e->launches : 0; }
int b200asr_qwen_set_option(b200asr_qwen* e, const char* key, int64_t value) {
  if (!e || !key) return B200ASR_E_INVALID;
  if (!strcmp(key, "graph")) { e->use_graph = value != 0; return B200ASR_OK; }
  if (!strcmp(key, "attn_tc")) { e->use_attn_tc = value != 0; return B200ASR_OK; }
  if (!strcmp(key, "persist")) { e->use_persist = value != 0; e->persist_per_sm = value == 2 ? 1 : 2; return B200ASR_OK; }   // 2: one CTA per SM
  if (!strcmp(key, "persist_timing")) {
    if (value && !e->p_timing) QCK(cudaMalloc(&e->p_timing, 512 * 8));
    if (!value && e->p_timing) { cudaStreamSynchronize(e->st); cudaFree(e->p_timing); e->p_timing = nullptr; }
    return B200ASR_OK;
  }
  if (!strcmp(key, "persist_dbg")) { e->persist_dbg = (int)value; return B200ASR_OK; }                                      // timing experiments
  if (!strcmp(key, "attn_tiled")) { e->use_attn_tiled = value != 0; return B200ASR_OK; }
  if (!strcmp(key, "pdl")) {
    e->use_pdl = value != 0; e->pdl_all = value == 2;
    if (e->step_graph) { cudaGraphExecDestroy(e->step_graph); e->step_graph = nullptr; }
    return B200ASR_OK;
  }
  if (!strcmp(key, "attn_split")) {
    e->use_attn_split = value != 0;
    if (e->step_graph) { cudaGraphExecDestroy(e->step_graph); e->step_graph = nullptr; }
    return B200ASR_OK;
  }
  return e->fail(B200ASR_E_INVALID, std::string("unknown option ") + key);
}
void* b200asr_qwen_stream(b200asr_qwen* e) { return e ? (void*)e->st : nullptr; }

}  // extern "C"
