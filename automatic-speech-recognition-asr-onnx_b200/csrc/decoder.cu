// Decoder-side kernels: weight-streaming skinny linears with fused LayerNorm,
// KV-cache append, causal self-attention over the resident cache, cross-attention
// over the encoder's cross-KV, and the on-device token-selection heads.
//
// Replaces the per-token DECODE_SESSION.run of
// /root/reference/Whisper/Inference_Whisper_ONNX.py:640 whose math is
// WHISPER_DECODER.forward (/root/reference/Whisper/Export_Whisper.py:614-667) plus
// the ARGMAX / GREEDY_SEARCH / APPLY_PENALTY / BEGIN_SUPPRESS heads (:228-331).
// The reference re-concatenates the KV cache every step (:640-641) and syncs to
// the host per token (:645); here the cache is resident ([L][B][H][max_target][64]),
// the step counter lives in device memory (DecState) and the loop never leaves
// the GPU.  These kernels are HBM-bound: one pass over the weights per step.
#include "common.cuh"

namespace b200asr {

// ---------------------------------------------------------------------------
// token + position embedding (Export_Whisper.py:450-458, 477-483, 494-497)
// ---------------------------------------------------------------------------
template <typename WT>
__global__ void dec_embed_kernel(const int* __restrict__ tokens, const WT* __restrict__ embed,
                                 const float* __restrict__ pos, int n_new, int d,
                                 const DecState* __restrict__ state, float* __restrict__ x) {
  const int r = blockIdx.x;
  const int i = r % n_new;
  const int tok = tokens[r];
  const int p = state->kv_len + i;
  for (int c = threadIdx.x; c < d; c += blockDim.x)
    x[(int64_t)r * d + c] = to_f<WT>(embed[(int64_t)tok * d + c]) + pos[(int64_t)p * d + c];
}

cudaError_t launch_dec_embed(const int* tokens, const void* embed, int w_dtype, const void* pos, int batch,
                             int n_new, int d, const DecState* state, float* x, cudaStream_t st) {
  const int rows = batch * n_new;
  if (w_dtype == kF32)
    dec_embed_kernel<float><<<rows, 256, 0, st>>>(tokens, (const float*)embed, (const float*)pos, n_new, d, state, x);
  else
    dec_embed_kernel<bf16><<<rows, 256, 0, st>>>(tokens, (const bf16*)embed, (const float*)pos, n_new, d, state, x);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// skinny linear: out[r][n] = act(LN(x[r]) . W[n] + bias[n]) + residual[r][n]
// one warp per output column (grid-stride), all rows of a chunk share each
// weight load; x rows are staged (and normalised) in shared memory.
// ---------------------------------------------------------------------------
constexpr int kDecThreads = 256;
constexpr int kRowChunk = 4;

template <typename WT> struct WVec;
template <> struct WVec<bf16> {
  static __device__ __forceinline__ void load8(const bf16* p, float* w) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); w[2 * i] = f.x; w[2 * i + 1] = f.y; }
  }
};
template <> struct WVec<float> {
  static __device__ __forceinline__ void load8(const float* p, float* w) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
  }
};

template <typename KT>
__device__ __forceinline__ void kv_store(void* cache, int64_t idx, float v) {
  reinterpret_cast<KT*>(cache)[idx] = from_f<KT>(v);
}

template <typename WT>
__global__ void __launch_bounds__(kDecThreads)
dec_linear_kernel(DecLinearArgs a) {
  extern __shared__ float xs[];   // [kRowChunk][K]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = kDecThreads >> 5;
  const int K = a.K;
  const WT* W = reinterpret_cast<const WT*>(a.W);
  for (int r0 = 0; r0 < a.rows; r0 += kRowChunk) {
    const int nr = min(kRowChunk, a.rows - r0);
    __syncthreads();
    for (int i = threadIdx.x; i < nr * K; i += kDecThreads) {
      const int r = i / K, k = i - r * K;
      xs[r * K + k] = a.x[(int64_t)(r0 + r) * a.ldx + k];
    }
    __syncthreads();
    if (a.ln_mode == 3 && warp < nr) {            // RMS norm without affine (SimplifiedLayerNormalization, Export_Qwen_ASR.py:1043-1077)
      float* xr = xs + warp * K;
      float q = 0.f;
      for (int k = lane; k < K; k += 32) q += xr[k] * xr[k];
      const float rstd = rsqrtf(warp_sum(q) / (float)K + a.eps);
      for (int k = lane; k < K; k += 32) xr[k] *= rstd;
    } else if (a.ln_mode != 0 && warp < nr) {
      float* xr = xs + warp * K;
      float s = 0.f;
      for (int k = lane; k < K; k += 32) s += xr[k];
      const float mean = warp_sum(s) / (float)K;
      float q = 0.f;
      for (int k = lane; k < K; k += 32) { const float t = xr[k] - mean; q += t * t; }
      const float rstd = rsqrtf(warp_sum(q) / (float)K + a.eps);
      for (int k = lane; k < K; k += 32) {
        float y = (xr[k] - mean) * rstd;
        if (a.ln_mode == 2) y = y * a.gamma[k] + a.beta[k];
        xr[k] = y;
      }
    }
    __syncthreads();
    for (int n = blockIdx.x * nwarps + warp; n < a.N; n += gridDim.x * nwarps) {
      float acc[kRowChunk];
#pragma unroll
      for (int r = 0; r < kRowChunk; ++r) acc[r] = 0.f;
      const WT* wr = W + (int64_t)n * K;
#pragma unroll 4
      for (int k = lane * 8; k < K; k += 256) {
        float w[8];
        WVec<WT>::load8(wr + k, w);
#pragma unroll
        for (int r = 0; r < kRowChunk; ++r) {
          if (r < nr) {
            const float4 x0 = *reinterpret_cast<const float4*>(xs + r * K + k);
            const float4 x1 = *reinterpret_cast<const float4*>(xs + r * K + k + 4);
            acc[r] = fmaf(w[0], x0.x, acc[r]); acc[r] = fmaf(w[1], x0.y, acc[r]);
            acc[r] = fmaf(w[2], x0.z, acc[r]); acc[r] = fmaf(w[3], x0.w, acc[r]);
            acc[r] = fmaf(w[4], x1.x, acc[r]); acc[r] = fmaf(w[5], x1.y, acc[r]);
            acc[r] = fmaf(w[6], x1.z, acc[r]); acc[r] = fmaf(w[7], x1.w, acc[r]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < kRowChunk; ++r) acc[r] = warp_sum(acc[r]);
      if (lane < nr) {
        float v = 0.f;
#pragma unroll
        for (int r = 0; r < kRowChunk; ++r) if (lane == r) v = acc[r];
        const int row = r0 + lane;
        if (a.bias) v += a.bias[n];
        if (a.act == kActGelu) v = gelu_erf(v);
        if (a.residual) v += a.residual[(int64_t)row * a.ldr + n];
        if (a.mode == 0) {
          a.out[(int64_t)row * a.ldo + n] = v;
        } else {
          const int d = a.n_heads * a.head_dim;
          if (n < d) {
            a.out[(int64_t)row * a.ldo + n] = v;
          } else {
            const int c = (n - d) % d;
            const int h = c / a.head_dim, dd = c - h * a.head_dim;
            const int b = row / a.n_new, i = row - b * a.n_new;
            const int64_t idx = (((int64_t)b * a.n_heads + h) * a.max_target + (a.state->kv_len + i)) * a.head_dim + dd;
            void* cache = (n < 2 * d) ? a.kcache : a.vcache;
            if (a.kv_dtype == kF32) kv_store<float>(cache, idx, v); else kv_store<bf16>(cache, idx, v);
          }
        }
      }
    }
  }
}

cudaError_t launch_dec_linear(const DecLinearArgs& a, cudaStream_t st) {
  if (a.K % 8 != 0) return cudaErrorInvalidValue;
  const size_t smem = (size_t)kRowChunk * a.K * sizeof(float);
  const int nwarps = kDecThreads / 32;
  int grid = (a.N + nwarps - 1) / nwarps;
  if (grid > 148 * 8) grid = 148 * 8;
  static AttrOnce attr;
  if (a.w_dtype == kF32) {
    if (attr.need(0)) cudaFuncSetAttribute(dec_linear_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    dec_linear_kernel<float><<<grid, kDecThreads, smem, st>>>(a);
  } else {
    if (attr.need(1)) cudaFuncSetAttribute(dec_linear_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    dec_linear_kernel<bf16><<<grid, kDecThreads, smem, st>>>(a);
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// causal self-attention over the resident cache: one warp per (row, head).
// Positions > kv_len+i are excluded, which equals the reference's additive
// -128 mask (Export_Whisper.py:471-474) whenever exp(-128 + delta) underflows.
// ---------------------------------------------------------------------------
template <typename KT> struct KRow;
template <> struct KRow<bf16> {
  static __device__ __forceinline__ float dot64(const bf16* k, const float* q) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 u = *reinterpret_cast<const uint4*>(k + j * 8);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        s = fmaf(f.x, q[j * 8 + 2 * i], s);
        s = fmaf(f.y, q[j * 8 + 2 * i + 1], s);
      }
    }
    return s;
  }
  static __device__ __forceinline__ float2 load2(const bf16* v) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(v));
  }
};
template <> struct KRow<float> {
  static __device__ __forceinline__ float dot64(const float* k, const float* q) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 u = *reinterpret_cast<const float4*>(k + j * 4);
      s = fmaf(u.x, q[j * 4], s); s = fmaf(u.y, q[j * 4 + 1], s);
      s = fmaf(u.z, q[j * 4 + 2], s); s = fmaf(u.w, q[j * 4 + 3], s);
    }
    return s;
  }
  static __device__ __forceinline__ float2 load2(const float* v) { return *reinterpret_cast<const float2*>(v); }
};

constexpr int kSelfWarps = 4;

template <typename KT>
__global__ void __launch_bounds__(kSelfWarps * 32)
dec_self_attn_kernel(const float* __restrict__ q, const KT* __restrict__ kc, const KT* __restrict__ vc,
                     int n_new, int n_heads, int max_target, int total, const DecState* __restrict__ state,
                     float* __restrict__ ctx) {
  extern __shared__ float sm[];            // per warp: q[64] + scores[max_target]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * kSelfWarps + warp;     // (row, head)
  if (item >= total) return;
  const int row = item / n_heads, h = item - row * n_heads;
  const int b = row / n_new, i = row - b * n_new;
  const int d = n_heads * 64;
  float* qs = sm + warp * (64 + max_target);
  float* sc = qs + 64;
  qs[lane] = q[(int64_t)row * d + h * 64 + lane];
  qs[lane + 32] = q[(int64_t)row * d + h * 64 + lane + 32];
  __syncwarp();
  const int kv = state->kv_len + i + 1;
  const KT* kbase = kc + ((int64_t)b * n_heads + h) * max_target * 64;
  const KT* vbase = vc + ((int64_t)b * n_heads + h) * max_target * 64;
  float m = -INFINITY;
  for (int p = lane; p < kv; p += 32) {
    const float s = KRow<KT>::dot64(kbase + (int64_t)p * 64, qs);
    sc[p] = s;
    m = fmaxf(m, s);
  }
  m = warp_max(m);
  float sum = 0.f;
  for (int p = lane; p < kv; p += 32) { const float e = expf(sc[p] - m); sc[p] = e; sum += e; }
  sum = warp_sum(sum);
  __syncwarp();
  float o0 = 0.f, o1 = 0.f;
  for (int p = 0; p < kv; ++p) {
    const float w = sc[p];
    const float2 v = KRow<KT>::load2(vbase + (int64_t)p * 64 + 2 * lane);
    o0 = fmaf(w, v.x, o0); o1 = fmaf(w, v.y, o1);
  }
  const float inv = 1.0f / sum;
  ctx[(int64_t)row * d + h * 64 + 2 * lane] = o0 * inv;
  ctx[(int64_t)row * d + h * 64 + 2 * lane + 1] = o1 * inv;
}

cudaError_t launch_dec_self_attn(const float* q, const void* kcache, const void* vcache, int kv_dtype, int batch,
                                 int n_new, int n_heads, int head_dim, int max_target, const DecState* state,
                                 float* ctx, cudaStream_t st) {
  if (head_dim != 64) return cudaErrorInvalidValue;
  const int total = batch * n_new * n_heads;
  const int grid = (total + kSelfWarps - 1) / kSelfWarps;
  const size_t smem = (size_t)kSelfWarps * (64 + max_target) * sizeof(float);
  if (kv_dtype == kF32)
    dec_self_attn_kernel<float><<<grid, kSelfWarps * 32, smem, st>>>(q, (const float*)kcache, (const float*)vcache,
                                                                     n_new, n_heads, max_target, total, state, ctx);
  else
    dec_self_attn_kernel<bf16><<<grid, kSelfWarps * 32, smem, st>>>(q, (const bf16*)kcache, (const bf16*)vcache,
                                                                    n_new, n_heads, max_target, total, state, ctx);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// cross-attention: one CTA per (row, head) over T encoder positions.
// cross_kv = [2L][B][T][d]: the fused cross-KV projection (Export_Whisper.py:393-447) written
// per (kind, layer) so one layer's K (z = l) and V (z = L + l) are contiguous in HBM.
// ---------------------------------------------------------------------------
constexpr int kCrossThreads = 128;

template <typename KT>
__global__ void __launch_bounds__(kCrossThreads)
dec_cross_attn_kernel(const float* __restrict__ q, const KT* __restrict__ ckv, int layer, int n_layers,
                      int n_new, int n_heads, int T_ld, const int* __restrict__ t_valid, float* __restrict__ ctx) {
  extern __shared__ float sm[];            // q[64] + scores[T] + red[kCrossThreads/32] + part[4][64]
  float* qs = sm;
  float* sc = sm + 64;
  float* red = sc + T_ld;
  float* part = red + 8;
  const int row = blockIdx.x, h = blockIdx.y;
  const int b = row / n_new;
  const int d = n_heads * 64;
  const int64_t ld = d;
  const int B = gridDim.x / n_new;
  const int T = t_valid ? min(T_ld, max(1, t_valid[b])) : T_ld;     // ragged batch: this clip's own encoder positions
  const KT* kb = ckv + (((int64_t)layer * B + b) * T_ld) * d + h * 64;
  const KT* vb = ckv + (((int64_t)(n_layers + layer) * B + b) * T_ld) * d + h * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 64) qs[threadIdx.x] = q[(int64_t)row * d + h * 64 + threadIdx.x];
  __syncthreads();
  float m = -INFINITY;
  for (int t = threadIdx.x; t < T; t += kCrossThreads) {
    const float s = KRow<KT>::dot64(kb + t * ld, qs);
    sc[t] = s;
    m = fmaxf(m, s);
  }
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int t = threadIdx.x; t < T; t += kCrossThreads) { const float e = expf(sc[t] - m); sc[t] = e; sum += e; }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = red[0] + red[1] + red[2] + red[3];
  float o0 = 0.f, o1 = 0.f;
  for (int t = warp; t < T; t += kCrossThreads / 32) {
    const float w = sc[t];
    const float2 v = KRow<KT>::load2(vb + t * ld + 2 * lane);
    o0 = fmaf(w, v.x, o0); o1 = fmaf(w, v.y, o1);
  }
  part[warp * 64 + 2 * lane] = o0;
  part[warp * 64 + 2 * lane + 1] = o1;
  __syncthreads();
  if (threadIdx.x < 64) {
    const float o = part[threadIdx.x] + part[64 + threadIdx.x] + part[128 + threadIdx.x] + part[192 + threadIdx.x];
    ctx[(int64_t)row * d + h * 64 + threadIdx.x] = o / sum;
  }
}

cudaError_t launch_dec_cross_attn(const float* q, const void* cross_kv, int kv_dtype, int layer, int n_layers,
                                  int batch, int n_new, int n_heads, int head_dim, int T, float* ctx,
                                  cudaStream_t st, const int* t_valid) {
  if (head_dim != 64) return cudaErrorInvalidValue;
  dim3 grid(batch * n_new, n_heads);
  const size_t smem = (size_t)(64 + T + 8 + 4 * 64) * sizeof(float);
  if (kv_dtype == kF32)
    dec_cross_attn_kernel<float><<<grid, kCrossThreads, smem, st>>>(q, (const float*)cross_kv, layer, n_layers, n_new,
                                                                    n_heads, T, t_valid, ctx);
  else
    dec_cross_attn_kernel<bf16><<<grid, kCrossThreads, smem, st>>>(q, (const bf16*)cross_kv, layer, n_layers, n_new,
                                                                   n_heads, T, t_valid, ctx);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// token selection: [begin-suppress] -> [sliding-window penalty] -> argmax ->
// history append -> stop latch -> advance DecState.  One CTA per utterance.
// BEGIN_SUPPRESS :228-240, APPLY_PENALTY :318-331, GREEDY_SEARCH/ARGMAX :243-260;
// host bookkeeping mirrored from Inference_Whisper_ONNX.py:584-663.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
select_token_kernel(SelectArgs a) {
  const float* __restrict__ begin_bias = a.begin_bias;
  __shared__ float sv[8];
  __shared__ int si[8];
  const int b = blockIdx.x;
  float* lg = a.logits + (int64_t)b * a.vocab;
  const int gen = a.n_gen[b];
  if (a.penalty_value != 1.0f && begin_bias == nullptr && a.temperature <= 0.f) {
    // merged decode graph: penalty_value input is 1.0 until generated_count >= PENALTY_RANGE (:629-633)
    if (threadIdx.x == 0 && gen >= a.penalty_range) {
      const int ns = a.n_save[b];
      const int first = max(0, ns - a.penalty_range);
      // gather-then-scatter semantics of :329-331: every target reads the ORIGINAL logit
      for (int j = first; j < ns; ++j) {
        const int id = a.save_id[(int64_t)b * a.save_ld + j];
        bool seen = false;
        for (int k = first; k < j; ++k) seen |= (a.save_id[(int64_t)b * a.save_ld + k] == id);
        if (!seen) lg[id] *= a.penalty_value;
      }
    }
    __syncthreads();
  }
  float best = -INFINITY;
  int besti = 0x7fffffff;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (a.temperature > 0.f) {
    // ---- TOPK_TOPP_SAMPLING: repetition penalty on every selected id -> / temperature -> sorted top-k ->
    //      softmax -> keep while (cumsum - p) <= top_p -> Gumbel-max (Export_Whisper.py:281-307) ----
    __shared__ float tk_v[64];
    __shared__ int tk_i[64];
    __shared__ float s_prev_v; __shared__ int s_prev_i;
    const int ns0 = a.n_save[b];
    const float rp = a.rep_penalty, irp = 1.0f / a.rep_penalty;
    // membership of "previously selected" as a bitmap (vocab <= 65536; larger vocabularies search the list)
    __shared__ unsigned hitmap[2048];
    const bool use_map = a.vocab <= 65536;
    if (rp != 1.0f && use_map) {
      for (int i = threadIdx.x; i < 2048; i += blockDim.x) hitmap[i] = 0u;
      __syncthreads();
      for (int j = threadIdx.x; j < ns0; j += blockDim.x) {
        const int id = a.save_id[(int64_t)b * a.save_ld + j];
        atomicOr(&hitmap[id >> 5], 1u << (id & 31));
      }
      __syncthreads();
    }
    auto score = [&](int i) -> float {
      float v = lg[i];
      if (begin_bias) v += begin_bias[i];
      if (rp != 1.0f) {
        bool hit = false;
        if (use_map) hit = (hitmap[i >> 5] >> (i & 31)) & 1u;
        else for (int j = 0; j < ns0; ++j) hit |= (a.save_id[(int64_t)b * a.save_ld + j] == i);
        if (hit) v = v < 0.f ? v * rp : v * irp;
      }
      return v;
    };
    const int K = min(a.top_k, 64);
    if (threadIdx.x == 0) { s_prev_v = INFINITY; s_prev_i = -1; }
    __syncthreads();
    for (int k = 0; k < K; ++k) {
      const float pv = s_prev_v; const int pi = s_prev_i;
      best = -INFINITY; besti = 0x7fffffff;
      for (int i = threadIdx.x; i < a.vocab; i += blockDim.x) {
        const float v = score(i);
        const bool after = (v < pv) || (v == pv && i > pi);      // strictly after the previous pick in (value desc, index asc)
        if (after && (v > best || (v == best && i < besti))) { best = v; besti = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
      }
      if (lane == 0) { sv[warp] = best; si[warp] = besti; }
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
          if (sv[w] > best || (sv[w] == best && si[w] < besti)) { best = sv[w]; besti = si[w]; }
        tk_v[k] = best; tk_i[k] = besti == 0x7fffffff ? 0 : besti;
        s_prev_v = best; s_prev_i = besti;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const float it = 1.0f / a.temperature;
      float mx = tk_v[0] * it, den = 0.f;
      float pr[64];
      for (int k = 0; k < K; ++k) { pr[k] = expf(tk_v[k] * it - mx); den += pr[k]; }
      const int launch = a.state->step;
      float cum = 0.f, gbest = -INFINITY; int win = 0;
      for (int k = 0; k < K; ++k) {
        const float p = pr[k] / den;
        cum += p;
        const bool keep = (cum - p) <= a.top_p;
        float u;
        if (a.noise && launch < a.noise_rows) {
          u = a.noise[((int64_t)launch * a.noise_batch + b) * a.noise_ld + k];   // rows are laid out [launch][max_batch][top_k]
        } else {                                       // counter-based hash (splitmix64) keyed by (seed, launch, b, k)
          unsigned long long z = a.seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(((int64_t)launch * a.batch + b) * 64 + k + 1);
          z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
          u = (float)((z >> 40) + 0.5) * (1.0f / 16777216.0f);
        }
        u = fminf(fmaxf(u, 1.0e-7f), 1.0f - 1.0e-7f);
        const float g = keep ? tk_v[k] * it - logf(-logf(u)) : -INFINITY;
        if (g > gbest) { gbest = g; win = k; }
      }
      best = tk_v[win]; besti = tk_i[win];
    }
  } else {
  for (int i = threadIdx.x; i < a.vocab; i += blockDim.x) {
    float v = lg[i];
    if (begin_bias) v += begin_bias[i];
    if (v > best || (v == best && i < besti)) { best = v; besti = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
  }
  if (lane == 0) { sv[warp] = best; si[warp] = besti; }
  __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (a.temperature <= 0.f)
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (sv[w] > best || (sv[w] == best && si[w] < besti)) { best = sv[w]; besti = si[w]; }
    if (besti == 0x7fffffff) besti = 0;       // all -inf / NaN row: torch.argmax returns 0
    if (a.cand_idx) besti = a.cand_idx[(int64_t)b * a.vocab + besti];
    const int step = a.state->step;
    a.cur_token[b] = besti;
    if (a.selected_hist && step < a.sel_ld) a.selected_hist[(int64_t)b * a.sel_ld + step] = besti;
    const int ns = a.n_save[b];
    if (ns < a.save_ld) { a.save_id[(int64_t)b * a.save_ld + ns] = besti; a.n_save[b] = ns + 1; }
    if (!a.finished[b]) {
      bool stop = false;
      for (int s = 0; s < a.n_stop; ++s) stop |= (a.stop_ids[s] == besti);
      const int limit = a.limit_v ? min(a.limit, a.limit_v[b]) : a.limit;
      if (stop || limit <= 0) {
        a.finished[b] = 1;
      } else {
        a.tokens[(int64_t)b * a.tokens_ld + gen] = besti;
        a.n_gen[b] = gen + 1;
        if (gen + 1 >= limit) a.finished[b] = 1;
      }
    }
  }
}

__global__ void advance_state_kernel(DecState* state, int n_new, const int* finished, int batch) {
  if (threadIdx.x == 0) {
    int done = 1;
    for (int b = 0; b < batch; ++b) done &= (finished[b] != 0);
    state->kv_len += n_new;
    state->step += 1;
    state->all_done = done;
  }
}

cudaError_t launch_select_token(const SelectArgs& a, cudaStream_t st) {
  select_token_kernel<<<a.batch, 256, 0, st>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  advance_state_kernel<<<1, 32, 0, st>>>(a.state, a.n_new, a.finished, a.batch);
  return cudaGetLastError();
}

// out[b] = softmax(logits[b] + add_bias)[index]   (NO_SPEECH_DETECTION, Export_Whisper.py:334-348)
__global__ void __launch_bounds__(256)
softmax_pick_kernel(const float* __restrict__ logits, const float* __restrict__ add_bias, int vocab, int index,
                    float* __restrict__ out) {
  __shared__ float red[8];
  const int b = blockIdx.x;
  const float* lg = logits + (int64_t)b * vocab;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < vocab; i += blockDim.x) m = fmaxf(m, lg[i] + (add_bias ? add_bias[i] : 0.f));
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < vocab; i += blockDim.x) s += expf(lg[i] + (add_bias ? add_bias[i] : 0.f) - m);
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    out[b] = expf(lg[index] + (add_bias ? add_bias[index] : 0.f) - m) / t;
  }
}

cudaError_t launch_softmax_pick(const float* logits, const float* add_bias, int vocab, int batch, int index,
                                float* out, cudaStream_t st) {
  softmax_pick_kernel<<<batch, 256, 0, st>>>(logits, add_bias, vocab, index, out);
  return cudaGetLastError();
}

}  // namespace b200asr
