// Persistent decoder kernel: prefill + the whole greedy loop in ONE cooperative launch.
//
// Replaces the reference's one-graph-launch-per-token loop
// (/root/reference/Whisper/Inference_Whisper_ONNX.py:584-663: rebind 64 KV tensors,
// run DECODE_SESSION, `.numpy()` sync per token) with a single grid of one CTA per SM
// that walks every phase of WHISPER_DECODER.forward
// (/root/reference/Whisper/Export_Whisper.py:614-667) for every token:
//
//   per layer:  LN+QKV(+KV append) | self-attn | out_proj+res | LN+cross-q |
//               cross-attn | out_proj+res | LN+fc1+GELU | fc2+res
//   per token:  LN + tied lm-head + suppress bias -> per-CTA argmax candidates ->
//               grid-wide argmax, stop latch, next-token embedding
//
// Phases are separated by a grid barrier (one L2 atomic + acquire polling, ~0.5 us) instead
// of a kernel boundary (~3-10 us).  The step is HBM-bound on the 1.6 GB bf16 weight set, so
// a prefetch lane per CTA streams the weights (and the layer's cross-KV) into L2 a fixed
// number of bytes ahead of the consuming phase with cp.async.bulk.prefetch.L2; the phases
// themselves then read L2-resident weights and the HBM stream never waits on a barrier.
#include "common.cuh"
#include <cooperative_groups.h>
#include <cstdio>

namespace b200asr {

constexpr int kMegaThreads = 512;
constexpr int kMegaWarps = kMegaThreads / 32;
constexpr int kRMax = 8;                 // activation rows staged per pass
constexpr int kPfPiece = 8192;           // bytes per bulk prefetch
constexpr int kAttnScratch = 2048;       // floats: scores (<= max(T, max_target)) live in the row buffer instead

// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// All CTAs are co-resident (cooperative launch); the counter is monotonic and zeroed by the host.
// Arrive = red.release.gpu after the CTA barrier (cumulative over the CTA's earlier writes), wait =
// ld.acquire.gpu polling by one thread, then a CTA barrier: no full fences on the critical path.
__device__ __forceinline__ void grid_arrive(unsigned* bar, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
  }
}
__device__ __forceinline__ void grid_wait(unsigned* bar, unsigned target) {
  if (threadIdx.x == 0) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    if (v < target) {
      const long long t0 = clock64();
      do {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        if (clock64() - t0 > 8000000000LL) {
          printf("b200asr decoder_mega: grid barrier timed out (block %d)\n", blockIdx.x);
          __trap();
        }
      } while (v < target);
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  __syncthreads();
}

__device__ __forceinline__ void prefetch_l2(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// The per-step read stream (weights in phase order + each layer's cross-KV) is cut into
// super-chunks of gridDim.x * kPfPiece bytes; CTA c owns the c-th piece of every super-chunk and
// the 32 lanes of its prefetch warp each take every 32nd super-chunk, so one `advance` costs a
// handful of instructions per lane.  Positions are unwrapped (they keep growing across tokens).
struct Prefetcher {
  long long done;        // next super-chunk position (unwrapped, padded-stream bytes) this lane will request
  long long blk_start;   // unwrapped start of the cursor block
  int blk;
  __device__ void init(int lane) { done = (long long)lane * gridDim.x * kPfPiece; blk_start = 0; blk = 0; }
  __device__ void advance(const MegaArgs& a, long long until) {
    const long long super = (long long)gridDim.x * kPfPiece;
    while (done < until) {
      while (done >= blk_start + a.pf_blocks[blk].padded) {
        blk_start += a.pf_blocks[blk].padded;
        blk = (blk + 1 == a.n_pf_blocks) ? 0 : blk + 1;
      }
      const PfBlock& pb = a.pf_blocks[blk];
      const long long off = done - blk_start + (long long)blockIdx.x * kPfPiece;
      if (off < pb.bytes) {
        const long long n = min((long long)kPfPiece, pb.bytes - off) & ~15LL;
        if (n > 0) prefetch_l2(pb.ptr + off, (unsigned)n);
      }
      done += 32 * super;
    }
  }
};

template <typename WT> struct MW;
template <> struct MW<bf16> {
  typedef uint4 Raw;
  static constexpr int kUnroll = 5;
  static __device__ __forceinline__ Raw load_raw(const bf16* p) {
    uint4 u;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
    return u;
  }
  static __device__ __forceinline__ float fma8(const Raw& u, const float4& x0, const float4& x1, float acc) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    float2 f = __bfloat1622float2(h[0]); acc = fmaf(f.x, x0.x, acc); acc = fmaf(f.y, x0.y, acc);
    f = __bfloat1622float2(h[1]); acc = fmaf(f.x, x0.z, acc); acc = fmaf(f.y, x0.w, acc);
    f = __bfloat1622float2(h[2]); acc = fmaf(f.x, x1.x, acc); acc = fmaf(f.y, x1.y, acc);
    f = __bfloat1622float2(h[3]); acc = fmaf(f.x, x1.z, acc); acc = fmaf(f.y, x1.w, acc);
    return acc;
  }
  static __device__ __forceinline__ void load8(const bf16* p, float* w) {
    uint4 u;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); w[2 * i] = f.x; w[2 * i + 1] = f.y; }
  }
  static __device__ __forceinline__ float dot64(const bf16* k, const float* q) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 u = __ldcg(reinterpret_cast<const uint4*>(k + j * 8));
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
    const unsigned u = __ldcg(reinterpret_cast<const unsigned*>(v));
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
  }
  static __device__ __forceinline__ float load1(const bf16* v) { return __bfloat162float(*v); }
  static __device__ __forceinline__ void store1(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
};
template <> struct MW<float> {
  struct Raw { float4 a, b; };
  static constexpr int kUnroll = 2;
  static __device__ __forceinline__ Raw load_raw(const float* p) {
    Raw r;
    r.a = __ldg(reinterpret_cast<const float4*>(p));
    r.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    return r;
  }
  static __device__ __forceinline__ float fma8(const Raw& u, const float4& x0, const float4& x1, float acc) {
    acc = fmaf(u.a.x, x0.x, acc); acc = fmaf(u.a.y, x0.y, acc); acc = fmaf(u.a.z, x0.z, acc); acc = fmaf(u.a.w, x0.w, acc);
    acc = fmaf(u.b.x, x1.x, acc); acc = fmaf(u.b.y, x1.y, acc); acc = fmaf(u.b.z, x1.z, acc); acc = fmaf(u.b.w, x1.w, acc);
    return acc;
  }
  static __device__ __forceinline__ void load8(const float* p, float* w) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
  }
  static __device__ __forceinline__ float dot64(const float* k, const float* q) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 u = __ldcg(reinterpret_cast<const float4*>(k + j * 4));
      s = fmaf(u.x, q[j * 4], s); s = fmaf(u.y, q[j * 4 + 1], s);
      s = fmaf(u.z, q[j * 4 + 2], s); s = fmaf(u.w, q[j * 4 + 3], s);
    }
    return s;
  }
  static __device__ __forceinline__ float2 load2(const float* v) { return __ldcg(reinterpret_cast<const float2*>(v)); }
  static __device__ __forceinline__ float load1(const float* v) { return *v; }
  static __device__ __forceinline__ void store1(float* p, float v) { *p = v; }
};

// weights (first U chunks) + bias of this warp's first column of the NEXT linear phase, fetched
// between barrier-arrive and barrier-wait so the HBM latency hides behind the barrier
template <typename WT>
struct Pre {
  typename MW<WT>::Raw w[MW<WT>::kUnroll];
  float bias;
  int valid;
};

// per-iteration geometry shared by the phase functions
struct Iter {
  int n_new, rows, kv_len;
  const int* tokens;       // [B][n_new]
};

enum InMode { kInRows = 0, kInEmbed = 1 };
enum OutMode { kOutStore = 0, kOutAccum = 1, kOutQkv = 2, kOutArgmax = 3 };

struct Lin {
  const float* in; long long ld_in;      // kInRows: fp32 rows
  int in_mode; int ln_mode; const float* gamma; const float* beta;
  const void* W; const float* bias; int N, K; int act;
  float* out; long long ld_out; int out_mode;
  int rows;
  int layer;                             // kOutQkv: cache layer
};

// columns n = gw, gw + G, ... of one pass over NR staged rows (rows beyond the pass read zero-filled smem)
template <typename WT, int NR>
__device__ __forceinline__ void linear_columns(const MegaArgs& a, const Iter& it, const Lin& L, const float* xs, int r0,
                                               float* best_v, int* best_i, bool begin_bias_on, bool penalty_on,
                                               const int* pen_ids, int pen_n, const Pre<WT>& pre_in) {
  Pre<WT> pre = pre_in;
  if (r0 != 0) pre.valid = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kMegaWarps + warp;
  const int G = gridDim.x * kMegaWarps;
  const int K = L.K;
  const int nr = min(NR, L.rows - r0);
  const WT* W = reinterpret_cast<const WT*>(L.W);
  constexpr int U = MW<WT>::kUnroll;
  const int nchunk = K >> 8;                 // K % 256 == 0 (checked on the host)
  bool have = pre.valid != 0;
  typename MW<WT>::Raw w[U];
#pragma unroll
  for (int u = 0; u < U; ++u) w[u] = pre.w[u];
  for (int n = gw; n < L.N; n += G) {
    float acc[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) acc[r] = 0.f;
    const WT* wr = W + (long long)n * K + lane * 8;
    // issue the epilogue's loads now so they overlap the dot product
    float bias_v = 0.f, old_v = 0.f;
    if (lane < nr) {
      if (L.bias) bias_v = have ? pre.bias : L.bias[n];
      if (L.out_mode == kOutAccum) old_v = __ldcg(L.out + (long long)(r0 + lane) * L.ld_out + n);
    }
    for (int c0 = 0; c0 < nchunk; c0 += U) {
      if (!have) {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (c0 + u < nchunk) w[u] = MW<WT>::load_raw(wr + (c0 + u) * 256);
      }
      have = false;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (c0 + u < nchunk) {
          const float* xb = xs + (c0 + u) * 256 + lane * 8;
#pragma unroll
          for (int r = 0; r < NR; ++r) {
            const float4 x0 = *reinterpret_cast<const float4*>(xb + r * K);
            const float4 x1 = *reinterpret_cast<const float4*>(xb + r * K + 4);
            acc[r] = MW<WT>::fma8(w[u], x0, x1, acc[r]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) acc[r] = warp_sum(acc[r]);
    if (lane < nr) {
      float v = 0.f;
#pragma unroll
      for (int r = 0; r < NR; ++r) if (lane == r) v = acc[r];
      const int row = r0 + lane;
      v += bias_v;
      if (L.act == kActGelu) v = gelu_erf(v);
      if (L.out_mode == kOutStore) {
        L.out[(long long)row * L.ld_out + n] = v;
      } else if (L.out_mode == kOutAccum) {
        L.out[(long long)row * L.ld_out + n] = old_v + v;
      } else if (L.out_mode == kOutQkv) {
        const int d = a.d;
        if (n < d) {
          L.out[(long long)row * L.ld_out + n] = v;
        } else {
          const int c = (n - d) % d;
          const int h = c >> 6, dd = c & 63;
          const int b = row / it.n_new, i = row - b * it.n_new;
          const long long idx = ((((long long)L.layer * a.batch + b) * a.n_heads + h) * a.max_target + (it.kv_len + i)) * 64 + dd;
          MW<WT>::store1(reinterpret_cast<WT*>(n < 2 * d ? a.kcache : a.vcache) + idx, v);
        }
      } else {   // kOutArgmax: row = utterance index (host keeps batch <= kRMax, so there is a single pass)
        if (a.logits) a.logits[(long long)row * a.vocab + n] = v;
        float hv = v;
        if (penalty_on) {
          bool hit = false;
          for (int j = 0; j < pen_n; ++j) hit |= (pen_ids[row * 32 + j] == n);
          if (hit) { hv = v * a.penalty_value; if (a.logits) a.logits[(long long)row * a.vocab + n] = hv; }
        }
        if (begin_bias_on) hv += a.begin_bias[n];
        const int slot = warp * kRMax + lane;
        if (hv > best_v[slot] || (hv == best_v[slot] && n < best_i[slot])) { best_v[slot] = hv; best_i[slot] = n; }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// one skinny-linear phase: every warp of the grid owns output columns n = gw, gw + G, ...
// ---------------------------------------------------------------------------
template <typename WT>
__device__ __forceinline__ void linear_phase(const MegaArgs& a, const Iter& it, const Lin& L, float* xs, float* red,
                                             float* best_v, int* best_i, bool begin_bias_on, bool penalty_on,
                                             const int* pen_ids, int pen_n, const Pre<WT>& pre) {
  // red: [2][kMegaWarps][kRMax] cross-warp LayerNorm partials
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = L.K, K4 = K >> 2;
  for (int r0 = 0; r0 < L.rows; r0 += kRMax) {
    const int nr = min(kRMax, L.rows - r0);
    const int nr_pad = nr <= 2 ? nr : (nr <= 4 ? 4 : kRMax);     // NR the column loop is instantiated for
    __syncthreads();
    // ---- stage rows: every thread moves float4s (all loads in flight at once), LN partial sums on the fly ----
    for (int r = 0; r < nr_pad; ++r) {
      float part = 0.f;
      float* xr = xs + r * K;
      if (r >= nr) {
        for (int k4 = threadIdx.x; k4 < K4; k4 += kMegaThreads) *reinterpret_cast<float4*>(xr + 4 * k4) = make_float4(0.f, 0.f, 0.f, 0.f);
      } else if (L.in_mode == kInEmbed) {
        const int row = r0 + r;
        const int b = row / it.n_new, i = row - b * it.n_new;
        const WT* er = reinterpret_cast<const WT*>(a.embed) + (long long)it.tokens[row] * K;
        const float* pr = a.pos + (long long)(it.kv_len + i) * K;
        for (int k4 = threadIdx.x; k4 < K4; k4 += kMegaThreads) {
          float4 v = *reinterpret_cast<const float4*>(pr + 4 * k4);
          v.x += MW<WT>::load1(er + 4 * k4); v.y += MW<WT>::load1(er + 4 * k4 + 1);
          v.z += MW<WT>::load1(er + 4 * k4 + 2); v.w += MW<WT>::load1(er + 4 * k4 + 3);
          *reinterpret_cast<float4*>(xr + 4 * k4) = v;
          if (blockIdx.x == 0) *reinterpret_cast<float4*>(a.x + (long long)row * K + 4 * k4) = v;   // residual stream copy
          part += (v.x + v.y) + (v.z + v.w);
        }
      } else {
        const float* src = L.in + (long long)(r0 + r) * L.ld_in;
        for (int k4 = threadIdx.x; k4 < K4; k4 += kMegaThreads) {
          const float4 v = __ldcg(reinterpret_cast<const float4*>(src + 4 * k4));
          *reinterpret_cast<float4*>(xr + 4 * k4) = v;
          part += (v.x + v.y) + (v.z + v.w);
        }
      }
      if (L.ln_mode != 0 && r < nr) {
        part = warp_sum(part);
        if (lane == 0) red[warp * kRMax + r] = part;
      }
    }
    __syncthreads();
    if (L.ln_mode != 0) {
      // two-pass LayerNorm: each thread revisits exactly the float4s it staged
      for (int r = 0; r < nr; ++r) {
        float mean = 0.f;
#pragma unroll
        for (int w = 0; w < kMegaWarps; ++w) mean += red[w * kRMax + r];
        mean /= (float)K;
        float qv = 0.f;
        const float* xr = xs + r * K;
        for (int k4 = threadIdx.x; k4 < K4; k4 += kMegaThreads) {
          const float4 v = *reinterpret_cast<const float4*>(xr + 4 * k4);
          const float t0 = v.x - mean, t1 = v.y - mean, t2 = v.z - mean, t3 = v.w - mean;
          qv += (t0 * t0 + t1 * t1) + (t2 * t2 + t3 * t3);
        }
        qv = warp_sum(qv);
        if (lane == 0) red[(kMegaWarps + warp) * kRMax + r] = qv;
      }
      __syncthreads();
      for (int r = 0; r < nr; ++r) {
        float mean = 0.f, var = 0.f;
#pragma unroll
        for (int w = 0; w < kMegaWarps; ++w) { mean += red[w * kRMax + r]; var += red[(kMegaWarps + w) * kRMax + r]; }
        mean /= (float)K;
        const float rstd = rsqrtf(var / (float)K + a.eps);
        float* xr = xs + r * K;
        for (int k4 = threadIdx.x; k4 < K4; k4 += kMegaThreads) {
          float4 v = *reinterpret_cast<float4*>(xr + 4 * k4);
          v.x = (v.x - mean) * rstd; v.y = (v.y - mean) * rstd; v.z = (v.z - mean) * rstd; v.w = (v.w - mean) * rstd;
          if (L.ln_mode == 2) {
            const float4 g = *reinterpret_cast<const float4*>(L.gamma + 4 * k4);
            const float4 bt = *reinterpret_cast<const float4*>(L.beta + 4 * k4);
            v.x = v.x * g.x + bt.x; v.y = v.y * g.y + bt.y; v.z = v.z * g.z + bt.z; v.w = v.w * g.w + bt.w;
          }
          *reinterpret_cast<float4*>(xr + 4 * k4) = v;
        }
      }
      __syncthreads();
    }
    // ---- columns ----
    switch (nr_pad) {
      case 1: linear_columns<WT, 1>(a, it, L, xs, r0, best_v, best_i, begin_bias_on, penalty_on, pen_ids, pen_n, pre); break;
      case 2: linear_columns<WT, 2>(a, it, L, xs, r0, best_v, best_i, begin_bias_on, penalty_on, pen_ids, pen_n, pre); break;
      case 4: linear_columns<WT, 4>(a, it, L, xs, r0, best_v, best_i, begin_bias_on, penalty_on, pen_ids, pen_n, pre); break;
      default: linear_columns<WT, 8>(a, it, L, xs, r0, best_v, best_i, begin_bias_on, penalty_on, pen_ids, pen_n, pre); break;
    }
  }
}

// ---------------------------------------------------------------------------
// attention phases: one CTA per (row, head)
// ---------------------------------------------------------------------------
template <typename WT>
__device__ void attn_task(const float* qrow, const WT* kbase, const WT* vbase, long long kv_stride, int npos,
                          float* out, float* sm) {
  // sm: q[64] | red[32] | part[16][64] | scores[npos]
  float* qs = sm; float* red = sm + 64; float* part = sm + 96; float* sc = sm + 96 + kMegaWarps * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (threadIdx.x < 64) qs[threadIdx.x] = __ldcg(qrow + threadIdx.x);
  __syncthreads();
  float m = -INFINITY;
  for (int p = threadIdx.x; p < npos; p += kMegaThreads) {
    const float s = MW<WT>::dot64(kbase + (long long)p * kv_stride, qs);
    sc[p] = s;
    m = fmaxf(m, s);
  }
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int w = 1; w < kMegaWarps; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int p = threadIdx.x; p < npos; p += kMegaThreads) { const float e = expf(sc[p] - m); sc[p] = e; sum += e; }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < kMegaWarps; ++w) sum += red[w];
  float o0 = 0.f, o1 = 0.f;
  {
    int p = warp;
    for (; p + 3 * kMegaWarps < npos; p += 4 * kMegaWarps) {
      float2 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = MW<WT>::load2(vbase + (long long)(p + j * kMegaWarps) * kv_stride + 2 * lane);
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float w = sc[p + j * kMegaWarps]; o0 = fmaf(w, v[j].x, o0); o1 = fmaf(w, v[j].y, o1); }
    }
    for (; p < npos; p += kMegaWarps) {
      const float w = sc[p];
      const float2 v = MW<WT>::load2(vbase + (long long)p * kv_stride + 2 * lane);
      o0 = fmaf(w, v.x, o0); o1 = fmaf(w, v.y, o1);
    }
  }
  part[warp * 64 + 2 * lane] = o0;
  part[warp * 64 + 2 * lane + 1] = o1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float o = 0.f;
#pragma unroll
    for (int w = 0; w < kMegaWarps; ++w) o += part[w * 64 + threadIdx.x];
    out[threadIdx.x] = o / sum;
  }
}

template <typename WT>
__device__ void self_attn_phase(const MegaArgs& a, const Iter& it, int layer, float* sm) {
  const int H = a.n_heads;
  for (int task = blockIdx.x; task < it.rows * H; task += gridDim.x) {
    const int row = task / H, h = task - row * H;
    const int b = row / it.n_new, i = row - b * it.n_new;
    const long long base = ((((long long)layer * a.batch + b) * H + h) * a.max_target) * 64;
    attn_task<WT>(a.q + (long long)row * a.d + h * 64, reinterpret_cast<const WT*>(a.kcache) + base,
                  reinterpret_cast<const WT*>(a.vcache) + base, 64, it.kv_len + i + 1,
                  a.ctx + (long long)row * a.d + h * 64, sm);
  }
}

template <typename WT>
__device__ void cross_attn_phase(const MegaArgs& a, const Iter& it, int layer, float* sm) {
  const int H = a.n_heads;
  for (int task = blockIdx.x; task < it.rows * H; task += gridDim.x) {
    const int row = task / H, h = task - row * H;
    const int b = row / it.n_new;
    const WT* ck = reinterpret_cast<const WT*>(a.cross_kv);
    const WT* kb = ck + (((long long)layer * a.batch + b) * a.T) * a.d + h * 64;
    const WT* vb = ck + (((long long)(a.n_layers + layer) * a.batch + b) * a.T) * a.d + h * 64;
    attn_task<WT>(a.q + (long long)row * a.d + h * 64, kb, vb, a.d, a.t_valid ? min(a.T, max(1, a.t_valid[b])) : a.T,
                  a.ctx + (long long)row * a.d + h * 64, sm);
  }
}

// ---------------------------------------------------------------------------
// linear phase `ph` (0 qkv, 2 out, 3 cq, 5 cout, 6 fc1, 7 fc2) of layer l, or the lm head (ph = 8)
__device__ __forceinline__ Lin make_lin(const MegaArgs& a, const Iter& it, int l, int ph) {
  const int d = a.d;
  if (ph == 8)
    return Lin{a.x + (long long)(it.n_new - 1) * d, (long long)it.n_new * d, kInRows, 2, a.ln_g, a.ln_b, a.embed,
               a.suppress_bias, a.vocab, d, kActNone, nullptr, 0, kOutArgmax, a.batch, 0};
  const MegaLayer& w = a.layers[l];
  switch (ph) {
    case 0: return Lin{a.x, d, (l == 0) ? kInEmbed : kInRows, 1, nullptr, nullptr, w.qkv_w, w.qkv_b, 3 * d, d, kActNone,
                       a.q, d, kOutQkv, it.rows, l};
    case 2: return Lin{a.ctx, d, kInRows, 0, nullptr, nullptr, w.out_w, w.out_b, d, d, kActNone, a.x, d, kOutAccum, it.rows, l};
    case 3: return Lin{a.x, d, kInRows, 1, nullptr, nullptr, w.cq_w, w.cq_b, d, d, kActNone, a.q, d, kOutStore, it.rows, l};
    case 5: return Lin{a.ctx, d, kInRows, 0, nullptr, nullptr, w.cout_w, w.cout_b, d, d, kActNone, a.x, d, kOutAccum, it.rows, l};
    case 6: return Lin{a.x, d, kInRows, 1, nullptr, nullptr, w.fc1_w, w.fc1_b, a.ffn, d, kActGelu, a.f, a.ffn, kOutStore, it.rows, l};
    default: return Lin{a.f, a.ffn, kInRows, 0, nullptr, nullptr, w.fc2_w, w.fc2_b, d, a.ffn, kActNone, a.x, d, kOutAccum, it.rows, l};
  }
}

template <typename WT>
__global__ void __launch_bounds__(kMegaThreads, 1)
decoder_mega_kernel(const __grid_constant__ MegaArgs a) {
  extern __shared__ float smem[];
  float* xs = smem;                               // [kRMax][max(d, ffn)]  (also attention scratch)
  __shared__ float best_v[kMegaWarps * kRMax];
  __shared__ int best_i[kMegaWarps * kRMax];
  __shared__ int s_tok[kRMax], s_ngen[kRMax], s_fin[kRMax], s_nsave[kRMax];
  __shared__ int s_pen[kRMax * 32];
  __shared__ int s_pen_n, s_pen_on, s_all_done;
  __shared__ float s_red[2 * kMegaWarps * kRMax];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int B = a.batch;
  unsigned bar_target = 0;
  int t_idx = 0;
  auto stamp = [&]() {
    if (a.timing && blockIdx.x == 0 && threadIdx.x == 0 && t_idx < a.timing_cap) {
      unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      a.timing[t_idx++] = t;
    }
  };
  stamp();
  Prefetcher pf; pf.init(lane);
  const bool pf_warp = (warp == kMegaWarps - 1) && a.pf_total > 0;
  long long pf_base = 0;                          // unwrapped stream position of this iteration's block 0

  if (threadIdx.x < B) {
    s_ngen[threadIdx.x] = a.n_gen[threadIdx.x];
    s_fin[threadIdx.x] = a.finished[threadIdx.x];
    s_nsave[threadIdx.x] = a.n_save[threadIdx.x];
  }
  int kv_len = a.state->kv_len;
  int step = a.state->step;
  __syncthreads();
  const int n_phases = 8 * a.n_layers + 1;
  Pre<WT> pre;
  pre.valid = 0; pre.bias = 0.f;

  for (int iter = 0; iter < a.n_iters; ++iter) {
    if (threadIdx.x == 0) {
      int done = 1;
      for (int b = 0; b < B; ++b) done &= (s_fin[b] != 0);
      s_all_done = done;
    }
    __syncthreads();
    if (s_all_done && !(iter == 0 && a.first_is_prefill)) break;    // uniform across the grid

    Iter it;
    it.n_new = (iter == 0) ? a.first_n_new : 1;
    it.rows = B * it.n_new;
    it.kv_len = kv_len;
    it.tokens = (iter == 0) ? a.first_tokens : s_tok;     // later tokens: the CTA-local argmax result (no extra barrier)
    const bool begin_on = (iter == 0) && a.first_is_prefill && a.begin_bias != nullptr;
    int pf_blk = 0;                                        // next stream block the phases will consume

    for (int idx = 0; idx < n_phases; ++idx) {
      const int l = idx >> 3;
      const int ph = (idx == n_phases - 1) ? 8 : (idx & 7);
      // stream blocks this phase reads: one weight matrix, or the layer's cross K and V
      pf_blk += (ph == 4) ? 2 : (ph == 1 ? 0 : 1);
      if (pf_warp) {
        const long long c = pf_base + (pf_blk < a.n_pf_blocks ? a.pf_blocks[pf_blk].start : a.pf_total);
        pf.advance(a, c + a.pf_ahead);
      }
      if (ph == 1) {
        self_attn_phase<WT>(a, it, l, xs);
      } else if (ph == 4) {
        cross_attn_phase<WT>(a, it, l, xs);
      } else {
        bool pen_on = false;
        if (ph == 8) {
          // sliding-window penalty ids (APPLY_PENALTY): active once generated_count >= penalty_range, decode launches only
          if (threadIdx.x == 0) {
            const bool on = (a.penalty_value != 1.0f) && !begin_on;
            s_pen_on = 0; s_pen_n = 0;
            if (on) {
              int nmax = 0;
              for (int b = 0; b < B; ++b) {
                const bool act = s_ngen[b] >= a.penalty_range;
                const int ns = s_nsave[b];
                const int first = max(0, ns - a.penalty_range);
                int cnt = 0;
                if (act) for (int j = first; j < ns && cnt < 32; ++j) s_pen[b * 32 + cnt++] = a.save_id[(long long)b * a.save_ld + j];
                for (int j = cnt; j < 32; ++j) s_pen[b * 32 + j] = -1;
                nmax = max(nmax, cnt);
              }
              s_pen_n = nmax; s_pen_on = nmax > 0;
            }
          }
          for (int i = threadIdx.x; i < kMegaWarps * kRMax; i += kMegaThreads) { best_v[i] = -INFINITY; best_i[i] = 0x7fffffff; }
          __syncthreads();
          pen_on = s_pen_on != 0;
        }
        const Lin p = make_lin(a, it, l, ph);
        linear_phase<WT>(a, it, p, xs, s_red, best_v, best_i, begin_on && ph == 8, pen_on, s_pen, s_pen_n, pre);
        pre.valid = 0;
        if (ph == 8) {
          __syncthreads();
          if (threadIdx.x < B) {
            float bv = -INFINITY; int bi = 0x7fffffff;
            for (int w = 0; w < kMegaWarps; ++w) {
              const float v = best_v[w * kRMax + threadIdx.x]; const int i = best_i[w * kRMax + threadIdx.x];
              if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
            }
            a.cand_val[(long long)blockIdx.x * B + threadIdx.x] = bv;
            a.cand_idx[(long long)blockIdx.x * B + threadIdx.x] = bi;
          }
        }
      }
      grid_arrive(a.bar, bar_target);
      if (ph != 1 && ph != 4) {
        // next linear phase (skipping an attention phase; wrapping into the next token's layer 0)
        int nidx = idx + 1;
        if (nidx < n_phases && ((nidx & 7) == 1 || (nidx & 7) == 4) && nidx != n_phases - 1) ++nidx;
        bool ok = true;
        if (nidx >= n_phases) { nidx = 0; ok = (iter + 1 < a.n_iters); }
        if (ok) {
          const int l2 = nidx >> 3;
          const int ph2 = (nidx == n_phases - 1) ? 8 : (nidx & 7);
          const Lin q = make_lin(a, it, l2, ph2);
          const int gw = blockIdx.x * kMegaWarps + warp;
          if (gw < q.N) {
            const WT* wr = reinterpret_cast<const WT*>(q.W) + (long long)gw * q.K + lane * 8;
            const int nchunk = q.K >> 8;
#pragma unroll
            for (int u = 0; u < MW<WT>::kUnroll; ++u)
              if (u < nchunk) pre.w[u] = MW<WT>::load_raw(wr + u * 256);
            pre.bias = q.bias ? q.bias[gw] : 0.f;
            pre.valid = 1;
          }
        }
      }
      grid_wait(a.bar, bar_target); stamp();
    }

    // ---- every CTA reduces the per-CTA argmax candidates identically (no extra barrier) ----
    if (warp < B) {
      float bv = -INFINITY; int bi = 0x7fffffff;
      for (int c = lane; c < (int)gridDim.x; c += 32) {
        const float v = __ldcg(a.cand_val + (long long)c * B + warp); const int i = __ldcg(a.cand_idx + (long long)c * B + warp);
        if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) {
        if (bi == 0x7fffffff) bi = 0;
        const int b = warp;
        s_tok[b] = bi;
        const int gen = s_ngen[b];
        const int ns = s_nsave[b];
        const bool g0 = blockIdx.x == 0;
        if (g0) {
          a.cur_token[b] = bi;
          if (step < a.sel_ld) a.selected_hist[(long long)b * a.sel_ld + step] = bi;
          if (ns < a.save_ld) a.save_id[(long long)b * a.save_ld + ns] = bi;
        }
        if (ns < a.save_ld) s_nsave[b] = ns + 1;
        if (!s_fin[b]) {
          bool stop = false;
          for (int s = 0; s < a.n_stop; ++s) stop |= (a.stop_ids[s] == bi);
          if (stop || a.limit <= 0) {
            s_fin[b] = 1;
          } else {
            if (g0) a.tokens[(long long)b * a.tokens_ld + gen] = bi;
            s_ngen[b] = gen + 1;
            if (gen + 1 >= a.limit) s_fin[b] = 1;
          }
        }
      }
    }
    __syncthreads();
    kv_len += it.n_new;
    step += 1;
    pf_base += a.pf_total;
    // the candidate buffers are rewritten only after 8*L more barriers: no hazard with slow readers
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x < B) {
    a.n_gen[threadIdx.x] = s_ngen[threadIdx.x];
    a.finished[threadIdx.x] = s_fin[threadIdx.x];
    a.n_save[threadIdx.x] = s_nsave[threadIdx.x];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int done = 1;
    for (int b = 0; b < B; ++b) done &= (s_fin[b] != 0);
    a.state->kv_len = kv_len; a.state->step = step; a.state->all_done = done;
  }
}

// ---------------------------------------------------------------------------
size_t mega_smem_bytes(int d, int ffn, int T, int max_target) {
  size_t rows = (size_t)kRMax * (size_t)(d > ffn ? d : ffn);
  size_t attn = 96 + kMegaWarps * 64 + (size_t)(T > max_target ? T : max_target) + 64;
  return (rows > attn ? rows : attn) * sizeof(float);
}

int mega_pf_piece() { return kPfPiece; }

bool mega_supported(int batch, int first_n_new, int d, int ffn) {
  return batch >= 1 && batch <= kRMax && (d % 256 == 0) && (ffn % 256 == 0) && first_n_new >= 0;
}

cudaError_t launch_decoder_mega(const MegaArgs& a, int w_dtype, int num_sms, cudaStream_t st) {
  const size_t smem = mega_smem_bytes(a.d, a.ffn, a.T, a.max_target);
  void* fn = w_dtype == kF32 ? (void*)decoder_mega_kernel<float> : (void*)decoder_mega_kernel<bf16>;
  static AttrOnce attr;
  if (attr.need(w_dtype == kF32 ? 0 : 1)) {
    cudaError_t r = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (r != cudaSuccess) return r;
  }
  if (smem > 220 * 1024) return cudaErrorInvalidValue;
  MegaArgs args = a;
  void* params[] = {&args};
  return cudaLaunchCooperativeKernel(fn, dim3(num_sms), dim3(kMegaThreads), params, smem, st);
}

}  // namespace b200asr
