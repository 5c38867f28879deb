// Qwen3-ASR decode step, the 5 x n_layers decoder-layer launches as ONE cooperative kernel (included by qwen.cu).
//
// A decode step of QWEN3_ASR_DECODER_MAIN.forward (/root/reference/Qwen_ASR/Export_Qwen_ASR.py:1265-1336) is, per layer,
// RMS norm + fused QKV | QK-norm + RoPE + cache append + attention | o_proj + residual | RMS norm + gate_up | SwiGLU + down +
// residual: five dependent memory-bound phases of 1-2 us of weight traffic each.  As separate launches (CUDA graph, qwen.cu) a
// phase costs 8-10 us at 4 clips, mostly launch ramp, drain and an exposed HBM round trip.  Here 2 CTAs per SM stay resident
// for the whole step; a phase boundary is one grid barrier (red + poll on a monotonic counter), and every CTA requests its
// weight rows (raw registers, the same layout as qwen_gemv_kernel) or its cache rows BEFORE it waits on the barrier, so the
// HBM round trip of phase p+1 overlaps the tail of phase p.  The arithmetic of a phase is that of qwen_gemv_kernel /
// qwen_attn_split_kernel<DH, 256, false> instruction for instruction (same k order, same reductions): the step's logits are
// bit-identical to the per-launch path whenever that path uses the same attention split (3-4 clips).
//
// Measured on B200 (Qwen3-ASR-0.6B, tools/qwen_persist_probe.py, tools/qwen_persist_timing.py): a grid barrier alone costs 1.3 us,
// but a layer takes 50 us here against 27 us for the five programmatic-dependent launches at one clip (1.55 vs 0.84 ms per step;
// 1.97 vs 1.52 at four clips): per phase 2-4 us of barrier + arrival skew, 1.7 us of activation staging, 2 us of math, and a
// 11-12 us attention phase (V rows are not prefetched at 128 registers).  The launch-per-phase path already overlaps the next
// phase's weight requests with the previous phase's tail, so a barrier per phase buys nothing: the kernel is kept behind the
// "persist" option (default off) as the measured baseline for a flag-based design with fewer synchronisation points.
//
// Activations cross CTAs only through global memory and only across a grid barrier; they are read with ld.global.cg, which is
// what makes the barrier's missing acquire fence sound on this hardware (see qp_arrive).

constexpr int kPersistThreads = 256;
constexpr int kPersistSK = 256;          // keys per attention task

struct QPersistLayer {
  const bf16 *qkv_w, *o_w, *gu_w, *down_w;
  const float* qk_g;                     // [2][DH] QK-norm weights (d^-0.25 folded)
};
struct QPersistArgs {
  const QPersistLayer* layers; int n_layers;
  float *x, *qkvf, *actx, *gu;           // [rows][hidden] | [rows][(H+2KH)*DH] | [rows][H*DH] | [rows][2*inter]
  int rows, hidden, heads, kv_heads, inter, max_seq;
  float eps;
  const float *cosT, *sinT;
  bf16 *kc, *vc; long long cache_layer_stride;
  const DecState* state; const int* kv_off;
  float* att_part; int* att_counter; int S;      // S key ranges per (clip, head)
  unsigned long long* timing; int timing_cap;   // optional: block 0 stamps %globaltimer at entry / after wait / after staging / after math of every phase
  int dbg;                                       // timing experiments only: 1 = linear phases are barriers only, 2 = attention phase is a barrier only
  unsigned* bar; unsigned bar_base;              // monotonic arrival counter; this launch's barriers complete at bar_base + k * gridDim.x
};

__device__ __forceinline__ void qp_stamp(unsigned long long* tm, int cap, int& ti) {
  if (tm && blockIdx.x == 0 && threadIdx.x == 0 && ti < cap) {
    unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    tm[ti++] = t;
  }
}
__device__ __forceinline__ void qp_arrive(unsigned* bar) {
  __syncthreads();
  // release only: the CTA's stores (ordered before thread 0 by the CTA barrier) are visible at L2 before the arrival counts.  No
  // acquire fence anywhere: on sm_100 it is MEMBAR + CCTL.IVALL, and an L1 invalidate in the middle of the next phase's
  // in-flight weight requests cost ~9 us per phase (measured with the stamps below).  Consumers read everything another CTA wrote
  // with ld.global.cg (L2 is the coherence point); L1 only ever holds weights, tables and this thread's own spills.
  if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
}
__device__ __forceinline__ void qp_wait(unsigned* bar, unsigned target) {
  if (threadIdx.x == 0) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    if ((int)(v - target) < 0) {
      const long long t0 = clock64();
      do {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        if (clock64() - t0 > 4000000000LL) __trap();      // ~2 s: a lost arrival fails the launch instead of hanging the GPU (no printf: a call here spills every live weight register)
      } while ((int)(v - target) < 0);
    }
  }
  __syncthreads();
}

// ---- one weight-streaming phase: out[r][n] = (rms ? rstd[r] : 1) * sum_k W[n][k] * x'[r][k] (+ residual); rows <= 4.
//      The weight registers of the CTA's first column pass are requested, THEN the barrier of the previous phase is awaited ----
template <int KS>
__device__ __forceinline__ void qp_gemv(const float* x, long long ldx, int rms, int swiglu, float eps, const bf16* W, const float* residual,
                                     long long ldr, float* out, long long ldo, int rows, int N, int K, float* gx, float (*red)[kGemvRows],
                                     float (*psum)[2][kGemvRows], unsigned* bar, unsigned wait_target, int do_wait, int dbg,
                                     unsigned long long* tm, int tcap, int& ti) {
  qp_stamp(tm, tcap, ti);
  if (dbg & 1) { if (do_wait) qp_wait(bar, wait_target); qp_arrive(bar); return; }
  constexpr int NC = 2;
  constexpr int CPB = (8 / KS) * NC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ks = warp % KS, cw = warp / KS;
  const int k0 = (ks * 32 + lane) * 8, kstep = KS * 256;
  const int n_pass = (N + CPB - 1) / CPB;
  WRaw<bf16> cur[NC][kGemvCH];      // no second buffer: a spilled weight register is reloaded from L2 after the barrier's acquire fence
  auto fetch = [&](auto& buf, int pass) {
    const int c0 = pass * CPB + cw * NC;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      if (pass < n_pass && c0 + j < N) {
        const bf16* wr = W + (long long)(c0 + j) * K;
#pragma unroll
        for (int i = 0; i < kGemvCH; ++i) {
          const int k = k0 + i * kstep;
          if (k < K) buf[j][i].load(wr + k);
        }
      }
    }
  };
  fetch(cur, blockIdx.x);
  if (do_wait) qp_wait(bar, wait_target); else __syncthreads();
  qp_stamp(tm, tcap, ti);
  float ss[kGemvRows];
#pragma unroll
  for (int r = 0; r < kGemvRows; ++r) {
    ss[r] = 0.f;
    if (r < rows) {
      const float* xr = x + (long long)r * ldx;
      for (int k = threadIdx.x * 4; k < K; k += 1024) {
        float4 v = __ldcg(reinterpret_cast<const float4*>(xr + k));
        if (swiglu) {
          const float4 u = __ldcg(reinterpret_cast<const float4*>(xr + K + k));
          v.x = v.x / (1.0f + expf(-v.x)) * u.x; v.y = v.y / (1.0f + expf(-v.y)) * u.y;
          v.z = v.z / (1.0f + expf(-v.z)) * u.z; v.w = v.w / (1.0f + expf(-v.w)) * u.w;
        }
        *reinterpret_cast<float4*>(gx + r * K + k) = v;
        ss[r] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
    }
  }
  if (rms) {
#pragma unroll
    for (int r = 0; r < kGemvRows; ++r) { const float t = warp_sum(ss[r]); if (lane == 0) red[warp][r] = t; }
  }
  __syncthreads();
  float rstd[kGemvRows];
#pragma unroll
  for (int r = 0; r < kGemvRows; ++r) {
    rstd[r] = 1.f;
    if (rms) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][r];
      rstd[r] = rsqrtf(t / (float)K + eps);
    }
  }
  qp_stamp(tm, tcap, ti);
  for (int pass = blockIdx.x; pass < n_pass; pass += gridDim.x) {
    if (pass != (int)blockIdx.x) fetch(cur, pass);              // matrices wider than one resident wave: later passes pay their own round trip
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
          if (r < rows) {
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
      if (r < rows && c < N) {
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
        if (residual) v += __ldcg(residual + (long long)r * ldr + c);
        out[(long long)r * ldo + c] = v;
      }
    }
    if (KS > 1) __syncthreads();
  }
  qp_stamp(tm, tcap, ti);
  qp_arrive(bar);
}

__host__ __device__ inline int qp_ks_for(int K) {
  int ks = 1;
  while (ks < 8 && K > ks * 256 * kGemvCH) ks *= 2;
  return ks;
}

// ---- attention phase: task = (clip x query head, key range); the arithmetic of qwen_attn_split_kernel<DH, 256, false> ----
template <int DH>
__device__ __forceinline__ void qp_attention(const QPersistArgs& a, int layer, const float* g, unsigned wait_target, int& ti) {
  constexpr int SK = kPersistSK;
  constexpr int M = DH / 32, half = DH / 2, EPL = DH / 32, VK = SK / 8;
  __shared__ float qs[DH], kn[DH], vn[DH], sc[SK], red[8], pw[8][DH];
  __shared__ int s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = a.heads, KH = a.kv_heads, S = a.S, max_seq = a.max_seq;
  const int NHD = H + 2 * KH;
  const int n_bh = a.rows * H, n_tasks = n_bh * S;
  const long long layer_off = (long long)layer * a.cache_layer_stride;
  bool waited = false;
  qp_stamp(a.timing, a.timing_cap, ti);
  if (a.dbg & 2) { qp_wait(a.bar, wait_target); qp_arrive(a.bar); return; }
  for (int task = blockIdx.x; task < n_tasks; task += gridDim.x) {
    const int sp = task / n_bh, bh = task - sp * n_bh;
    const int b = bh / H, h = bh - b * H;
    const int pos = min(a.state->kv_len + a.kv_off[b], max_seq - 1), n_keys = pos + 1;
    const int per = (n_keys + S - 1) / S;
    const int lo = sp * per, hi = min(n_keys, lo + per);
    const int kh = h / (H / KH);
    bf16* K = a.kc + layer_off + ((long long)b * KH + kh) * max_seq * DH;
    bf16* V = a.vc + layer_off + ((long long)b * KH + kh) * max_seq * DH;
    uint4 kreg[DH / 8];
    const int jk = lo + threadIdx.x;
    const bool k_cached = jk < hi && jk != pos;
    if (k_cached) {
      const uint4* kr = reinterpret_cast<const uint4*>(K + (long long)jk * DH);
#pragma unroll
      for (int c = 0; c < DH / 8; ++c) kreg[c] = __ldcg(kr + c);
    }
    if (!waited) { qp_wait(a.bar, wait_target); waited = true; qp_stamp(a.timing, a.timing_cap, ti); } else __syncthreads();
    if (warp < 3) {
      const int head = warp == 0 ? h : (warp == 1 ? H + kh : H + KH + kh);
      const float* src = a.qkvf + ((long long)b * NHD + head) * DH;
      float v[M];
#pragma unroll
      for (int m = 0; m < M; ++m) v[m] = __ldcg(src + lane + 32 * m);
      if (warp < 2) {
        float ss = 0.f;
#pragma unroll
        for (int m = 0; m < M; ++m) ss += v[m] * v[m];
        const float r = rsqrtf(warp_sum(ss) / (float)DH + a.eps);
        const float* gg = g + (warp == 0 ? 0 : DH);
#pragma unroll
        for (int m = 0; m < M; ++m) v[m] *= r * gg[lane + 32 * m];
#pragma unroll
        for (int m = 0; m < M / 2; ++m) {
          const int j = lane + 32 * m;
          const float c = a.cosT[(long long)pos * half + j], sn = a.sinT[(long long)pos * half + j];
          const float aa = v[m], bb = v[m + M / 2];
          v[m] = aa * c - bb * sn;
          v[m + M / 2] = bb * c + aa * sn;
        }
      }
      if (warp == 0) {
#pragma unroll
        for (int m = 0; m < M; ++m) qs[lane + 32 * m] = v[m];
      } else {
        bf16* dst = (warp == 1 ? K : V) + (long long)pos * DH;
        float* keep = warp == 1 ? kn : vn;
#pragma unroll
        for (int m = 0; m < M; ++m) {
          const bf16 rv = __float2bfloat16_rn(v[m]);
          if (sp == 0) dst[lane + 32 * m] = rv;
          keep[lane + 32 * m] = __bfloat162float(rv);
        }
      }
    }
    __syncthreads();
    float s = -INFINITY;
    if (jk < hi) {
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
    if (jk < hi) { p = expf(s - mx); sc[threadIdx.x] = p; }
    float sum = warp_sum(p);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w];
    float acc[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) acc[e] = 0.f;
#pragma unroll 8
    for (int i = 0; i < VK; ++i) {
      const int j = lo + warp + 8 * i;
      if (j < hi) {
        const float pj = sc[j - lo];
        if (j == pos) {
#pragma unroll
          for (int e = 0; e < EPL; ++e) acc[e] = fmaf(pj, vn[lane * EPL + e], acc[e]);
        } else {
          const uint32_t* vr = reinterpret_cast<const uint32_t*>(V + (long long)j * DH + lane * EPL);
#pragma unroll
          for (int e = 0; e < EPL / 2; ++e) {
            const uint32_t raw = __ldcg(vr + e);
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw));
            acc[2 * e] = fmaf(pj, f.x, acc[2 * e]);
            acc[2 * e + 1] = fmaf(pj, f.y, acc[2 * e + 1]);
          }
        }
      }
    }
#pragma unroll
    for (int e = 0; e < EPL; ++e) pw[warp][lane * EPL + e] = acc[e];
    __syncthreads();
    float* mine = a.att_part + ((long long)bh * S + sp) * (DH + 2);
    if (threadIdx.x < DH) {
      float o = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) o += pw[w][threadIdx.x];
      mine[threadIdx.x] = o;
    }
    if (threadIdx.x == 0) { mine[DH] = mx; mine[DH + 1] = sum; }
    __syncthreads();
    if (threadIdx.x == 0) {            // release: this CTA's partial is at L2 before the ticket counts; the merge reads with ld.global.cg
      int old;
      asm volatile("atom.release.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(a.att_counter + bh), "r"(1) : "memory");
      s_last = (old == S - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
      if (threadIdx.x < DH) {
        const float* pb = a.att_part + (long long)bh * S * (DH + 2);
        float Mx = -INFINITY;
        for (int q = 0; q < S; ++q) Mx = fmaxf(Mx, __ldcg(pb + q * (DH + 2) + DH));
        float num = 0.f, den = 0.f;
        for (int q = 0; q < S; ++q) {
          const float wq = expf(__ldcg(pb + q * (DH + 2) + DH) - Mx);
          num = fmaf(wq, __ldcg(pb + q * (DH + 2) + threadIdx.x), num);
          den = fmaf(wq, __ldcg(pb + q * (DH + 2) + DH + 1), den);
        }
        a.actx[(long long)bh * DH + threadIdx.x] = num / den;
      }
      if (threadIdx.x == 0) a.att_counter[bh] = 0;
    }
  }
  if (!waited) qp_wait(a.bar, wait_target);
  qp_stamp(a.timing, a.timing_cap, ti);
  qp_arrive(a.bar);
}

// KSH / KSQ / KSI = warps per column (qp_ks_for) of the phases whose reduction length is hidden / heads * DH / inter
template <int DH, int KSH, int KSQ, int KSI>
__global__ void __launch_bounds__(kPersistThreads, 2)
qwen_persist_kernel(const __grid_constant__ QPersistArgs a) {
  extern __shared__ float gx[];                  // [kGemvRows][max(hidden, H*DH, inter)]
  __shared__ float red[8][kGemvRows];
  __shared__ float psum[8][2][kGemvRows];
  const int Hd = a.hidden, QD = a.heads * DH, NQ = (a.heads + 2 * a.kv_heads) * DH, I = a.inter;
  const unsigned G = gridDim.x;
  int ti = 0;
  unsigned target = a.bar_base;                  // completion count of the barrier the next phase waits on
  for (int l = 0; l < a.n_layers; ++l) {
    const QPersistLayer L = a.layers[l];
    qp_gemv<KSH>(a.x, Hd, 1, 0, a.eps, L.qkv_w, nullptr, 0, a.qkvf, NQ, a.rows, NQ, Hd, gx, red, psum, a.bar, target, l > 0, a.dbg, a.timing, a.timing_cap, ti);
    target += G;
    qp_attention<DH>(a, l, L.qk_g, target, ti);
    target += G;
    qp_gemv<KSQ>(a.actx, QD, 0, 0, a.eps, L.o_w, a.x, Hd, a.x, Hd, a.rows, Hd, QD, gx, red, psum, a.bar, target, 1, a.dbg, a.timing, a.timing_cap, ti);
    target += G;
    qp_gemv<KSH>(a.x, Hd, 1, 0, a.eps, L.gu_w, nullptr, 0, a.gu, 2 * I, a.rows, 2 * I, Hd, gx, red, psum, a.bar, target, 1, a.dbg, a.timing, a.timing_cap, ti);
    target += G;
    qp_gemv<KSI>(a.gu, 2 * I, 0, 1, a.eps, L.down_w, a.x, Hd, a.x, Hd, a.rows, Hd, I, gx, red, psum, a.bar, target, 1, a.dbg, a.timing, a.timing_cap, ti);
    target += G;
  }
}
