// Streaming greedy-decode kernel ("ring"): the whole decode loop of
// /root/reference/Whisper/Inference_Whisper_ONNX.py:584-663 (one DECODE_SESSION launch
// per token, 64 KV rebinds, `.numpy()` sync) as ONE cooperative launch, like
// decoder_mega.cu, but organised so that the HBM stream never waits on the token's
// dependency chain:
//
//   * one producer lane per CTA walks the CTA's static share of the per-token read
//     stream (its rows of every weight matrix of WHISPER_DECODER.forward,
//     /root/reference/Whisper/Export_Whisper.py:614-667, plus the cross-attention K/V
//     tiles of the (utterance, head) tasks it owns) and copies it into a shared-memory
//     ring with cp.async.bulk / cp.async.bulk.tensor (TMA) + mbarrier complete_tx,
//     running ahead of the consumers by the ring's capacity, across phase and token
//     boundaries;
//   * 16 consumer warps take stages off the ring (half a d-wide weight segment per
//     warp), so a phase's critical path is: poll the input vector -> LayerNorm in
//     shared memory -> dot products against weights that are already on chip ->
//     publish;
//   * phases exchange their (tiny) activation vectors through L2 with flag-in-data
//     words (value, sequence number) written by one 64-bit store and polled by one
//     64-bit load -- no grid barrier, no fence on the critical path (decoder_mega.cu
//     pays ~2 us barrier + ~0.7 us reload per phase, 257 phases per token).
//
// bf16 weights / KV, fp32 activations and accumulation; batch <= 4 utterances, one new
// token per utterance per iteration (the multi-token prefill stays in decoder_mega.cu).
#include "common.cuh"
#include "ptx.cuh"
#include <algorithm>
#include <cstdio>

namespace b200asr {

constexpr int kRingConsumerWarps = 16;
constexpr int kRingConsumers = kRingConsumerWarps * 32;     // 512
constexpr int kRingThreads = kRingConsumers + 32;           // + producer warp
constexpr int kRingMaxStages = 16;
constexpr int kSegPerStage = 8;                             // d-wide weight segments per ring stage
constexpr int kTcRound = 8;                                 // TC path: chunks whose per-warp partials are reduced together
constexpr long long kSpinLimit = 6000000000LL;              // ~3 s at 2 GHz: a protocol bug traps instead of hanging

// ---------------------------------------------------------------------------
// PTX wrappers (local to this file)
// ---------------------------------------------------------------------------
namespace {

using namespace ptx;

// named barrier over the 512 consumer threads (the producer warp never joins it)
__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, %0;" ::"n"(kRingConsumers) : "memory"); }

// flag-in-data exchange word: low 32 bits = fp32 (or int) payload, high 32 bits = sequence number
__device__ __forceinline__ void ll_store(unsigned long long* p, unsigned payload, unsigned seq) {
  const unsigned long long v = ((unsigned long long)seq << 32) | payload;
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ll_load(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __noinline__ unsigned ll_spin(const unsigned long long* p, unsigned seq) {
  const long long t0 = clock64();
  for (;;) {
    const unsigned long long v = ll_load(p);
    if ((unsigned)(v >> 32) == seq) return (unsigned)v;
    if (clock64() - t0 > kSpinLimit) {
      printf("b200asr decoder_ring: exchange wait timed out (block %d thread %d seq %u saw %u)\n", blockIdx.x,
             threadIdx.x, seq, (unsigned)(v >> 32));
      __trap();
    }
  }
}
// gather n exchange words of sequence `seq` into fp32 shared memory; `tid` in [0, nthr)
__device__ __forceinline__ void ll_gather(const unsigned long long* src, int n, unsigned seq, float* dst, int tid, int nthr,
                                          bool nospin = false) {
  constexpr int U = 4;
  for (int i0 = tid; i0 < n; i0 += nthr * U) {
    unsigned long long v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { const int i = i0 + u * nthr; if (i < n) v[u] = ll_load(src + i); }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * nthr;
      if (i < n) {
        unsigned pay = (unsigned)v[u];
        if ((unsigned)(v[u] >> 32) != seq && !nospin) pay = ll_spin(src + i, seq);
        dst[i] = __uint_as_float(pay);
      }
    }
  }
}

// warp-level tensor-core pieces of the TC dot-product path (weights = A, activations = B)
__device__ __forceinline__ void ldsm_x2(uint32_t& r0, uint32_t& r1, uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(saddr));
}
__device__ __forceinline__ void mma_bf16_16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct Pipe { int stage; uint32_t phase; };
__device__ __forceinline__ void pipe_advance(Pipe& p, int n_stages) {
  if (++p.stage == n_stages) { p.stage = 0; p.phase ^= 1; }
}

// balanced split of N output columns over the grid: every CTA owns floor(N/G) or ceil(N/G) of them
__device__ __forceinline__ void col_split(int N, int& n0, int& cnt) {
  const int G = gridDim.x, c = blockIdx.x;
  const int q = N / G, r = N - q * G;
  cnt = q + (c < r ? 1 : 0);
  n0 = c * q + min(c, r);
}

enum RIn { kRInLocal = 0, kRInX = 1, kRInVec = 2 };
enum ROut { kROutVec = 0, kROutResid = 1, kROutHead = 2 };

struct RLin {
  const bf16* W; const float* bias; int N, K;
  int in_mode, ln_mode; const float* gamma; const float* beta;
  int act, out_mode;
};

__device__ __forceinline__ RLin ring_lin(const MegaArgs& a, int l, int ph) {
  const int d = a.d;
  if (ph == 8)
    return RLin{reinterpret_cast<const bf16*>(a.embed), a.suppress_bias, a.vocab, d, kRInX, 2, a.ln_g, a.ln_b, kActNone, kROutHead};
  const MegaLayer& w = a.layers[l];
  switch (ph) {
    case 0: return RLin{(const bf16*)w.qkv_w, w.qkv_b, 3 * d, d, l == 0 ? kRInLocal : kRInX, 1, nullptr, nullptr, kActNone, kROutVec};
    case 2: return RLin{(const bf16*)w.out_w, w.out_b, d, d, kRInVec, 0, nullptr, nullptr, kActNone, kROutResid};
    case 3: return RLin{(const bf16*)w.cq_w, w.cq_b, d, d, kRInX, 1, nullptr, nullptr, kActNone, kROutVec};
    case 5: return RLin{(const bf16*)w.cout_w, w.cout_b, d, d, kRInVec, 0, nullptr, nullptr, kActNone, kROutResid};
    case 6: return RLin{(const bf16*)w.fc1_w, w.fc1_b, a.ffn, d, kRInX, 1, nullptr, nullptr, kActGelu, kROutVec};
    default: return RLin{(const bf16*)w.fc2_w, w.fc2_b, d, a.ffn, kRInVec, 0, nullptr, nullptr, kActNone, kROutResid};
  }
}

// (utterance, head) attention task owned by this CTA in layer l (kind 0 self, 1 cross), or -1
__device__ __forceinline__ int my_task(const RingArgs& ra, int l, int kind, int ntask) {
  const int G = gridDim.x;
  const int off = (l * 37 + (kind ? G / 2 : 0)) % G;
  const int rel = ((int)blockIdx.x - off + G) % G;
  const int t = (int)(((long long)rel * ra.task_inv) % G);
  return t < ntask ? t : -1;
}

}  // namespace

// ---------------------------------------------------------------------------
template <int NR, bool DBG, bool TC>
__global__ void __launch_bounds__(kRingThreads, 1)
decoder_ring_kernel(const __grid_constant__ CUtensorMap cross_map, const __grid_constant__ RingArgs ra) {
  const MegaArgs& a = ra.m;
  extern __shared__ __align__(128) uint8_t smem_raw[];     // no integer round trip: keeps the shared address space visible
  uint8_t* ring = smem_raw;
  const int d = a.d, ffn = a.ffn, B = a.batch, H = a.n_heads, T = a.T;
  const int SB = ra.stage_bytes, NS = ra.n_stages;
  const int kmax = d > ffn ? d : ffn;
  float* xs = reinterpret_cast<float*>(ring + (size_t)NS * SB);      // [NR][kmax]   staged (normalised) input rows
  float* xloc = xs + (size_t)NR * kmax;                              // [NR][d]      this CTA's copy of the residual stream
  float* part = xloc + (size_t)NR * d;                               // [part_cap][NR] per-(segment, half) partial dots
  float* sc = part + (size_t)ra.part_cap * NR;                       // [sc_cap]     attention scores
  float* qs = sc + ra.sc_cap;                                        // [64]
  float* opart = qs + 64;                                            // [16][64]
  float* cand = opart + kRingConsumerWarps * 64;                     // [G][NR][2]
  float* sbias2 = cand + (size_t)gridDim.x * NR * 2;                 // [2][part_cap] bias slice of the current / next phase (TC path)
  float* tcpart = sbias2 + 2 * ra.part_cap;                             // [kTcRound][16 warps][8 rows][NR] (TC path)

  __shared__ uint64_t full_bar[kRingMaxStages], empty_bar[kRingMaxStages];
  __shared__ float wbest_v[kRingConsumerWarps * 4];
  __shared__ int wbest_i[kRingConsumerWarps * 4];
  __shared__ int s_tok[4], s_ngen[4], s_fin[4], s_nsave[4];
  __shared__ int s_pen[4 * 32];
  __shared__ int s_hist[4 * 32];                   // last 32 selected ids per utterance (circular, index = n_save % 32)
  __shared__ int s_pen_n, s_all_done;
  __shared__ volatile int s_stop;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x;
  const int L = a.n_layers;
  const int R = ra.box_rows;                       // cross-attention rows per ring stage
  const int P = TC ? 2 * d + 16 : 2 * d;           // TC path: weight rows land 16 bytes askew so ldmatrix is conflict-free
  const int nbox = (T + R - 1) / R;
  const int ntask = B * H;
  const int dbg = DBG ? ra.debug : 0;              // timing experiments only (results are garbage when set)

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kRingConsumerWarps); }
    mbar_fence_init();
    prefetch_tensormap(&cross_map);
    s_stop = 0;
  }
  if (tid < B) {
    s_ngen[tid] = a.n_gen[tid]; s_fin[tid] = a.finished[tid]; s_nsave[tid] = a.n_save[tid]; s_tok[tid] = a.first_tokens[tid];
    // selection history written by earlier launches; ids selected inside this launch are tracked in shared memory
    // by every CTA (other CTAs' plain global stores are not visible through this SM's L1)
    const int ns = a.n_save[tid];
    for (int j = max(0, ns - 32); j < ns; ++j) s_hist[tid * 32 + (j & 31)] = a.save_id[(long long)tid * a.save_ld + j];
  }
  __syncthreads();

  // =========================================================================
  // producer: one lane streams this CTA's share of every token's read stream
  // =========================================================================
  if (warp == kRingConsumerWarps) {
    {                                             // the whole warp walks the schedule; lane 0 owns the barriers
      Pipe p{0, 0};
      long long issued = 0;
      bool stop = false;
      auto acquire = [&]() -> bool {           // wait until the consumers released ring slot p.stage (warp-uniform result)
        int ok = 1;
        if (lane == 0) {
          if (s_stop) ok = 0;
          else if (!mbar_try_wait(&empty_bar[p.stage], p.phase ^ 1)) {
            const long long t0 = clock64();
            while (!mbar_try_wait(&empty_bar[p.stage], p.phase ^ 1)) {
              if (s_stop) { ok = 0; break; }
              if (clock64() - t0 > kSpinLimit) { printf("b200asr decoder_ring: producer stalled (block %d)\n", blockIdx.x); __trap(); }
            }
          }
        }
        return __shfl_sync(0xffffffffu, ok, 0) != 0;
      };
      auto stream_rows = [&](const bf16* W, int N, int K) {
        int n0, cnt; col_split(N, n0, cnt);
        if (TC) {
          // chunk = up to 8 weight rows x one d-wide K part, one bulk copy per row (rows land P bytes apart);
          // K parts outermost so a consumer warp re-reads its activation fragment only when the part changes
          const int kparts = K / d, ngroups = (cnt + 7) >> 3;
          for (int part = 0; part < kparts && !stop; ++part)
            for (int grp = 0; grp < ngroups && !stop; ++grp) {
              if (!acquire()) { stop = true; break; }
              const int nrows = min(8, cnt - grp * 8);
              if (lane == 0) mbar_expect_tx(&full_bar[p.stage], (uint32_t)(nrows * 2 * d));
              if (ra.debug & 32) {                  // timing experiment: one copy per chunk (rows land unskewed: garbage results)
                if (lane == 0) bulk_g2s(ring + (size_t)p.stage * SB, W + ((long long)(n0 + grp * 8) * K + (long long)part * d),
                                        (uint32_t)(nrows * 2 * d), &full_bar[p.stage]);
              } else
              if (lane < nrows)                     // one copy per lane: the eight row copies of a chunk issue in parallel
                bulk_g2s(ring + (size_t)p.stage * SB + (size_t)lane * P, W + ((long long)(n0 + grp * 8 + lane) * K + (long long)part * d),
                         (uint32_t)(2 * d), &full_bar[p.stage]);
              ++issued; pipe_advance(p, NS);
            }
          return;
        }
        const char* src = reinterpret_cast<const char*>(W + (long long)n0 * K);
        const long long bytes = (long long)cnt * K * 2;
        for (long long off = 0; off < bytes && !stop; off += SB) {
          if (!acquire()) { stop = true; break; }
          const uint32_t n = (uint32_t)min((long long)SB, bytes - off);
          if (lane == 0) {
            mbar_expect_tx(&full_bar[p.stage], n);
            bulk_g2s(ring + (size_t)p.stage * SB, src + off, n, &full_bar[p.stage]);
          }
          ++issued; pipe_advance(p, NS);
        }
      };
      // one call site per helper: the phases run one after the other, so every inlined copy of a phase body
      // would be a separate, cold stretch of the instruction stream
      const int n_phases = 8 * L + 1;
      for (int iter = 0; iter < a.n_iters && !stop; ++iter) {
        for (int idx = 0; idx < n_phases && !stop; ++idx) {
          const int l = idx >> 3;
          const int ph = (idx == n_phases - 1) ? 8 : (idx & 7);
          if (ph == 1) continue;                    // self-attention reads the resident cache directly
          if (ph == 4) {
            const int t = (dbg & 16) ? -1 : my_task(ra, l, 1, ntask);
            if (t < 0) continue;
            const int b = t / H, h = t - b * H;
            for (int i = 0; i < 2 * nbox; ++i) {
              const int kind = i >= nbox ? 1 : 0;
              const int row0 = ((kind * L + l) * B + b) * T + (i - kind * nbox) * R;
              if (!acquire()) { stop = true; break; }
              if (lane == 0) {
                mbar_expect_tx(&full_bar[p.stage], (uint32_t)(R * 128));
                tma_load_2d(ring + (size_t)p.stage * SB, &cross_map, h * 64, row0, &full_bar[p.stage]);
              }
              ++issued; pipe_advance(p, NS);
            }
            continue;
          }
          if (dbg & 8) continue;
          const RLin Lp = ring_lin(a, l, ph);
          stream_rows(Lp.W, Lp.N, Lp.K);
        }
      }
      // drain: every copy that was issued must have landed before the CTA may exit
      __syncwarp();
      for (int s = 0; s < NS && lane == 0; ++s) {
        if (issued > s) {
          const uint32_t par = (s < p.stage) ? p.phase : (p.phase ^ 1);
          mbar_wait(&full_bar[s], par, "decoder_ring", kSpinLimit);
        }
      }
    }
    __syncthreads();
    return;
  }

  // =========================================================================
  // consumers
  // =========================================================================
  Pipe pipe{0, 0};
  unsigned seq = 1;                                // sequence number of the NEXT exchange to be published
  int kv_len = a.state->kv_len;
  int step = a.state->step;
  int t_idx = 0;
  auto stamp = [&]() {
    if (DBG && a.timing && blockIdx.x == 0 && tid == 0 && t_idx < a.timing_cap) {
      unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      a.timing[t_idx++] = t;
    }
  };
  auto fstamp = [&]() { if (DBG && ra.fine_timing) stamp(); };
  auto exbuf = [&](unsigned e) -> unsigned long long* { return ra.ll + (size_t)(e & 3u) * ra.ll_stride; };
  stamp();

  // token embedding + learned position of the fed-back token -> xloc (every CTA, redundantly)
  auto embed_rows = [&]() {
    for (int r = 0; r < B; ++r) {
      const bf16* er = reinterpret_cast<const bf16*>(a.embed) + (long long)s_tok[r] * d;
      const float* pr = a.pos + (long long)kv_len * d;
      for (int k = tid * 2; k < d; k += kRingConsumers * 2) {
        const float2 e2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(er + k));
        const float2 p2 = *reinterpret_cast<const float2*>(pr + k);
        xloc[r * d + k] = p2.x + e2.x;
        xloc[r * d + k + 1] = p2.y + e2.y;
      }
    }
  };

  // ---- one skinny linear phase ------------------------------------------------------------
  auto linear_phase = [&](const RLin& Lp, bool pen_on) {
    const int K = Lp.K;
    int n0, cnt; col_split(Lp.N, n0, cnt);
    const int kq = K / d;                          // d-wide segments per weight row
    const unsigned long long* in = exbuf(seq - 1);
    // epilogue operands fetched before the wait so their latency hides behind it
    float bias_v = 0.f;
    if (!TC && Lp.out_mode != kROutHead && tid < cnt * NR && Lp.bias) bias_v = Lp.bias[n0 + tid / NR];
    float* sbias = sbias2 + (seq & 1u) * ra.part_cap;       // alternates per phase: the previous phase's epilogue may still read its copy
    if (TC && tid < cnt) sbias[tid] = Lp.bias ? Lp.bias[n0 + tid] : 0.f;
    // ---- input ----
    constexpr int GV = 4;                          // values of one d-wide row a thread stages (d <= 2048)
    if (Lp.in_mode == kRInVec) {
      for (int r = 0; r < B; ++r) ll_gather(in + (size_t)r * ra.ld_vec, K, seq - 1, xs + (size_t)r * K, tid, kRingConsumers, dbg & 1);
      for (int r = B; r < NR; ++r) for (int k = tid; k < K; k += kRingConsumers) xs[(size_t)r * K + k] = 0.f;
      csync();
      fstamp();
    } else {
      // residual-stream rows: gather (or reuse the local copy) into registers, two-pass LayerNorm with one
      // cross-warp exchange per pass, normalised rows into xs.  K == d for every phase that takes this path.
      float xv[NR][GV];
      if (Lp.in_mode == kRInX) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          if (r < B) {
            const unsigned long long* src = in + (size_t)r * ra.ld_vec;
            unsigned long long w[GV];
#pragma unroll
            for (int u = 0; u < GV; ++u) { const int k = tid + u * kRingConsumers; if (k < d) w[u] = ll_load(src + k); }
#pragma unroll
            for (int u = 0; u < GV; ++u) {
              const int k = tid + u * kRingConsumers;
              xv[r][u] = 0.f;
              if (k < d) {
                unsigned pay = (unsigned)w[u];
                if ((unsigned)(w[u] >> 32) != seq - 1 && !(dbg & 1)) pay = ll_spin(src + k, seq - 1);
                xv[r][u] = __uint_as_float(pay);
                xloc[r * d + k] = xv[r][u];
              }
            }
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < NR; ++r)
#pragma unroll
          for (int u = 0; u < GV; ++u) { const int k = tid + u * kRingConsumers; xv[r][u] = (r < B && k < d) ? xloc[r * d + k] : 0.f; }
      }
      float mean[NR], rstd[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) { mean[r] = 0.f; rstd[r] = 1.f; }
      if (Lp.ln_mode != 0 && !(dbg & 2)) {
        // one cross-warp exchange: per-warp fp32 sums of x and x^2 (<= 96 values each), combined in double so the
        // E[x^2] - mean^2 form loses nothing that matters next to the two-pass form of the other decoder kernels
        float* red = opart;                        // [2][NR][16] cross-warp partials (opart is idle outside attention)
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          float sv = (xv[r][0] + xv[r][1]) + (xv[r][2] + xv[r][3]);
          float qv = fmaf(xv[r][0], xv[r][0], xv[r][1] * xv[r][1]) + fmaf(xv[r][2], xv[r][2], xv[r][3] * xv[r][3]);
          sv = warp_sum(sv);
          qv = warp_sum(qv);
          if (lane == 0) { red[r * kRingConsumerWarps + warp] = sv; red[(NR + r) * kRingConsumerWarps + warp] = qv; }
        }
        csync();
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          double sm = 0.0, sq = 0.0;
#pragma unroll
          for (int w = 0; w < kRingConsumerWarps; ++w) { sm += (double)red[r * kRingConsumerWarps + w]; sq += (double)red[(NR + r) * kRingConsumerWarps + w]; }
          const double mu = sm / (double)d;
          const double var = fmax(sq / (double)d - mu * mu, 0.0);
          mean[r] = (float)mu;
          rstd[r] = rsqrtf((float)var + a.eps);
        }
      }
      fstamp();
#pragma unroll
      for (int r = 0; r < NR; ++r) {
#pragma unroll
        for (int u = 0; u < GV; ++u) {
          const int k = tid + u * kRingConsumers;
          if (k < d) {
            float v = (xv[r][u] - mean[r]) * rstd[r];
            if (Lp.ln_mode == 2) v = v * Lp.gamma[k] + Lp.beta[k];
            xs[(size_t)r * K + k] = (r < B) ? v : 0.f;
          }
        }
      }
      csync();
    }
    fstamp();
    unsigned long long* out = exbuf(seq);
    if constexpr (TC) {
      // ---- weights off the ring on the tensor cores: a chunk is an [8 rows][d] bf16 tile; warp w multiplies its
      //      d/16-wide K slice (mma.sync m16n8k16, rows 8-15 of the A tile zero) by the activation fragment it keeps in
      //      registers (B operand: up to 8 utterances), and leaves an [8][NR] partial for the cross-warp reduction ----
      const int kparts = K / d, ngroups = (cnt + 7) >> 3, nchunk = kparts * ngroups;
      const int ksl = d >> 4;                      // K elements per warp
      const int nsteps = ksl >> 4;                 // k16 steps per warp (d = 1280: 5)
      constexpr int XS = 5;
      uint32_t bfr[XS][2];
      int cur_part = -1;
      const int g8 = lane >> 2, q4 = lane & 3;
      for (int c0 = 0; c0 < nchunk; c0 += kTcRound) {
        const int cend = min(nchunk, c0 + kTcRound);
        for (int ch = c0; ch < cend; ++ch) {
          const int part_i = kparts > 1 ? ch / ngroups : 0;
          if (part_i != cur_part) {
            cur_part = part_i;
            const float* xb = xs + part_i * d + warp * ksl + q4 * 2;        // B[k][n] = x_hat[n][k], n = lane / 4
#pragma unroll
            for (int sidx = 0; sidx < XS; ++sidx) {
              float2 lo = make_float2(0.f, 0.f), hi = make_float2(0.f, 0.f);
              if (sidx < nsteps && g8 < NR) {
                lo = *reinterpret_cast<const float2*>(xb + (size_t)g8 * K + sidx * 16);
                hi = *reinterpret_cast<const float2*>(xb + (size_t)g8 * K + sidx * 16 + 8);
              }
              const __nv_bfloat162 l2 = __floats2bfloat162_rn(lo.x, lo.y), h2 = __floats2bfloat162_rn(hi.x, hi.y);
              bfr[sidx][0] = *reinterpret_cast<const uint32_t*>(&l2);
              bfr[sidx][1] = *reinterpret_cast<const uint32_t*>(&h2);
            }
          }
          mbar_wait(&full_bar[pipe.stage], pipe.phase, "decoder_ring", kSpinLimit);
          // independent accumulators: the k16 steps of a warp are not chained through one C fragment (the phase is
          // latency-bound, a dependent mma.sync chain would cost ~35 cycles per step)
          float cacc[XS][4];
          const uint32_t abase = smem_u32(ring + (size_t)pipe.stage * SB) + (uint32_t)((lane & 7) * P) +
                                 (uint32_t)((warp * ksl + ((lane >> 3) & 1) * 8) * 2);
          uint32_t afr[XS][2];
#pragma unroll
          for (int sidx = 0; sidx < XS; ++sidx) {
            afr[sidx][0] = afr[sidx][1] = 0u;
            if (sidx < nsteps) ldsm_x2(afr[sidx][0], afr[sidx][1], abase + (uint32_t)(sidx * 32));
          }
#pragma unroll
          for (int sidx = 0; sidx < XS; ++sidx) {
            cacc[sidx][0] = cacc[sidx][1] = cacc[sidx][2] = cacc[sidx][3] = 0.f;
            if (sidx < nsteps) mma_bf16_16816(cacc[sidx], afr[sidx][0], 0u, afr[sidx][1], 0u, bfr[sidx][0], bfr[sidx][1]);
          }
          const float s0 = ((cacc[0][0] + cacc[1][0]) + (cacc[2][0] + cacc[3][0])) + cacc[4][0];
          const float s1 = ((cacc[0][1] + cacc[1][1]) + (cacc[2][1] + cacc[3][1])) + cacc[4][1];
          float* tp = tcpart + ((size_t)((ch - c0) * kRingConsumerWarps + warp) * 8 + g8) * NR;
          if (q4 * 2 < NR) tp[q4 * 2] = s0;
          if (q4 * 2 + 1 < NR) tp[q4 * 2 + 1] = s1;
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[pipe.stage]);
          pipe_advance(pipe, NS);
        }
        csync();
        // ---- reduce the 16 (x kparts) per-warp partials of every output of this round: 16 lanes per output ----
        const int j_lo = kparts > 1 ? 0 : c0 * 8;
        const int j_hi = kparts > 1 ? cnt : min(cnt, cend * 8);
        const int n_out = (j_hi - j_lo) * NR;
        for (int base = 0; base < n_out * 16; base += kRingConsumers) {
          const int t = base + tid;
          const int o = t >> 4, wsub = t & 15;
          float v = 0.f;
          int j = 0, r = 0;
          if (o < n_out) {
            j = j_lo + o / NR; r = o - (o / NR) * NR;
            const int grp = j >> 3, g = j & 7;
            for (int pi = 0; pi < kparts; ++pi) {
              const int cc = (kparts > 1 ? pi * ngroups + grp : grp) - c0;
              v += tcpart[((size_t)(cc * kRingConsumerWarps + wsub) * 8 + g) * NR + r];
            }
          }
#pragma unroll
          for (int sh = 8; sh > 0; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
          if (o < n_out && wsub == 0) {
            if (Lp.out_mode == kROutHead) {
              part[(size_t)j * NR + r] = v;          // summed logits; the head epilogue below adds bias / penalty
            } else if (r < B) {
              v += sbias[j];
              if (Lp.act == kActGelu) v = gelu_erf(v);
              if (Lp.out_mode == kROutResid) v += xloc[r * d + n0 + j];
              ll_store(out + (size_t)r * ra.ld_vec + n0 + j, __float_as_uint(v), seq);
            }
          }
        }
        if (cend < nchunk || Lp.out_mode == kROutHead) csync();       // partials are reused by the next round
      }
    } else {
    // ---- weights off the ring: warp w takes half (w & 1) of segment (w >> 1) of every stage.  8 % kq == 0, so a
    //      warp always meets the same d/2-wide slice of the input row: it lives in registers for the whole phase ----
    const int nseg = cnt * kq;
    const long long bytes = (long long)cnt * K * 2;
    const int nchunk = (int)((bytes + SB - 1) / SB);
    const int hseg = warp >> 1, hhalf = warp & 1;
    const int half_elems = d >> 1;
    constexpr int XJ = 5;                          // float4 fragments per lane: d/2 <= 5 * 128
    constexpr bool kRegX = NR <= 2;                // 4 rows would need 80 registers: those re-read shared memory per chunk
    float4 xf[kRegX ? NR : 1][XJ];
    const float* xb = xs + (hseg % kq) * d + hhalf * half_elems + lane * 4;
    if (kRegX) {
#pragma unroll
      for (int r = 0; r < (kRegX ? NR : 1); ++r)
#pragma unroll
        for (int j = 0; j < XJ; ++j)
          xf[r][j] = (j * 128 + lane * 4 < half_elems) ? *reinterpret_cast<const float4*>(xb + (size_t)r * K + j * 128)
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    constexpr int PB = 4;                          // chunks whose warp reductions are batched (independent shuffle chains)
    float pend[PB][NR];
    for (int ch0 = 0; ch0 < nchunk; ch0 += PB) {
#pragma unroll
      for (int c = 0; c < PB; ++c) {
        const int ch = ch0 + c;
#pragma unroll
        for (int r = 0; r < NR; ++r) pend[c][r] = 0.f;
        if (ch < nchunk) {
          if (!(dbg & 8)) mbar_wait(&full_bar[pipe.stage], pipe.phase, "decoder_ring", kSpinLimit);
          const int g = ch * kSegPerStage + hseg;
          if (g < nseg && !(dbg & 4)) {
            const bf16* wseg = reinterpret_cast<const bf16*>(ring + (size_t)pipe.stage * SB) + hseg * d + hhalf * half_elems + lane * 4;
            uint2 wu[XJ];
#pragma unroll
            for (int j = 0; j < XJ; ++j)
              wu[j] = (j * 128 + lane * 4 < half_elems) ? *reinterpret_cast<const uint2*>(wseg + j * 128) : make_uint2(0u, 0u);
            float a0[NR], a1[NR];
#pragma unroll
            for (int r = 0; r < NR; ++r) { a0[r] = 0.f; a1[r] = 0.f; }
#pragma unroll
            for (int j = 0; j < XJ; ++j) {
              const float2 w01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wu[j].x));
              const float2 w23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wu[j].y));
#pragma unroll
              for (int r = 0; r < NR; ++r) {
                float4 x;
                if (kRegX) x = xf[kRegX ? r : 0][j];
                else x = (j * 128 + lane * 4 < half_elems) ? *reinterpret_cast<const float4*>(xb + (size_t)r * K + j * 128)
                                                           : make_float4(0.f, 0.f, 0.f, 0.f);
                a0[r] = fmaf(w01.x, x.x, a0[r]); a1[r] = fmaf(w01.y, x.y, a1[r]);
                a0[r] = fmaf(w23.x, x.z, a0[r]); a1[r] = fmaf(w23.y, x.w, a1[r]);
              }
            }
#pragma unroll
            for (int r = 0; r < NR; ++r) pend[c][r] = a0[r] + a1[r];
          }
          __syncwarp();
          if (lane == 0 && !(dbg & 8)) mbar_arrive(&empty_bar[pipe.stage]);
          if (!(dbg & 8)) pipe_advance(pipe, NS);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int c = 0; c < PB; ++c)
#pragma unroll
          for (int r = 0; r < NR; ++r) pend[c][r] += __shfl_xor_sync(0xffffffffu, pend[c][r], o);
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < PB; ++c) {
          const int g = (ch0 + c) * kSegPerStage + hseg;
          if (ch0 + c < nchunk && g < nseg) {
#pragma unroll
            for (int r = 0; r < NR; ++r) part[(size_t)(g * 2 + hhalf) * NR + r] = pend[c][r];
          }
        }
      }
    }
    csync();
    fstamp();
    }
    // ---- epilogue + publish ----
    if (Lp.out_mode != kROutHead) {
      if (!TC && tid < cnt * NR) {
        const int j = tid / NR, r = tid - j * NR;
        if (r < B) {
          float v = 0.f;
          for (int s = 0; s < 2 * kq; ++s) v += part[(size_t)(j * 2 * kq + s) * NR + r];
          v += bias_v;
          if (Lp.act == kActGelu) v = gelu_erf(v);
          if (Lp.out_mode == kROutResid) v += xloc[r * d + n0 + j];
          ll_store(out + (size_t)r * ra.ld_vec + n0 + j, __float_as_uint(v), seq);
        }
      }
    } else {
      // tied lm head: suppress bias, sliding-window penalty, per-CTA argmax candidate per utterance
      const int r = tid % NR;                       // kRingConsumers % NR == 0: a thread keeps one row
      float bv = -INFINITY; int bi = 0x7fffffff;
      if (r < B) {
        for (int t = tid; t < cnt * NR; t += kRingConsumers) {
          const int j = t / NR, n = n0 + j;
          float v = TC ? part[(size_t)j * NR + r] + sbias[j]
                       : part[(size_t)(j * 2) * NR + r] + part[(size_t)(j * 2 + 1) * NR + r] + Lp.bias[n];
          if (pen_on) {
            bool hit = false;
            for (int q = 0; q < s_pen_n; ++q) hit |= (s_pen[r * 32 + q] == n);
            if (hit) v *= a.penalty_value;
          }
          if (a.logits) a.logits[(long long)r * a.vocab + n] = v;
          if (v > bv || (v == bv && n < bi)) { bv = v; bi = n; }
        }
      }
#pragma unroll
      for (int o = 16; o >= NR; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane < NR) { wbest_v[warp * 4 + lane] = bv; wbest_i[warp * 4 + lane] = bi; }
      csync();
      if (tid < NR) {                              // rows >= B publish (-inf, none) so the gather below completes
        bv = -INFINITY; bi = 0x7fffffff;
        for (int w = 0; w < kRingConsumerWarps; ++w) {
          const float v = wbest_v[w * 4 + tid]; const int i = wbest_i[w * 4 + tid];
          if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
        }
        ll_store(out + (size_t)(blockIdx.x * NR + tid) * 2, __float_as_uint(bv), seq);
        ll_store(out + (size_t)(blockIdx.x * NR + tid) * 2 + 1, (unsigned)bi, seq);
      }
    }
    ++seq;
    stamp();
  };

  // ---- self-attention of one (utterance, head): q/k/v of the new position arrive through the exchange,
  //      older positions come from the resident cache (appended by this same CTA in earlier iterations) ----
  auto self_attn_phase = [&](int l) {
    const int t = (dbg & 16) ? -1 : my_task(ra, l, 0, ntask);
    if (t >= 0) {
      const int b = t / H, h = t - b * H;
      const unsigned long long* in = exbuf(seq - 1) + (size_t)b * ra.ld_vec;
      bf16* kc = reinterpret_cast<bf16*>(a.kcache) + ((((long long)l * B + b) * H + h) * a.max_target) * 64;
      bf16* vc = reinterpret_cast<bf16*>(a.vcache) + ((((long long)l * B + b) * H + h) * a.max_target) * 64;
      // rows already in the cache do not depend on this phase's input: start copying them into the (idle) input-row
      // buffer before polling for q / k / v, so the score and PV loops below run out of shared memory
      const int cap_pos = (int)(((size_t)NR * kmax * sizeof(float)) / 256);
      const int npre = min(kv_len, cap_pos);
      bf16* sk = reinterpret_cast<bf16*>(xs);
      bf16* sv = sk + (size_t)npre * 64;
      for (int idx = tid; idx < npre * 16; idx += kRingConsumers) {
        const int which = idx >= npre * 8 ? 1 : 0;
        const int j = idx - which * npre * 8;
        const bf16* src = (which ? vc : kc) + (long long)(j >> 3) * 64 + (j & 7) * 8;
        bf16* dst = (which ? sv : sk) + (size_t)(j >> 3) * 64 + (j & 7) * 8;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      float* knv = cand;                           // [2][64] the new position's k and v, rounded through bf16 like the cache
      if (tid < 192) {
        const int which = tid >> 6, dd = tid & 63;
        const unsigned long long* p = in + which * d + h * 64 + dd;
        unsigned long long v = ll_load(p);
        unsigned pay = (unsigned)v;
        if ((unsigned)(v >> 32) != seq - 1 && !(dbg & 1)) pay = ll_spin(p, seq - 1);
        const float f = __uint_as_float(pay);
        if (which == 0) {
          qs[dd] = f;
        } else {
          const bf16 hb = __float2bfloat16_rn(f);
          (which == 1 ? kc : vc)[(long long)kv_len * 64 + dd] = hb;      // append for the later tokens
          knv[(which - 1) * 64 + dd] = __bfloat162float(hb);
        }
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
      csync();
      const int npos = kv_len + 1;
      float m = -INFINITY;
      for (int p = tid; p < npos; p += kRingConsumers) {
        float s = 0.f;
        if (p == kv_len) {
#pragma unroll 8
          for (int i = 0; i < 64; ++i) s = fmaf(knv[i], qs[i], s);
        } else if (p < npre) {
          const bf16* kr = sk + (size_t)p * 64;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int jj = (j + p) & 7;              // 16-byte chunks rotated by position: conflict-free
            const uint4 u = *reinterpret_cast<const uint4*>(kr + jj * 8);
            const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
            const float4 qa = *reinterpret_cast<const float4*>(qs + jj * 8);
            const float4 qb = *reinterpret_cast<const float4*>(qs + jj * 8 + 4);
            float2 f = __bfloat1622float2(hh[0]); s = fmaf(f.x, qa.x, s); s = fmaf(f.y, qa.y, s);
            f = __bfloat1622float2(hh[1]); s = fmaf(f.x, qa.z, s); s = fmaf(f.y, qa.w, s);
            f = __bfloat1622float2(hh[2]); s = fmaf(f.x, qb.x, s); s = fmaf(f.y, qb.y, s);
            f = __bfloat1622float2(hh[3]); s = fmaf(f.x, qb.z, s); s = fmaf(f.y, qb.w, s);
          }
        } else {
          const bf16* kr = kc + (long long)p * 64;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 u = __ldcg(reinterpret_cast<const uint4*>(kr + j * 8));
            const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = __bfloat1622float2(hh[i]);
              s = fmaf(f.x, qs[j * 8 + 2 * i], s);
              s = fmaf(f.y, qs[j * 8 + 2 * i + 1], s);
            }
          }
        }
        sc[p] = s;
        m = fmaxf(m, s);
      }
      m = warp_max(m);
      if (lane == 0) opart[warp] = m;
      csync();
      m = opart[0];
#pragma unroll
      for (int w = 1; w < kRingConsumerWarps; ++w) m = fmaxf(m, opart[w]);
      csync();
      float sum = 0.f;
      for (int p = tid; p < npos; p += kRingConsumers) { const float e = expf(sc[p] - m); sc[p] = e; sum += e; }
      sum = warp_sum(sum);
      if (lane == 0) opart[kRingConsumerWarps * 64 - 32 + warp] = sum;     // tail of opart: not touched by the PV partials of warp < 15.5
      csync();
      float o0 = 0.f, o1 = 0.f;
      for (int p = warp; p < npos; p += kRingConsumerWarps) {
        const float w = sc[p];
        float2 v;
        if (p == kv_len) {
          v = make_float2(knv[64 + 2 * lane], knv[64 + 2 * lane + 1]);
        } else if (p < npre) {
          v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sv + (size_t)p * 64 + 2 * lane));
        } else {
          const unsigned u = __ldcg(reinterpret_cast<const unsigned*>(vc + (long long)p * 64 + 2 * lane));
          v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
        }
        o0 = fmaf(w, v.x, o0); o1 = fmaf(w, v.y, o1);
      }
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < kRingConsumerWarps; ++w) tot += opart[kRingConsumerWarps * 64 - 32 + w];
      csync();                                     // everyone has read the sums before the partials overwrite the area
      opart[warp * 64 + 2 * lane] = o0;
      opart[warp * 64 + 2 * lane + 1] = o1;
      csync();
      if (tid < 64) {
        float o = 0.f;
#pragma unroll
        for (int w = 0; w < kRingConsumerWarps; ++w) o += opart[w * 64 + tid];
        ll_store(exbuf(seq) + (size_t)b * ra.ld_vec + h * 64 + tid, __float_as_uint(o / tot), seq);
      }
      csync();
    }
    ++seq;
    stamp();
  };

  // ---- cross-attention of one (utterance, head): K and V tiles come off the ring ----
  auto cross_attn_phase = [&](int l) {
    const int t = (dbg & 16) ? -1 : my_task(ra, l, 1, ntask);
    if (t >= 0) {
      const int b = t / H, h = t - b * H;
      if (tid < 64) {
        const unsigned long long* p = exbuf(seq - 1) + (size_t)b * ra.ld_vec + h * 64 + tid;
        unsigned long long v = ll_load(p);
        unsigned pay = (unsigned)v;
        if ((unsigned)(v >> 32) != seq - 1 && !(dbg & 1)) pay = ll_spin(p, seq - 1);
        qs[tid] = __uint_as_float(pay);
      }
      csync();
      // scores: two threads per position (32 dims each), 16-byte chunks rotated by position -> conflict-free
      for (int i = 0; i < nbox; ++i) {
        mbar_wait(&full_bar[pipe.stage], pipe.phase, "decoder_ring", kSpinLimit);
        const bf16* kb = reinterpret_cast<const bf16*>(ring + (size_t)pipe.stage * SB);
        for (int idx = tid; idx < R * 2; idx += kRingConsumers) {
          const int p = idx >> 1, hf = idx & 1;
          const bf16* kr = kb + p * 64 + hf * 32;
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int jj = (j + p) & 3;
            const uint4 u = *reinterpret_cast<const uint4*>(kr + jj * 8);
            const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
            const float4 qa = *reinterpret_cast<const float4*>(qs + hf * 32 + jj * 8);
            const float4 qb = *reinterpret_cast<const float4*>(qs + hf * 32 + jj * 8 + 4);
            float2 f = __bfloat1622float2(hh[0]); s = fmaf(f.x, qa.x, s); s = fmaf(f.y, qa.y, s);
            f = __bfloat1622float2(hh[1]); s = fmaf(f.x, qa.z, s); s = fmaf(f.y, qa.w, s);
            f = __bfloat1622float2(hh[2]); s = fmaf(f.x, qb.x, s); s = fmaf(f.y, qb.y, s);
            f = __bfloat1622float2(hh[3]); s = fmaf(f.x, qb.z, s); s = fmaf(f.y, qb.w, s);
          }
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          const int pos = i * R + p;
          if (hf == 0 && pos < T) sc[pos] = s;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[pipe.stage]);
        pipe_advance(pipe, NS);
      }
      csync();
      float m = -INFINITY;
      for (int p = lane; p < T; p += 32) m = fmaxf(m, sc[p]);
      m = warp_max(m);                             // every warp computes the same maximum
      csync();
      for (int p = tid; p < T; p += kRingConsumers) sc[p] = expf(sc[p] - m);
      csync();
      float tot = 0.f;
      for (int p = lane; p < T; p += 32) tot += sc[p];
      tot = warp_sum(tot);
      float o0 = 0.f, o1 = 0.f;
      for (int i = 0; i < nbox; ++i) {
        mbar_wait(&full_bar[pipe.stage], pipe.phase, "decoder_ring", kSpinLimit);
        const bf16* vb = reinterpret_cast<const bf16*>(ring + (size_t)pipe.stage * SB);
        const int pmax = min(R, T - i * R);
        for (int p = warp; p < pmax; p += kRingConsumerWarps) {
          const float w = sc[i * R + p];
          const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(vb + p * 64 + 2 * lane));
          o0 = fmaf(w, v.x, o0); o1 = fmaf(w, v.y, o1);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[pipe.stage]);
        pipe_advance(pipe, NS);
      }
      opart[warp * 64 + 2 * lane] = o0;
      opart[warp * 64 + 2 * lane + 1] = o1;
      csync();
      if (tid < 64) {
        float o = 0.f;
#pragma unroll
        for (int w = 0; w < kRingConsumerWarps; ++w) o += opart[w * 64 + tid];
        ll_store(exbuf(seq) + (size_t)b * ra.ld_vec + h * 64 + tid, __float_as_uint(o / tot), seq);
      }
      csync();
    }
    ++seq;
    stamp();
  };

  // =========================================================================
  for (int iter = 0; iter < a.n_iters; ++iter) {
    if (tid == 0) {
      int done = 1;
      for (int b = 0; b < B; ++b) done &= (s_fin[b] != 0);
      s_all_done = done;
    }
    csync();
    if (s_all_done) break;                         // identical in every CTA
    embed_rows();
    csync();
    const int n_phases = 8 * L + 1;
    for (int idx = 0; idx < n_phases; ++idx) {      // one call site per phase body (instruction-cache footprint)
      const int l = idx >> 3;
      const int ph = (idx == n_phases - 1) ? 8 : (idx & 7);
      if (ph == 1) { self_attn_phase(l); continue; }
      if (ph == 4) { cross_attn_phase(l); continue; }
      bool pen_on = false;
      if (ph == 8) {
        // sliding-window penalty ids (APPLY_PENALTY, Export_Whisper.py:318-331): active once generated >= penalty_range
        if (tid == 0) {
          int nmax = 0;
          if (a.penalty_value != 1.0f) {
            for (int b = 0; b < B; ++b) {
              const bool act = s_ngen[b] >= a.penalty_range;
              const int ns = s_nsave[b];
              const int first = max(0, ns - a.penalty_range);
              int cntp = 0;
              if (act) for (int j = first; j < ns && cntp < 32; ++j) s_pen[b * 32 + cntp++] = s_hist[b * 32 + (j & 31)];
              for (int j = cntp; j < 32; ++j) s_pen[b * 32 + j] = -1;
              nmax = max(nmax, cntp);
            }
          }
          s_pen_n = nmax;
        }
        csync();
        pen_on = s_pen_n > 0;
      }
      linear_phase(ring_lin(a, l, ph), pen_on);
    }
    // ---- every CTA reduces the per-CTA candidates identically ----
    ll_gather(exbuf(seq - 1), G * NR * 2, seq - 1, cand, tid, kRingConsumers, dbg & 1);
    csync();
    if (warp < B) {
      float bv = -INFINITY; int bi = 0x7fffffff;
      for (int c = lane; c < G; c += 32) {
        const float v = cand[(c * NR + warp) * 2];
        const int i = __float_as_int(cand[(c * NR + warp) * 2 + 1]);
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
        if (dbg) bi = (int)((unsigned)bi % (unsigned)a.vocab);   // timing experiments feed back garbage: keep the gather in range
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
        if (ns < a.save_ld) { s_hist[b * 32 + (ns & 31)] = bi; s_nsave[b] = ns + 1; }
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
    csync();
    kv_len += 1;
    step += 1;
  }
  csync();
  if (tid == 0) s_stop = 1;
  if (blockIdx.x == 0 && tid < B) {
    a.n_gen[tid] = s_ngen[tid];
    a.finished[tid] = s_fin[tid];
    a.n_save[tid] = s_nsave[tid];
  }
  if (blockIdx.x == 0 && tid == 0) {
    int done = 1;
    for (int b = 0; b < B; ++b) done &= (s_fin[b] != 0);
    a.state->kv_len = kv_len; a.state->step = step; a.state->all_done = done;
  }
  __syncthreads();                                 // joins the producer warp after its drain
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int ring_nr(int batch) { return batch <= 1 ? 1 : (batch <= 2 ? 2 : 4); }

bool ring_supported(int batch, int d, int ffn, int n_heads, int vocab, int num_sms) {
  return batch >= 1 && batch <= 4 && d % 256 == 0 && d <= 1280 && ffn % d == 0 && (kSegPerStage % (ffn / d)) == 0 &&
         d >= num_sms && vocab >= num_sms && batch * n_heads <= num_sms && d == n_heads * 64;
}

// shared-memory plan for one launch; returns false when it does not fit
bool ring_plan(const MegaArgs& a, int num_sms, bool tc, RingArgs* ra, size_t* smem_bytes) {
  const int NR = ring_nr(a.batch);
  const int d = a.d, ffn = a.ffn;
  const int kmax = d > ffn ? d : ffn;
  ra->tc = tc ? 1 : 0;
  ra->stage_bytes = tc ? kSegPerStage * (2 * d + 16) : kSegPerStage * d * 2;
  ra->box_rows = tc ? ((ra->stage_bytes / 128) & ~31) : ra->stage_bytes / 128;
  auto per_cta = [&](int N) { return (N + num_sms - 1) / num_sms; };
  int max_cnt = std::max(std::max(per_cta(3 * d), per_cta(ffn)), std::max(per_cta(d), per_cta(a.vocab)));
  int cap = per_cta(3 * d) * 2;
  cap = std::max(cap, per_cta(ffn) * 2);
  cap = std::max(cap, per_cta(d) * 2 * (ffn / d));
  cap = std::max(cap, per_cta(a.vocab) * 2);
  ra->part_cap = tc ? ((max_cnt + 3) & ~3) : ((cap + kSegPerStage * 2 + 3) & ~3);     // keeps the arrays carved after it 16-byte aligned
  ra->sc_cap = ((a.T > a.max_target ? a.T : a.max_target) + 8 + 3) & ~3;
  if (tc) {      // a K = ffn phase must reduce all its K parts in one round
    const int ngroups = (per_cta(d) + 7) / 8;
    if (ngroups * (ffn / d) > kTcRound || ra->box_rows < 32) return false;
  }
  const size_t fixed = 128 /*alignment slack*/ +
                       sizeof(float) * ((size_t)NR * kmax + (size_t)NR * d + (size_t)ra->part_cap * NR + ra->sc_cap + 64 +
                                        kRingConsumerWarps * 64 + (size_t)num_sms * NR * 2 + 2 * (size_t)ra->part_cap +
                                        (tc ? (size_t)kTcRound * kRingConsumerWarps * 8 * NR : 0));
  const size_t budget = 227 * 1024 - 2048;        // static __shared__ + slack
  if (fixed + 2 * (size_t)ra->stage_bytes > budget) return false;
  int ns = (int)((budget - fixed) / ra->stage_bytes);
  if (ns > kRingMaxStages) ns = kRingMaxStages;
  ra->n_stages = ns;
  *smem_bytes = fixed + (size_t)ns * ra->stage_bytes;
  return true;
}

size_t ring_exchange_words(int batch, int d, int ffn, int num_sms) {
  const int NR = ring_nr(batch);
  const size_t vec = (size_t)NR * (size_t)std::max(3 * d, ffn);
  const size_t cand = (size_t)num_sms * NR * 2;
  return std::max(vec, cand);
}

cudaError_t launch_decoder_ring(const RingArgs& ra_in, const CUtensorMap& cross_map, int num_sms, size_t smem_bytes,
                                cudaStream_t st) {
  const int NR = ring_nr(ra_in.m.batch);
  // the instrumented build (phase stamps, ablation switches) is a separate instantiation so the product path
  // carries none of its branches
  const bool dbg = ra_in.debug != 0 || ra_in.fine_timing != 0 || ra_in.m.timing != nullptr;
  // three builds per row count: tensor-core dot products (product path), CUDA-core dot products (cross-check),
  // CUDA-core + instrumentation
  void* fns[9] = {(void*)decoder_ring_kernel<1, false, true>,  (void*)decoder_ring_kernel<2, false, true>,  (void*)decoder_ring_kernel<4, false, true>,
                  (void*)decoder_ring_kernel<1, false, false>, (void*)decoder_ring_kernel<2, false, false>, (void*)decoder_ring_kernel<4, false, false>,
                  (void*)decoder_ring_kernel<1, true, false>,  (void*)decoder_ring_kernel<2, true, false>,  (void*)decoder_ring_kernel<4, true, false>};
  static AttrOnce attr;
  if (dbg && ra_in.tc && ra_in.debug != 32) return cudaErrorInvalidValue;   // the instrumented build exists for the CUDA-core path only
  const int slot = (NR == 1 ? 0 : (NR == 2 ? 1 : 2)) + (ra_in.tc ? 0 : (dbg ? 6 : 3));
  void* fn = fns[slot];
  if (attr.need(slot)) {
    cudaError_t r = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
    if (r != cudaSuccess) return r;
  }
  RingArgs ra = ra_in;
  CUtensorMap tm = cross_map;
  void* params[] = {&tm, &ra};
  return cudaLaunchCooperativeKernel(fn, dim3(num_sms), dim3(kRingThreads), params, smem_bytes, st);
}

}  // namespace b200asr
