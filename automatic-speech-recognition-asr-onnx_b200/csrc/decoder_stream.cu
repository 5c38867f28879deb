// Split-K tensor-core streaming decode kernel: the whole greedy loop of
// /root/reference/Whisper/Inference_Whisper_ONNX.py:584-663 (one DECODE_SESSION launch per token, 64 KV rebinds, a
// `.numpy()` sync per token) and the prompt prefill (:437-583) as ONE cooperative launch of 148 CTAs.  The math is
// WHISPER_DECODER.forward, /root/reference/Whisper/Export_Whisper.py:614-667, heads :228-240,318-331.
//
// Why this shape.  A decode step reads 1.6 GB of bf16 weights for a handful of activation rows; the step is a chain of
// 257 dependent phases (6 skinny linears + 2 attentions per layer, then the tied head).  The predecessor
// (decoder_ring.cu) split every linear by output columns and did the dot products on CUDA cores: ~900 instructions per
// warp per phase, cost proportional to the number of utterances, 22 % of the HBM roofline.  Here:
//
//   * every linear is split along K as well as N: the unit of work is an "atom" = 128 weight rows x 64 k (one 16 KB
//     SWIZZLE_128B TMA box).  A CTA owns a contiguous run of atoms of each phase (host-built schedule, balanced to one
//     atom cumulatively), so every byte the tensor core pulls out of shared memory is a useful weight byte;
//   * the dot products run on tcgen05: A = the weight box straight off the TMA ring (M = 128), B = the activation
//     slice as a 16-row bf16 tile (each utterance contributes a hi and a lo row: x = hi + lo keeps ~16 mantissa bits,
//     so the activations are not rounded to bf16), D = fp32 in TMEM.  One elected lane issues; the cost of a phase is
//     independent of the number of utterances (up to 8);
//   * partial sums meet in L2: an epilogue thread converts its row's sum to 52-bit fixed point and adds
//     (1 << 52 | value) to the output word with one RED.  Integer adds commute, so the result is bit-reproducible
//     whatever the arrival order; the top 12 bits count contributors, which is how a reader knows the word is complete
//     (flag-in-data, no grid barrier, no fence).  The residual stream lives in such words for the whole step: out /
//     cross-out / fc2 add straight into it;
//   * LayerNorm is folded around the GEMM: sum_k W[n][k] (x[k] - mu) rstd = rstd (sum_k W[n][k] x[k] - mu ws[n]) with
//     ws = row sums of W (computed once at load).  The statistics are accumulated by the readers of x into two more
//     words per utterance and are only needed one phase later, so no phase waits on a reduction;
//   * attention (one CTA per (utterance, head) task, rotating over the grid by layer) streams K / V boxes through the
//     same ring (cross K/V and the resident self-KV cache alike) and runs a warp-local online soft-max;
//   * one producer lane walks the CTA's share of the step's read stream in consumption order and runs ahead of the
//     consumers by the ring's capacity, across phase and token boundaries, so HBM never waits on the token's chain.
//
// Accumulator words are double-buffered by step parity; while a step runs on one set every CTA zeroes its share of
// the other (last read one step earlier, all CTAs pass the per-step arg-max exchange in between).
//
// The one place where a word is written more than once per step is the residual stream (out / cross-out / fc2 add into it
// in place).  A reader that fell a phase behind would poll for a contributor count the words have already passed, so the
// in-place writers wait for a per-(layer, version) "readers done" word that the MMA lane of every reading CTA bumps once the
// CTA's B operand is complete (three 8-byte words per layer; found by tools/stress_transcribe.py).
// With 4+ activation rows the B operand is handed to the MMA lanes slot by slot (one mbarrier per k-atom slot), with 1-2
// rows once per phase.
#include "common.cuh"
#include "ptx.cuh"
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <algorithm>
#include <cstdio>
#include <vector>

namespace b200asr {

constexpr int kStWorkerWarps = 8;
constexpr int kStWorkers = kStWorkerWarps * 32;       // 256
constexpr int kStThreads = kStWorkers + 96;           // warp 0: TMA producer, warp 1: MMA lane 0, warps 2..9: workers, warp 10: MMA lane 1
constexpr int kStStage = 16384;                       // one atom / one 128-row K or V box
constexpr int kStMaxStages = 12;
constexpr int kStSlot = 2048;                         // B operand of one k-atom: 16 rows x 128 bytes
constexpr int kStPartLd = 68;
constexpr float kFixScale = 16777216.0f;              // 2^24: accumulator resolution 6e-8, range +-1.3e8
constexpr float kFixInv = 1.0f / 16777216.0f;
constexpr float kSqScale = 4096.0f;                   // 2^12 for sums of squares (range 5e11)
constexpr long long kStSpin = 6000000000LL;           // ~3 s: a protocol bug traps instead of hanging the box

typedef unsigned long long u64;

namespace {

using namespace ptx;

__device__ __forceinline__ void wbar() { asm volatile("bar.sync 1, %0;" ::"n"(kStWorkers) : "memory"); }

__device__ __forceinline__ u64 ld_w(const u64* p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void ld_w2(const u64* p, u64& a, u64& b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_w(u64* p, u64 v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void st_zero2(u64* p) {
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %1};" ::"l"(p), "l"(0ull) : "memory");
}
__device__ __forceinline__ void red_add(u64* p, u64 v) { asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }

// accumulator word = (contributors << 52) + two's-complement fixed-point sum
__device__ __forceinline__ u64 enc_fix(float v, float scale, unsigned n = 1u) { return ((u64)n << 52) + (u64)__float2ll_rn(v * scale); }
__device__ __forceinline__ unsigned acc_cnt(u64 w) { return (unsigned)((w + (1ull << 51)) >> 52); }
__device__ __forceinline__ long long acc_raw(u64 w, unsigned c) { return (long long)(w - ((u64)c << 52)); }
__device__ __forceinline__ float acc_val(u64 w, unsigned c) { return (float)acc_raw(w, c) * kFixInv; }

// Everything that is not on the per-phase path is kept out of line: the step is a chain of ~260 short phases and the
// instruction stream of a phase has to stay cache-resident (a 264 KB first version of this kernel ran every phase cold).
__device__ __noinline__ void st_timeout(int where, int x, int y) {
  printf("b200asr decoder_stream: wait timed out (where %d, block %d thread %d, %d %d)\n", where, blockIdx.x, threadIdx.x, x, y);
  __trap();
}
__device__ __noinline__ void swait_slow(uint64_t* bar, uint32_t parity, int where) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > kStSpin) st_timeout(where, (int)parity, 0);
}
__device__ __forceinline__ void swait(uint64_t* bar, uint32_t parity, int where) {
  if (!mbar_try_wait(bar, parity)) swait_slow(bar, parity, where);
}
// in-place writers: wait (rare) until every reader CTA of the previous residual-stream version has reported
__device__ __noinline__ void rd_wait_slow(const u64* p, unsigned ex) {
  const long long t0 = clock64();
  for (;;) {
    const u64 w = ld_w(p);
    if (acc_cnt(w) == ex) return;
    if (clock64() - t0 > kStSpin) st_timeout(21, (int)acc_cnt(w), (int)ex);
  }
}
__device__ __noinline__ float2 gelu2(float a, float b) { return make_float2(gelu_erf(a), gelu_erf(b)); }

// (mean, rstd) from the two statistics words of a row (sum x in 2^-24, sum x^2 in 2^-12 fixed point)
__device__ __noinline__ float2 ln_stats(u64 w0, u64 w1, unsigned cnt, float inv_d, float eps) {
  const double sm = (double)acc_raw(w0, cnt) * (1.0 / 16777216.0), sq = (double)acc_raw(w1, cnt) * (1.0 / 4096.0);
  const double mu = sm * (double)inv_d;
  return make_float2((float)mu, rsqrtf((float)fmax(sq * (double)inv_d - mu * mu, 0.0) + eps));
}

__device__ __forceinline__ void tma_2d_hint(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_3d_hint(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(0), "l"(pol) : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* r) {
  uint32_t u[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = __uint_as_float(u[j]);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* r) {
  uint32_t u[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                 "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 16; ++j) r[j] = __uint_as_float(u[j]);
}
// FP8 weights (E4M3 A operand, per-row scale) x activations split into four E5M2 rows (B operand): K = 32 per instruction
__device__ __forceinline__ void tc_mma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// instruction descriptor for kind::f8f6f4: D = f32 (bit 4), A = E4M3 (0 @7), B = E5M2 (1 @10), K-major both, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t idesc_f8_e4m3_e5m2(int m, int n) {
  return (1u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// x ~= sum of four E5M2 values (3 significant bits each): byte j of the result is component j
__device__ __forceinline__ uint32_t split_e5m2x4(float x) {
  uint32_t out = 0;
  float r = x;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_fp8_storage_t q = __nv_cvt_float_to_fp8(r, __NV_SATFINITE, __NV_E5M2);
    const float back = __half2float(__half(__nv_cvt_fp8_to_halfraw(q, __NV_E5M2)));
    r -= back;
    out |= (uint32_t)q << (8 * j);
  }
  return out;
}

struct RingPos { int stage; uint32_t phase; };
__device__ __forceinline__ void ring_adv(RingPos& p, int n, int NS) {
  int s = p.stage + n;
  while (s >= NS) { s -= NS; p.phase ^= 1u; }
  p.stage = s;
}

// linear phase p6 (0 qkv, 1 out, 2 cq, 3 cout, 4 fc1, 5 fc2): k-atoms per weight row
__device__ __forceinline__ int phase_ka(int p6, int d, int ffn, int ksh) { return (p6 == 5 ? ffn : d) >> ksh; }
__device__ __forceinline__ int phase_p6(int ph) { return ph == 0 ? 0 : (ph == 2 ? 1 : (ph == 3 ? 2 : ph - 2)); }

// first attention task (utterance * H + head) owned by this CTA in layer l (kind 0 self, 1 cross); further tasks at + grid
__device__ __forceinline__ int first_task(int l, int kind, int task_inv) {
  const int G = gridDim.x;
  const int off = (l * 37 + (kind ? G / 2 : 0)) % G;
  const int rel = ((int)blockIdx.x - off + G) % G;
  return (int)(((long long)rel * task_inv) % G);
}

}  // namespace

// ---------------------------------------------------------------------------
// DBG: per-phase time stamps into a.timing (block 0), compiled only into the instrumented instantiation
// F8: the weight ring carries E4M3 bytes (atom = 128 rows x 128 k), activations are staged as four E5M2 rows per utterance,
//     the epilogue multiplies by the per-row weight scale (NRT <= 4)
// LEAN: the plain greedy decode launch (one new token per clip per iteration, no prompt rows, no begin-suppress, no penalty, no
//       logits dump, no per-clip key counts): the same arithmetic with those branches compiled out -- the phase loop's instruction
//       footprint decides its speed, and this is the launch that runs 32 of the 33 heads of a clip
template <int NRT, bool DBG, bool F8, bool LEAN>
__global__ void __launch_bounds__(kStThreads, 1)
decoder_stream_kernel(const __grid_constant__ CUtensorMap cross_map, const __grid_constant__ CUtensorMap kc_map,
                      const __grid_constant__ CUtensorMap vc_map, const __grid_constant__ StreamArgs sa) {
  const MegaArgs& a = sa.m;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int NS = sa.n_stages, L = a.n_layers, G = gridDim.x, d = a.d, ffn = a.ffn, B = a.batch, H = a.n_heads, T = a.T;
  const int Bs = a.batch_stride > 0 ? a.batch_stride : B, b0c = a.clip0;   // K/V strides and first clip of a sub-batch launch
  const int n_sched = 6 * L + 1;
  const int n_cnt = L * 3 * sa.cnt_ld, n_xexp = L * 3 * sa.xt;
  uint8_t* ring = base;
  uint8_t* bbuf = ring + (size_t)NS * kStStage;
  auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  int4* s_sched = reinterpret_cast<int4*>(bbuf + (size_t)sa.n_slots * kStSlot);     // {atom begin, atom end, first tile, first k-atom}
  unsigned char* s_cnt = reinterpret_cast<unsigned char*>(s_sched) + (size_t)n_sched * 16;
  unsigned short* s_xexp = reinterpret_cast<unsigned short*>(s_cnt + up16((size_t)n_cnt));
  float* s_cand = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(s_xexp) + up16((size_t)(n_xexp + 3 * L) * 2));   // [G][NRT][2]
  float* s_part = s_cand + (((size_t)G * NRT * 2 + 3) & ~(size_t)3);             // [8][68] per-warp (max, sum, o[64]) + [8*68] new-token score
  float* s_qs = s_part + kStWorkerWarps * kStPartLd + 8;                         // [64]
  float* s_knv = s_qs + 64;                                                      // [2][64]
  float* s_best = s_knv + 1024;                                                  // [8][NRT][2]   (s_knv: [k | v][8 new positions][64])
  float* s_sc = s_best + kStWorkerWarps * NRT * 2 + ((4 - ((kStWorkerWarps * NRT * 2) & 3)) & 3);   // [8 warps][64] attention scores / probabilities

  __shared__ uint64_t full_bar[kStMaxStages], empty_bar[kStMaxStages];
  // slot_bar[s] (4+ rows): the B operand of k-atom slot s is staged (every staging lane of every active row arrives), so the MMA lanes
  // start on the slots whose input words have landed while the stagers of the others are still polling -- with 4+ rows staging
  // takes several poll rounds (0.982 -> 0.935 ms per step at batch 4); with 1-2 rows one hand-over per phase through b_ready is
  // faster (0.817 vs 0.840 ms at batch 1)
  __shared__ uint64_t slot_bar[32], b_ready, acc_full[2], acc_empty[2];
  __shared__ int box_cnt[kStMaxStages];
  __shared__ uint32_t tmem_slot;
  __shared__ int s_tok[8], s_ngen[8], s_fin[8], s_nsave[8];
  __shared__ int s_pen[8 * 32], s_hist[8 * 32];
  __shared__ int s_pen_n;
  __shared__ volatile int s_go, s_stop, s_prog;      // s_prog: self-attention phases this CTA has completed (it * L + l + 1)
  // monotonic hand-off of the step loop to the MMA lanes (a parity barrier would alias: on a CTA without atoms nothing else
  // keeps the workers from running two steps ahead of an MMA lane): steps started so far, and the iteration the loop ended at
  __shared__ volatile int s_step, s_break_it;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int KSH = F8 ? 7 : 6;                          // log2(k per atom): a 128-byte swizzle row holds 64 bf16 or 128 fp8 weights
  const int KAd = d >> KSH;                                // k-atoms of a d-wide row
  const int KUd = d >> 6;                                  // 64-wide activation units of a d-wide row (who stages a unit contributes its statistics)
  const int n_first = LEAN ? 1 : (a.first_n_new > 0 ? a.first_n_new : 1);
  // sa.multi: the n_first prompt positions of every utterance are processed together in ONE iteration, as NF "virtual
  // utterances" per clip (row vu = utterance * NF + position; causal self-attention among a clip's rows).  Such a launch has
  // a single iteration (the host follows it with a plain decode launch), so the row count is a launch constant.
  const int NF = LEAN ? 1 : (sa.multi ? n_first : 1);
  const int R = B * NF;                                    // activation rows of this launch
  const int total_iters = (!LEAN && sa.multi) ? 1 : a.n_iters + n_first - 1;   // otherwise: leading forced (prompt) tokens, then n_iters heads
  const int ntask = R * H;
  const long long xreg = sa.set_words - (long long)L * sa.layer_words;   // residual-stream words + head statistics
  const int kv0 = a.state->kv_len;
  constexpr int kEpiWarps = NRT > 4 ? 8 : 4;
  constexpr bool kSlotHandOver = NRT >= 4;                 // B operand handed to the MMA lanes slot by slot instead of once per phase

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); box_cnt[s] = 0; }
    mbar_init(&b_ready, kStWorkers);
    for (int s = 0; s < 32; ++s) mbar_init(&slot_bar[s], (uint32_t)(32 * R * (F8 ? 2 : 1)));
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 2); mbar_init(&acc_empty[s], kStWorkerWarps); }
    mbar_fence_init();
    prefetch_tensormap(&cross_map); prefetch_tensormap(&kc_map); prefetch_tensormap(&vc_map);
    s_stop = 0; s_go = 0; s_prog = 0; s_step = 0; s_break_it = 0x7fffffff;
  }
#pragma unroll 1
  for (int i = tid; i < n_sched; i += kStThreads) s_sched[i] = sa.sched[(size_t)blockIdx.x * n_sched + i];
#pragma unroll 1
  for (int i = tid; i < n_cnt; i += kStThreads) s_cnt[i] = sa.cnt[i];
#pragma unroll 1
  for (int i = tid; i < n_xexp + 3 * L; i += kStThreads) s_xexp[i] = sa.xexp[i];     // + [L][3] CTAs that read the residual stream in qkv / cq / fc1
#pragma unroll 1
  for (int i = tid; i < sa.n_slots * kStSlot / 16; i += kStThreads) reinterpret_cast<uint4*>(bbuf)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid < B) {
    s_ngen[tid] = a.n_gen[tid]; s_fin[tid] = a.finished[tid]; s_nsave[tid] = a.n_save[tid]; s_tok[tid] = 0;
    const int ns = a.n_save[tid];
#pragma unroll 1
    for (int j = max(0, ns - 32); j < ns; ++j) s_hist[tid * 32 + (j & 31)] = a.save_id[(long long)tid * a.save_ld + j];
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 64u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    // =======================================================================
    // producer: one lane streams this CTA's share of every step's read stream, in consumption order
    // =======================================================================
    if (lane == 0) {
      uint64_t pol = 0;
      if (sa.l2_hint) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
      else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
      RingPos p{0, 0};
      long long issued = 0;
      bool stop = false;
      auto acquire = [&]() -> bool {
        if (s_stop) return false;
        if (!mbar_try_wait(&empty_bar[p.stage], p.phase ^ 1u)) {
          const long long t0 = clock64();
          while (!mbar_try_wait(&empty_bar[p.stage], p.phase ^ 1u)) {
            if (s_stop) return false;
            if (clock64() - t0 > kStSpin) st_timeout(1, p.stage, 0);
          }
        }
        return true;
      };
      auto weights = [&](const CUtensorMap* tm, int4 r, int KA) {
        int tile = r.z, ka = r.w;
        for (int at = r.x; at < r.y; ++at) {
          if (!acquire()) { stop = true; return; }
          mbar_expect_tx(&full_bar[p.stage], kStStage);
          tma_3d_hint(ring + (size_t)p.stage * kStStage, tm, ka << KSH, tile * 128, &full_bar[p.stage], pol);
          ++issued; ring_adv(p, 1, NS);
          if (++ka == KA) { ka = 0; ++tile; }
        }
      };
      auto kvbox = [&](const CUtensorMap* tm, int c0, int row) {
        if (!acquire()) { stop = true; return; }
        mbar_expect_tx(&full_bar[p.stage], kStStage);
        tma_2d_hint(ring + (size_t)p.stage * kStStage, tm, c0, row, &full_bar[p.stage], pol);
        ++issued; ring_adv(p, 1, NS);
      };
      for (int it = 0; it < total_iters && !stop; ++it) {
        const int kv = kv0 + it;
        const bool head_on = LEAN || sa.multi || it >= n_first - 1;
        for (int l = 0; l < L && !stop; ++l) {
          for (int ph = 0; ph < 8 && !stop; ++ph) {
            if (ph == 1) {                      // resident self-KV rows [0, kv) of my (utterance, head) tasks: K_i, V_i interleaved
              const int nb = (kv + 127) >> 7;
              const int t0 = first_task(l, 0, sa.task_inv);
              if (t0 < ntask && it > 0) {
                // the newest cache row was appended by this CTA's workers in the previous step's phase: a short model lets
                // the ring run more than a step ahead, so wait until they are past it
                const int need = (it - 1) * L + l + 1;
                const long long tw = clock64();
                while (s_prog < need && !s_stop)
                  if (clock64() - tw > kStSpin) st_timeout(7, need, s_prog);
              }
              for (int t = t0; t < ntask && !stop; t += G) {
                const int vu = t / H, h = t - vu * H, b = vu / NF;
                const int row0 = ((l * Bs + b0c + b) * H + h) * a.max_target;
                for (int g0 = 0; g0 < nb && !stop; g0 += 4) {          // groups of <= 4 boxes: the K boxes, then the matching V boxes
                  const int ge = min(nb, g0 + 4);
                  for (int i = g0; i < ge && !stop; ++i) kvbox(&kc_map, 0, row0 + i * 128);
                  for (int i = g0; i < ge && !stop; ++i) kvbox(&vc_map, 0, row0 + i * 128);
                }
              }
            } else if (ph == 4) {               // cross K / V of my tasks
              const int nb = (T + 127) >> 7;
              for (int t = first_task(l, 1, sa.task_inv); t < ntask && !stop; t += G) {
                const int vu = t / H, h = t - vu * H, b = vu / NF;
                const int rk = (l * Bs + b0c + b) * T, rv = ((L + l) * Bs + b0c + b) * T;
                for (int g0 = 0; g0 < nb && !stop; g0 += 4) {
                  const int ge = min(nb, g0 + 4);
                  for (int i = g0; i < ge && !stop; ++i) kvbox(&cross_map, h * 64, rk + i * 128);
                  for (int i = g0; i < ge && !stop; ++i) kvbox(&cross_map, h * 64, rv + i * 128);
                }
              }
            } else {
              const int p6 = phase_p6(ph);
              weights(&sa.wmaps[l * 6 + p6], s_sched[l * 6 + p6], phase_ka(p6, d, ffn, KSH));
            }
          }
        }
        if (head_on && !stop) weights(&sa.wmaps[6 * L], s_sched[6 * L], KAd);
      }
      // drain: every copy that was issued must have landed before the CTA may exit
      for (int s = 0; s < NS; ++s) {
        if (issued > s) {
          const uint32_t par = (s < p.stage) ? p.phase : (p.phase ^ 1u);
          swait(&full_bar[s], par, 2);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 10) {
    // =======================================================================
    // MMA issuers: two lanes (warps 1 and 10) on alternate atoms of the CTA's run, each with its own accumulator; D[128 weight rows][16] (+)= A[128][64] (ring stage) . B[16][64] (activation slot)
    // =======================================================================
    if (lane == 0) {
      constexpr uint32_t idesc = F8 ? idesc_f8_e4m3_e5m2(128, 16) : idesc_bf16(128, 16);
      RingPos p{0, 0};
      uint32_t bpar = 0, spar = 0;                        // parity of b_ready; bit s: parity of slot s's next hand-over
      int tctr = 0;
      const uint32_t ring_u = smem_u32(ring), bbuf_u = smem_u32(bbuf);
      const int ml = warp == 10;                          // my lane: atoms whose offset in the run has my parity
      auto linear = [&](int4 r, int KA, u64* rd_sig) {
        // everything that does not depend on this phase's activations happens before the wait for them: my weight stages of
        // the phase (streamed ahead by the producer) and the first accumulator buffer are confirmed here, so the section
        // between "B operand staged" and "accumulator ready" is MMA issue only (measured: 0.91 -> 0.86 ms per step)
        const int npre = min(r.y - r.x, NS - 2);           // (the head's tiles exceed the ring: the rest is confirmed in the loop)
        {
          RingPos q = p;
          for (int i = 0; i < npre; ++i) { if ((i & 1) == ml) swait(&full_bar[q.stage], q.phase, 5); ring_adv(q, 1, NS); }
          if (r.x < r.y) swait(&acc_empty[tctr & 1], (uint32_t)(((tctr >> 1) & 1) ^ 1), 4);
        }
        if (r.x >= r.y) return;                            // no atoms of this phase here: no hand-shake either (workers skip it too)
        const int nsl = min(r.y - r.x, KA);                // slots the workers stage in this phase
        uint32_t waited = 0;
        if (!kSlotHandOver) {
          swait(&b_ready, bpar, 3); bpar ^= 1u;
          tc_fence_after();
          // every worker of this CTA has consumed its input words: report the CTA as a finished reader of the residual stream
          // (qkv / cq / fc1; the in-place writers of the next version wait for all of them, see the epilogue)
          if (rd_sig && ml == 0) red_add(rd_sig, 1ull << 52);
        }
        int ka = r.w;
        bool tile_start = true, fresh = true;
        for (int at = r.x; at < r.y; ++at) {
          const int buf = tctr & 1;
          if (tile_start && at != r.x) { swait(&acc_empty[buf], (uint32_t)(((tctr >> 1) & 1) ^ 1), 4); tc_fence_after(); }
          tile_start = false;
          if (((at - r.x) & 1) == ml) {
            if (at - r.x >= npre) { swait(&full_bar[p.stage], p.phase, 5); tc_fence_after(); }
            int slot = ka - r.w; if (slot < 0) slot += KA;
            if (kSlotHandOver) {
              if (!((waited >> slot) & 1u)) { swait(&slot_bar[slot], (spar >> slot) & 1u, 3); waited |= 1u << slot; }
              tc_fence_after();
            }
            const uint64_t adesc = smem_desc_sw128(ring_u + (uint32_t)p.stage * kStStage);
            const uint64_t bdesc = smem_desc_sw128(bbuf_u + (uint32_t)slot * kStSlot);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (F8) tc_mma_f8(tmem + (uint32_t)(buf * 32 + ml * 16), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (fresh && k == 0) ? 0u : 1u);
              else tc_mma_bf16(tmem + (uint32_t)(buf * 32 + ml * 16), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (fresh && k == 0) ? 0u : 1u);
            }
            fresh = false;
            tc_commit(&empty_bar[p.stage]);
          }
          ring_adv(p, 1, NS);
          ++ka;
          if (ka == KA || at + 1 == r.y) {
            // my MMAs of this tile are tracked by the commit; a lane that had no atom in the tile just arrives (the epilogue
            // skips its accumulator: same parity rule)
            if (!fresh) tc_commit(&acc_full[buf]); else mbar_arrive(&acc_full[buf]);
            ++tctr; fresh = true; tile_start = true;
            if (ka == KA) ka = 0;
          }
        }
        if (kSlotHandOver) {
          if (rd_sig && ml == 0) {                         // reader report: once EVERY slot of the phase is staged (the other lane's too)
            for (int sl = 0; sl < nsl; ++sl)
              if (!((waited >> sl) & 1u)) swait(&slot_bar[sl], (spar >> sl) & 1u, 3);
            red_add(rd_sig, 1ull << 52);
          }
          spar ^= nsl >= 32 ? 0xffffffffu : ((1u << nsl) - 1u);    // every staged slot completed one hand-over, waited on by me or not
        }
      };
      for (int it = 0; it < total_iters; ++it) {
        {
          const long long t0 = clock64();
          while (s_step <= it && s_break_it > it)
            if (clock64() - t0 > kStSpin) st_timeout(6, it, s_step);
          if (s_break_it <= it) break;
        }
        const int kv = kv0 + it;
        const bool head_on = LEAN || sa.multi || it >= n_first - 1;
        for (int l = 0; l < L; ++l) {
          for (int ph = 0; ph < 8; ++ph) {
            if (ph == 1 || ph == 4) {             // attention stages are consumed by the workers: skip over them
              const int nb = ph == 1 ? ((kv + 127) >> 7) : ((T + 127) >> 7);
              int n = 0;
              for (int t = first_task(l, ph == 4, sa.task_inv); t < ntask; t += G) n += 2 * nb;
              ring_adv(p, n, NS);
            } else {
              const int p6 = phase_p6(ph);
              u64* lay_w = sa.acc + (size_t)(it & 1) * sa.set_words + xreg + (long long)l * sa.layer_words;
              linear(s_sched[l * 6 + p6], phase_ka(p6, d, ffn, KSH),
                     (p6 == 0 || p6 == 2 || p6 == 4) ? lay_w + (long long)NRT * (6 * d + ffn) + 6 * NRT + (p6 >> 1) : nullptr);
            }
          }
        }
        if (head_on) linear(s_sched[6 * L], KAd, nullptr);
      }
    }
    __syncwarp();
  } else {
    // =======================================================================
    // workers (256 threads): activation gather -> B tiles, TMEM epilogue -> RED, attention, arg-max
    // =======================================================================
    const int wt = tid - 64, ww = wt >> 5;
    const int r_mine = ww % NRT;                         // the utterance this warp stages
    const int g_mine = ww / NRT;                         // slot group
    constexpr int NG = kStWorkerWarps / NRT;             // slot groups
    constexpr int U = NRT <= 2 ? 1 : (NRT == 4 ? 2 : 4); // k-atoms a thread polls together
    constexpr int RR = NRT < 4 ? NRT : 4;                // utterances an epilogue thread handles
    constexpr int NC = F8 ? 16 : 8, CPR = F8 ? 4 : 2;    // accumulator columns an epilogue thread reads; columns per utterance
    const bool row_ok = r_mine < R;
    const int q_tm = warp & 3;                           // TMEM lane quarter this warp may read
    const int set_tm = ww >> 2;                          // 0: utterances 0-3 (columns 0-7), 1: utterances 4-7 (columns 8-15)
    const bool epi_warp = ww < kEpiWarps;
    const float inv_d = 1.0f / (float)d;
    RingPos pos{0, 0};
    int tctr = 0;
    int step = a.state->step;
    int it_done = 0;
    int t_idx = 0;
    auto stamp = [&]() {
      if (DBG && a.timing && blockIdx.x == 0 && wt == 0 && t_idx < a.timing_cap) {
        unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        a.timing[t_idx++] = t;
      }
    };
    stamp();

    for (int it = 0; it < total_iters; ++it) {
      const int kv = kv0 + it;
      const bool head_on = LEAN || sa.multi || it >= n_first - 1;
      const bool begin_on = !LEAN && head_on && (sa.multi || it == n_first - 1) && a.first_is_prefill && a.begin_bias != nullptr;
      if (wt == 0) {
        int done = 1;
        for (int b = 0; b < B; ++b) done &= (s_fin[b] != 0);
        if (!LEAN && it < n_first && a.first_is_prefill) done = 0;          // the prefill always runs
        s_go = done ? 2 : 1;
      }
      if (wt < R && it < n_first) s_tok[wt] = a.first_tokens[(!LEAN && sa.multi) ? wt : wt * n_first + it];
      wbar();
      if (wt == 0) { if (s_go == 2) s_break_it = it; else s_step = it + 1; }
      if (s_go == 2) break;
      u64* set = sa.acc + (size_t)(it & 1) * sa.set_words;
      u64* oset = sa.acc + (size_t)((it + 1) & 1) * sa.set_words;
      u64* xw = set;                                               // [NRT][d] residual stream, then [NRT][2] head statistics
      u64* hstats = xw + (long long)NRT * d;

      const int n_idx = 8 * L + (head_on ? 1 : 0);
#pragma unroll 1
      for (int idx = 0; idx < n_idx; ++idx) {
        const int l = idx >> 3, ph = idx & 7;
        const bool is_head = idx == 8 * L;
        u64* lay = set + xreg + (long long)(is_head ? 0 : l) * sa.layer_words;
        u64* stats = lay + (long long)NRT * (6 * d + ffn);
        const StreamLayer& slr = sa.sl[is_head ? 0 : l];

        if (!is_head && (ph == 1 || ph == 4)) {
          // =================================================================
          // attention of my (utterance, head) tasks of layer l.  ph 1: self (resident cache rows [0, kv) off the ring + the
          // new position from the exchange), ph 4: cross (T rows off the ring)
          // =================================================================
          const int kind = ph == 4;
          if (!kind) {
            // zero this CTA's share of the other set's layer-l words (and, at layer 0, of its residual-stream words)
            const long long ngran = sa.layer_words >> 1;
            const long long g0 = ngran * blockIdx.x / G, g1 = ngran * (blockIdx.x + 1) / G;
            u64* zb = oset + xreg + (long long)l * sa.layer_words;
#pragma unroll 1
            for (long long i = g0 + wt; i < g1; i += kStWorkers) st_zero2(zb + 2 * i);
            if (l == 0) {
              const long long xg = xreg >> 1;
              const long long x0 = xg * blockIdx.x / G, x1 = xg * (blockIdx.x + 1) / G;
#pragma unroll 1
              for (long long i = x0 + wt; i < x1; i += kStWorkers) st_zero2(oset + 2 * i);
            }
          }
          const int nb = ((kind ? T : kv) + 127) >> 7;         // boxes are streamed for the batch maximum; a clip's own length masks
#pragma unroll 1
          for (int t = first_task(l, kind, sa.task_inv); t < ntask; t += G) {
            const int vu = t / H, h = t - vu * H;            // row (virtual utterance) and head
            const int ub = vu / NF, pi = vu - ub * NF;       // clip and position within this launch's new rows
            const int nvalid = kind ? ((!LEAN && a.t_valid) ? a.t_valid[ub] : T) : kv;
            // q of my row, and (self-attention) k / v of the clip's new rows 0..pi: the causal part that is not in the cache yet
            const int nitems = kind ? 64 : 64 + (pi + 1) * 128;
#pragma unroll 1
            for (int idx = wt; idx < nitems; idx += kStWorkers) {
              const int e = idx - 64;
              const int which = idx < 64 ? 0 : 1 + ((e >> 6) & 1), dd = idx & 63, j = idx < 64 ? pi : e >> 7;
              const int row = ub * NF + j;
              const int n = which * d + h * 64 + dd;
              const u64* p = lay + (kind ? (long long)NRT * 4 * d + (long long)row * d : (long long)row * 3 * d) + n;
              const u64* stp = stats + (kind * NRT + row) * 2;
              const unsigned ex = s_cnt[(l * 3 + kind) * sa.cnt_ld + (n >> 7)];
              const float wsn = (kind ? slr.cq_ws : slr.qkv_ws)[n], bn = (kind ? slr.cq_b : slr.qkv_b)[n];
              u64 w0, w1, wv;
              long long t0 = 0;
              for (;;) {
                ld_w2(stp, w0, w1);
                wv = ld_w(p);
                if (acc_cnt(w0) == (unsigned)KUd && acc_cnt(w1) == (unsigned)KUd && acc_cnt(wv) == ex) break;
                if (t0 == 0) t0 = clock64();
                else if (clock64() - t0 > kStSpin) st_timeout(30, (int)acc_cnt(w0), (int)acc_cnt(wv));
              }
              const float2 ms = ln_stats(w0, w1, (unsigned)KUd, inv_d, a.eps);
              const float val = fmaf(ms.y, acc_val(wv, ex) - ms.x * wsn, bn);
              if (which == 0) {
                s_qs[dd] = val;
              } else {
                const bf16 hb = __float2bfloat16_rn(val);
                if (j == pi) {                               // my own position: append for the later tokens
                  bf16* cache = reinterpret_cast<bf16*>(which == 1 ? a.kcache : a.vcache);
                  cache[((((long long)l * Bs + b0c + ub) * H + h) * a.max_target + kv + pi) * 64 + dd] = hb;
                }
                s_knv[((which - 1) * 8 + j) * 64 + dd] = __bfloat162float(hb);
              }
            }
            wbar();
            // warp ww takes rows [16 ww, 16 ww + 16) of every 128-row box; two lanes per row (32 dims each), the row's 16-byte
            // chunks rotated by the row index so the quarter-warp phases of the 128-bit loads are conflict-free
            const int rowl = lane >> 1, half = lane & 1;
            float qh[32];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int cc = (c + rowl) & 3;
              const float4 qa = *reinterpret_cast<const float4*>(s_qs + half * 32 + cc * 8);
              const float4 qb = *reinterpret_cast<const float4*>(s_qs + half * 32 + cc * 8 + 4);
              qh[c * 8 + 0] = qa.x; qh[c * 8 + 1] = qa.y; qh[c * 8 + 2] = qa.z; qh[c * 8 + 3] = qa.w;
              qh[c * 8 + 4] = qb.x; qh[c * 8 + 5] = qb.y; qh[c * 8 + 6] = qb.z; qh[c * 8 + 7] = qb.w;
            }
            float m = -INFINITY, lsum = 0.f, o0 = 0.f, o1 = 0.f;
            const int row = ww * 16 + rowl;
            float* sc_w = s_sc + ww * 64;                    // this warp's scores / probabilities: [4 boxes][16 rows]
#pragma unroll 1
            for (int g0 = 0; g0 < nb; g0 += 4) {
              // a group of <= 4 key boxes, then their value boxes.  Pass 1 only leaves raw scores in the warp's strip of shared
              // memory, so the max / exp / sum chain runs once per group instead of once per box
              const int nbg = min(4, nb - g0);
#pragma unroll 1
              for (int j = 0; j < nbg; ++j) {
                if (DBG && a.timing && wt == 0) {            // instrumented build: how long do attention stages make us wait?
                  const long long c0 = clock64();
                  const bool ready = mbar_try_wait(&full_bar[pos.stage], pos.phase);
                  swait(&full_bar[pos.stage], pos.phase, 33);
                  atomicAdd(&a.timing[a.timing_cap - 4 + kind * 2], (unsigned long long)(clock64() - c0));
                  atomicAdd(&a.timing[a.timing_cap - 3 + kind * 2], ready ? 1ull : (1ull << 32) + 1ull);
                }
                swait(&full_bar[pos.stage], pos.phase, 33);
                const bf16* kr = reinterpret_cast<const bf16*>(ring + (size_t)pos.stage * kStStage) + row * 64 + half * 32;
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const int cc = (c + rowl) & 3;
                  const uint4 u = *reinterpret_cast<const uint4*>(kr + cc * 8);
                  const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
                  float2 f = __bfloat1622float2(hh[0]); s0 = fmaf(f.x, qh[c * 8 + 0], s0); s1 = fmaf(f.y, qh[c * 8 + 1], s1);
                  f = __bfloat1622float2(hh[1]); s0 = fmaf(f.x, qh[c * 8 + 2], s0); s1 = fmaf(f.y, qh[c * 8 + 3], s1);
                  f = __bfloat1622float2(hh[2]); s0 = fmaf(f.x, qh[c * 8 + 4], s0); s1 = fmaf(f.y, qh[c * 8 + 5], s1);
                  f = __bfloat1622float2(hh[3]); s0 = fmaf(f.x, qh[c * 8 + 6], s0); s1 = fmaf(f.y, qh[c * 8 + 7], s1);
                }
                float sv = s0 + s1;
                sv += __shfl_xor_sync(0xffffffffu, sv, 1);
                if ((g0 + j) * 128 + row >= nvalid) sv = -INFINITY;
                if (half == 0) sc_w[j * 16 + rowl] = sv;
                __syncwarp();
                if (lane == 0 && atomicAdd(&box_cnt[pos.stage], 1) == kStWorkerWarps - 1) { box_cnt[pos.stage] = 0; mbar_arrive(&empty_bar[pos.stage]); }
                ring_adv(pos, 1, NS);
              }
              // the group's soft-max statistics: lanes 0..31 and 32..63 of the strip
              const float sa0 = lane < nbg * 16 ? sc_w[lane] : -INFINITY;
              const float sa1 = lane + 32 < nbg * 16 ? sc_w[lane + 32] : -INFINITY;
              const float mnew = fmaxf(m, warp_max(fmaxf(sa0, sa1)));
              const float msafe = mnew == -INFINITY ? 0.f : mnew;
              const float scl = __expf(m - msafe);
              const float p0 = __expf(sa0 - msafe), p1 = __expf(sa1 - msafe);
              __syncwarp();
              sc_w[lane] = p0; sc_w[lane + 32] = p1;
              lsum = lsum * scl + warp_sum(p0 + p1);
              o0 *= scl; o1 *= scl; m = mnew;
              __syncwarp();
#pragma unroll 1
              for (int j = 0; j < nbg; ++j) {
                swait(&full_bar[pos.stage], pos.phase, 34);
                const bf16* vb = reinterpret_cast<const bf16*>(ring + (size_t)pos.stage * kStStage) + (ww * 16) * 64 + 2 * lane;
                float oa0 = 0.f, oa1 = 0.f, ob0 = 0.f, ob1 = 0.f;
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                  const float pa_ = sc_w[j * 16 + i], pb_ = sc_w[j * 16 + i + 1];
                  const float2 va = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(vb + i * 64));
                  const float2 vb2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(vb + (i + 1) * 64));
                  oa0 = fmaf(pa_, va.x, oa0); oa1 = fmaf(pa_, va.y, oa1);
                  ob0 = fmaf(pb_, vb2.x, ob0); ob1 = fmaf(pb_, vb2.y, ob1);
                }
                o0 += oa0 + ob0; o1 += oa1 + ob1;
                __syncwarp();
                if (lane == 0 && atomicAdd(&box_cnt[pos.stage], 1) == kStWorkerWarps - 1) { box_cnt[pos.stage] = 0; mbar_arrive(&empty_bar[pos.stage]); }
                ring_adv(pos, 1, NS);
              }
            }
            float* pw = s_part + ww * kStPartLd;
            if (lane == 0) { pw[0] = m; pw[1] = lsum; }
            pw[4 + 2 * lane] = o0; pw[5 + 2 * lane] = o1;
            if (kind == 0 && ww <= pi) {                     // warp j: score of new position j (j <= pi < 8 warps)
              const float* kn = s_knv + ww * 64;
              const float sn = warp_sum(fmaf(kn[lane], s_qs[lane], kn[lane + 32] * s_qs[lane + 32]));
              if (lane == 0) s_part[kStWorkerWarps * kStPartLd + ww] = sn;
            }
            wbar();
            if (wt < 64) {
              const int nnew = kind == 0 ? pi + 1 : 0;
              float M = -INFINITY;
#pragma unroll 1
              for (int j = 0; j < nnew; ++j) M = fmaxf(M, s_part[kStWorkerWarps * kStPartLd + j]);
#pragma unroll
              for (int w = 0; w < kStWorkerWarps; ++w) M = fmaxf(M, s_part[w * kStPartLd]);
              float Lt = 0.f, o = 0.f;
#pragma unroll
              for (int w = 0; w < kStWorkerWarps; ++w) {
                const float e = __expf(s_part[w * kStPartLd] - M);
                Lt = fmaf(e, s_part[w * kStPartLd + 1], Lt);
                o = fmaf(e, s_part[w * kStPartLd + 4 + wt], o);
              }
#pragma unroll 1
              for (int j = 0; j < nnew; ++j) {
                const float e = __expf(s_part[kStWorkerWarps * kStPartLd + j] - M);
                Lt += e; o = fmaf(e, s_knv[(8 + j) * 64 + wt], o);
              }
              u64* dst = lay + (long long)NRT * (kind ? 5 : 3) * d + (long long)vu * d + h * 64 + wt;
              st_w(dst, enc_fix(o / Lt, kFixScale));
            }
            if (kind == 0) asm volatile("fence.proxy.async;" ::: "memory");   // appended cache rows -> visible to later TMA reads
            wbar();
          }
          if (kind == 0 && wt == 0) s_prog = it * L + l + 1;
          stamp();
          continue;
        }

        // ===================================================================
        // linear phase (or the head): stage the B operand, then reduce the accumulator tiles into the output words
        // ===================================================================
        const int p6 = phase_p6(ph);
        const int4 r = s_sched[is_head ? 6 * L : l * 6 + p6];
        // mode 0: x0 = embedding + position (layer 0 qkv), 1: residual words, 2: attention context words,
        //      3: fc1 words -> LN fold + GELU, 4: residual words x gamma (head)
        int KA = KAd, KU = KUd, mode, exp_row = 0, Nrows = d;     // atoms / 64-wide activation units per weight row
        const float* wscale = nullptr;                            // F8: per-row weight scale
        const u64* src = xw; long long src_ld = d;
        u64* stat_out = nullptr; const u64* stat_in = nullptr;
        const float* fold_ws = nullptr; const float* fold_b = nullptr;
        u64* dst = xw; long long dst_ld = d; const float* bias = nullptr; bool add_x0 = false;
        if (is_head) {
          mode = 4; exp_row = (L - 1) * 3 + 2; stat_out = hstats; fold_ws = a.ln_g; wscale = sa.head_s;
          if (wt == 0) {
            // sliding-window penalty ids (APPLY_PENALTY, Export_Whisper.py:318-331): active once generated >= penalty_range
            int nmax = 0;
            if (!LEAN && a.penalty_value != 1.0f && !begin_on) {
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
        } else if (p6 == 0) {
          mode = l == 0 ? 0 : 1; exp_row = l > 0 ? (l - 1) * 3 + 2 : 0; stat_out = stats;
          Nrows = 3 * d; dst = lay; dst_ld = 3 * d; wscale = slr.qkv_s;
        } else if (p6 == 1) {
          mode = 2; src = lay + (long long)NRT * 3 * d; bias = slr.out_b; add_x0 = l == 0; wscale = slr.out_s;
        } else if (p6 == 2) {
          mode = 1; exp_row = l * 3; stat_out = stats + 2 * NRT; dst = lay + (long long)NRT * 4 * d; wscale = slr.cq_s;
        } else if (p6 == 3) {
          mode = 2; src = lay + (long long)NRT * 5 * d; bias = slr.cout_b; wscale = slr.cout_s;
        } else if (p6 == 4) {
          mode = 1; exp_row = l * 3 + 1; stat_out = stats + 4 * NRT; Nrows = ffn; dst = lay + (long long)NRT * 6 * d; dst_ld = ffn;
          wscale = slr.fc1_s;
        } else {
          KA = ffn >> KSH; KU = ffn >> 6; mode = 3; wscale = slr.fc2_s; src = lay + (long long)NRT * 6 * d; src_ld = ffn; exp_row = l * 3 + 2; stat_in = stats + 4 * NRT;
          fold_ws = slr.fc1_ws; fold_b = slr.fc1_b; bias = slr.fc2_b;
        }

        // ---- B-operand staging ----
        {
          const int nat = r.y - r.x;
          const int nslot = (nat < KA ? nat : KA) * (F8 ? 2 : 1);     // in 64-wide units: an fp8 slot (128 k) is two of them
          const int ka0 = r.w * (F8 ? 2 : 1);
          // LayerNorm statistics: the k-atoms of tile 0 (head: the k-atoms whose index is one of my vocabulary tiles) are
          // reduced by whoever stages them -- exactly one contributor per atom, d / 64 contributions per word
          const int st_n = (stat_out && mode != 4 && r.z == 0) ? min(nat, KA - r.x) * (F8 ? 2 : 1) : 0;   // my units [0, st_n) are tile-0 atoms
          const int ht0 = r.z, ht1 = mode == 4 ? r.z + nat / KA : 0;                          // head: my vocabulary tiles
          float mean = 0.f, rstd = 1.f;
          bool need_stats = mode == 3 && row_ok && g_mine < nslot;
          const u64* srow = src + (long long)r_mine * src_ld;
          float ssum = 0.f, ssq = 0.f;
          unsigned nstat = 0;
#pragma unroll 1
          for (int s0 = g_mine; s0 < nslot && row_ok; s0 += U * NG) {
            u64 w[U][2];
            int kk[U];
            unsigned ex[U];
            float2 pa[U], pb[U];                              // per-unit parameters, requested before the poll
            unsigned pend = 0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int s = s0 + u * NG;
              kk[u] = 0; ex[u] = 0; pa[u] = make_float2(0.f, 0.f); pb[u] = make_float2(0.f, 0.f);
              if (s < nslot) {
                int ka = ka0 + s; if (ka >= KU) ka -= KU;
                kk[u] = ka * 64 + 2 * lane;
                if (mode == 1 || mode == 4) ex[u] = s_xexp[exp_row * sa.xt + (kk[u] >> 7)];
                else if (mode == 3) ex[u] = s_cnt[exp_row * sa.cnt_ld + (kk[u] >> 7)];
                else ex[u] = 1;
                pend |= 1u << u;
                if (mode == 0) {
                  pa[u] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(
                      reinterpret_cast<const bf16*>(a.embed) + (long long)s_tok[r_mine] * d + kk[u]));
                  pb[u] = *reinterpret_cast<const float2*>(a.pos + (long long)(kv + r_mine % NF) * d + kk[u]);
                } else if (mode >= 3) {
                  pa[u] = *reinterpret_cast<const float2*>(fold_ws + kk[u]);
                  if (mode == 3) pb[u] = *reinterpret_cast<const float2*>(fold_b + kk[u]);
                }
              }
            }
            float fv[U][2];
            if (mode != 0) {
              // two poll rounds in flight, half a round trip apart: a word that completes is seen sooner than with one
              // load-check-reload chain (the L2 round trip is most of an exchange's latency).  A unit is decoded the moment
              // either round shows it complete, so no register that still has a load in flight is read afterwards.
              unsigned todo = pend | (need_stats ? 16u : 0u);
              u64 wb[U][2], sa0 = 0, sa1 = 0, sb0 = 0, sb1 = 0;
              long long t0 = 0;
#pragma unroll
              for (int u = 0; u < U; ++u) if (todo & (1u << u)) ld_w2(srow + kk[u], w[u][0], w[u][1]);
              if (todo & 16u) ld_w2(stat_in + r_mine * 2, sa0, sa1);
              for (;;) {
#pragma unroll
                for (int u = 0; u < U; ++u) if (todo & (1u << u)) ld_w2(srow + kk[u], wb[u][0], wb[u][1]);
                if (todo & 16u) ld_w2(stat_in + r_mine * 2, sb0, sb1);
#pragma unroll
                for (int u = 0; u < U; ++u)
                  if ((todo & (1u << u)) && acc_cnt(w[u][0]) == ex[u] && acc_cnt(w[u][1]) == ex[u]) {
                    fv[u][0] = acc_val(w[u][0], ex[u]); fv[u][1] = acc_val(w[u][1], ex[u]); todo &= ~(1u << u);
                  }
                if ((todo & 16u) && acc_cnt(sa0) == (unsigned)KUd && acc_cnt(sa1) == (unsigned)KUd) {
                  const float2 ms = ln_stats(sa0, sa1, (unsigned)KUd, inv_d, a.eps);
                  mean = ms.x; rstd = ms.y; todo &= ~16u;
                }
                if (!todo) break;
#pragma unroll
                for (int u = 0; u < U; ++u) if (todo & (1u << u)) ld_w2(srow + kk[u], w[u][0], w[u][1]);
                if (todo & 16u) ld_w2(stat_in + r_mine * 2, sa0, sa1);
#pragma unroll
                for (int u = 0; u < U; ++u)
                  if ((todo & (1u << u)) && acc_cnt(wb[u][0]) == ex[u] && acc_cnt(wb[u][1]) == ex[u]) {
                    fv[u][0] = acc_val(wb[u][0], ex[u]); fv[u][1] = acc_val(wb[u][1], ex[u]); todo &= ~(1u << u);
                  }
                if ((todo & 16u) && acc_cnt(sb0) == (unsigned)KUd && acc_cnt(sb1) == (unsigned)KUd) {
                  const float2 ms = ln_stats(sb0, sb1, (unsigned)KUd, inv_d, a.eps);
                  mean = ms.x; rstd = ms.y; todo &= ~16u;
                }
                if (!todo) break;
                if (t0 == 0) t0 = clock64();
                else if (clock64() - t0 > kStSpin) st_timeout(12 + mode, kk[0], (int)todo);
              }
              need_stats = false;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (!(pend & (1u << u))) continue;
              const int s = s0 + u * NG;
              float f0, f1;
              if (mode == 0) {
                f0 = pb[u].x + pa[u].x; f1 = pb[u].y + pa[u].y;
              } else {
                f0 = fv[u][0]; f1 = fv[u][1];
              }
              const int ka = kk[u] >> 6;
              if (s < st_n || (mode == 4 && ka >= ht0 && ka < ht1)) { ssum += f0 + f1; ssq += fmaf(f0, f0, f1 * f1); nstat += 1; }
              if (mode == 3) {
                const float2 g2 = gelu2(fmaf(rstd, f0 - mean * pa[u].x, pb[u].x), fmaf(rstd, f1 - mean * pa[u].y, pb[u].y));
                f0 = g2.x; f1 = g2.y;
              } else if (mode == 4) {
                f0 *= pa[u].x; f1 *= pa[u].y;
              }
              if (F8) {
                // x = sum of four E5M2 values: rows 4r .. 4r + 3 of the slot's 16-row K-major SWIZZLE_128B tile (one byte per k);
                // unit s is the (s & 1) half of slot s >> 1
                const uint32_t q0 = split_e5m2x4(f0), q1 = split_e5m2x4(f1);
                uint8_t* slotp = bbuf + (size_t)(s >> 1) * kStSlot;
                const int colb = (s & 1) * 64 + 2 * lane, chunk = colb >> 4, within = colb & 15;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int rw = 4 * r_mine + j;
                  *reinterpret_cast<unsigned short*>(slotp + (rw >> 3) * 1024 + (rw & 7) * 128 + ((chunk ^ (rw & 7)) << 4) + within) =
                      (unsigned short)(((q0 >> (8 * j)) & 0xffu) | (((q1 >> (8 * j)) & 0xffu) << 8));
                }
              } else {
              // x = hi + lo, both bf16: rows 2r (hi) and 2r + 1 (lo) of the slot's 16-row K-major SWIZZLE_128B tile
              const __nv_bfloat162 hi = __floats2bfloat162_rn(f0, f1);
              const float2 hf = __bfloat1622float2(hi);
              const __nv_bfloat162 lo = __floats2bfloat162_rn(f0 - hf.x, f1 - hf.y);
              uint8_t* slotp = bbuf + (size_t)s * kStSlot;
              const int rh = 2 * r_mine, rl = rh + 1;
              const int chunk = lane >> 2, within = (lane & 3) * 4;
              *reinterpret_cast<__nv_bfloat162*>(slotp + (rh >> 3) * 1024 + (rh & 7) * 128 + ((chunk ^ (rh & 7)) << 4) + within) = hi;
              *reinterpret_cast<__nv_bfloat162*>(slotp + (rl >> 3) * 1024 + (rl & 7) * 128 + ((chunk ^ (rl & 7)) << 4) + within) = lo;
              }
              if (kSlotHandOver) {
                fence_proxy_async_smem();
                mbar_arrive(&slot_bar[F8 ? (s >> 1) : s]);    // this lane's share of the slot is visible to the tensor core
              }
            }
          }
          if (!kSlotHandOver && nat > 0) {                   // phases without atoms here have no MMA side to hand over to
            fence_proxy_async_smem();
            mbar_arrive(&b_ready);
            // (The residual stream is accumulated IN PLACE by out / cross-out / fc2, so those writers may only start once every
            // reader of the version before has consumed it: a reader that fell a phase behind would otherwise poll for a
            // contributor count the words have already passed -- observed as a hang once in ~500 clips under stress.  The MMA lane
            // reports the CTA as done when b_ready completes; the writers check in their epilogue.)
          }
          // statistics after the MMA has been released: the readers need them one phase later.  One RED pair per warp,
          // counting the k-atoms it covers (nstat is warp-uniform)
          if (nstat) {
            const float sv = warp_sum(ssum), qv = warp_sum(ssq);
            if (lane == 0) red_add(stat_out + r_mine * 2, enc_fix(sv, kFixScale, nstat));
            if (lane == 1) red_add(stat_out + r_mine * 2 + 1, enc_fix(qv, kSqScale, nstat));
          }
        }

        if (!is_head) {
          // ---- TMEM epilogue: row sums (hi + lo) -> fixed point -> RED into the output words.  `bias` (and x0 at layer 0) is
          //      added by the contributor that owns k-atom 0 of the tile. ----
          int tile = r.z, at = r.x;
          // in-place writers (out, cross-out, fc2): all readers of the previous version of the residual stream must be done; the
          // word is requested here, before the wait for the tensor core, so the check costs no latency in the common case
          const bool inplace = p6 == 1 || p6 == 3 || p6 == 5;
          const u64* rdp = stats + 6 * NRT + (p6 >> 1);
          const unsigned rd_ex = inplace ? (unsigned)s_xexp[n_xexp + l * 3 + (p6 >> 1)] : 0u;
          u64 rdw = (inplace && epi_warp && at < r.y) ? ld_w(rdp) : 0ull;
          bool rd_ok = !inplace;
#pragma unroll 1
          while (at < r.y) {
            const int tend = min(r.y, at + KA - (at == r.x ? r.w : 0));
            const bool desig = at != r.x || r.w == 0;
            const int n = tile * 128 + q_tm * 32 + lane;
            float bv = 0.f, x0v[RR], wsc = 1.f;
#pragma unroll
            for (int rr = 0; rr < RR; ++rr) x0v[rr] = 0.f;
            if (F8 && epi_warp && n < Nrows) wsc = wscale[n];
            if (epi_warp && desig && n < Nrows) {
              if (bias) bv = bias[n];
              if (add_x0) {
#pragma unroll
                for (int rr = 0; rr < RR; ++rr) {
                  const int rq = set_tm * 4 + rr;
                  if (rq < R) x0v[rr] = a.pos[(long long)(kv + rq % NF) * d + n] +
                                        __bfloat162float(reinterpret_cast<const bf16*>(a.embed)[(long long)s_tok[rq] * d + n]);
                }
              }
            }
            const int buf = tctr & 1;
            swait(&acc_full[buf], (uint32_t)((tctr >> 1) & 1), 20);     // every worker: the B slots are free again after this
            if (epi_warp) {
              tc_fence_after();
              // the tile's atoms alternate between the two MMA lanes' accumulators; a lane without an atom in this tile left its
              // accumulator untouched
              const int o0 = at - r.x, nt = tend - at;
              const bool use0 = nt >= 2 || (o0 & 1) == 0, use1 = nt >= 2 || (o0 & 1) == 1;
              float v[NC], v1[NC];
#pragma unroll
              for (int j = 0; j < NC; ++j) { v[j] = 0.f; v1[j] = 0.f; }
              if (F8) {
                if (use0) tmem_ld16(tmem + ((uint32_t)(q_tm * 32) << 16) + (uint32_t)(buf * 32), v);
                if (use1) tmem_ld16(tmem + ((uint32_t)(q_tm * 32) << 16) + (uint32_t)(buf * 32 + 16), v1);
              } else {
                if (use0) tmem_ld8(tmem + ((uint32_t)(q_tm * 32) << 16) + (uint32_t)(buf * 32 + set_tm * 8), v);
                if (use1) tmem_ld8(tmem + ((uint32_t)(q_tm * 32) << 16) + (uint32_t)(buf * 32 + 16 + set_tm * 8), v1);
              }
#pragma unroll
              for (int j = 0; j < NC; ++j) v[j] += v1[j];
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&acc_empty[buf]);
              if (!rd_ok) {
                if (acc_cnt(rdw) != rd_ex) rd_wait_slow(rdp, rd_ex);
                rd_ok = true;
              }
              if (n < Nrows) {
#pragma unroll
                for (int rr = 0; rr < RR; ++rr) {
                  const int rq = set_tm * 4 + rr;
                  if (rq < R) {
                    float val = 0.f;
#pragma unroll
                    for (int cc = 0; cc < CPR; ++cc) val += v[CPR * rr + cc];
                    if (F8) val *= wsc;
                    if (desig) val += bv + x0v[rr];
                    red_add(dst + (long long)rq * dst_ld + n, enc_fix(val, kFixScale));
                  }
                }
              }
            }
            // every worker warp releases the buffer (not only the ones that read it): a warp that merely waited could otherwise be
            // lapped by two more tiles on the same buffer and wait on an aliased parity
            if (!epi_warp) { __syncwarp(); if (lane == 0) mbar_arrive(&acc_empty[buf]); }
            ++tctr; at = tend; ++tile;
          }
          ring_adv(pos, r.y - r.x, NS);
          stamp();
        }
      }

      // ---- tied lm head epilogue: whole 128-row vocabulary tiles per CTA; final LayerNorm folded around the GEMM ----
      float cv = -INFINITY; int ci = 0x7fffffff;              // this thread's candidate (threads wt < NRT publish)
      if (head_on) {
        const int4 r = s_sched[6 * L];
        const int t_end = r.z + (r.y - r.x) / KAd;
        const float* wscale = sa.head_s;
        wbar();                                             // s_pen / s_pen_n visible
        const bool pen_on = !LEAN && s_pen_n > 0;
        float mean[RR], rstd[RR], bvv[RR]; int bii[RR];
#pragma unroll
        for (int rr = 0; rr < RR; ++rr) { mean[rr] = 0.f; rstd[rr] = 1.f; bvv[rr] = -INFINITY; bii[rr] = 0x7fffffff; }
        if (epi_warp && r.x < r.y) {
#pragma unroll
          for (int rr = 0; rr < RR; ++rr) {
            const int rq = set_tm * 4 + rr;
            if (rq < R && rq % NF == NF - 1) {              // only a clip's last row feeds the head
              u64 w0, w1;
              long long t0 = 0;
              for (;;) {
                ld_w2(hstats + rq * 2, w0, w1);
                if (acc_cnt(w0) == (unsigned)KUd && acc_cnt(w1) == (unsigned)KUd) break;
                if (t0 == 0) t0 = clock64();
                else if (clock64() - t0 > kStSpin) st_timeout(40, (int)acc_cnt(w0), (int)acc_cnt(w1));
              }
              const float2 ms = ln_stats(w0, w1, (unsigned)KUd, inv_d, a.eps);
              mean[rr] = ms.x; rstd[rr] = ms.y;
            }
          }
        }
#pragma unroll 1
        for (int tile = r.z; tile < t_end; ++tile) {
          const int n = tile * 128 + q_tm * 32 + lane;
          float gn = 0.f, btn = 0.f, bg = 0.f, wsc = 1.f;
          if (epi_warp && n < a.vocab) {
            gn = sa.head_g[n]; btn = sa.head_b[n] + a.suppress_bias[n];
            if (F8) wsc = wscale[n];
            if (begin_on) bg = a.begin_bias[n];
          }
          const int buf = tctr & 1;
          swait(&acc_full[buf], (uint32_t)((tctr >> 1) & 1), 42);
          if (epi_warp) {
            tc_fence_after();
            float v[NC], v1[NC];                             // head tiles are whole (KAd >= 2 atoms): both lanes contribute
            if (F8) tmem_ld16(tmem + ((uint32_t)(q_tm * 32) << 16) + (uint32_t)(buf * 32), v);
            else tmem_ld8(tmem + ((uint32_t)(q_tm * 32) << 16) + (uint32_t)(buf * 32 + set_tm * 8), v);
            if (KAd >= 2) {
              if (F8) tmem_ld16(tmem + ((uint32_t)(q_tm * 32) << 16) + (uint32_t)(buf * 32 + 16), v1);
              else tmem_ld8(tmem + ((uint32_t)(q_tm * 32) << 16) + (uint32_t)(buf * 32 + 16 + set_tm * 8), v1);
#pragma unroll
              for (int j = 0; j < NC; ++j) v[j] += v1[j];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            if (n < a.vocab) {
#pragma unroll
              for (int rr = 0; rr < RR; ++rr) {
                const int rq = set_tm * 4 + rr;
                if (rq < R && rq % NF == NF - 1) {
                  const int ub = rq / NF;
                  float dot = 0.f;
#pragma unroll
                  for (int cc = 0; cc < CPR; ++cc) dot += v[CPR * rr + cc];
                  if (F8) dot *= wsc;
                  float val = fmaf(rstd[rr], dot - mean[rr] * gn, btn);
                  if (pen_on) {
                    bool hit = false;
#pragma unroll 1
                    for (int qq = 0; qq < s_pen_n; ++qq) hit |= (s_pen[ub * 32 + qq] == n);
                    if (hit) val *= a.penalty_value;
                  }
                  if (!LEAN && a.logits) a.logits[(long long)ub * a.vocab + n] = val;
                  val += bg;
                  if (val > bvv[rr] || (val == bvv[rr] && n < bii[rr])) { bvv[rr] = val; bii[rr] = n; }
                }
              }
            }
          }
          if (!epi_warp) { __syncwarp(); if (lane == 0) mbar_arrive(&acc_empty[buf]); }
          ++tctr;
        }
        ring_adv(pos, r.y - r.x, NS);
        if (epi_warp) {
#pragma unroll
          for (int rr = 0; rr < RR; ++rr) {
            float bv = bvv[rr]; int bi = bii[rr];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
              const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
              if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            const int rq = set_tm * 4 + rr;
            if (lane == 0 && rq < NRT) { s_best[(ww * NRT + rq) * 2] = bv; s_best[(ww * NRT + rq) * 2 + 1] = __int_as_float(bi); }   // by row
          }
        }
      }
      __threadfence();                                     // the zero stores of this step before the candidate goes out
      asm volatile("fence.proxy.async;" ::: "memory");    // cache rows appended this step, read by TMA in the next one
      wbar();
      if (head_on && wt < B) {                             // utterance wt: its last row's candidates from the four warps that read it
        const int rl = wt * NF + NF - 1;
        for (int w = (rl >> 2) * 4; w < (rl >> 2) * 4 + 4; ++w) {
          const float v = s_best[(w * NRT + rl) * 2]; const int i = __float_as_int(s_best[(w * NRT + rl) * 2 + 1]);
          if (v > cv || (v == cv && i < ci)) { cv = v; ci = i; }
        }
      }
      stamp();

      // ---- per-step exchange: every CTA publishes its candidate per utterance (flag-in-data, sequence = iteration + 1),
      //      gathers all of them and reduces identically.  Also the step's full synchronisation point. ----
      {
        u64* cbuf = sa.cand + (size_t)(it & 1) * G * NRT * 2;
        const unsigned seq = (unsigned)it + 1u;
        if (wt < NRT) {
          st_w(cbuf + (size_t)(blockIdx.x * NRT + wt) * 2, ((u64)seq << 32) | (u64)__float_as_uint(cv));
          st_w(cbuf + (size_t)(blockIdx.x * NRT + wt) * 2 + 1, ((u64)seq << 32) | (u64)(unsigned)ci);
        }
        const int nw = G * NRT * 2;
#pragma unroll 1
        for (int i = wt; i < nw; i += kStWorkers) {
          u64 v = ld_w(cbuf + i);
          if ((unsigned)(v >> 32) != seq) {
            const long long t0 = clock64();
            do {
              v = ld_w(cbuf + i);
              if (clock64() - t0 > kStSpin) st_timeout(50, i, (int)(v >> 32));
            } while ((unsigned)(v >> 32) != seq);
          }
          s_cand[i] = __uint_as_float((unsigned)v);
        }
        __threadfence();
        wbar();
        if (head_on && ww < B) {
          float bv = -INFINITY; int bi = 0x7fffffff;
#pragma unroll 1
          for (int c = lane; c < G; c += 32) {
            const float v = s_cand[(c * NRT + ww) * 2];
            const int i = __float_as_int(s_cand[(c * NRT + ww) * 2 + 1]);
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
            const int b = ww;
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
              bool stopf = false;
              for (int s = 0; s < a.n_stop; ++s) stopf |= (a.stop_ids[s] == bi);
              if (stopf || a.limit <= 0) {
                s_fin[b] = 1;
              } else {
                if (g0) a.tokens[(long long)b * a.tokens_ld + gen] = bi;
                s_ngen[b] = gen + 1;
                if (gen + 1 >= a.limit) s_fin[b] = 1;
              }
            }
          }
        }
        if (head_on) step += 1;
        wbar();
      }
      ++it_done;
      stamp();
    }
    // ---- wind down: stop the producer, write the loop state back ----
    wbar();
    if (wt == 0) s_stop = 1;
    if (blockIdx.x == 0 && wt < B) {
      a.n_gen[wt] = s_ngen[wt];
      a.finished[wt] = s_fin[wt];
      a.n_save[wt] = s_nsave[wt];
    }
    if (blockIdx.x == 0 && wt == 0) {
      int done = 1;
      for (int b = 0; b < B; ++b) done &= (s_fin[b] != 0);
      if (!sa.keep_state) { a.state->kv_len = kv0 + it_done * NF; a.state->step = step; a.state->all_done = done; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 64u);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
// out[n] = sum_k bf16(W[n][k]) * vec[k]   (vec == nullptr: plain row sums) -- the LayerNorm-fold operands, computed once
__global__ void rowdot_bf16_kernel(const bf16* __restrict__ W, const float* __restrict__ vec, float* __restrict__ out, int N, int K) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= N) return;
  const bf16* wr = W + (long long)row * K;
  double acc = 0.0;
  for (int k = lane * 2; k < K; k += 64) {
    const float2 w2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(wr + k));
    acc += (double)w2.x * (double)(vec ? vec[k] : 1.f) + (double)w2.y * (double)(vec ? vec[k + 1] : 1.f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[row] = (float)acc;
}
cudaError_t launch_rowdot_bf16(const void* W, const float* vec, float* out, int N, int K, cudaStream_t st) {
  rowdot_bf16_kernel<<<(N + 7) / 8, 256, 0, st>>>(reinterpret_cast<const bf16*>(W), vec, out, N, K);
  return cudaGetLastError();
}

// ---- FP8 weight path: W[n][k] ~= scale[n] * e4m3(W8[n][k]), scale[n] = max_k |W[n][k]| / 448 (one warp per row) ----
__global__ void quant_rows_e4m3_kernel(const bf16* __restrict__ W, uint8_t* __restrict__ W8, float* __restrict__ scale, int N, int K) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= N) return;
  const bf16* wr = W + (long long)row * K;
  float amax = 0.f;
  for (int k = lane; k < K; k += 32) amax = fmaxf(amax, fabsf(__bfloat162float(wr[k])));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  const float sc = amax > 0.f ? amax / 448.0f : 1.0f;
  const float inv = 1.0f / sc;
  for (int k = lane; k < K; k += 32)
    W8[(long long)row * K + k] = (uint8_t)__nv_cvt_float_to_fp8(__bfloat162float(wr[k]) * inv, __NV_SATFINITE, __NV_E4M3);
  if (lane == 0) scale[row] = sc;
}
cudaError_t launch_quant_rows_e4m3(const void* W, void* W8, float* scale, int N, int K, cudaStream_t st) {
  quant_rows_e4m3_kernel<<<(N + 7) / 8, 256, 0, st>>>(reinterpret_cast<const bf16*>(W), reinterpret_cast<uint8_t*>(W8), scale, N, K);
  return cudaGetLastError();
}
// out[n] = scale[n] * sum_k e4m3(W8[n][k]) * vec[k]: the LayerNorm-fold operands of the quantised matrices
__global__ void rowdot_e4m3_kernel(const uint8_t* __restrict__ W8, const float* __restrict__ scale, const float* __restrict__ vec,
                                   float* __restrict__ out, int N, int K) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= N) return;
  const uint8_t* wr = W8 + (long long)row * K;
  double acc = 0.0;
  for (int k = lane; k < K; k += 32) {
    const float w = __half2float(__half(__nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)wr[k], __NV_E4M3)));
    acc += (double)w * (double)(vec ? vec[k] : 1.f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[row] = (float)(acc * (double)scale[row]);
}
cudaError_t launch_rowdot_e4m3(const void* W8, const float* scale, const float* vec, float* out, int N, int K, cudaStream_t st) {
  rowdot_e4m3_kernel<<<(N + 7) / 8, 256, 0, st>>>(reinterpret_cast<const uint8_t*>(W8), scale, vec, out, N, K);
  return cudaGetLastError();
}

static int stream_nrt(int batch) { return batch <= 1 ? 1 : (batch <= 2 ? 2 : (batch <= 4 ? 4 : 8)); }

bool stream_supported(int batch, int d, int ffn, int n_heads, int vocab, int T, int num_sms) {
  return batch >= 1 && batch <= kStreamMaxBatch && d % 64 == 0 && ffn % 64 == 0 && d == n_heads * 64 && vocab >= 2 * d && T >= 1 &&
         num_sms >= 8 && num_sms % kRingTaskMul != 0;
}

// Schedule: for every linear phase the atoms (128 rows x 64 k, tile-major then k) are dealt to the CTAs as contiguous runs
// of floor / ceil(A / G) atoms; the CTAs with the smallest cumulative load take the ceil, so every CTA's share of the
// step's read stream stays within one atom of the mean.  The head deals whole vocabulary tiles the same way.
bool stream_plan(int batch, int d, int ffn, int n_heads, int vocab, int n_layers, int T, int max_target, int num_sms,
                 StreamPlan* plan, void* sched_out, void* cnt_out, void* xexp_out, int fp8) {
  (void)T; (void)max_target; (void)n_heads;
  const int G = num_sms, L = n_layers;
  const int nrt = stream_nrt(batch);
  const int kpa = fp8 ? 128 : 64;                  // k per atom (one 128-byte swizzle row of weights)
  if (fp8 && (nrt > 4 || d % 128 || ffn % 128)) return false;
  const int KAd = d / kpa, KAf = ffn / kpa;
  auto tiles = [](int n) { return (n + 127) / 128; };
  const int rowsN[6] = {3 * d, d, d, d, ffn, d};
  const int kas[6] = {KAd, KAd, KAd, KAd, KAd, KAf};
  const int n_sched = 6 * L + 1;
  std::vector<int4> sched((size_t)G * n_sched);
  const int cnt_ld = std::max(tiles(3 * d), tiles(ffn)), xt = tiles(d);
  std::vector<unsigned char> cnt((size_t)L * 3 * cnt_ld, 0);
  std::vector<unsigned short> xexp((size_t)L * 3 * xt + (size_t)L * 3, 0);      // + [L][3]: CTAs with atoms in qkv / cq / fc1 (readers of the residual stream)
  std::vector<long long> load(G, 0);
  std::vector<int> order(G);
  std::vector<int> xcum(xt, 0);
  int max_slots = 1;
  int rot = 0;
  auto deal = [&](int A, int unit, int col, int KA, std::vector<int>& n_of) {
    // n_of[c] = units for CTA c; the `rem` least-loaded CTAs (ties: rotating start) take one more
    const int q = A / G, rem = A % G;
    for (int c = 0; c < G; ++c) { order[c] = (c + rot) % G; n_of[c] = q; }
    rot = (rot + 53) % G;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return load[x] < load[y]; });
    for (int i = 0; i < rem; ++i) n_of[order[i]] += 1;
    int start = 0;
    for (int c = 0; c < G; ++c) {
      sched[(size_t)c * n_sched + col] = make_int4(start, start + n_of[c], start / KA, start % KA);
      start += n_of[c];
      load[c] += (long long)n_of[c] * unit;
    }
  };
  std::vector<int> n_of(G);
  for (int l = 0; l < L; ++l) {
    for (int p6 = 0; p6 < 6; ++p6) {
      const int KA = kas[p6], nt = tiles(rowsN[p6]);
      deal(nt * KA, 1, l * 6 + p6, KA, n_of);
      std::vector<int> c_of(nt, 0);
      for (int c = 0; c < G; ++c) {
        const int4 r = sched[(size_t)c * n_sched + l * 6 + p6];
        if (r.y <= r.x) continue;
        for (int t = r.x / KA; t <= (r.y - 1) / KA; ++t) c_of[t] += 1;
        max_slots = std::max(max_slots, std::min(KA, r.y - r.x));
      }
      for (int t = 0; t < nt; ++t) if (c_of[t] > 255) return false;
      if (p6 == 0 || p6 == 2 || p6 == 4) {
        const int row = l * 3 + (p6 >> 1);
        int readers = 0;
        for (int c = 0; c < G; ++c) { const int4 rr = sched[(size_t)c * n_sched + l * 6 + p6]; readers += rr.y > rr.x; }
        xexp[(size_t)L * 3 * xt + row] = (unsigned short)readers;
        for (int t = 0; t < nt; ++t) cnt[(size_t)row * cnt_ld + t] = (unsigned char)c_of[t];
      } else {
        const int row = l * 3 + (p6 == 1 ? 0 : (p6 == 3 ? 1 : 2));
        for (int t = 0; t < xt; ++t) {
          xcum[t] += c_of[t];
          if (xcum[t] > 4000) return false;                 // 12-bit contributor count
          xexp[(size_t)row * xt + t] = (unsigned short)xcum[t];
        }
      }
    }
  }
  deal(tiles(vocab), KAd, 6 * L, 1, n_of);
  for (int c = 0; c < G; ++c) {            // head: whole tiles -> atom range [t0 * KAd, t1 * KAd), first tile t0, first k-atom 0
    int4& r = sched[(size_t)c * n_sched + 6 * L];
    r = make_int4(r.x * KAd, r.y * KAd, r.x, 0);
  }
  max_slots = std::max(max_slots, KAd);
  if (plan) {
    plan->nrt = nrt; plan->cnt_ld = cnt_ld; plan->xt = xt; plan->n_slots = max_slots;
    long long lw = (long long)nrt * (6 * d + ffn) + 6 * nrt + 4;        // + 3 reader-done counters
    lw = (lw + 15) / 16 * 16;
    long long xr = (long long)nrt * d + 2 * nrt;
    xr = (xr + 15) / 16 * 16;
    plan->layer_words = lw; plan->set_words = xr + (long long)L * lw;
    plan->cand_words = (size_t)2 * G * nrt * 2;
    const int n_cnt = L * 3 * cnt_ld, n_xexp = L * 3 * xt + L * 3;
    auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t fixed = 1024 /*alignment*/ + (size_t)max_slots * kStSlot + (size_t)n_sched * 16 + up16((size_t)n_cnt) +
                         up16((size_t)n_xexp * 2) +
                         4 * ((((size_t)G * nrt * 2 + 3) & ~(size_t)3) + kStWorkerWarps * kStPartLd + 8 + 64 + 1024 + (size_t)kStWorkerWarps * nrt * 2 + 16 + kStWorkerWarps * 64 + 4);
    const size_t budget = 227 * 1024 - 3072;           // static __shared__ + slack
    if (fixed + 6 * (size_t)kStStage > budget) return false;   // an attention group holds up to 4 key boxes before releasing any
    int ns = (int)((budget - fixed) / kStStage);
    if (ns > kStMaxStages) ns = kStMaxStages;
    plan->n_stages = ns;
    plan->smem_bytes = fixed + (size_t)ns * kStStage;
  }
  if (sched_out) *reinterpret_cast<std::vector<int4>*>(sched_out) = sched;
  if (cnt_out) *reinterpret_cast<std::vector<unsigned char>*>(cnt_out) = cnt;
  if (xexp_out) *reinterpret_cast<std::vector<unsigned short>*>(xexp_out) = xexp;
  return true;
}

cudaError_t launch_decoder_stream(const StreamArgs& sa_in, const CUtensorMap& cross_map, const CUtensorMap& kc_map,
                                  const CUtensorMap& vc_map, int nrt, int num_sms, size_t smem_bytes, cudaStream_t st) {
  void* fns[15] = {(void*)decoder_stream_kernel<1, false, false, false>, (void*)decoder_stream_kernel<2, false, false, false>,
                   (void*)decoder_stream_kernel<4, false, false, false>, (void*)decoder_stream_kernel<8, false, false, false>,
                   (void*)decoder_stream_kernel<1, true, false, false>, (void*)decoder_stream_kernel<2, true, false, false>,
                   (void*)decoder_stream_kernel<4, true, false, false>, (void*)decoder_stream_kernel<8, true, false, false>,
                   (void*)decoder_stream_kernel<1, false, true, false>, (void*)decoder_stream_kernel<2, false, true, false>,
                   (void*)decoder_stream_kernel<4, false, true, false>,
                   (void*)decoder_stream_kernel<1, false, false, true>, (void*)decoder_stream_kernel<2, false, false, true>,
                   (void*)decoder_stream_kernel<4, false, false, true>, (void*)decoder_stream_kernel<8, false, false, true>};
  if (sa_in.fp8 && (nrt > 4 || sa_in.m.timing)) return cudaErrorInvalidValue;
  const MegaArgs& ma = sa_in.m;
  const bool lean = sa_in.lean && !sa_in.fp8 && !ma.timing && !sa_in.multi && ma.first_n_new <= 1 && !ma.first_is_prefill && !ma.t_valid &&
                    ma.penalty_value == 1.0f && !ma.logits;
  const int cls = nrt == 1 ? 0 : (nrt == 2 ? 1 : (nrt == 4 ? 2 : 3));
  const int slot = sa_in.fp8 ? 8 + cls : (lean ? 11 + cls : cls + (ma.timing ? 4 : 0));
  void* fn = fns[slot];
  static AttrOnce attr;
  if (attr.need(slot)) {
    cudaError_t r = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 3072);
    if (r != cudaSuccess) return r;
  }
  StreamArgs sa = sa_in;
  CUtensorMap m0 = cross_map, m1 = kc_map, m2 = vc_map;
  void* params[] = {&m0, &m1, &m2, &sa};
  return cudaLaunchCooperativeKernel(fn, dim3(num_sms), dim3(kStThreads), params, smem_bytes, st);
}

}  // namespace b200asr
