// Fused encoder self-attention for sm_100a: softmax(Q K^T) V of one (utterance, head, 128-query tile) per CTA,
// both contractions on tcgen05 tensor cores with the accumulators in TMEM.
//
// Replaces, for the encoder layer loop of /root/reference/Whisper/Export_Whisper.py:430-437 (per-head
// softmax(q k^T) v, no mask, the d^-0.25 scale pre-folded into q and k by the exporter, :380-388), the three
// launches of the first engine version (CUDA-core Q K^T, row softmax, CUDA-core P V through a T x T score matrix
// in HBM).  Nothing T x T ever leaves the SM here:
//
//   TMA   : Q tile [128][64], K [T][64], V [T][64] straight out of the fused-QKV activation buffer [M][3d]
//           (one tensor map, SWIZZLE_128B), landing in shared memory in UMMA's canonical layouts
//   MMA 1 : S[128][T] = Q K^T   (A = Q K-major, B = K K-major), fp32 in TMEM columns [0, T)
//   warps : two threads per query row (= one TMEM lane, 32-key chunks of alternating parity): row max, exp2, row sum in
//           registers, combined through shared memory once each --
//           P written as bf16 into shared memory in the K-major SWIZZLE_128B layout, 64 keys at a time,
//           double-buffered against
//   MMA 2 : O[128][64] += P V   (A = P K-major, B = V MN-major: V stays [key][dh] exactly as the QKV GEMM
//           wrote it, no transpose), fp32 in TMEM columns [448, 512)
//   store : O / rowsum -> bf16 context rows [M][d] at column h*64
//
// Single pass (the whole score row lives in TMEM), so T <= 448 keys (8.96 s of audio); longer clips keep the
// unfused path.
#include "common.cuh"
#include "ptx.cuh"
#include <cstdio>

namespace b200asr {

constexpr int kAttSoftmaxWarps = 8;       // two warps per TMEM lane quarter: each takes every other 32-key chunk of its rows
constexpr int kAttThreads = (kAttSoftmaxWarps + 1) * 32;   // + warp 8: TMA + MMA
constexpr int kAttBM = 128;
constexpr int kAttTile = 128 * 64 * 2;    // one [128][64] bf16 tile = 16 KB

using namespace ptx;

struct AttArgs {
  bf16* ctx; int64_t ld_ctx;      // [M][d] context rows
  int T, d, n_heads;
  const int* kv_valid;            // optional [batch]: keys >= kv_valid[b] get `mask_add` added to their score
  float mask_add;                 // (the reference's additive -128 key mask, Qwen_ASR/Export_Qwen_ASR.py:768-775)
};

// DH = head dimension (64: Whisper; 128: SenseVoice / Paraformer).  A 128-wide head is two 64-column SWIZZLE_128B tiles per
// operand: the Q K^T contraction walks both (8 k16 steps), V is an MN-major B operand with two 64-wide N blocks LBO apart,
// O takes 128 TMEM columns (scores then fit T <= 384; shared memory holds K and V for T <= 256).
// MASK = the optional additive key mask (kv_valid) is compiled in; the unmasked instantiation keeps the soft-max loops lean
template <int DH, bool MASK>
__global__ void __launch_bounds__(kAttThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttArgs a) {
  constexpr int NH = DH / 64;                         // 64-column tiles per operand row
  constexpr uint32_t kOCol = 512 - DH;                // TMEM column of the O accumulator
  extern __shared__ __align__(1024) uint8_t att_smem[];
  const int T = a.T;
  const int nkb = (T + 127) / 128;                    // 128-key boxes of K and of V
  const uint32_t base_addr = smem_u32(att_smem);
  uint8_t* sQ = att_smem + ((1024u - (base_addr & 1023u)) & 1023u);      // SWIZZLE_128B atoms need 1024-byte alignment
  uint8_t* sK = sQ + NH * kAttTile;                   // [NH][nkb] tiles of [128 keys][64 dims]
  uint8_t* sV = sK + (size_t)NH * nkb * kAttTile;     // [NH][nkb] tiles of [128 keys][64 dims]
  uint8_t* sP = sV + (size_t)NH * nkb * kAttTile;     // 2 x [128][64] bf16
  __shared__ uint64_t bar_qk, bar_v, bar_s, bar_o, bar_pfull[2], bar_pempty[2];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int row0 = b * T;                             // first row of this utterance in [M][3d]
  const int m0 = mt * kAttBM;
  const int Tp = (T + 15) & ~15;                      // keys rounded up to the MMA K / N granule
  const int nblk = (Tp + 63) / 64;                    // 64-key P blocks

  if (threadIdx.x == 0) {
    mbar_init(&bar_qk, 1); mbar_init(&bar_v, 1); mbar_init(&bar_s, 1); mbar_init(&bar_o, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_pfull[i], kAttSoftmaxWarps * 32); mbar_init(&bar_pempty[i], 1); }
    mbar_fence_init();
    prefetch_tensormap(&tmQKV);
  }
  if (warp == kAttSoftmaxWarps) {
    tmem_alloc(&tmem_slot, 512u);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == kAttSoftmaxWarps) {
    if (lane == 0) {
      // ---- TMA: Q + K on one barrier (needed first), V on its own ----
      mbar_expect_tx(&bar_qk, (uint32_t)(NH * (1 + nkb) * kAttTile));
#pragma unroll
      for (int hf = 0; hf < NH; ++hf) {
        tma_load_3d(sQ + (size_t)hf * kAttTile, &tmQKV, h * DH + hf * 64, row0 + m0, 0, &bar_qk);
        for (int j = 0; j < nkb; ++j)
          tma_load_3d(sK + (size_t)(hf * nkb + j) * kAttTile, &tmQKV, a.d + h * DH + hf * 64, row0 + j * 128, 0, &bar_qk);
      }
      mbar_expect_tx(&bar_v, (uint32_t)(NH * nkb * kAttTile));
#pragma unroll
      for (int hf = 0; hf < NH; ++hf)
        for (int j = 0; j < nkb; ++j)
          tma_load_3d(sV + (size_t)(hf * nkb + j) * kAttTile, &tmQKV, 2 * a.d + h * DH + hf * 64, row0 + j * 128, 0, &bar_v);
      // ---- S = Q K^T, 128 keys per instruction group ----
      mbar_wait(&bar_qk, 0, "attention_tc");
      tc_fence_after();
      for (int n0 = 0; n0 < Tp; n0 += 128) {
        const int n = min(128, Tp - n0);
        const uint32_t idesc = idesc_bf16(kAttBM, n, 0);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) {
          const int hf = k >> 2;
          const uint64_t qdesc = smem_desc_sw128(smem_u32(sQ) + (uint32_t)hf * kAttTile);
          const uint64_t kdesc = smem_desc_sw128(smem_u32(sK) + (uint32_t)(hf * nkb) * kAttTile + (uint32_t)n0 * 128u);
          tc_mma_bf16(tmem + (uint32_t)n0, qdesc + (uint64_t)(2 * (k & 3)), kdesc + (uint64_t)(2 * (k & 3)), idesc, k != 0);
        }
      }
      tc_commit(&bar_s);
      // ---- O += P V, one 64-key block at a time ----
      mbar_wait(&bar_v, 0, "attention_tc");
      const uint32_t idesc_pv = idesc_bf16(kAttBM, DH, 1);
      const uint64_t v_lbo = (uint64_t)(((uint32_t)nkb * kAttTile) >> 4) << 16;      // pitch of the 64-wide N blocks of V
      for (int j = 0; j < nblk; ++j) {
        const int buf = j & 1;
        mbar_wait(&bar_pfull[buf], (uint32_t)((j >> 1) & 1), "attention_tc");
        tc_fence_after();
        const int ksteps = min(4, (Tp - j * 64) / 16);
        const uint64_t pdesc = smem_desc_sw128(smem_u32(sP) + (uint32_t)buf * kAttTile);
#pragma unroll 1
        for (int k = 0; k < ksteps; ++k) {
          // V rows (keys) j*64 + k*16 ..+15: two 8-key groups, 2048 bytes per step
          uint64_t vdesc = smem_desc_sw128(smem_u32(sV) + (uint32_t)(j * 64 + k * 16) * 128u);
          if (NH > 1) vdesc = (vdesc & ~(0x3FFFull << 16)) | v_lbo;
          tc_mma_bf16(tmem + kOCol, pdesc + (uint64_t)(2 * k), vdesc, idesc_pv, (j | k) != 0);
        }
        tc_commit(&bar_pempty[buf]);                   // P buffer reusable once these MMAs have read it
      }
      tc_commit(&bar_o);
    }
    __syncwarp();
  } else {
    // ---- softmax: a query row = a TMEM lane, shared by two threads (warps w and w + 4 may both read lane quarter
    //      w % 4): thread `half` takes the 32-key chunks with that parity, i.e. its half of every 64-key P block ----
    __shared__ float s_mx[2][kAttBM], s_sum[2][kAttBM];
    const int q4 = warp & 3, half = warp >> 2;
    const int r = q4 * 32 + lane;                     // 0..127
    const uint32_t lane_base = tmem + ((uint32_t)(q4 * 32) << 16);
    const int kvl = MASK ? a.kv_valid[b] : T;
    const float madd = MASK ? a.mask_add : 0.f;
    mbar_wait(&bar_s, 0, "attention_tc");
    tc_fence_after();
    float m = -INFINITY;
    for (int c0 = half * 32; c0 < T; c0 += 64) {
      uint32_t v[32];
      tmem_ld32(lane_base + (uint32_t)c0, v);
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c0 + i < T) m = fmaxf(m, __uint_as_float(v[i]) + ((MASK && c0 + i >= kvl) ? madd : 0.f));
    }
    s_mx[half][r] = m;
    asm volatile("bar.sync 1, %0;" ::"n"(kAttSoftmaxWarps * 32) : "memory");
    m = fmaxf(s_mx[0][r], s_mx[1][r]);
    const float ml2 = m * 1.4426950408889634f;
    float sum = 0.f;
    for (int j = 0; j < nblk; ++j) {
      const int buf = j & 1;
      if (j >= 2) mbar_wait(&bar_pempty[buf], (uint32_t)(((j >> 1) - 1) & 1), "attention_tc");
      uint8_t* prow = sP + (size_t)buf * kAttTile + (size_t)r * 128;
      {
        const int c0 = j * 64 + half * 32;
        uint32_t packed[16];
        if (c0 < Tp) {
          uint32_t v[32];
          tmem_ld32(lane_base + (uint32_t)c0, v);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float e0 = 0.f, e1 = 0.f;
            if (c0 + i < T) e0 = exp2f(fmaf(__uint_as_float(v[i]) + ((MASK && c0 + i >= kvl) ? madd : 0.f), 1.4426950408889634f, -ml2));
            if (c0 + i + 1 < T) e1 = exp2f(fmaf(__uint_as_float(v[i + 1]) + ((MASK && c0 + i + 1 >= kvl) ? madd : 0.f), 1.4426950408889634f, -ml2));
            sum += e0 + e1;
            const __nv_bfloat162 p2 = __floats2bfloat162_rn(e0, e1);
            packed[i >> 1] = *reinterpret_cast<const uint32_t*>(&p2);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) packed[i] = 0u;
        }
        // 32 keys = four 16-byte chunks; chunk index XOR (row & 7) = SWIZZLE_128B
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int chunk = (half * 4 + c) ^ (r & 7);
          *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();                     // generic-proxy writes -> visible to the MMA (async proxy)
      mbar_arrive(&bar_pfull[buf]);
    }
    s_sum[half][r] = sum;
    asm volatile("bar.sync 1, %0;" ::"n"(kAttSoftmaxWarps * 32) : "memory");
    // ---- O / rowsum -> context: each of the row's two threads stores 32 of the 64 output dims ----
    mbar_wait(&bar_o, 0, "attention_tc");
    tc_fence_after();
    const float inv = 1.0f / (s_sum[0][r] + s_sum[1][r]);
    const int t = m0 + r;
    bf16* dst = a.ctx + (int64_t)(row0 + t) * a.ld_ctx + h * DH + half * (DH / 2);
#pragma unroll
    for (int part = 0; part < DH / 64; ++part) {
      uint32_t o[32];
      tmem_ld32(lane_base + kOCol + half * (DH / 2) + part * 32, o);
      if (t < T) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 p2 = __floats2bfloat162_rn(__uint_as_float(o[c * 8 + 2 * i]) * inv, __uint_as_float(o[c * 8 + 2 * i + 1]) * inv);
            w[i] = *reinterpret_cast<const uint32_t*>(&p2);
          }
          *reinterpret_cast<uint4*>(dst + part * 32 + c * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kAttSoftmaxWarps) {
    tc_fence_after();
    tmem_dealloc(tmem, 512u);
  }
}

// ---------------------------------------------------------------------------
// Long sequences (T > 448, or 128-wide heads beyond 256 keys): the score row no longer fits TMEM / K and V no longer fit
// shared memory, so the keys are streamed in 128-key boxes through a 2-stage ring and the soft-max is done in two passes:
//   pass 1: S_j = Q K_j^T per box -> running row maximum (nothing else is kept);
//   pass 2: S_j again -> P_j = exp2((S_j - max) log2 e) as bf16 through swizzled shared memory -> O += P_j V_j in TMEM.
// Recomputing Q K^T costs one more K read (from L2) and 0.5 x the MMA work; it avoids rescaling the O accumulator in TMEM.
// Warps 0-7 soft-max (two threads per query row), warp 8 MMA lane, warp 9 TMA lane.
constexpr int kAttLongThreads = (kAttSoftmaxWarps + 2) * 32;

template <int DH, bool MASK>
__global__ void __launch_bounds__(kAttLongThreads, 1)
attention_tc_long_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttArgs a) {
  constexpr int NH = DH / 64;
  constexpr uint32_t kOCol = 256;                      // S buffers at TMEM columns [0,128) and [128,256), O at [256, 256 + DH)
  constexpr int kStageT = NH * kAttTile;               // one 128-key box of K (or V): NH tiles of [128][64]
  extern __shared__ __align__(1024) uint8_t att_smem[];
  const int T = a.T;
  const int nkb = (T + 127) / 128;
  const uint32_t base_addr = smem_u32(att_smem);
  uint8_t* sQ = att_smem + ((1024u - (base_addr & 1023u)) & 1023u);
  uint8_t* sK = sQ + kStageT;                          // [2 stages][NH tiles]
  uint8_t* sV = sK + 2 * kStageT;
  uint8_t* sP = sV + 2 * kStageT;                      // 2 x [128][64] bf16
  __shared__ uint64_t bar_q, bar_o, kfull[2], kempty[2], vfull[2], vempty[2], sfull[2], sempty[2], pfull[2], pempty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_mx[2][kAttBM], s_sum[2][kAttBM];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int row0 = b * T;
  const int m0 = mt * kAttBM;
  const int Tp = (T + 15) & ~15;

  if (threadIdx.x == 0) {
    mbar_init(&bar_q, 1); mbar_init(&bar_o, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kfull[i], 1); mbar_init(&kempty[i], 1); mbar_init(&vfull[i], 1); mbar_init(&vempty[i], 1);
      mbar_init(&sfull[i], 1); mbar_init(&sempty[i], kAttSoftmaxWarps);
      mbar_init(&pfull[i], kAttSoftmaxWarps * 32); mbar_init(&pempty[i], 1);
    }
    mbar_fence_init();
    prefetch_tensormap(&tmQKV);
  }
  if (warp == kAttSoftmaxWarps) tmem_alloc(&tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == kAttSoftmaxWarps + 1) {
    // ---- TMA lane: Q once, K boxes twice (both passes), V boxes in pass 2 ----
    if (lane == 0) {
      mbar_expect_tx(&bar_q, (uint32_t)kStageT);
#pragma unroll
      for (int hf = 0; hf < NH; ++hf) tma_load_3d(sQ + (size_t)hf * kAttTile, &tmQKV, h * DH + hf * 64, row0 + m0, 0, &bar_q);
      for (int i = 0; i < 2 * nkb; ++i) {
        const int j = i < nkb ? i : i - nkb, st = i & 1;
        mbar_wait(&kempty[st], (uint32_t)(((i >> 1) & 1) ^ 1), "attention_tc_long");
        mbar_expect_tx(&kfull[st], (uint32_t)kStageT);
#pragma unroll
        for (int hf = 0; hf < NH; ++hf)
          tma_load_3d(sK + (size_t)st * kStageT + (size_t)hf * kAttTile, &tmQKV, a.d + h * DH + hf * 64, row0 + j * 128, 0, &kfull[st]);
        if (i >= nkb) {
          const int vs = j & 1;
          mbar_wait(&vempty[vs], (uint32_t)(((j >> 1) & 1) ^ 1), "attention_tc_long");
          mbar_expect_tx(&vfull[vs], (uint32_t)kStageT);
#pragma unroll
          for (int hf = 0; hf < NH; ++hf)
            tma_load_3d(sV + (size_t)vs * kStageT + (size_t)hf * kAttTile, &tmQKV, 2 * a.d + h * DH + hf * 64, row0 + j * 128, 0, &vfull[vs]);
        }
      }
    }
    __syncwarp();
  } else if (warp == kAttSoftmaxWarps) {
    // ---- MMA lane ----
    if (lane == 0) {
      mbar_wait(&bar_q, 0, "attention_tc_long");
      tc_fence_after();
      const uint32_t idesc_pv = idesc_bf16(kAttBM, DH, 1);
      const uint64_t v_lbo = (uint64_t)(((uint32_t)kAttTile) >> 4) << 16;
      int g = 0;                                         // running 64-key P block counter (buffer g & 1)
      auto issue_s = [&](int i) {                        // S[i & 1] = Q K_j^T for ring index i
        const int j = i < nkb ? i : i - nkb, st = i & 1;
        const int n = min(128, Tp - j * 128);
        mbar_wait(&kfull[st], (uint32_t)((i >> 1) & 1), "attention_tc_long");
        mbar_wait(&sempty[st], (uint32_t)(((i >> 1) & 1) ^ 1), "attention_tc_long");
        tc_fence_after();
        const uint32_t idesc = idesc_bf16(kAttBM, n, 0);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) {
          const int hf = k >> 2;
          const uint64_t qdesc = smem_desc_sw128(smem_u32(sQ) + (uint32_t)hf * kAttTile);
          const uint64_t kdesc = smem_desc_sw128(smem_u32(sK) + (uint32_t)st * kStageT + (uint32_t)hf * kAttTile);
          tc_mma_bf16(tmem + (uint32_t)(st * 128), qdesc + (uint64_t)(2 * (k & 3)), kdesc + (uint64_t)(2 * (k & 3)), idesc, k != 0);
        }
        tc_commit(&sfull[st]);
        tc_commit(&kempty[st]);
      };
      auto issue_pv = [&](int j) {                       // O += P V for key box j (its 64-key blocks)
        const int vs = j & 1;
        mbar_wait(&vfull[vs], (uint32_t)((j >> 1) & 1), "attention_tc_long");
        for (int hb = 0; hb < 2; ++hb) {
          const int key0 = j * 128 + hb * 64;
          if (key0 >= Tp) break;
          const int buf = g & 1;
          mbar_wait(&pfull[buf], (uint32_t)((g >> 1) & 1), "attention_tc_long");
          tc_fence_after();
          const int ksteps = min(4, (Tp - key0) / 16);
          const uint64_t pdesc = smem_desc_sw128(smem_u32(sP) + (uint32_t)buf * kAttTile);
#pragma unroll 1
          for (int k = 0; k < ksteps; ++k) {
            uint64_t vdesc = smem_desc_sw128(smem_u32(sV) + (uint32_t)vs * kStageT + (uint32_t)(hb * 64 + k * 16) * 128u);
            if (NH > 1) vdesc = (vdesc & ~(0x3FFFull << 16)) | v_lbo;
            tc_mma_bf16(tmem + kOCol, pdesc + (uint64_t)(2 * k), vdesc, idesc_pv, (g | k) != 0);
          }
          tc_commit(&pempty[buf]);
          ++g;
        }
        tc_commit(&vempty[vs]);
      };
      for (int i = 0; i < nkb; ++i) issue_s(i);          // pass 1
      for (int j = 0; j < nkb; ++j) {                    // pass 2: S of box j goes out before P V of box j - 1
        issue_s(nkb + j);
        if (j > 0) issue_pv(j - 1);
      }
      issue_pv(nkb - 1);
      tc_commit(&bar_o);
    }
    __syncwarp();
  } else {
    // ---- soft-max warps ----
    const int q4 = warp & 3, half = warp >> 2;
    const int r = q4 * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(q4 * 32) << 16);
    const int kvl = MASK ? a.kv_valid[b] : T;
    const float madd = MASK ? a.mask_add : 0.f;
    float m = -INFINITY;
    for (int i = 0; i < nkb; ++i) {                      // pass 1: row maximum
      const int st = i & 1;
      mbar_wait(&sfull[st], (uint32_t)((i >> 1) & 1), "attention_tc_long");
      tc_fence_after();
      for (int c0 = half * 32; c0 < 128; c0 += 64) {
        const int key0 = i * 128 + c0;
        if (key0 >= T) break;
        uint32_t v[32];
        tmem_ld32(lane_base + (uint32_t)(st * 128 + c0), v);
#pragma unroll
        for (int t = 0; t < 32; ++t)
          if (key0 + t < T) m = fmaxf(m, __uint_as_float(v[t]) + ((MASK && key0 + t >= kvl) ? madd : 0.f));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sempty[st]);
    }
    s_mx[half][r] = m;
    asm volatile("bar.sync 1, %0;" ::"n"(kAttSoftmaxWarps * 32) : "memory");
    m = fmaxf(s_mx[0][r], s_mx[1][r]);
    const float ml2 = m * 1.4426950408889634f;
    float sum = 0.f;
    int g = 0;
    for (int j = 0; j < nkb; ++j) {                      // pass 2: probabilities -> P blocks
      const int i = nkb + j, st = i & 1;
      mbar_wait(&sfull[st], (uint32_t)((i >> 1) & 1), "attention_tc_long");
      tc_fence_after();
      for (int hb = 0; hb < 2; ++hb) {
        const int blk0 = j * 128 + hb * 64;
        if (blk0 >= Tp) break;
        const int buf = g & 1;
        if (g >= 2) mbar_wait(&pempty[buf], (uint32_t)(((g >> 1) - 1) & 1), "attention_tc_long");
        uint8_t* prow = sP + (size_t)buf * kAttTile + (size_t)r * 128;
        const int key0 = blk0 + half * 32;
        uint32_t packed[16];
        if (key0 < Tp) {
          uint32_t v[32];
          tmem_ld32(lane_base + (uint32_t)(st * 128 + hb * 64 + half * 32), v);
#pragma unroll
          for (int t = 0; t < 32; t += 2) {
            float e0 = 0.f, e1 = 0.f;
            if (key0 + t < T) e0 = exp2f(fmaf(__uint_as_float(v[t]) + ((MASK && key0 + t >= kvl) ? madd : 0.f), 1.4426950408889634f, -ml2));
            if (key0 + t + 1 < T) e1 = exp2f(fmaf(__uint_as_float(v[t + 1]) + ((MASK && key0 + t + 1 >= kvl) ? madd : 0.f), 1.4426950408889634f, -ml2));
            sum += e0 + e1;
            const __nv_bfloat162 p2 = __floats2bfloat162_rn(e0, e1);
            packed[t >> 1] = *reinterpret_cast<const uint32_t*>(&p2);
          }
        } else {
#pragma unroll
          for (int t = 0; t < 16; ++t) packed[t] = 0u;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int chunk = (half * 4 + c) ^ (r & 7);
          *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(&pfull[buf]);
        ++g;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sempty[st]);
    }
    s_sum[half][r] = sum;
    asm volatile("bar.sync 1, %0;" ::"n"(kAttSoftmaxWarps * 32) : "memory");
    mbar_wait(&bar_o, 0, "attention_tc_long");
    tc_fence_after();
    const float inv = 1.0f / (s_sum[0][r] + s_sum[1][r]);
    const int t = m0 + r;
    bf16* dst = a.ctx + (int64_t)(row0 + t) * a.ld_ctx + h * DH + half * (DH / 2);
#pragma unroll
    for (int part = 0; part < DH / 64; ++part) {
      uint32_t o[32];
      tmem_ld32(lane_base + kOCol + half * (DH / 2) + part * 32, o);
      if (t < T) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t w[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const __nv_bfloat162 p2 = __floats2bfloat162_rn(__uint_as_float(o[c * 8 + 2 * q]) * inv, __uint_as_float(o[c * 8 + 2 * q + 1]) * inv);
            w[q] = *reinterpret_cast<const uint32_t*>(&p2);
          }
          *reinterpret_cast<uint4*>(dst + part * 32 + c * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kAttSoftmaxWarps) {
    tc_fence_after();
    tmem_dealloc(tmem, 512u);
  }
}

// ---------------------------------------------------------------------------
static bool attention_tc_single_pass(int T, int dh) {
  if (dh == 64) return T <= 448;            // 448 score columns + 64 output columns of TMEM
  return T <= 256;                          // 128-wide heads: K and V (2 x 2 tiles per 128 keys) must fit shared memory
}

bool attention_tc_supported(int T, int d, int n_heads) {
  if (T < 1 || n_heads <= 0 || d % n_heads || (d % 8)) return false;
  const int dh = d / n_heads;
  return dh == 64 || dh == 128;             // any length: single pass when the score row fits TMEM, two-pass streaming beyond
}

// qkv: bf16 [M = batch*T][3d] (q | k | v per row), ctx: bf16 [M][d]
cudaError_t launch_attention_tc(const void* qkv, void* ctx, int batch, int T, int d, int n_heads, cudaStream_t st,
                                std::string* err, const int* kv_valid, float mask_add) {
  if (!attention_tc_supported(T, d, n_heads)) { if (err) *err = "attention_tc: unsupported shape"; return cudaErrorInvalidValue; }
  const int dh = d / n_heads, nh = dh / 64;
  const int nkb = (T + 127) / 128;
  static AttrOnce attr;
  if (attr.need()) {
    cudaError_t r = cudaFuncSetAttribute(attention_tc_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 11 * kAttTile + 1024);
    if (r == cudaSuccess) r = cudaFuncSetAttribute(attention_tc_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 11 * kAttTile + 1024);
    if (r == cudaSuccess) r = cudaFuncSetAttribute(attention_tc_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * kAttTile + 1024);
    if (r == cudaSuccess) r = cudaFuncSetAttribute(attention_tc_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * kAttTile + 1024);
    if (r == cudaSuccess) r = cudaFuncSetAttribute(attention_tc_long_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * kAttTile + 1024);
    if (r == cudaSuccess) r = cudaFuncSetAttribute(attention_tc_long_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * kAttTile + 1024);
    if (r == cudaSuccess) r = cudaFuncSetAttribute(attention_tc_long_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * kAttTile + 1024);
    if (r == cudaSuccess) r = cudaFuncSetAttribute(attention_tc_long_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * kAttTile + 1024);
    if (r != cudaSuccess) return r;
  }
  CUtensorMap tm;
  if (!make_tmap_rows_sw128(&tm, qkv, 3 * (int64_t)d, (int64_t)batch * T, 3 * (int64_t)d, 128, err)) return cudaErrorNotSupported;
  AttArgs a;
  a.ctx = reinterpret_cast<bf16*>(ctx); a.ld_ctx = d; a.T = T; a.d = d; a.n_heads = n_heads; a.kv_valid = kv_valid; a.mask_add = mask_add;
  dim3 grid((T + kAttBM - 1) / kAttBM, n_heads, batch);
  if (!attention_tc_single_pass(T, dh)) {
    const size_t smem = (size_t)(5 * nh + 2) * kAttTile + 1024;      // Q + 2 K stages + 2 V stages (nh tiles each) + 2 P blocks
    if (dh == 64) {
      if (kv_valid) attention_tc_long_kernel<64, true><<<grid, kAttLongThreads, smem, st>>>(tm, a);
      else attention_tc_long_kernel<64, false><<<grid, kAttLongThreads, smem, st>>>(tm, a);
    } else {
      if (kv_valid) attention_tc_long_kernel<128, true><<<grid, kAttLongThreads, smem, st>>>(tm, a);
      else attention_tc_long_kernel<128, false><<<grid, kAttLongThreads, smem, st>>>(tm, a);
    }
    return cudaGetLastError();
  }
  const size_t smem = (size_t)(nh + 2 * nkb * nh + 2) * kAttTile + 1024;
  if (dh == 64) {
    if (kv_valid) attention_tc_kernel<64, true><<<grid, kAttThreads, smem, st>>>(tm, a);
    else attention_tc_kernel<64, false><<<grid, kAttThreads, smem, st>>>(tm, a);
  } else {
    if (kv_valid) attention_tc_kernel<128, true><<<grid, kAttThreads, smem, st>>>(tm, a);
    else attention_tc_kernel<128, false><<<grid, kAttThreads, smem, st>>>(tm, a);
  }
  return cudaGetLastError();
}

}  // namespace b200asr
