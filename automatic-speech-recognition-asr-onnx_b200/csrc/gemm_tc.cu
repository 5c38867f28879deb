// bf16 tensor-core GEMM for sm_100a: TMA -> swizzled smem ring -> tcgen05.mma with
// fp32 accumulators in TMEM -> tcgen05.ld epilogue (bias, erf-GELU, fp32 residual,
// bf16/fp32 store).  Persistent, warp-specialised:
//   warp 0      TMA producer   (one elected lane)
//   warp 1      TMEM allocator + MMA issuer (one elected lane)
//   warps 2..5  epilogue (warp w owns TMEM lanes 32*(w&3)..+31)
// Three pipelines: smem full/empty (TMA<->MMA), TMEM full/empty (MMA<->epilogue,
// double-buffered accumulator so tile i's epilogue overlaps tile i+1's mainloop),
// static tile scheduler (m fastest so concurrent CTAs share the weight tile in L2).
//
// This is the engine's kernel for every Linear of the Whisper encoder
// (/root/reference/Whisper/Export_Whisper.py:428-447: conv stem as strided-view
// GEMMs, fused QKV, out_proj, fc1/fc2, fused cross-KV) in bf16 mode.
#include "common.cuh"
#include <cstdio>
#include <mutex>
#include <unordered_map>

namespace b200asr {

constexpr int BM = 128;
constexpr int BK = 64;                 // 64 bf16 = 128 B = one SWIZZLE_128B atom row
constexpr int UMMA_K = 16;
constexpr int kTcThreads = 192;
constexpr int kEpiScratchBytes = 4 * 32 * 33 * 4;
constexpr uint32_t kStageBytesA = BM * BK * 2;
constexpr int kRingBytes = 196608;

struct EpiArgs {
  void* C; int64_t ldc, sC; int c_dtype;
  const float* bias;
  const float* residual; int64_t ldr, sR;
  int act;
  int M, N, K, batch;
  int tiles_m, tiles_n;
  int a_batched, b_batched;
  int64_t sBias;
};

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (launch failure) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("b200asr gemm_tc: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"):
// start>>4 | LBO=1 (unused for swizzled K-major) | SBO=1024B>>4 (8-row group pitch) | layout=SWIZZLE_128B(2)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D=f32 (bit4), A=B=bf16 (bits 7,10), K-major both, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int BN>
struct TcCfg {
  static constexpr uint32_t kStageBytesB = BN * BK * 2;
  static constexpr uint32_t kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kStages = kRingBytes / kStageBytes;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kRingBytes + kEpiScratchBytes + 256 + 1024;   // +1024 alignment slack
};

template <int BN>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const EpiArgs e) {
  using Cfg = TcCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* scratch = reinterpret_cast<float*>(smem + kRingBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kRingBytes + kEpiScratchBytes);
  uint64_t* full_bar = bars;                   // [kStages]
  uint64_t* empty_bar = bars + kStages;        // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;    // [2]
  uint64_t* tempty_bar = tfull_bar + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (e.K + BK - 1) / BK;
  const int tiles_per_batch = e.tiles_m * e.tiles_n;
  const int num_tiles = tiles_per_batch * e.batch;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int z = tile / tiles_per_batch;
        const int rem = tile - z * tiles_per_batch;
        const int nt = rem / e.tiles_m, mt = rem - nt * e.tiles_m;
        const int m0 = mt * BM, n0 = nt * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * Cfg::kStageBytes;
          uint8_t* sb = sa + kStageBytesA;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_3d(sa, &tmA, kb * BK, m0, e.a_batched ? z : 0, &full_bar[stage]);
          tma_load_3d(sb, &tmB, kb * BK, n0, e.b_batched ? z : 0, &full_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * Cfg::kStageBytes);
          const uint64_t adesc = make_smem_desc_sw128(sa);
          const uint64_t bdesc = make_smem_desc_sw128(sa + kStageBytesA);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 elements (32 B) along K inside the 128B swizzle atom: +2 in the (addr>>4) field
            tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);          // frees the smem slot when these MMAs have read it
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull_bar[acc]);              // accumulator complete -> epilogue
        acc ^= 1; if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;                      // TMEM lane quarter this warp may read
    float* my = scratch + (warp - 2) * (32 * 33);
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int z = tile / tiles_per_batch;
      const int rem = tile - z * tiles_per_batch;
      const int nt = rem / e.tiles_m, mt = rem - nt * e.tiles_m;
      const int m0 = mt * BM + q * 32, n0 = nt * BN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int rows_here = min(32, e.M - m0);   // may be <= 0 for a ragged last M tile
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= e.N || rows_here <= 0) break;
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), r);
#pragma unroll
        for (int j = 0; j < 32; ++j) my[lane * 33 + j] = __uint_as_float(r[j]);
        __syncwarp();
        const int col = n0 + c0 + lane;
        if (col < e.N) {
          const float bv = e.bias ? e.bias[(int64_t)z * e.sBias + col] : 0.f;
          // residual may alias C (in-place residual add), so the compiler cannot hoist these loads past the
          // stores below: fetch the whole column strip first (32 independent loads in flight, not 32 round trips)
          float res[32];
          if (e.residual) {
            const float* rp = e.residual + (int64_t)z * e.sR + (int64_t)m0 * e.ldr + col;
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) res[rr] = rr < rows_here ? rp[(int64_t)rr * e.ldr] : 0.f;
          } else {
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) res[rr] = 0.f;
          }
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            if (rr < rows_here) {
              const int64_t row = m0 + rr;
              float v = my[rr * 33 + lane] + bv;
              if (e.act == kActGelu) v = gelu_erf(v);
              else if (e.act == kActRelu) v = fmaxf(v, 0.f);
              v += res[rr];
              const int64_t o = (int64_t)z * e.sC + row * e.ldc + col;
              if (e.c_dtype == kF32) reinterpret_cast<float*>(e.C)[o] = v;
              else reinterpret_cast<bf16*>(e.C)[o] = __float2bfloat16_rn(v);
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// dims {K, rows, batch}; box {64, box_rows, 1}; bf16; 128B swizzle
static bool make_tmap(CUtensorMap* tm, const void* base, int64_t K, int64_t rows, int64_t batch, int64_t ld,
                      int64_t batch_stride, int box_rows, std::string* err) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { if (err) *err = "cuTensorMapEncodeTiled entry point not found"; return false; }
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
  if (batch_stride <= 0) batch_stride = ld * rows;
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)batch_stride * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
    return false;
  }
  return true;
}

bool make_tmap_rows_sw128(CUtensorMap* tm, const void* base, int64_t cols, int64_t rows, int64_t ld, int box_rows,
                          std::string* err) {
  return make_tmap(tm, base, cols, rows, 1, ld, 0, box_rows, err);
}

bool make_tmap_2d_plain(CUtensorMap* tm, const void* base, int64_t cols, int64_t rows, int64_t ld, int box_cols,
                        int box_rows, std::string* err) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { if (err) *err = "cuTensorMapEncodeTiled entry point not found"; return false; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled (2d plain) failed with CUresult " + std::to_string((int)r);
    return false;
  }
  return true;
}

bool gemm_tc_supported(const GemmArgs& g) {
  if (g.a_dtype != kBF16 || g.b_dtype != kBF16 || g.transB) return false;
  if (g.batch_inner != 1) return false;
  if ((g.lda % 8) || (g.ldb % 8) || (g.sAo % 8) || (g.sBo % 8)) return false;
  if ((reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.B) & 15)) return false;
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return false;
  return true;
}

template <int BN>
static cudaError_t launch_bn(const GemmArgs& g, int num_sms, cudaStream_t st, std::string* err) {
  using Cfg = TcCfg<BN>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  CUtensorMap tmA, tmB;
  const bool a_batched = g.sAo != 0 && g.batch > 1;
  if (!make_tmap(&tmA, g.A, g.K, g.M, a_batched ? g.batch : 1, g.lda, g.sAo, BM, err)) return cudaErrorNotSupported;
  const bool b_batched = g.sBo != 0 && g.batch > 1;
  if (!make_tmap(&tmB, g.B, g.K, g.N, b_batched ? g.batch : 1, g.ldb, g.sBo, BN, err)) return cudaErrorNotSupported;
  EpiArgs e;
  e.C = g.C; e.ldc = g.ldc; e.sC = g.sCo; e.c_dtype = g.c_dtype;
  e.bias = g.bias; e.residual = g.residual; e.ldr = g.ldr; e.sR = g.sRo; e.act = g.act;
  e.M = g.M; e.N = g.N; e.K = g.K; e.batch = g.batch;
  e.tiles_m = (g.M + BM - 1) / BM; e.tiles_n = (g.N + BN - 1) / BN; e.a_batched = a_batched ? 1 : 0; e.b_batched = b_batched ? 1 : 0; e.sBias = g.sBias;
  const int tiles = e.tiles_m * e.tiles_n * g.batch;
  const int grid = tiles < num_sms ? tiles : num_sms;
  gemm_tc_kernel<BN><<<grid, kTcThreads, Cfg::kSmemBytes, st>>>(tmA, tmB, e);
  return cudaGetLastError();
}

cudaError_t launch_gemm_tc(const GemmArgs& g, int num_sms, cudaStream_t st, std::string* err) {
  if (!gemm_tc_supported(g)) { if (err) *err = "gemm_tc: unsupported operand layout"; return cudaErrorInvalidValue; }
  const int64_t tm = (g.M + BM - 1) / BM;
  auto tiles = [&](int bn) { return tm * ((g.N + bn - 1) / bn) * g.batch; };
  // largest N tile that still gives (nearly) one tile per SM; small problems take the narrow tile
  // up to four row tiles (one short clip): every tile is latency-bound (pipeline fill + epilogue tail), and narrow tiles
  // spread that tail over more SMs -- measured 5.06 ms vs 5.26 ms for the large-v3 encoder at batch 1
  if (g.batch == 1 && tm <= 4) return launch_bn<64>(g, num_sms, st, err);
  const int64_t want = (int64_t)num_sms * 9 / 10;
  if (tiles(256) >= want) return launch_bn<256>(g, num_sms, st, err);
  if (tiles(128) >= want) return launch_bn<128>(g, num_sms, st, err);
  return launch_bn<64>(g, num_sms, st, err);
}

}  // namespace b200asr
