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
#include "ptx.cuh"
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
  int pdl_b_early;                // launched as a programmatic dependent and B is a device-constant weight matrix
};

using namespace ptx;

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
    mbar_fence_init();
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)Cfg::kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  if (warp != 0) pdl_wait();                     // epilogue reads bias / residual and writes C; the producer waits after its B requests

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      // Programmatic dependent launch: when B is a weight matrix (never written on the device), the first ring pass of B
      // tiles is requested before waiting for the kernel that produces A -- the ring slots start out empty
      int pre = 0;
      if (e.pdl_b_early && blockIdx.x < num_tiles) {
        const int z = blockIdx.x / tiles_per_batch;
        const int rem = blockIdx.x - z * tiles_per_batch;
        const int n0 = (rem / e.tiles_m) * BN;
        pre = num_kb < kStages ? num_kb : kStages;
        for (int kb = 0; kb < pre; ++kb) {
          mbar_expect_tx(&full_bar[kb], Cfg::kStageBytes);
          tma_load_3d(smem + (size_t)kb * Cfg::kStageBytes + kStageBytesA, &tmB, kb * BK, n0, e.b_batched ? z : 0, &full_bar[kb]);
        }
      }
      pdl_wait();
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int z = tile / tiles_per_batch;
        const int rem = tile - z * tiles_per_batch;
        const int nt = rem / e.tiles_m, mt = rem - nt * e.tiles_m;
        const int m0 = mt * BM, n0 = nt * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          uint8_t* sa = smem + (size_t)stage * Cfg::kStageBytes;
          uint8_t* sb = sa + kStageBytesA;
          if (tile == (int)blockIdx.x && kb < pre) {
            tma_load_3d(sa, &tmA, kb * BK, m0, e.a_batched ? z : 0, &full_bar[stage]);
          } else {
            mbar_wait(&empty_bar[stage], phase ^ 1, "gemm_tc");
            mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            tma_load_3d(sa, &tmA, kb * BK, m0, e.a_batched ? z : 0, &full_bar[stage]);
            tma_load_3d(sb, &tmB, kb * BK, n0, e.b_batched ? z : 0, &full_bar[stage]);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_bf16(BM, BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1, "gemm_tc");
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase, "gemm_tc");
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * Cfg::kStageBytes);
          const uint64_t adesc = smem_desc_sw128(sa);
          const uint64_t bdesc = smem_desc_sw128(sa + kStageBytesA);
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
      mbar_wait(&tfull_bar[acc], acc_phase, "gemm_tc");
      tc_fence_after();
      const int rows_here = min(32, e.M - m0);   // may be <= 0 for a ragged last M tile
      // vector path: every row pointer is 16-byte aligned, so a thread (= one accumulator row) streams its 32 columns
      // with 128-bit loads / stores; otherwise (odd leading dimensions, ragged last N chunk) the transposing path below
      const bool vec_ok = ((e.ldc * (e.c_dtype == kF32 ? 4 : 2)) & 15) == 0 && ((e.sC * (e.c_dtype == kF32 ? 4 : 2)) & 15) == 0 &&
                          (reinterpret_cast<uintptr_t>(e.C) & 15) == 0 &&
                          (!e.residual || (((e.ldr * 4) & 15) == 0 && ((e.sR * 4) & 15) == 0 && (reinterpret_cast<uintptr_t>(e.residual) & 15) == 0)) &&
                          (!e.bias || ((reinterpret_cast<uintptr_t>(e.bias) & 15) == 0 && ((e.sBias * 4) & 15) == 0));
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= e.N || rows_here <= 0) break;
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), r);
        if (vec_ok && n0 + c0 + 32 <= e.N) {
          if (lane < rows_here) {
            const int64_t row = m0 + lane;
            const int colb = n0 + c0;
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (e.bias) {
              const float4* bp = reinterpret_cast<const float4*>(e.bias + (int64_t)z * e.sBias + colb);
#pragma unroll
              for (int j = 0; j < 8; ++j) { const float4 b4 = __ldg(bp + j); v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w; }
            }
            if (e.act == kActGelu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
            } else if (e.act == kActRelu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            } else if (e.act == kActGeluTanh) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gelu_tanh(v[j]);
            }
            if (e.residual) {      // may alias C (in-place residual add): all loads of the strip are issued before its stores
              const float4* rp = reinterpret_cast<const float4*>(e.residual + (int64_t)z * e.sR + row * e.ldr + colb);
              float4 r4[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) r4[j] = rp[j];
#pragma unroll
              for (int j = 0; j < 8; ++j) { v[4 * j] += r4[j].x; v[4 * j + 1] += r4[j].y; v[4 * j + 2] += r4[j].z; v[4 * j + 3] += r4[j].w; }
            }
            const int64_t o = (int64_t)z * e.sC + row * e.ldc + colb;
            if (e.c_dtype == kF32) {
              float4* cp = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.C) + o);
#pragma unroll
              for (int j = 0; j < 8; ++j) cp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
              uint4* cp = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(e.C) + o);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * j + 2 * i], v[8 * j + 2 * i + 1]);
                  w[i] = *reinterpret_cast<const uint32_t*>(&p2);
                }
                cp[j] = make_uint4(w[0], w[1], w[2], w[3]);
              }
            }
          }
          continue;
        }
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
              else if (e.act == kActGeluTanh) v = gelu_tanh(v);
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
    tmem_dealloc(tmem_base, (uint32_t)Cfg::kTmemCols);
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

bool make_tmap_rows_sw128_u8(CUtensorMap* tm, const void* base, int64_t cols, int64_t rows, int64_t ld, int box_rows,
                             std::string* err) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { if (err) *err = "cuTensorMapEncodeTiled entry point not found"; return false; }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 1};
  cuuint64_t strides[2] = {(cuuint64_t)ld, (cuuint64_t)ld * (cuuint64_t)rows};
  cuuint32_t box[3] = {128, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled (u8) failed with CUresult " + std::to_string((int)r);
    return false;
  }
  return true;
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
  static AttrOnce smem_attr;
  if (smem_attr.need()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
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
  e.pdl_b_early = g.pdl ? 1 : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = Cfg::kSmemBytes; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN>, tmA, tmB, e);
}

cudaError_t launch_gemm_tc(const GemmArgs& g, int num_sms, cudaStream_t st, std::string* err) {
  if (!gemm_tc_supported(g)) { if (err) *err = "gemm_tc: unsupported operand layout"; return cudaErrorInvalidValue; }
  const int64_t tm = (g.M + BM - 1) / BM;
  auto tiles = [&](int bn) { return tm * ((g.N + bn - 1) / bn) * g.batch; };
  // largest N tile that still gives (nearly) one tile per SM; small problems take the narrow tile
  // up to four row tiles (one short clip): every tile is latency-bound (pipeline fill + epilogue tail), and narrow tiles
  // spread that tail over more SMs -- measured 5.06 ms vs 5.26 ms for the large-v3 encoder at batch 1
  if (g.batch == 1 && tm <= 4) {
    // ... except when 128-wide tiles fill the machine in about one round (400 rows: QKV 4 x 30 = 120 tiles, fc1 160 tiles, instead
    // of 240 / 320 in two / three rounds: encoder 3.49 -> 3.33 ms; 256-wide tiles for the 5120-wide fc1 measured slower)
    if (tiles(128) <= num_sms + num_sms / 8 && tiles(128) * 4 >= (int64_t)num_sms * 3) return launch_bn<128>(g, num_sms, st, err);
    return launch_bn<64>(g, num_sms, st, err);
  }
  const int64_t want = (int64_t)num_sms * 8 / 10;      // (1600 rows x 1280 columns: 130 tiles of 128 in one round beat 260 tiles of 64 in two)
  if (tiles(256) >= want) return launch_bn<256>(g, num_sms, st, err);
  if (tiles(128) >= want) return launch_bn<128>(g, num_sms, st, err);
  return launch_bn<64>(g, num_sms, st, err);
}

}  // namespace b200asr
