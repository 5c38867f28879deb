// CUDA-core GEMM with the engine's fused epilogue (bias, exact-erf GELU, fp32 residual).
//
// Role: (1) the fp32 "parity" precision mode -- every contraction of the Whisper
// graphs (Export_Whisper.py:428-447, 614-667) in plain fp32 FMA so logits can be
// held to the 1e-3 bar; (2) small / oddly-strided batched products (per-head
// QK^T and PV) until the fused tcgen05 attention kernel takes them; (3) the
// cross-check for gemm_tc.cu in tests.  The tensor-core path is gemm_tc.cu.
#include "common.cuh"

namespace b200asr {

constexpr int SBM = 64, SBN = 64, SBK = 16;

template <typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(GemmArgs g) {
  __shared__ float As[SBK][SBM + 4];
  __shared__ float Bs[SBK][SBN + 4];
  const int z = blockIdx.z;
  const int zo = z / g.batch_inner, zi = z - zo * g.batch_inner;
  const TA* A = reinterpret_cast<const TA*>(g.A) + zo * g.sAo + zi * g.sAi;
  const TB* B = reinterpret_cast<const TB*>(g.B) + zo * g.sBo + zi * g.sBi;
  TC* C = reinterpret_cast<TC*>(g.C) + zo * g.sCo + zi * g.sCi;
  const float* R = g.residual ? g.residual + zo * g.sRo + zi * g.sRi : nullptr;
  const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < g.K; k0 += SBK) {
    // A tile: 64 rows x 16 k; thread loads 4 elements (k fastest across threads)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = threadIdx.x + i * 256;
      const int r = idx >> 4, k = idx & 15;
      const int gm = m0 + r, gk = k0 + k;
      As[k][r] = (gm < g.M && gk < g.K) ? to_f<TA>(A[(int64_t)gm * g.lda + gk]) : 0.f;
    }
    if (!g.transB) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = threadIdx.x + i * 256;
        const int r = idx >> 4, k = idx & 15;
        const int gn = n0 + r, gk = k0 + k;
        Bs[k][r] = (gn < g.N && gk < g.K) ? to_f<TB>(B[(int64_t)gn * g.ldb + gk]) : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = threadIdx.x + i * 256;
        const int k = idx >> 6, r = idx & 63;
        const int gn = n0 + r, gk = k0 + k;
        Bs[k][r] = (gn < g.N && gk < g.K) ? to_f<TB>(B[(int64_t)gk * g.ldb + gn]) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SBK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= g.N) continue;
      float v = acc[i][j];
      if (g.bias) v += g.bias[(int64_t)z * g.sBias + gn];
      if (g.act == kActGelu) v = gelu_erf(v);
      else if (g.act == kActRelu) v = fmaxf(v, 0.f);
      else if (g.act == kActGeluTanh) v = gelu_tanh(v);
      if (R) v += R[(int64_t)gm * g.ldr + gn];
      C[(int64_t)gm * g.ldc + gn] = from_f<TC>(v);
    }
  }
}

cudaError_t launch_gemm_simt(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0 || g.batch <= 0) return cudaSuccess;
  dim3 grid((g.N + SBN - 1) / SBN, (g.M + SBM - 1) / SBM, g.batch);
  if (g.a_dtype != g.b_dtype) return cudaErrorInvalidValue;
  if (g.a_dtype == kF32 && g.c_dtype == kF32) gemm_simt_kernel<float, float, float><<<grid, 256, 0, st>>>(g);
  else if (g.a_dtype == kF32 && g.c_dtype == kBF16) gemm_simt_kernel<float, float, bf16><<<grid, 256, 0, st>>>(g);
  else if (g.a_dtype == kBF16 && g.c_dtype == kF32) gemm_simt_kernel<bf16, bf16, float><<<grid, 256, 0, st>>>(g);
  else gemm_simt_kernel<bf16, bf16, bf16><<<grid, 256, 0, st>>>(g);
  return cudaGetLastError();
}

}  // namespace b200asr
