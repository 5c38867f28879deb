// Row kernels of the encoder: LayerNorm (affine-less or affine) and row softmax.
//
// The reference folds every in-layer LayerNorm affine into the next Linear
// (Export_Whisper.py:215-225), so the layer-body norms are affine-less; only
// encoder.layer_norm / decoder.layer_norm keep gamma/beta (:438, :663).
#include "common.cuh"

namespace b200asr {

// one warp per row; two-pass (mean, then centred variance) in registers.
template <typename OutT, int kMaxPerLane>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
                 const float* __restrict__ beta, OutT* __restrict__ out, int64_t ldo, int rows, int d, float eps) {
  // programmatic dependent launch (both calls are no-ops for a plain launch): let the next kernel become resident now (a
  // tcgen05 GEMM requests its weight tiles before it waits for us), and wait for the kernel that produces x
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + (int64_t)row * ldx;
  float v[kMaxPerLane];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + i * 32;
    v[i] = c < d ? xr[c] : 0.f;
    s += v[i];
  }
  const float mean = warp_sum(s) / (float)d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + i * 32;
    const float t = c < d ? v[i] - mean : 0.f;
    q += t * t;
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
  OutT* orow = out + (int64_t)row * ldo;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + i * 32;
    if (c < d) {
      float y = (v[i] - mean) * rstd;
      if (gamma) y = y * gamma[c] + beta[c];
      orow[c] = from_f<OutT>(y);
    }
  }
}

cudaError_t launch_layernorm(const float* x, int64_t ldx, const float* gamma, const float* beta, void* out,
                             int out_dtype, int64_t ldo, int rows, int d, float eps, cudaStream_t st, int pdl) {
  if (d > 32 * 64) return cudaErrorInvalidValue;
  const int wpb = 8;
  dim3 grid((rows + wpb - 1) / wpb);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(wpb * 32); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
#define LN_LAUNCH(T, N) do { cudaError_t le = cudaLaunchKernelEx(&cfg, layernorm_kernel<T, N>, x, ldx, gamma, beta, (T*)out, ldo, rows, d, eps); if (le != cudaSuccess) return le; } while (0)
  if (d <= 32 * 8) { if (out_dtype == kF32) LN_LAUNCH(float, 8); else LN_LAUNCH(bf16, 8); }
  else if (d <= 32 * 40) { if (out_dtype == kF32) LN_LAUNCH(float, 40); else LN_LAUNCH(bf16, 40); }
  else { if (out_dtype == kF32) LN_LAUNCH(float, 64); else LN_LAUNCH(bf16, 64); }
#undef LN_LAUNCH
  return cudaGetLastError();
}

// p[row] = softmax(s[row]) ; one warp per row, fp32 math (encoder self-attention; optional per-entry key count).
template <typename OutT>
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ s, OutT* __restrict__ p, int64_t rows, int cols,
                    const int* __restrict__ valid, int64_t rows_per_entry) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* sr = s + row * cols;
  // ragged batch: only the first valid[entry] columns are keys of this entry; the rest get probability 0
  const int nv = valid ? min(cols, max(1, valid[row / rows_per_entry])) : cols;
  float m = -INFINITY;
  for (int c = lane; c < nv; c += 32) m = fmaxf(m, sr[c]);
  m = warp_max(m);
  float sum = 0.f;
  for (int c = lane; c < nv; c += 32) sum += expf(sr[c] - m);
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  OutT* pr = p + row * cols;
  for (int c = lane; c < cols; c += 32) pr[c] = from_f<OutT>(c < nv ? expf(sr[c] - m) * inv : 0.f);
}

cudaError_t launch_softmax_rows(const float* s, void* p, int p_dtype, int64_t rows, int cols, cudaStream_t st,
                                const int* valid, int64_t rows_per_entry) {
  const int wpb = 8;
  const unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
  if (rows_per_entry <= 0) rows_per_entry = 1;
  if (p_dtype == kF32) softmax_rows_kernel<float><<<grid, wpb * 32, 0, st>>>(s, (float*)p, rows, cols, valid, rows_per_entry);
  else softmax_rows_kernel<bf16><<<grid, wpb * 32, 0, st>>>(s, (bf16*)p, rows, cols, valid, rows_per_entry);
  return cudaGetLastError();
}

}  // namespace b200asr
