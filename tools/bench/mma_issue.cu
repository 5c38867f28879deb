// Micro-benchmark: issue cost of small-N tcgen05.mma (M = 128, N = 16, K = 16, bf16, A and B from shared memory), as the
// streaming decode kernel issues them: per atom 4 MMAs + one commit.  Variants: one accumulator (dependent chain) vs four.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../automatic-speech-recognition-asr-onnx_b200/csrc/ptx.cuh"
using namespace b200asr::ptx;

template <int NACC, int N>
__global__ void __launch_bounds__(128, 1) k(long long* out, int atoms) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar[8], done;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (8 * 16384 + 8192) / 4; i += 128) reinterpret_cast<uint32_t*>(base)[i] = 0x3f803f80u;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); mbar_init(&done, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 256u);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = idesc_bf16(128, N);
    const uint32_t a_u = smem_u32(base), b_u = a_u + 8 * 16384;
    const long long t0 = clock64();
    for (int at = 0; at < atoms; ++at) {
      const uint64_t adesc = smem_desc_sw128(a_u + (uint32_t)(at & 7) * 16384);
      const uint64_t bdesc = smem_desc_sw128(b_u);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        tc_mma_bf16(tmem + (uint32_t)((NACC == 1 ? 0 : kk) * N), adesc + 2 * kk, bdesc + 2 * kk, idesc, at > 0 || (NACC == 1 && kk > 0));
      tc_commit(&bar[at & 7]);
    }
    const long long t1 = clock64();
    tc_commit(&done);
    mbar_wait(&done, 0, "bench");
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 256u); }
}

template <int NACC, int N>
void run(const char* name, int grid, int atoms, long long* d) {
  cudaFuncSetAttribute(k<NACC, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384 + 8192 + 1024);
  k<NACC, N><<<grid, 128, 8 * 16384 + 8192 + 1024>>>(d, atoms);
  cudaDeviceSynchronize();
  k<NACC, N><<<grid, 128, 8 * 16384 + 8192 + 1024>>>(d, atoms);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-28s grid %3d atoms %3d: issue %6lld cyc (%5.1f / mma), done %6lld cyc (%5.1f / mma)  %s\n", name, grid, atoms, h[0],
         (double)h[0] / (4.0 * atoms), h[1], (double)h[1] / (4.0 * atoms), cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  for (int grid : {1, 148}) {
    for (int atoms : {1, 4, 16, 64}) {
      run<1, 16>("N=16 one accumulator", grid, atoms, d);
      run<4, 16>("N=16 four accumulators", grid, atoms, d);
    }
    run<1, 64>("N=64 one accumulator", grid, 16, d);
    run<1, 256>("N=256 one accumulator", grid, 16, d);
  }
  return 0;
}
