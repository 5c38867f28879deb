// Fused PCM -> reflect-pad -> windowed DFT power -> mel -> log10 front end.
//
// Replaces, for the Whisper graphs, the Conv1d-as-DFT STFT of
// /root/reference/Whisper/STFT_Process.py:224-246 (pad rule :96-102) and the
// log-mel lines of /root/reference/Whisper/Export_Whisper.py:424-427.  The DFT
// basis and the mel filterbank are model constants handed over as tensors
// ("stft_kernel", "mel_fbank"), exactly like the ONNX initialisers.
//
// HBM traffic per 8 s clip: 256 KB int16 in, 410 KB fp32 log-mel out; the
// 643 KB basis stays L2-resident.  All arithmetic is fp32 (power reaches 1e11
// in the Kaldi variants; see SURVEY "Hard parts").
#include "common.cuh"

namespace b200asr {

constexpr int kFramesPerCta = 16;
constexpr int kFrontThreads = 256;

// grid (ceil(T/16), B), block 256.
// smem: span of PCM covering 16 frames, then 16 x F power values.
template <bool kF32In>
__global__ void __launch_bounds__(kFrontThreads)
logmel_kernel(const void* __restrict__ pcm_v, int n_samples_max, const int* __restrict__ n_per_clip, int64_t pcm_stride,
              const float* __restrict__ basis_t, const float* __restrict__ fbank,
              const int* __restrict__ fb_start, const int* __restrict__ fb_len,
              int n_fft, int hop, int n_mels, int T,
              float* __restrict__ mel_raw, int* __restrict__ max_key) {
  extern __shared__ float smem[];
  const int F = n_fft / 2 + 1;
  const int span = (kFramesPerCta - 1) * hop + n_fft;
  float* xs = smem;                 // [span]
  float* pw = smem + ((span + 3) & ~3);   // [16][F]
  const int b = blockIdx.y;
  const int n_samples = n_per_clip ? n_per_clip[b] : n_samples_max;     // ragged batch: this clip's own length (reflect pad at ITS end)
  const int Tb = n_samples / hop;                                       // frames of this clip; later rows of the batch grid are padding
  const int f0 = blockIdx.x * kFramesPerCta;
  const int half = n_fft / 2;
  const int64_t start = (int64_t)f0 * hop - half;   // first original-sample index of the span

  // ---- stage PCM with reflect padding (coalesced 2-byte / 4-byte reads) ----
  for (int i = threadIdx.x; i < span; i += kFrontThreads) {
    int64_t s = start + i;
    if (s < 0) s = -s;                                   // left reflect: x[1..half] flipped
    if (s >= n_samples) s = 2 * (int64_t)n_samples - 2 - s;   // right reflect
    float v = 0.f;
    if (s >= 0 && s < n_samples) {
      if (kF32In) v = reinterpret_cast<const float*>(pcm_v)[b * pcm_stride + s];
      else v = (float)reinterpret_cast<const int16_t*>(pcm_v)[b * pcm_stride + s] * (1.0f / 32768.0f);
    }
    xs[i] = v;
  }
  __syncthreads();

  // ---- DFT: thread f accumulates re/im of bin f for the 16 frames ----
  const int nfr = min(kFramesPerCta, T - f0);
  for (int f = threadIdx.x; f < F; f += kFrontThreads) {
    float re[kFramesPerCta], im[kFramesPerCta];
#pragma unroll
    for (int r = 0; r < kFramesPerCta; ++r) { re[r] = 0.f; im[r] = 0.f; }
    const float* bc = basis_t + f;
    const float* bs = basis_t + F + f;
    const bool vec_ok = (hop % 4 == 0) && (n_fft % 4 == 0);
    int t = 0;
    if (vec_ok) {
      // 4 taps per step: one broadcast LDS.128 per frame feeds 8 FMAs (in-order over t, like a plain loop)
      for (; t < n_fft; t += 4) {
        float c[4], s[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          c[j] = __ldg(bc + (int64_t)(t + j) * 2 * F);
          s[j] = __ldg(bs + (int64_t)(t + j) * 2 * F);
        }
#pragma unroll
        for (int r = 0; r < kFramesPerCta; ++r) {
          const float4 x = *reinterpret_cast<const float4*>(xs + r * hop + t);
          re[r] = fmaf(x.x, c[0], re[r]); im[r] = fmaf(x.x, s[0], im[r]);
          re[r] = fmaf(x.y, c[1], re[r]); im[r] = fmaf(x.y, s[1], im[r]);
          re[r] = fmaf(x.z, c[2], re[r]); im[r] = fmaf(x.z, s[2], im[r]);
          re[r] = fmaf(x.w, c[3], re[r]); im[r] = fmaf(x.w, s[3], im[r]);
        }
      }
    }
    for (; t < n_fft; ++t) {
      const float c = __ldg(bc + (int64_t)t * 2 * F);
      const float s = __ldg(bs + (int64_t)t * 2 * F);
#pragma unroll
      for (int r = 0; r < kFramesPerCta; ++r) {
        const float x = xs[r * hop + t];
        re[r] = fmaf(x, c, re[r]);
        im[r] = fmaf(x, s, im[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < kFramesPerCta; ++r) pw[r * F + f] = re[r] * re[r] + im[r] * im[r];
  }
  __syncthreads();

  // ---- mel + log10, per-utterance running max ----
  float local_max = -INFINITY;
  for (int i = threadIdx.x; i < nfr * n_mels; i += kFrontThreads) {
    const int r = i / n_mels, m = i - r * n_mels;
    const int s0 = fb_start[m], len = fb_len[m];
    const float* w = fbank + (int64_t)m * F + s0;
    const float* p = pw + r * F + s0;
    float acc = 0.f;
    for (int k = 0; k < len; ++k) acc = fmaf(w[k], p[k], acc);
    const float v = log10f(fmaxf(acc, 1e-10f));
    mel_raw[((int64_t)b * T + (f0 + r)) * n_mels + m] = v;
    if (f0 + r < Tb) local_max = fmaxf(local_max, v);
  }
  local_max = warp_max(local_max);
  if ((threadIdx.x & 31) == 0 && local_max > -INFINITY) atomicMax(max_key + b, float_to_key(local_max));
}

// y = (max(x, max-8) + 4) / 4  written time-major into the zero-padded conv input
// (row 0 and row T+1 are the conv "padding=1" rows).  Export_Whisper.py:426-427.
template <typename OutT>
__global__ void mel_finalize_kernel(const float* __restrict__ mel_raw, const int* __restrict__ max_key, const int* __restrict__ n_per_clip,
                                    int hop, int T, int n_mels, OutT* __restrict__ mel_pad) {
  const int b = blockIdx.y;
  const int64_t nvalid = n_per_clip ? (int64_t)(n_per_clip[b] / hop) * n_mels : (int64_t)T * n_mels;   // rows past the clip's end stay zero
  const float floor_v = key_to_float(max_key[b]) - 8.0f;
  const int64_t n = (int64_t)T * n_mels;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = mel_raw[b * n + i];
    v = i < nvalid ? (fmaxf(v, floor_v) + 4.0f) * 0.25f : 0.f;
    mel_pad[(int64_t)b * (T + 2) * n_mels + n_mels + i] = from_f<OutT>(v);
  }
}

__global__ void fill_i32_kernel(int* p, int v, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

cudaError_t launch_fill_i32(int* p, int v, int n, cudaStream_t st) {
  fill_i32_kernel<<<(n + 255) / 256, 256, 0, st>>>(p, v, n);
  return cudaGetLastError();
}

// rows [first[b], T) of a [B][T + 2][d] time-major, zero-padded activation buffer (row 0 = left pad) set to zero: in a ragged batch
// the conv stem must see zeros past each clip's end, exactly like the `padding=1` of the reference's per-clip convolution
template <typename OutT>
__global__ void zero_tail_rows_kernel(OutT* __restrict__ buf, const int* __restrict__ n_per_clip, int hop, int T, int d) {
  const int b = blockIdx.y;
  const int first = n_per_clip[b] / hop;
  const int64_t n = (int64_t)(T - first) * d;
  OutT* p = buf + ((int64_t)b * (T + 2) + 1 + first) * d;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = from_f<OutT>(0.f);
}
cudaError_t launch_zero_tail_rows(void* buf, int dtype, const int* n_per_clip, int hop, int batch, int T, int d, cudaStream_t st) {
  dim3 grid(32, batch);
  if (dtype == kF32) zero_tail_rows_kernel<float><<<grid, 256, 0, st>>>((float*)buf, n_per_clip, hop, T, d);
  else zero_tail_rows_kernel<bf16><<<grid, 256, 0, st>>>((bf16*)buf, n_per_clip, hop, T, d);
  return cudaGetLastError();
}

cudaError_t launch_logmel(const void* pcm, int pcm_is_f32, int batch, int n_samples, int64_t pcm_stride,
                          const float* basis_t, const float* fbank, const int* fb_start, const int* fb_len,
                          int n_fft, int hop, int n_mels, float* mel_raw, int* max_key, cudaStream_t st, const int* n_per_clip) {
  const int T = n_samples / hop;
  const int F = n_fft / 2 + 1;
  const int span = (kFramesPerCta - 1) * hop + n_fft;
  const size_t smem = (((span + 3) & ~3) + (size_t)kFramesPerCta * F) * sizeof(float);
  dim3 grid((T + kFramesPerCta - 1) / kFramesPerCta, batch);
  if (pcm_is_f32)
    logmel_kernel<true><<<grid, kFrontThreads, smem, st>>>(pcm, n_samples, n_per_clip, pcm_stride, basis_t, fbank, fb_start,
                                                            fb_len, n_fft, hop, n_mels, T, mel_raw, max_key);
  else
    logmel_kernel<false><<<grid, kFrontThreads, smem, st>>>(pcm, n_samples, n_per_clip, pcm_stride, basis_t, fbank, fb_start,
                                                             fb_len, n_fft, hop, n_mels, T, mel_raw, max_key);
  return cudaGetLastError();
}

cudaError_t launch_mel_finalize(const float* mel_raw, const int* max_key, int batch, int T, int n_mels,
                                void* mel_pad, int out_dtype, cudaStream_t st, const int* n_per_clip, int hop) {
  dim3 grid((unsigned)min((int64_t)64, ((int64_t)T * n_mels + 255) / 256), batch);
  if (out_dtype == kF32)
    mel_finalize_kernel<float><<<grid, 256, 0, st>>>(mel_raw, max_key, n_per_clip, hop, T, n_mels, (float*)mel_pad);
  else
    mel_finalize_kernel<bf16><<<grid, 256, 0, st>>>(mel_raw, max_key, n_per_clip, hop, T, n_mels, (bf16*)mel_pad);
  return cudaGetLastError();
}

}  // namespace b200asr
