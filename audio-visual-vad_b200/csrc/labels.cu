// Label generation and dataset statistics on the GPU.
//
//   clean_speech_VAD  (packages/processing/target.py:5-48):  frame energy over nfft-sample frames every hop samples
//       (conditional pad-at-end, center=False as every script passes), label = energy > 10^threshold * min(energy).
//       The reference sums np.power(frames, 2) over axis 0 of an (nfft, T) fp32 array, i.e. sequentially over the sample
//       index in fp32; the kernel keeps that order (one thread per frame) so the energies are bit-identical, and the
//       comparison runs in fp64 exactly like numpy's float32-array > float64-scalar.
//   clean_speech_IBM  (target.py:50-70):  20*log10(|X| + eps) > max - ibm_threshold over the whole utterance.
//   statistics        (scripts/create_audio_train_files.py:273-280, 365-368):  per-bin sum and sum of squares over the
//       valid frames of a batch, accumulated in fp64; mean = S/n, std = sqrt((Q - n*mean^2) / (n - 1)).
#include <math.h>

#include "common.cuh"

namespace avvad {

__global__ void frame_energy_kernel(const float* __restrict__ wave, int64_t wave_stride,
                                    const int32_t* __restrict__ n_samples, const int32_t* __restrict__ n_frames, int t_max,
                                    int nfft, int hop, float* __restrict__ energy) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= t_max) return;
  float acc = 0.f;
  if (t < n_frames[b]) {
    const float* x = wave + (int64_t)b * wave_stride;
    const int n = n_samples[b];
    const int i0 = t * hop;
    for (int i = 0; i < nfft; ++i) {  // sequential fp32 sum over the sample index, as numpy's axis-0 reduction
      const float v = (i0 + i < n) ? x[i0 + i] : 0.f;  // pad-at-end zeros
      acc = __fadd_rn(acc, __fmul_rn(v, v));
    }
  }
  energy[(int64_t)b * t_max + t] = acc;
}

// one CTA per utterance: min over the valid frames, then the threshold compare
__global__ void __launch_bounds__(256) vad_threshold_kernel(const float* __restrict__ energy,
                                                            const int32_t* __restrict__ n_frames, int t_max,
                                                            double factor, float* __restrict__ labels) {
  __shared__ float red[256];
  const int b = blockIdx.x;
  const int T = min(n_frames[b], t_max);
  const float* e = energy + (int64_t)b * t_max;
  float lo = INFINITY;
  for (int t = threadIdx.x; t < T; t += 256) lo = fminf(lo, e[t]);
  red[threadIdx.x] = lo;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = fminf(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  const double thr = factor * (double)red[0];
  for (int t = threadIdx.x; t < t_max; t += 256)
    labels[(int64_t)b * t_max + t] = (t < T && (double)e[t] > thr) ? 1.f : 0.f;
}

// IBM: pass 1 -- max of 20*log10(|X|+eps) over the utterance (bins x valid frames); pass 2 -- compare
__global__ void __launch_bounds__(256) ibm_max_kernel(const float2* __restrict__ stft, const int32_t* __restrict__ n_frames,
                                                      int t_max, int bins, float eps, float* __restrict__ db,
                                                      float* __restrict__ umax) {
  __shared__ float red[256];
  const int b = blockIdx.y;
  const int T = min(n_frames[b], t_max);
  float hi = -INFINITY;
  const int64_t base = (int64_t)b * bins * t_max;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < (int64_t)bins * t_max; i += (int64_t)gridDim.x * 256) {
    const int t = (int)(i % t_max);
    const float2 z = stft[base + i];
    const float v = 20.f * log10f(hypotf(z.x, z.y) + eps);
    db[base + i] = v;
    if (t < T) hi = fmaxf(hi, v);
  }
  red[threadIdx.x] = hi;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    // float max through the monotone int mapping (values may be negative)
    const float v = red[0];
    int iv = __float_as_int(v);
    iv = (iv >= 0) ? iv : (iv ^ 0x7FFFFFFF);
    atomicMax(reinterpret_cast<int*>(umax) + b, iv);
  }
}
__global__ void ibm_apply_kernel(const float* __restrict__ db, const float* __restrict__ umax,
                                 const int32_t* __restrict__ n_frames, int t_max, int bins, float threshold,
                                 float* __restrict__ mask) {
  const int b = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)bins * t_max) return;
  int iv = reinterpret_cast<const int*>(umax)[b];
  iv = (iv >= 0) ? iv : (iv ^ 0x7FFFFFFF);
  const float mx = __int_as_float(iv);
  const int t = (int)(i % t_max);
  const int64_t o = (int64_t)b * bins * t_max + i;
  mask[o] = (t < min(n_frames[b], t_max) && db[o] > mx - threshold) ? 1.f : 0.f;
}

// per-bin running sums over the valid frames of x (B, t_max, bins): one CTA per bin, fixed order, fp64
__global__ void __launch_bounds__(256) stats_accumulate_kernel(const float* __restrict__ x,
                                                               const int32_t* __restrict__ n_frames, int B, int t_max,
                                                               int bins, double* __restrict__ sum,
                                                               double* __restrict__ sumsq) {
  __shared__ double s1[256], s2[256];
  const int k = blockIdx.x;
  double a = 0.0, q = 0.0;
  for (int64_t r = threadIdx.x; r < (int64_t)B * t_max; r += 256) {
    const int b = (int)(r / t_max), t = (int)(r - (int64_t)b * t_max);
    if (t < n_frames[b]) {
      const double v = (double)x[r * bins + k];
      a += v;
      q += v * v;
    }
  }
  s1[threadIdx.x] = a;
  s2[threadIdx.x] = q;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      s1[threadIdx.x] += s1[threadIdx.x + s];
      s2[threadIdx.x] += s2[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    sum[k] += s1[0];
    sumsq[k] += s2[0];
  }
}
__global__ void stats_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, double n,
                                      int bins, float* __restrict__ mean, float* __restrict__ stdv) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= bins) return;
  const double m = sum[k] / n;
  mean[k] = (float)m;
  stdv[k] = (float)sqrt((1.0 / (n - 1.0)) * (sumsq[k] - n * m * m));  // the reference's "empirical std"
}

}  // namespace avvad

using namespace avvad;

extern "C" int avvad_vad_labels(const float* wave, int64_t wave_stride, const int32_t* n_samples,
                                const int32_t* n_frames, int32_t B, int32_t t_max, int32_t nfft, int32_t hop,
                                double vad_threshold, float* energy_scratch, float* labels, void* stream) {
  AVVAD_CHECK_ARG(wave && n_samples && n_frames && energy_scratch && labels, "null pointer");
  AVVAD_CHECK_ARG(B > 0 && t_max > 0 && nfft > 0 && hop > 0 && B <= 65535, "bad size");
  cudaStream_t st = (cudaStream_t)stream;
  frame_energy_kernel<<<dim3((unsigned)ceil_div(t_max, 128), B), 128, 0, st>>>(wave, wave_stride, n_samples, n_frames,
                                                                              t_max, nfft, hop, energy_scratch);
  AVVAD_LAUNCHED();
  vad_threshold_kernel<<<B, 256, 0, st>>>(energy_scratch, n_frames, t_max, pow(10.0, vad_threshold), labels);
  AVVAD_LAUNCHED();
  return AVVAD_OK;
}

extern "C" int avvad_ibm_labels(const float* stft_ft2, const int32_t* n_frames, int32_t B, int32_t t_max, int32_t bins,
                                float eps, float ibm_threshold, float* db_scratch, float* max_scratch, float* mask,
                                void* stream) {
  AVVAD_CHECK_ARG(stft_ft2 && n_frames && db_scratch && max_scratch && mask, "null pointer");
  AVVAD_CHECK_ARG(B > 0 && t_max > 0 && bins > 0 && B <= 65535, "bad size");
  cudaStream_t st = (cudaStream_t)stream;
  // -inf in the monotone int mapping: 0xFF800000 ^ 0x7FFFFFFF = 0x807FFFFF -> any finite value is larger
  AVVAD_CUDA(cudaMemsetAsync(max_scratch, 0x80, sizeof(float) * B, st));
  const int64_t total = (int64_t)bins * t_max;
  const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(total, 256), 256);
  ibm_max_kernel<<<dim3(blocks, B), 256, 0, st>>>((const float2*)stft_ft2, n_frames, t_max, bins, eps, db_scratch,
                                                  max_scratch);
  AVVAD_LAUNCHED();
  ibm_apply_kernel<<<dim3((unsigned)ceil_div(total, 256), B), 256, 0, st>>>(db_scratch, max_scratch, n_frames, t_max,
                                                                          bins, ibm_threshold, mask);
  AVVAD_LAUNCHED();
  return AVVAD_OK;
}

extern "C" int avvad_stats_accumulate(const float* x, const int32_t* n_frames, int32_t B, int32_t t_max, int32_t bins,
                                      double* sum, double* sumsq, void* stream) {
  AVVAD_CHECK_ARG(x && n_frames && sum && sumsq && B > 0 && t_max > 0 && bins > 0, "bad argument");
  stats_accumulate_kernel<<<bins, 256, 0, (cudaStream_t)stream>>>(x, n_frames, B, t_max, bins, sum, sumsq);
  AVVAD_LAUNCHED();
  return AVVAD_OK;
}

extern "C" int avvad_stats_finalize(const double* sum, const double* sumsq, double n, int32_t bins, float* mean,
                                    float* stdv, void* stream) {
  AVVAD_CHECK_ARG(sum && sumsq && mean && stdv && bins > 0 && n > 1.0, "bad argument");
  stats_finalize_kernel<<<(unsigned)ceil_div(bins, 128), 128, 0, (cudaStream_t)stream>>>(sum, sumsq, n, bins, mean, stdv);
  AVVAD_LAUNCHED();
  return AVVAD_OK;
}
