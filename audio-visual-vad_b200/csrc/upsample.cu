// 30 -> 62.5 fps frame-rate conversion as an index-exact gather, fused with u8->f32 conversion,
// standardisation and the collate zero-padding.
//
// Reference semantics: scripts/create_video_train_files_upsampled.py:116-173 (ffmpeg fps filter,
// nearest-timestamp duplication; index map pinned by tests/golden/golden_upsample.npz),
// scripts/evaluate_AV_net.py:176-182, packages/utils.py:157-166.
#include "common.cuh"

namespace avvad {

__host__ __device__ __forceinline__ int upsample_src_index(int k, int n_src, int num, int den) {
  // max{i : floor(i*num/den + 1/2) <= k}  ==  (den*(2k+1) - 1) / (2*num)
  long long v = ((long long)den * (2LL * k + 1) - 1) / (2LL * num);
  return v < n_src - 1 ? (int)v : n_src - 1;
}

template <typename SrcT>
__global__ void upsample_kernel(const SrcT* __restrict__ src, const int32_t* __restrict__ n_src,
                                const int32_t* __restrict__ n_out, int f_max, int t_max, int hw, int num,
                                int den, float mean, float inv_den, int standardise, float* __restrict__ out) {
  const int b = blockIdx.z;
  const int k = blockIdx.y;
  const int F = n_src[b];
  const int T = n_out[b];
  float* o = out + ((int64_t)b * t_max + k) * hw;
  const bool live = (k < T) && F > 0;
  const SrcT* s = live ? src + ((int64_t)b * f_max + upsample_src_index(k, F, num, den)) * hw : nullptr;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
    float v = live ? (float)s[i] : 0.f;
    if (standardise) v = (v - mean) / inv_den;
    o[i] = v;
  }
}

__global__ void upsample_index_kernel(int n_src, int n_out, int num, int den, int32_t* __restrict__ out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n_out) out[k] = upsample_src_index(k, n_src, num, den);
}

}  // namespace avvad

using namespace avvad;

extern "C" int64_t avvad_upsampled_length(int64_t n_src, int32_t num, int32_t den) {
  if (n_src <= 0 || num <= 0 || den <= 0) return 0;
  return (2 * n_src * num + den) / (2 * (int64_t)den);
}

extern "C" int avvad_upsample_gather(const void* src, int src_is_f32, const int32_t* n_src, const int32_t* n_out,
                                     int32_t B, int32_t f_max, int32_t t_max, int32_t hw, int32_t num, int32_t den,
                                     float mean, float stdv, float eps, int standardise, float* out, void* stream) {
  AVVAD_CHECK_ARG(src && n_src && n_out && out, "null pointer");
  AVVAD_CHECK_ARG(B > 0 && f_max > 0 && t_max > 0 && hw > 0 && num > 0 && den > 0, "non-positive size");
  AVVAD_CHECK_ARG(t_max <= 65535 && B <= 65535, "t_max and B must be <= 65535");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)ceil_div(hw, 1024), t_max, B);
  // the reference divides by (std + eps) in fp32 (evaluate_AV_net.py:182)
  const float inv_den = stdv + eps;
  if (src_is_f32)
    upsample_kernel<float><<<grid, 256, 0, st>>>((const float*)src, n_src, n_out, f_max, t_max, hw, num, den, mean,
                                                 inv_den, standardise, out);
  else
    upsample_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)src, n_src, n_out, f_max, t_max, hw, num, den,
                                                   mean, inv_den, standardise, out);
  AVVAD_LAUNCHED();
  return AVVAD_OK;
}

extern "C" int avvad_upsample_index(int32_t n_src, int32_t n_out, int32_t num, int32_t den, int32_t* out_idx,
                                    void* stream) {
  AVVAD_CHECK_ARG(out_idx && n_src > 0 && n_out > 0 && num > 0 && den > 0, "bad argument");
  upsample_index_kernel<<<(unsigned)ceil_div(n_out, 256), 256, 0, (cudaStream_t)stream>>>(n_src, n_out, num, den,
                                                                                          out_idx);
  AVVAD_LAUNCHED();
  return AVVAD_OK;
}
