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

// Feature-level form of the same gather: out[b][k][:] = feat_src[b][src(k)][:] for k < n_out[b], else feat_pad[:]
// (the trunk feature of the collate zero frame).  8 channels per thread.
__global__ void feature_gather_kernel(const float* __restrict__ feat_src, const float* __restrict__ feat_pad,
                                      const int32_t* __restrict__ n_src, const int32_t* __restrict__ n_out, int f_max,
                                      int t_max, int C, int num, int den, int64_t total, float* __restrict__ out_f32,
                                      __nv_bfloat16* __restrict__ out_bf16, int64_t ld_bf16, int64_t col_off) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int chunks = C / 8;
  const int ch = (int)(idx % chunks) * 8;
  const int64_t row = idx / chunks;  // b * t_max + k
  const int b = (int)(row / t_max), k = (int)(row - (int64_t)b * t_max);
  const int F = n_src[b], T = n_out[b];
  const float* s = (k < T && F > 0) ? feat_src + ((int64_t)b * f_max + upsample_src_index(k, F, num, den)) * C + ch
                                    : feat_pad + ch;
  const float4 a = *reinterpret_cast<const float4*>(s);
  const float4 c = *reinterpret_cast<const float4*>(s + 4);
  if (out_f32) {
    float4* o = reinterpret_cast<float4*>(out_f32 + row * C + ch);
    o[0] = a;
    o[1] = c;
  }
  if (out_bf16) {
    __nv_bfloat16* o = out_bf16 + row * ld_bf16 + col_off + ch;  // col_off may be odd (513): scalar stores
    o[0] = __float2bfloat16_rn(a.x); o[1] = __float2bfloat16_rn(a.y); o[2] = __float2bfloat16_rn(a.z);
    o[3] = __float2bfloat16_rn(a.w); o[4] = __float2bfloat16_rn(c.x); o[5] = __float2bfloat16_rn(c.y);
    o[6] = __float2bfloat16_rn(c.z); o[7] = __float2bfloat16_rn(c.w);
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

extern "C" int avvad_feature_gather(const float* feat_src, const float* feat_pad, const int32_t* n_src,
                                    const int32_t* n_out, int32_t B, int32_t f_max, int32_t t_max, int32_t C,
                                    int32_t num, int32_t den, float* out_f32, void* out_bf16, int64_t ld_bf16,
                                    int64_t col_off, void* stream) {
  AVVAD_CHECK_ARG(feat_src && feat_pad && n_src && n_out && (out_f32 || out_bf16), "null pointer");
  AVVAD_CHECK_ARG(B > 0 && f_max > 0 && t_max > 0 && C > 0 && C % 8 == 0 && num > 0 && den > 0, "bad size");
  const int64_t total = (int64_t)B * t_max * (C / 8);
  feature_gather_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      feat_src, feat_pad, n_src, n_out, f_max, t_max, C, num, den, total, out_f32, (__nv_bfloat16*)out_bf16, ld_bf16,
      col_off);
  AVVAD_LAUNCHED();
  return AVVAD_OK;
}
