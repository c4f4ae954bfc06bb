// DCT -> mouth-ROI image decode on the GPU: the NTCD-TIMIT `.mat` files hold 67x67 2-D DCT
// coefficients per 30 fps frame; the reference decodes them with scipy (`idct(idct(x).T).T`, type 2, norm=None),
// normalises to 8-bit grey levels and rotates by 270 degrees before handing the frames to ffmpeg
// (scripts/create_video_train_files_upsampled.py:137-162, packages/processing/video.py:5-24).  Here one CTA decodes one
// frame: both 67x67 matrix products run in fp64 out of shared memory (the oracle is float64), followed by the
// normalisation and np.rot90(., 3) written as u8 (per-frame min-max, the variant that produced the reference's shipped
// `*_upsampled.h5` files) or as fp32 (global normalisation, the script as shipped; two-pass with a workspace).
#include <math.h>

#include <mutex>

#include "common.cuh"

namespace avvad {

constexpr int kRoi = 67;
constexpr int kRoiHW = kRoi * kRoi;
constexpr int kDctThreads = 256;
constexpr size_t kDctSmem = 3 * kRoiHW * sizeof(double);

// y[k] = x[0] + 2 sum_{j>=1} x[j] cos(pi (2k+1) j / 2n): scipy.fftpack.idct(x) (type 2, norm=None) as a matrix
static const double* idct_matrix_device() {
  static double* table[kMaxDevices] = {};  // one copy per device (nn.DataParallel: several devices, one process)
  static PerDeviceOnce once;
  int cur = 0;
  if (cudaGetDevice(&cur) != cudaSuccess || cur < 0 || cur >= kMaxDevices) return nullptr;
  const cudaError_t e = once.run([cur] {
    static double h[kRoiHW];
    for (int k = 0; k < kRoi; ++k)
      for (int j = 0; j < kRoi; ++j)
        h[k * kRoi + j] = (j == 0) ? 1.0 : 2.0 * cos(M_PI * (2.0 * k + 1.0) * (double)j / (2.0 * kRoi));
    double* d = nullptr;
    cudaError_t err = cudaMalloc(&d, sizeof(h));
    if (err != cudaSuccess) return err;
    err = cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    if (err != cudaSuccess) {
      cudaFree(d);
      return err;
    }
    table[cur] = d;
    return cudaSuccess;
  });
  return e == cudaSuccess ? table[cur] : nullptr;
}

// step2 = C @ (a @ C^T)   (== idct(idct(a).T).T of the reference), left in `s2` (shared)
__device__ __forceinline__ void idct2d(const float* __restrict__ coeff, const double* __restrict__ Cg, double* a,
                                       double* C, double* s1) {
  const int tid = threadIdx.x;
  for (int i = tid; i < kRoiHW; i += kDctThreads) {
    a[i] = (double)coeff[i];
    C[i] = Cg[i];
  }
  __syncthreads();
  for (int i = tid; i < kRoiHW; i += kDctThreads) {  // s1[r][k] = sum_j a[r][j] C[k][j]
    const int r = i / kRoi, k = i - r * kRoi;
    double acc = 0.0;
    for (int j = 0; j < kRoi; ++j) acc = fma(a[r * kRoi + j], C[k * kRoi + j], acc);
    s1[i] = acc;
  }
  __syncthreads();
  for (int i = tid; i < kRoiHW; i += kDctThreads) {  // a[k][c] = sum_j C[k][j] s1[j][c]   (a is free now)
    const int k = i / kRoi, c = i - k * kRoi;
    double acc = 0.0;
    for (int j = 0; j < kRoi; ++j) acc = fma(C[k * kRoi + j], s1[j * kRoi + c], acc);
    a[i] = acc;
  }
  __syncthreads();
}

__device__ __forceinline__ void block_minmax(double& lo, double& hi, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    red[w] = lo;
    red[8 + w] = hi;
  }
  __syncthreads();
  lo = red[0];
  hi = red[8];
#pragma unroll
  for (int i = 1; i < kDctThreads / 32; ++i) {
    lo = fmin(lo, red[i]);
    hi = fmax(hi, red[8 + i]);
  }
  __syncthreads();
}

// mode 0: per-frame min-max -> u8 -> rot90(., 3); optional fp64->fp32 copy of the un-normalised decode
__global__ void __launch_bounds__(kDctThreads)
dct_roi_u8_kernel(const float* __restrict__ coeff, const double* __restrict__ Cg, uint8_t* __restrict__ out_u8,
                  float* __restrict__ idct_out) {
  extern __shared__ __align__(16) uint8_t dsm[];
  double* a = reinterpret_cast<double*>(dsm);
  double* C = a + kRoiHW;
  double* s1 = C + kRoiHW;
  __shared__ double red[16];
  const int64_t f = blockIdx.x;
  idct2d(coeff + f * kRoiHW, Cg, a, C, s1);
  double lo = INFINITY, hi = -INFINITY;
  for (int i = threadIdx.x; i < kRoiHW; i += kDctThreads) {
    lo = fmin(lo, a[i]);
    hi = fmax(hi, a[i]);
  }
  block_minmax(lo, hi, red);
  const double sc = 255.0 / (hi - lo);
  for (int i = threadIdx.x; i < kRoiHW; i += kDctThreads) {
    // np.rot90(m, 3)[r][c] = m[n-1-c][r]
    const int r = i / kRoi, c = i - r * kRoi;
    const double v = a[(kRoi - 1 - c) * kRoi + r];
    if (out_u8) {
      double q = rint((v - lo) * sc);  // np.rint: round half to even, like rint() in the default rounding mode
      q = fmin(fmax(q, 0.0), 255.0);
      out_u8[f * kRoiHW + i] = (uint8_t)q;
    }
    if (idct_out) idct_out[f * kRoiHW + i] = (float)a[i];  // un-rotated decode
  }
}

// mode 1, pass 1: decode to the fp64 workspace + per-frame global min and the largest per-row range
__global__ void __launch_bounds__(kDctThreads)
dct_roi_decode_kernel(const float* __restrict__ coeff, const double* __restrict__ Cg, double* __restrict__ work,
                      double* __restrict__ frame_min, double* __restrict__ frame_rowrange) {
  extern __shared__ __align__(16) uint8_t dsm[];
  double* a = reinterpret_cast<double*>(dsm);
  double* C = a + kRoiHW;
  double* s1 = C + kRoiHW;
  __shared__ double red[16];
  const int64_t f = blockIdx.x;
  idct2d(coeff + f * kRoiHW, Cg, a, C, s1);
  double lo = INFINITY, rr = -INFINITY;
  for (int i = threadIdx.x; i < kRoiHW; i += kDctThreads) {
    lo = fmin(lo, a[i]);
    work[f * kRoiHW + i] = a[i];
  }
  for (int r = threadIdx.x; r < kRoi; r += kDctThreads) {  // A.max(axis=-1) - A.min(axis=-1)
    double rl = a[r * kRoi], rh = rl;
    for (int c = 1; c < kRoi; ++c) {
      rl = fmin(rl, a[r * kRoi + c]);
      rh = fmax(rh, a[r * kRoi + c]);
    }
    rr = fmax(rr, rh - rl);
  }
  block_minmax(lo, rr, red);
  if (threadIdx.x == 0) {
    frame_min[f] = lo;
    frame_rowrange[f] = rr;
  }
}
__global__ void __launch_bounds__(1024) dct_roi_stats_kernel(const double* __restrict__ frame_min,
                                                             const double* __restrict__ frame_rowrange, int64_t F,
                                                             double* __restrict__ stats) {
  __shared__ double slo[1024], shi[1024];
  double lo = INFINITY, hi = -INFINITY;
  for (int64_t i = threadIdx.x; i < F; i += 1024) {
    lo = fmin(lo, frame_min[i]);
    hi = fmax(hi, frame_rowrange[i]);
  }
  slo[threadIdx.x] = lo;
  shi[threadIdx.x] = hi;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      slo[threadIdx.x] = fmin(slo[threadIdx.x], slo[threadIdx.x + s]);
      shi[threadIdx.x] = fmax(shi[threadIdx.x], shi[threadIdx.x + s]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    stats[0] = slo[0];
    stats[1] = shi[0];
  }
}
// mode 1, pass 2: (A - A.min()) / max_row_range * 255, rot90(., 3), fp32 out
__global__ void dct_roi_global_kernel(const double* __restrict__ work, const double* __restrict__ stats, int64_t F,
                                      float* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= F * kRoiHW) return;
  const int64_t f = idx / kRoiHW;
  const int i = (int)(idx - f * kRoiHW);
  const int r = i / kRoi, c = i - r * kRoi;
  const double v = work[f * kRoiHW + (kRoi - 1 - c) * kRoi + r];
  out[idx] = (float)((v - stats[0]) / stats[1] * 255.0);
}

}  // namespace avvad

using namespace avvad;

extern "C" size_t avvad_dct_roi_workspace_bytes(int64_t n_frames) {
  if (n_frames <= 0) return 0;
  return (size_t)n_frames * kRoiHW * sizeof(double) + (size_t)(2 * n_frames + 2) * sizeof(double);
}

extern "C" int avvad_dct_roi_decode(const float* dct, int64_t n_frames, int mode, uint8_t* out_u8, float* out_f32,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  AVVAD_CHECK_ARG(dct && n_frames > 0 && (mode == 0 || mode == 1), "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const double* C = idct_matrix_device();
  if (!C) {
    set_error("idct matrix allocation failed (no CUDA device?)");
    return AVVAD_ERR_CUDA;
  }
  static PerDeviceOnce once;
  AVVAD_CUDA(once.run([] {
    cudaError_t e = cudaFuncSetAttribute(dct_roi_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDctSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(dct_roi_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDctSmem);
    return e;
  }));
  if (mode == 0) {
    AVVAD_CHECK_ARG(out_u8 || out_f32, "mode 0 needs out_u8 (normalised, rotated) and/or out_f32 (raw decode)");
    dct_roi_u8_kernel<<<(unsigned)n_frames, kDctThreads, kDctSmem, st>>>(dct, C, out_u8, out_f32);
    AVVAD_LAUNCHED();
    return AVVAD_OK;
  }
  AVVAD_CHECK_ARG(out_f32 && workspace, "mode 1 needs out_f32 and a workspace");
  if (workspace_bytes < avvad_dct_roi_workspace_bytes(n_frames)) {
    set_error("dct_roi: workspace too small");
    return AVVAD_ERR_WORKSPACE;
  }
  double* work = (double*)workspace;
  double* fmin_ = work + n_frames * kRoiHW;
  double* frr = fmin_ + n_frames;
  double* stats = frr + n_frames;
  dct_roi_decode_kernel<<<(unsigned)n_frames, kDctThreads, kDctSmem, st>>>(dct, C, work, fmin_, frr);
  AVVAD_LAUNCHED();
  dct_roi_stats_kernel<<<1, 1024, 0, st>>>(fmin_, frr, n_frames, stats);
  AVVAD_LAUNCHED();
  const int64_t total = n_frames * kRoiHW;
  dct_roi_global_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(work, stats, n_frames, out_f32);
  AVVAD_LAUNCHED();
  return AVVAD_OK;
}
