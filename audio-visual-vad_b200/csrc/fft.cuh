// In-shared-memory 1024-point complex FFT (radix-2 Stockham autosort), 256 threads per transform.
// Used by the audio front end (two real frames packed into one complex transform) and by the MCB
// fusion (circular convolution of two real count sketches).
#pragma once
#include "common.cuh"

namespace avvad {

constexpr int kFftN = 1024;
constexpr int kFftThreads = 256;

// exp(-2*pi*i*k/1024), k in [0,512): filled once on the host in double precision.
const float2* fft_twiddles_device();  // returns device pointer, initialising on first use (nullptr on error)

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// Transforms the 1024 complex values in `a` (shared memory); `b` is a 1024-entry scratch buffer and
// `tw` the 512-entry twiddle table (shared memory).  All kFftThreads threads of the group must call
// it; `tid` in [0,256).  Result is returned in `a` (10 ping-pong stages).  The caller must have
// synchronised after filling `a`; the function ends with a __syncthreads().
__device__ __forceinline__ void fft1024_smem(float2* __restrict__ a, float2* __restrict__ b,
                                             const float2* __restrict__ tw, int tid) {
  float2* in = a;
  float2* out = b;
#pragma unroll 1
  for (int s = 0; s < 10; ++s) {
    const int Ns = 1 << s;
    const int tw_stride = 512 >> s;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int j = tid + jj * kFftThreads;
      const int k = j & (Ns - 1);
      const float2 w = tw[k * tw_stride];
      const float2 u = in[j];
      const float2 v = cmul(in[j + 512], w);
      const int o = ((j - k) << 1) + k;
      out[o] = make_float2(u.x + v.x, u.y + v.y);
      out[o + Ns] = make_float2(u.x - v.x, u.y - v.y);
    }
    __syncthreads();
    float2* t = in;
    in = out;
    out = t;
  }
}

}  // namespace avvad
