// In-shared-memory 1024-point complex FFT (radix-4 Stockham autosort), 256 threads per transform.
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

// Transforms the 1024 complex values in `a` (shared memory); `b` is a 1024-entry scratch buffer and `tw` the
// 512-entry twiddle table exp(-2*pi*i*k/1024) (shared memory).  All kFftThreads threads of the group must call it;
// `tid` in [0,256).  Radix-4 Stockham autosort: five stages, one butterfly per thread and stage (half the
// shared-memory passes and barriers of the radix-2 form).  The result lands in `b` (odd number of ping-pong
// stages) and the function RETURNS the buffer holding it.  The caller must have synchronised after filling `a`;
// the function ends with a __syncthreads().
__device__ __forceinline__ float2 tw1024(const float2* __restrict__ tw, int i) {  // i in [0, 1024)
  const float2 w = tw[i & 511];
  return (i & 512) ? make_float2(-w.x, -w.y) : w;
}
__device__ __forceinline__ float2* fft1024_smem(float2* __restrict__ a, float2* __restrict__ b,
                                                const float2* __restrict__ tw, int tid) {
  float2* in = a;
  float2* out = b;
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int Ns = 1 << (2 * s);        // 1, 4, 16, 64, 256
    const int k = tid & (Ns - 1);
    const int tstep = k * (256 >> (2 * s));  // k * 1024 / (4 * Ns)
    float2 v0 = in[tid];
    float2 v1 = in[tid + 256];
    float2 v2 = in[tid + 512];
    float2 v3 = in[tid + 768];
    if (s > 0) {
      v1 = cmul(v1, tw1024(tw, tstep));
      v2 = cmul(v2, tw1024(tw, 2 * tstep));
      v3 = cmul(v3, tw1024(tw, 3 * tstep));
    }
    // 4-point DFT (forward, e^{-i...})
    const float2 b0 = make_float2(v0.x + v2.x, v0.y + v2.y);
    const float2 b1 = make_float2(v0.x - v2.x, v0.y - v2.y);
    const float2 b2 = make_float2(v1.x + v3.x, v1.y + v3.y);
    const float2 b3 = make_float2(v1.y - v3.y, v3.x - v1.x);  // (v1 - v3) * (-i)
    const int o = ((tid - k) << 2) + k;
    out[o] = make_float2(b0.x + b2.x, b0.y + b2.y);
    out[o + Ns] = make_float2(b1.x + b3.x, b1.y + b3.y);
    out[o + 2 * Ns] = make_float2(b0.x - b2.x, b0.y - b2.y);
    out[o + 3 * Ns] = make_float2(b1.x - b3.x, b1.y - b3.y);
    __syncthreads();
    float2* t = in;
    in = out;
    out = t;
  }
  return in;
}

}  // namespace avvad
