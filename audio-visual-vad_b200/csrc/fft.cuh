// In-shared-memory 1024-point complex FFT (radix-4 Stockham autosort), 256 threads per transform.
// Used by the audio front end (two real frames packed into one complex transform) and by the MCB
// fusion (circular convolution of two real count sketches).
#pragma once
#include "common.cuh"

namespace avvad {

constexpr int kFftN = 1024;
constexpr int kFftThreads = 256;

// Twiddle tables, filled once on the host in double precision.  Device layout (float2 entries):
//   [0, 512)      exp(-2*pi*i*k/1024), k in [0,512)  (the front end derives the periodic Hann window from it)
//   [512, 1536)   per-stage tables of the radix-4 Stockham FFT (kFftTwStage entries, 1020 used): stage s = 1..4 with
//                 Ns = 4^s holds w^1[Ns], w^2[Ns], w^3[Ns] with w = exp(-2*pi*i*k/(4*Ns)), k in [0,Ns), at
//                 fft_tw_off(s).  A thread of stage s reads entry k = tid & (Ns-1) of each: consecutive lanes read
//                 consecutive entries (or the same one), so the loads are bank-conflict free.  Indexing one 512-entry
//                 table with k*(256>>2s) instead put all lanes of a warp on 1-4 banks (ncu: 4.8-way conflicts on the
//                 shared loads, 64 % of all load wavefronts of the MCB kernel).
#ifndef AVVAD_FFT_ROT
#define AVVAD_FFT_ROT 1
#endif
constexpr int kFftTwHann = 512;
constexpr int kFftTwStage = 1024;
//   [1536, 1792)  stage-B table of the register FFT (fft_reg.cuh): exp(-2*pi*i*r*k/256) at [r*16 + k], r, k in [0,16)
constexpr int kFftTwB = 256;
constexpr int kFftTwTotal = kFftTwHann + kFftTwStage + kFftTwB;
__host__ __device__ __forceinline__ constexpr int fft_tw_off(int s) { return s == 1 ? 0 : s == 2 ? 12 : s == 3 ? 60 : 252; }
const float2* fft_twiddles_device();  // returns device pointer, initialising on first use (nullptr on error)

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// Transforms the 1024 complex values in `a` (shared memory); `b` is a 1024-entry scratch buffer and `tw` the
// kFftTwStage-entry per-stage twiddle table (shared memory copy of fft_twiddles_device() + kFftTwHann).  All
// kFftThreads threads of the group must call it; `tid` in [0,256).  Radix-4 Stockham autosort: five stages, one
// butterfly per thread and stage (half the shared-memory passes and barriers of the radix-2 form).  The result lands
// in `b` (odd number of ping-pong stages) and the function RETURNS the buffer holding it.  The caller must have
// synchronised after filling `a` and `tw`; the function ends with a __syncthreads().
__device__ __forceinline__ float2* fft1024_smem(float2* __restrict__ a, float2* __restrict__ b,
                                                const float2* __restrict__ tw, int tid) {
  float2* in = a;
  float2* out = b;
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int Ns = 1 << (2 * s);        // 1, 4, 16, 64, 256
    const int k = tid & (Ns - 1);
    float2 v0 = in[tid];
    float2 v1 = in[tid + 256];
    float2 v2 = in[tid + 512];
    float2 v3 = in[tid + 768];
    if (s > 0) {
      const float2* ts = tw + fft_tw_off(s) + k;
      v1 = cmul(v1, ts[0]);
      v2 = cmul(v2, ts[Ns]);
      v3 = cmul(v3, ts[2 * Ns]);
    }
    // 4-point DFT (forward, e^{-i...})
    const float2 b0 = make_float2(v0.x + v2.x, v0.y + v2.y);
    const float2 b1 = make_float2(v0.x - v2.x, v0.y - v2.y);
    const float2 b2 = make_float2(v1.x + v3.x, v1.y + v3.y);
    const float2 b3 = make_float2(v1.y - v3.y, v3.x - v1.x);  // (v1 - v3) * (-i)
    const int o = ((tid - k) << 2) + k;
    const float2 r0 = make_float2(b0.x + b2.x, b0.y + b2.y);
    const float2 r1 = make_float2(b1.x + b3.x, b1.y + b3.y);
    const float2 r2 = make_float2(b0.x - b2.x, b0.y - b2.y);
    const float2 r3 = make_float2(b1.x - b3.x, b1.y - b3.y);
    if (AVVAD_FFT_ROT && s < 2) {
      // Stages 0 and 1 scatter with a stride of 4 / 16 elements: for a fixed q all lanes of a warp fall on 4 bank
      // pairs (8-way conflict, 45 % of the store wavefronts of the MCB kernel).  Rotating the order in which a lane
      // writes its four outputs (q = i + lane for stage 0, i + lane/4 for stage 1) spreads every store instruction
      // over 16 bank pairs x 2 lanes = the conflict-free two wavefronts.
      const int rot = (s == 0) ? tid : (tid >> 2);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int q = (i + rot) & 3;
        const float2 lo = (q & 1) ? r1 : r0, hi = (q & 1) ? r3 : r2;
        out[o + q * Ns] = (q & 2) ? hi : lo;
      }
    } else {
      out[o] = r0;
      out[o + Ns] = r1;
      out[o + 2 * Ns] = r2;
      out[o + 3 * Ns] = r3;
    }
    __syncthreads();
    float2* t = in;
    in = out;
    out = t;
  }
  return in;
}

}  // namespace avvad
