// Register-resident 1024-point complex FFT for a GROUP of 64 threads (two warps), 16 values per thread.
//
// fft.cuh's radix-4 Stockham form makes five round trips through shared memory per transform (load 4, store 4 per
// thread and stage, plus three twiddle loads): ncu showed the MCB row kernel at 92 % L1/shared throughput, i.e. bound
// by shared-memory wavefronts at 4-5 % of the HBM bound SURVEY 8(d) assigns to the stage.  Here the same Stockham
// autosort runs with radices 16 x 16 x 4 and the radix-16 / radix-4 butterflies stay in registers, so a transform is
// two stores and three loads of the 1024 values instead of five of each, and 27 twiddle loads per thread instead of 60:
//
//   thread t holds x[t + 64 m] in v[m], m = 0..15  (input AND output layout: transforms chain without a reshuffle)
//   stage A (R = 16, Ns = 1)  : DFT16 in registers, store y[r] at logical index 16 t + r         (8 x STS.128)
//   stage B (R = 16, Ns = 16) : load index t + 64 r, twiddle exp(-2 pi i r k / 256) with k = t & 15, DFT16,
//                               store at 256 (t >> 4) + k + 16 r
//   stage C (R = 4, Ns = 256) : four butterflies j = t + 64 q: load index j + 256 r, twiddle exp(-2 pi i r j / 1024)
//                               (fft.cuh's stage-4 table), DFT4, result X[j + 256 r] = v[q + 4 r]
//
// The exchange buffer is ONE 1024-value array per group with a pitch of 18 entries per 16 (logical index i lives at
// (i >> 4) * 18 + (i & 15)): a thread's 16 consecutive stage-A outputs are 128 contiguous bytes and the 144-byte row
// pitch spreads the quarter-warps of every 128-bit store over all 32 banks; every other access pattern above is 16
// consecutive entries per half-warp.  Reads and in-place writes of a stage are separated by a group barrier
// (`bar.sync id, 64`), so groups of one CTA never wait for each other.
#pragma once
#include "fft.cuh"

namespace avvad {

constexpr int kRfThreads = 64;                    // threads per transform
constexpr int kRfPitch = 18;                      // entries per row of 16
constexpr int kRfBufEntries = 64 * kRfPitch;      // 1152 float2 = 9,216 bytes per group
constexpr int kFftTwC = 768;                      // stage-C table = fft.cuh's stage-4 table (w, w^2, w^3 for k < 256);
                                                  // stage-B table: kFftTwB entries behind the per-stage tables (fft.cuh)

__device__ __forceinline__ void rf_group_bar(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// forward 4-point DFT in place: (a0, a1, a2, a3) -> (X0, X1, X2, X3)
__device__ __forceinline__ void rf_dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 b0 = make_float2(a0.x + a2.x, a0.y + a2.y);
  const float2 b1 = make_float2(a0.x - a2.x, a0.y - a2.y);
  const float2 b2 = make_float2(a1.x + a3.x, a1.y + a3.y);
  const float2 b3 = make_float2(a1.y - a3.y, a3.x - a1.x);  // (a1 - a3) * (-i)
  a0 = make_float2(b0.x + b2.x, b0.y + b2.y);
  a1 = make_float2(b1.x + b3.x, b1.y + b3.y);
  a2 = make_float2(b0.x - b2.x, b0.y - b2.y);
  a3 = make_float2(b1.x - b3.x, b1.y - b3.y);
}

// position of output r of rf_dft16 in the register array
__host__ __device__ __forceinline__ constexpr int rf_perm(int r) { return 4 * (r & 3) + (r >> 2); }

// forward 16-point DFT in place, 4 x 4 Cooley-Tukey: input v[n], output y[r] at v[rf_perm(r)]
__device__ __forceinline__ void rf_dft16(float2 (&v)[16]) {
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) rf_dft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);  // v[4 r1 + n2] = A[n2][r1]
  // twiddles W16^(n2 r1), W16 = exp(-2 pi i / 16)
  constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
  v[5] = cmul(v[5], make_float2(c1, -s1));    // W^1
  v[6] = cmul(v[6], make_float2(h, -h));      // W^2
  v[7] = cmul(v[7], make_float2(s1, -c1));    // W^3
  v[9] = cmul(v[9], make_float2(h, -h));      // W^2
  v[10] = make_float2(v[10].y, -v[10].x);     // W^4 = -i
  v[11] = cmul(v[11], make_float2(-h, -h));   // W^6
  v[13] = cmul(v[13], make_float2(s1, -c1));  // W^3
  v[14] = cmul(v[14], make_float2(-h, -h));   // W^6
  v[15] = cmul(v[15], make_float2(-c1, s1));  // W^9 = -W^1
#pragma unroll
  for (int r1 = 0; r1 < 4; ++r1) rf_dft4(v[4 * r1], v[4 * r1 + 1], v[4 * r1 + 2], v[4 * r1 + 3]);
}

// Forward transform of the group's 1024 values (thread t of the group: v[m] = x[t + 64 m] in, X[t + 64 m] out).
// `buf`: the group's kRfBufEntries exchange buffer, free on entry (every thread of the group is past a barrier behind
// its last read) and free again on return.  twB / twC: shared-memory copies of the two tables above.
__device__ __forceinline__ void fft1024_reg(float2 (&v)[16], float2* __restrict__ buf, const float2* __restrict__ twB,
                                            const float2* __restrict__ twC, int t, int bar_id) {
  const int k = t & 15, a = t >> 4;
  // ---- stage A ----
  rf_dft16(v);
  {
    float4* o = reinterpret_cast<float4*>(buf + t * kRfPitch);
#pragma unroll
    for (int r = 0; r < 16; r += 2)
      o[r >> 1] = make_float4(v[rf_perm(r)].x, v[rf_perm(r)].y, v[rf_perm(r + 1)].x, v[rf_perm(r + 1)].y);
  }
  rf_group_bar(bar_id);
  // ---- stage B ----
#pragma unroll
  for (int r = 0; r < 16; ++r) v[r] = buf[(a + 4 * r) * kRfPitch + k];
  rf_group_bar(bar_id);  // all reads of the group done before the in-place writes below
#pragma unroll
  for (int r = 1; r < 16; ++r) v[r] = cmul(v[r], twB[r * 16 + k]);
  rf_dft16(v);
#pragma unroll
  for (int r = 0; r < 16; ++r) buf[(16 * a + r) * kRfPitch + k] = v[rf_perm(r)];
  rf_group_bar(bar_id);
  // ---- stage C ----
  float2 u[16];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int r = 0; r < 4; ++r) u[4 * q + r] = buf[(a + 4 * q + 16 * r) * kRfPitch + k];
  rf_group_bar(bar_id);  // the buffer is free again
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int j = t + 64 * q;
    u[4 * q + 1] = cmul(u[4 * q + 1], twC[j]);
    u[4 * q + 2] = cmul(u[4 * q + 2], twC[256 + j]);
    u[4 * q + 3] = cmul(u[4 * q + 3], twC[512 + j]);
    rf_dft4(u[4 * q], u[4 * q + 1], u[4 * q + 2], u[4 * q + 3]);
#pragma unroll
    for (int r = 0; r < 4; ++r) v[q + 4 * r] = u[4 * q + r];
  }
}

}  // namespace avvad
