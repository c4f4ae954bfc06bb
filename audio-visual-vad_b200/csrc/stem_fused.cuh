// Fused ResNet stem for single-channel 67x67 ROIs: conv 7x7 / stride 2 / pad 3 (3 identical input channels folded
// into one) + folded BN + ReLU + max-pool 3x3 / stride 2 / pad 1, one frame per CTA iteration (persistent CTAs).
// Only the fp32 frame (18 KB) is read from and the pooled 17x17x64 bf16 map (37 KB) written to global memory; the
// im2col rows and the 34x34x64 convolution output never leave the SM.
//
// Per frame:
//   1. frame -> SMEM as bf16 with a 3-pixel zero border
//   2. ten 128-row im2col tiles (K = 49 -> 64) are built in SMEM (SWIZZLE_128B K-major, two buffers) straight from
//      the SMEM frame; after each tile one thread issues 4 x tcgen05.mma (128 x 64 x 16) against the resident
//      weights into one of 8 TMEM accumulator slots; the previous tile's accumulator is drained meanwhile
//      (tcgen05.ld -> +bias -> ReLU -> bf16 -> XOR-swizzled SMEM conv buffer)
//   3. 3x3/2 max-pool over the SMEM conv buffer -> NHWC bf16 global
#pragma once
#include "gemm_tc.cuh"

namespace avvad {
namespace tc {

constexpr int kStemThreads = 512;
constexpr int kStemPrefetch = (67 * 67 + kStemThreads - 1) / kStemThreads;  // pixels per thread
constexpr int kImgPitchB = 80;                 // bf16 elements per padded image row (73 used)
constexpr int kImgRows = 73;
constexpr uint32_t kStemImgBytes = 12288;      // 73*80*2 = 11680, rounded
constexpr uint32_t kStemABytes = 2 * 16384;
constexpr uint32_t kStemWBytes = 8192;
constexpr uint32_t kStemConvBytes = 1156 * 128 + 512;  // + slack for the 4-row tail tile's unused rows
constexpr uint32_t kStemSmem = 1024 + kStemABytes + kStemWBytes + kStemImgBytes + kStemConvBytes + 256 /*bias*/ + 64;

template <int PART>
__device__ __forceinline__ void stem_build_quarter(const __nv_bfloat16* __restrict__ ip, uint32_t dst_row, int rsw) {
  // k = PART*16 + i, tap (fr, fs) = (k / 7, k % 7), value = ip[fr*80 + fs]; k >= 49 -> 0
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k0 = PART * 16 + 2 * i, k1 = k0 + 1;
    uint32_t lo = 0, hi = 0;
    if (k0 < 49) lo = *reinterpret_cast<const unsigned short*>(ip + (k0 / 7) * kImgPitchB + (k0 % 7));
    if (k1 < 49) hi = *reinterpret_cast<const unsigned short*>(ip + (k1 / 7) * kImgPitchB + (k1 % 7));
    w[i] = lo | (hi << 16);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int c = PART * 2 + j;
    const uint32_t dst = dst_row + (uint32_t)((c ^ rsw) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(w[4 * j]), "r"(w[4 * j + 1]),
                 "r"(w[4 * j + 2]), "r"(w[4 * j + 3])
                 : "memory");
  }
}

__global__ void __launch_bounds__(kStemThreads)
stem_fused_kernel(const float* __restrict__ frames, int64_t n_frames, const __nv_bfloat16* __restrict__ w1b,
                  const float* __restrict__ bias, __nv_bfloat16* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  // layout: A[2] (32 KB) | W (8 KB) | img | conv | bias | barriers
  const uint32_t sA = base, sW = base + kStemABytes;
  __nv_bfloat16* img = reinterpret_cast<__nv_bfloat16*>(smem + kStemABytes + kStemWBytes);
  uint8_t* conv = smem + kStemABytes + kStemWBytes + kStemImgBytes;
  float* bias_s = reinterpret_cast<float*>(conv + kStemConvBytes);
  const uint32_t bar0 = base + kStemABytes + kStemWBytes + kStemImgBytes + kStemConvBytes + 256;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + (bar0 - base) + 16);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  // one-off: weights [64 out][64 k] -> SW128 K-major tile, bias, zero the image (borders stay zero), barriers, TMEM
  for (int i = tid; i < 64 * 8; i += kStemThreads) {
    const int r = i >> 3, c = i & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(w1b + r * 64 + c * 8);
    *reinterpret_cast<uint4*>(smem + kStemABytes + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  if (tid < 64) bias_s[tid] = bias[tid];
  for (int i = tid; i < (int)(kStemImgBytes / 4); i += kStemThreads) reinterpret_cast<uint32_t*>(img)[i] = 0u;
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
    tmem_relinquish();
  }
  fence_proxy_async();  // weights were written through the generic proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  constexpr uint32_t idesc = make_idesc(64);
  const uint32_t w_lo = desc_lo(sW);

  uint32_t uses0 = 0, uses1 = 0;  // completed-phase counters of the two MMA barriers
  const int q = warp & 3, cgrp = warp >> 2;  // epilogue: TMEM lane quarter, 16-column group

  // the pixels of a frame are fetched into registers one frame ahead (while the previous frame is being pooled)
  float pf[kStemPrefetch];
  auto prefetch = [&](int64_t f) {
    const float* src = frames + f * (67 * 67);
#pragma unroll
    for (int j = 0; j < kStemPrefetch; ++j) {
      const int i = tid + j * kStemThreads;
      pf[j] = (i < 67 * 67) ? __ldg(src + i) : 0.f;
    }
  };
  if ((int64_t)blockIdx.x < n_frames) prefetch(blockIdx.x);

  for (int64_t f = blockIdx.x; f < n_frames; f += gridDim.x) {
    // ---- 1. prefetched frame -> padded bf16 image (the previous frame's last im2col tile has been built)
#pragma unroll
    for (int j = 0; j < kStemPrefetch; ++j) {
      const int i = tid + j * kStemThreads;
      if (i < 67 * 67) {
        const int r = i / 67, c = i - r * 67;
        img[(r + 3) * kImgPitchB + c + 3] = __float2bfloat16_rn(pf[j]);
      }
    }
    __syncthreads();

    // ---- 2. ten tiles, software pipelined: build A(t) | MMA(t) | drain accumulator(t-1)
#pragma unroll 1
    for (int t = 0; t <= 10; ++t) {
      if (t < 10) {
        const int r = tid >> 2;
        const int m = t * 128 + r;
        const int mm = m < 1156 ? m : 1155;  // tail rows: any valid address, results are ignored
        const int oh = mm / 34, ow = mm - oh * 34;
        const __nv_bfloat16* ip = img + (2 * oh) * kImgPitchB + 2 * ow;
        const uint32_t dst_row = sA + (uint32_t)(t & 1) * 16384u + (uint32_t)(((r >> 3) << 10) + ((r & 7) << 7));
        switch (tid & 3) {
          case 0: stem_build_quarter<0>(ip, dst_row, r & 7); break;
          case 1: stem_build_quarter<1>(ip, dst_row, r & 7); break;
          case 2: stem_build_quarter<2>(ip, dst_row, r & 7); break;
          default: stem_build_quarter<3>(ip, dst_row, r & 7); break;
        }
        fence_proxy_async();
      } else if (f + gridDim.x < n_frames) {
        prefetch(f + gridDim.x);  // t == 10: all tiles of this frame are built; fetch the next frame's pixels now
      }
      __syncthreads();
      if (t < 10 && tid == 0) {
        tc_fence_after();
        const uint32_t a_lo = desc_lo(sA + (uint32_t)(t & 1) * 16384u);
        const uint32_t d = tmem_acc + (uint32_t)(t & 7) * 64u;
        umma_f16_lo(d, a_lo, w_lo, idesc, 0);
        umma_f16_lo(d, a_lo + 2, w_lo + 2, idesc, 1);
        umma_f16_lo(d, a_lo + 4, w_lo + 4, idesc, 1);
        umma_f16_lo(d, a_lo + 6, w_lo + 6, idesc, 1);
        umma_commit(bar0 + 8 * (t & 1));
      }
      if (t >= 1) {
        const int tp = t - 1;
        // wait for MMA(tp): barrier (tp & 1), parity = number of earlier completions on it
        if (tp & 1) {
          mbar_wait(bar0 + 8, uses1 & 1u);
          ++uses1;
        } else {
          mbar_wait(bar0, uses0 & 1u);
          ++uses0;
        }
        tc_fence_after();
        uint32_t v[16];
        tmem_ld16(tmem_acc + (uint32_t)(tp & 7) * 64u + (uint32_t)cgrp * 16u + ((uint32_t)(q * 32) << 16), v);
        tmem_ld_wait();
        const int m = tp * 128 + q * 32 + lane;
        if (m < 1156) {
          uint8_t* crow = conv + m * 128;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              x[e] = fmaxf(__uint_as_float(v[8 * c + e]) + bias_s[cgrp * 16 + 8 * c + e], 0.f);
            uint4 o;
            o.x = pack_bf16x2(x[0], x[1]);
            o.y = pack_bf16x2(x[2], x[3]);
            o.z = pack_bf16x2(x[4], x[5]);
            o.w = pack_bf16x2(x[6], x[7]);
            const int chunk = cgrp * 2 + c;
            *reinterpret_cast<uint4*>(crow + ((chunk ^ (m & 7)) << 4)) = o;
          }
        }
        tc_fence_before();
      }
    }
    __syncthreads();

    // ---- 3. 3x3 / stride 2 / pad 1 max-pool (post-ReLU values: window clipping == -inf padding)
    __nv_bfloat16* o = out + f * (289 * 64);
    for (int item = tid; item < 289 * 8; item += kStemThreads) {
      const int pp = item >> 3, ch = item & 7;
      const int ph = pp / 17, pw = pp - ph * 17;
      float mx[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) mx[e] = 0.f;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const int y = 2 * ph + dy;
        if (y < 0 || y >= 34) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int x = 2 * pw + dx;
          if (x < 0 || x >= 34) continue;
          const int m = y * 34 + x;
          const uint4 vv = *reinterpret_cast<const uint4*>(conv + m * 128 + ((ch ^ (m & 7)) << 4));
          const float2 a = unpack_bf16x2(vv.x), b = unpack_bf16x2(vv.y), c2 = unpack_bf16x2(vv.z), d2 = unpack_bf16x2(vv.w);
          mx[0] = fmaxf(mx[0], a.x); mx[1] = fmaxf(mx[1], a.y); mx[2] = fmaxf(mx[2], b.x); mx[3] = fmaxf(mx[3], b.y);
          mx[4] = fmaxf(mx[4], c2.x); mx[5] = fmaxf(mx[5], c2.y); mx[6] = fmaxf(mx[6], d2.x); mx[7] = fmaxf(mx[7], d2.y);
        }
      }
      uint4 r;
      r.x = pack_bf16x2(mx[0], mx[1]);
      r.y = pack_bf16x2(mx[2], mx[3]);
      r.z = pack_bf16x2(mx[4], mx[5]);
      r.w = pack_bf16x2(mx[6], mx[7]);
      *reinterpret_cast<uint4*>(o + pp * 64 + ch * 8) = r;
    }
    __syncthreads();  // conv buffer and image are rewritten by the next frame
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem_acc, 512);
  }
}

}  // namespace tc
}  // namespace avvad
