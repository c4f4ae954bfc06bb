// Fused ResNet BasicBlock for the 17 x 17 x 64 maps of layer1 (eval mode, BatchNorm folded) on CTA PAIRS (sm_100a):
//     z = relu(conv_b(relu(conv_a(x) + bias_a)) + bias_b + x)                      (packages/models/AV_Net.py:78-94: resnet18.layer1)
// in ONE kernel: the output of conv_a never leaves the SM -- its epilogue writes it, zero border included, straight into
// the shared-memory slab that conv_b's row-shifted UMMA descriptors read.  Layer by layer, a block moves x, y, y, x, z
// through HBM (15 GB per block and step at the benchmark shape); fused it reads x and writes z (6 GB, the residual comes
// back out of L2).
//
// Why pairs: both weight sets (2 x 72 KB) and three slabs do not fit one SM.  A pair (tcgen05 cta_group::2) splits every
// weight tile over the two CTAs (2 x 36 KB each; both tensor cores read both halves) and computes two frames at once as
// M = 256 MMAs: CTA r stages frame 2i + r.  Per CTA: one x slab (46 KB, reloaded while conv_b runs), two y slabs (conv_b
// of frame n-1 reads one while conv_a's epilogue of frame n writes the other), 72 KB of weights.
//
// Tensor-pipe schedule (leader's MMA warp):  A(0) | A(1) B(0) | A(2) B(1) | ...   with A = conv_a, B = conv_b, 108 MMAs each
//   epilogue a(n) (accumulator A -> bias, ReLU, border mask -> y slab) runs under B(n-1),
//   epilogue b(n-1) (accumulator B + bias + residual -> ReLU -> global) runs under A(n+1):
// the same eight warps alternate between the two, and neither accumulator needs a second copy (2 x 192 TMEM columns).
//
//   warp 0  : TMA (resident weight halves once; the x slab of every frame, cp.async.bulk.tensor...cta_group::2)
//   warp 1  : TMEM alloc (cta_group::2); rank 0 issues all MMAs, commits are multicast to both CTAs
//   warps 2-9: epilogues (TMEM lane quarter q = warp & 3, channel half (warp - 2) >> 2)
// Geometry: GEMM rows are the positions p = y * 19 + x of the zero-padded 19 x 19 grid (3 blocks of 128 rows, 289 of 384
// valid); the pixel of output position p sits at slab row p + 20; tap (r, s) is a descriptor start (r * 19 + s) rows down.
#pragma once
#include "conv_slab.cuh"

namespace avvad {
namespace tc {

constexpr int kBlkThreads = 320;
constexpr int kBlkW = 17, kBlkWp = 19;
constexpr int kBlkRows = kBlkWp * kBlkWp;           // 361 slab rows carry data
constexpr uint32_t kBlkSlabBytes = 47104;           // 361 x 128 B rounded up to 1 KB (descriptors overrun into the next
                                                    // region for the discarded GEMM rows: any readable bytes will do)
constexpr uint32_t kBlkWTile = 32 * 128;            // one tap of one conv, this CTA's 32 output channels
constexpr uint32_t kBlkWBytes = 9 * kBlkWTile;      // 36,864 per conv
constexpr uint32_t kBlkOffY = kBlkSlabBytes;
constexpr uint32_t kBlkOffW = 3 * kBlkSlabBytes;
constexpr uint32_t kBlkOffBias = kBlkOffW + 2 * kBlkWBytes;
constexpr uint32_t kBlkOffBar = kBlkOffBias + 512;
constexpr uint32_t kBlkSmem = 1024 + kBlkOffBar + 256;

struct BlockGeom {
  int64_t n_frames;
  int n_pairs;
  const float* bias_a;  // folded BN biases [64] (null = 0)
  const float* bias_b;
  const __nv_bfloat16* x;  // block input = residual, NHWC [n][17][17][64]
  __nv_bfloat16* z;        // block output
};
struct BlockMaps {
  CUtensorMap x;   // (64, 17, 17, n) box {64, 19, 19, 1}
  CUtensorMap wa;  // [64][576] box {64, 32}
  CUtensorMap wb;
};

__global__ void __launch_bounds__(kBlkThreads, 1)
tc_block17_kernel(const __grid_constant__ BlockMaps maps, const BlockGeom g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sX = base, sY = base + kBlkOffY, sW = base + kBlkOffW;
  const uint32_t bar0 = base + kBlkOffBar;
  // barriers: x_full | x_empty | accA_full | accA_empty | accB_full | accB_empty | y_full[2] | y_empty[2] | w_full
  enum { XF = 0, XE, AF, AE, BF, BE, YF0, YF1, YE0, YE1, WF, NBAR };
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kBlkOffBar + 8 * (NBAR + 1));
  float* bias_s = reinterpret_cast<float*>(smem + kBlkOffBias);  // [a 64 | b 64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = pair_rank();
  const int pair = blockIdx.x >> 1;
  // frames of this CTA: f(i) = (i * n_pairs + pair) * 2 + rank, i = 0 .. n_iter-1 (frames past the end: TMA zero fill, no store)
  const int64_t per_round = 2ll * g.n_pairs;
  const int n_iter = (int)((g.n_frames + per_round - 1) / per_round);

  if (threadIdx.x == 0) {
    mbar_init(BAR(XF), 1);
    mbar_init(BAR(XE), 1);
    mbar_init(BAR(AF), 1);
    mbar_init(BAR(AE), 2 * 8);  // every epilogue warp of both CTAs (the leader's copy is used)
    mbar_init(BAR(BF), 1);
    mbar_init(BAR(BE), 2 * 8);
    mbar_init(BAR(YF0), 2 * 8);
    mbar_init(BAR(YF1), 2 * 8);
    mbar_init(BAR(YE0), 1);
    mbar_init(BAR(YE1), 1);
    mbar_init(BAR(WF), 1);
    fence_barrier_init();
    tma_prefetch_desc(&maps.x);
    tma_prefetch_desc(&maps.wa);
    tma_prefetch_desc(&maps.wb);
    // this CTA's half (32 output channels) of both weight sets, resident for the whole launch
    mbar_arrive_expect_tx(BAR(WF), 2 * kBlkWBytes);
    for (int tap = 0; tap < 9; ++tap) {
      tma_load_2d(sW + tap * kBlkWTile, &maps.wa, tap * 64, (int)rank * 32, BAR(WF));
      tma_load_2d(sW + kBlkWBytes + tap * kBlkWTile, &maps.wb, tap * 64, (int)rank * 32, BAR(WF));
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(const_cast<uint32_t*>(tmem_slot))),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 128; i += kBlkThreads)
    bias_s[i] = (i < 64) ? (g.bias_a ? g.bias_a[i] : 0.f) : (g.bias_b ? g.bias_b[i - 64] : 0.f);
  // y slabs: rows 0..19 (top border row and the left border of the first image row) are never written by an epilogue
  for (int i = threadIdx.x; i < 2 * 20 * 8; i += kBlkThreads) {
    const int slab = i / 160, r = i % 160;
    *reinterpret_cast<uint4*>(smem + kBlkOffY + slab * kBlkSlabBytes + r * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  if (warp == 0) mbar_wait(BAR(WF), 0);
  tc_fence_before();
  __syncthreads();
  pair_sync();  // weights of both halves resident, both CTAs' barriers initialised, TMEM allocated, borders zeroed
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  const uint32_t bar0_leader = pair_mapa(bar0, 0);
  auto LBAR = [&](int i) { return bar0_leader + 8u * (uint32_t)i; };

  if (warp == 0) {
    // ================= producer: the x slab of every frame =================
    for (int i = 0; i < n_iter; ++i) {
      mbar_wait(BAR(XE), ((uint32_t)i & 1u) ^ 1u);  // conv_a of frame i-1 has read the slab
      if (elect_one_sync()) {
        const int64_t f = ((int64_t)i * g.n_pairs + pair) * 2 + rank;
        if (rank == 0) mbar_arrive_expect_tx(BAR(XF), 2u * kBlkRows * 128u);
        tma_load_4d_2sm(sX, &maps.x, 0, -1, -1, (int)f, LBAR(XF));
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader) =================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_m256(64);
      // 9 taps x 3 blocks x 4 K16 steps of one convolution
      auto conv = [&](uint32_t slab, uint32_t wbase, uint32_t d0) {
        const uint32_t slab_lo = desc_lo(slab);
        // one tap per loop iteration (12 MMAs with immediate offsets): fully unrolled, the 2 x 108 descriptor pairs of
        // both convolutions were all computed up front and spilled
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t w_lo = desc_lo(wbase) + (uint32_t)tap * (kBlkWTile >> 4);
          const int fr = tap / 3, fs = tap - fr * 3;
          const uint32_t a0 = slab_lo + (uint32_t)(fr * kBlkWp + fs) * 8u;  // 16-byte units: one slab row = 8
#pragma unroll
          for (int m = 0; m < 3; ++m) {
            umma2_f16_lo2(d0 + m * 64, a0 + m * 1024u, w_lo, idesc, tap != 0, kDescHi);
            umma2_f16_lo2(d0 + m * 64, a0 + m * 1024u + 2, w_lo + 2, idesc, 1, kDescHi);
            umma2_f16_lo2(d0 + m * 64, a0 + m * 1024u + 4, w_lo + 4, idesc, 1, kDescHi);
            umma2_f16_lo2(d0 + m * 64, a0 + m * 1024u + 6, w_lo + 6, idesc, 1, kDescHi);
          }
        }
      };
      for (int i = 0; i <= n_iter; ++i) {
        if (i < n_iter) {  // A(i)
          mbar_wait(BAR(XF), (uint32_t)i & 1u);
          mbar_wait(BAR(AE), ((uint32_t)i & 1u) ^ 1u);  // epilogue a(i-1) has read accumulator A
          tc_fence_after();
          if (elect_one_sync()) {
            conv(sX, sW, tmem_acc);
            umma2_commit_mc2(BAR(XE));
            umma2_commit_mc2(BAR(AF));
          }
          __syncwarp();
        }
        if (i >= 1) {  // B(i-1)
          const uint32_t j = (uint32_t)(i - 1), slot = j & 1u;
          mbar_wait(BAR(YF0 + slot), (j >> 1) & 1u);   // epilogue a(i-1) has written the slab in both CTAs
          mbar_wait(BAR(BE), (j & 1u) ^ 1u);           // epilogue b(i-2) has read accumulator B
          tc_fence_after();
          if (elect_one_sync()) {
            conv(sY + slot * kBlkSlabBytes, sW + kBlkWBytes, tmem_acc + 192);
            umma2_commit_mc2(BAR(YE0 + slot));
            umma2_commit_mc2(BAR(BF));
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================= epilogues: warps 2..9 =================
    const int q = warp & 3;
    const int jh = (warp - 2) >> 2;  // channel half: columns 32*jh .. 32*jh+31 of every accumulator block
    // tile-invariant geometry of this thread's three rows p = m*128 + q*32 + lane
    int loc[3];    // element offset of the output pixel inside a frame (y*17 + x)*64 + 32*jh, -1 = border / discarded row
    int yrow[3];   // y-slab byte offset of the row's 64-byte half (pixel row p + 20), -1 = beyond the slab
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      const int p = m * 128 + q * 32 + lane;
      const int y = p / kBlkWp, x = p - y * kBlkWp;
      loc[m] = (y < kBlkW && x < kBlkW) ? (y * kBlkW + x) * 64 + 32 * jh : -1;
      const int r = p + kBlkWp + 1;
      yrow[m] = (r < kBlkRows) ? r : -1;
    }
    const float* ba = bias_s + 32 * jh;
    const float* bb = bias_s + 64 + 32 * jh;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    for (int i = 0; i <= n_iter; ++i) {
      if (i < n_iter) {
        // ---- epilogue a(i): accumulator A -> relu(. + bias_a), zero on the border -> y slab (SWIZZLE_128B K-major rows)
        const uint32_t slot = (uint32_t)i & 1u;
        mbar_wait(BAR(AF), (uint32_t)i & 1u);
        mbar_wait(BAR(YE0 + slot), (((uint32_t)i >> 1) & 1u) ^ 1u);  // conv_b of frame i-2 has read this slab
        tc_fence_after();
        uint8_t* ys = smem + kBlkOffY + slot * kBlkSlabBytes;
#pragma unroll
        for (int m = 0; m < 3; ++m) {
#pragma unroll
          for (int hc = 0; hc < 2; ++hc) {  // 16 accumulator columns at a time keeps the register count down
            uint32_t v[16];
            tmem_ld16(tmem_acc + (uint32_t)(m * 64 + 32 * jh + 16 * hc) + lane_sel, v);
            tmem_ld_wait();
            if (m == 2 && hc == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) pair_arrive(LBAR(AE));
            }
            if (yrow[m] >= 0) {
              const bool valid = loc[m] >= 0;
              const int r = yrow[m];
#pragma unroll
              for (int c = 0; c < 2; ++c) {  // two 16-byte chunks = 16 channels
                uint32_t w[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float f0 = __uint_as_float(v[8 * c + 2 * e]) + ba[16 * hc + 8 * c + 2 * e];
                  const float f1 = __uint_as_float(v[8 * c + 2 * e + 1]) + ba[16 * hc + 8 * c + 2 * e + 1];
                  w[e] = valid ? pack_relu_bf16x2(f0, f1) : 0u;
                }
                const int chunk = (4 * jh + 2 * hc + c) ^ (r & 7);
                *reinterpret_cast<uint4*>(ys + r * 128 + chunk * 16) = make_uint4(w[0], w[1], w[2], w[3]);
              }
            }
          }
        }
        fence_proxy_async();  // generic-proxy slab writes -> visible to the tensor core's (async proxy) operand reads
        __syncwarp();
        if (lane == 0) pair_arrive(LBAR(YF0 + slot));
      }
      if (i >= 1) {
        // ---- epilogue b(i-1): accumulator B + bias_b + x -> relu -> global
        const int j = i - 1;
        const int64_t f = ((int64_t)j * g.n_pairs + pair) * 2 + rank;
        const bool live = f < g.n_frames;
        const int64_t foff = f * (int64_t)(kBlkW * kBlkW * 64);
        u32x8 rb[3][2] = {};
#pragma unroll
        for (int m = 0; m < 3; ++m)
          if (live && loc[m] >= 0) {
            rb[m][0] = ld_global_256(g.x + foff + loc[m]);
            rb[m][1] = ld_global_256(g.x + foff + loc[m] + 16);
          }
        mbar_wait(BAR(BF), (uint32_t)j & 1u);
        tc_fence_after();
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int e = 0; e < 8; ++e) asm volatile("" : "+r"(rb[m][c].v[e]));  // keep the loads ahead of the wait
#pragma unroll
        for (int m = 0; m < 3; ++m) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {  // 16 channels = one 32-byte store
            uint32_t v[16];
            tmem_ld16(tmem_acc + 192u + (uint32_t)(m * 64 + 32 * jh + 16 * c) + lane_sel, v);
            tmem_ld_wait();
            if (m == 2 && c == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) pair_arrive(LBAR(BE));
            }
            if (live && loc[m] >= 0) {
              u32x8 o;
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const uint32_t rw = rb[m][c].v[e];
                const float f0 = __uint_as_float(v[2 * e]) + bb[16 * c + 2 * e] + __uint_as_float(rw << 16);
                const float f1 = __uint_as_float(v[2 * e + 1]) + bb[16 * c + 2 * e + 1] + __uint_as_float(rw & 0xFFFF0000u);
                o.v[e] = pack_relu_bf16x2(f0, f1);
              }
              st_global_256(g.z + foff + loc[m] + 16 * c, o);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  pair_sync();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(512) : "memory");
  }
}

// z = relu(conv_b(relu(conv_a(x) + bias_a)) + bias_b + x) for n frames of 17 x 17 x 64 (weights packed [64][576])
int launch_block17(const __nv_bfloat16* x, const __nv_bfloat16* wa, const float* bias_a, const __nv_bfloat16* wb,
                   const float* bias_b, __nv_bfloat16* z, int64_t n, cudaStream_t st);
bool block17_enabled();

}  // namespace tc
}  // namespace avvad
