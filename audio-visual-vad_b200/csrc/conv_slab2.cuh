// Packed two-frame slab convolution for the 17x17x64 -> 17x17x64 layers (layer1), sm_100a.
//
// conv_slab.cuh spends 384 accumulator rows on the 289 outputs of a frame (75 %): every frame carries its own halo
// rows and columns and three 128-row blocks per frame leave the last block half empty.  Here a slab holds TWO frames
// in a tighter grid:
//   * row pitch Wp = OW + 1: the single zero column in front of a row (TMA out-of-bounds fill at x = -1) is also the
//     right halo of the row above;
//   * frame pitch OH + 1 rows: the zero row in front of a frame (y = -1) is also the bottom halo of the frame above;
//     the rows behind the second frame are zeroed once per CTA and never written again.
// Output (f, y, x) is GEMM row p = f*(OH+1)*Wp + y*Wp + x and tap (r, s) reads slab row p + r*Wp + s, as before.
// Two frames need rows 0..628 = five 128-row blocks for 578 outputs (90 %): one sixth fewer MMAs than 2 x 3 blocks.
// The five blocks are issued as two groups (3 + 2 blocks) so that the accumulators stay double-buffered in TMEM
// (2 x 192 columns): the epilogue of one group overlaps the MMAs of the next.  Two 85 KB slabs leave no room for
// resident weights, so the weights stream through a two-stage ring of filter rows (3 taps = 24 KB per stage; once per group; L2 hits).
//
//   warp 0   : TMA producer (one slab box per tile; tap tiles)
//   warp 1   : TMEM alloc + MMA issue
//   warps 2-9: epilogue (bias / residual / ReLU -> NHWC bf16)
#pragma once
#include "conv_slab.cuh"

namespace avvad {
namespace tc {

constexpr int kSlab2Frames = 2;
constexpr int kSlab2Blocks = 5;
constexpr int kSlab2GroupA = 3;              // blocks 0..2, then blocks 3..4
constexpr int kSlab2WStages = 2;              // ring of filter ROWS (three taps each)
constexpr uint32_t kSlab2WTap = 64 * 128;    // one tap: [Cout = 64][Cin = 64] bf16
constexpr uint32_t kSlab2WTile = 3 * kSlab2WTap;  // one ring stage: the three taps of a filter row
constexpr uint32_t kSlab2AccCols = kSlab2GroupA * 64;

struct Slab2Geom {
  int64_t n_frames, total_tiles;
  int OH, OW, Wp, per_frame;  // per_frame = (OH + 1) * Wp positions
  uint32_t box_bytes;         // bytes one slab TMA box delivers (2 frames x (OH+1) x Wp x 128)
  uint32_t slab_bytes;        // allocation per slab buffer (1 KB multiple, >= (5*128 + 2*Wp + 2) * 128)
};

// dynamic smem: [slab 0][slab 1][weight ring][barriers | tmem slot | bias]
__global__ void __launch_bounds__(kSlabThreads)
tc_slab2_kernel(const __grid_constant__ SlabMaps maps, const Slab2Geom g, const EpiParams ep) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t w_base = base + 2 * g.slab_bytes;
  const uint32_t bar0 = w_base + kSlab2WStages * kSlab2WTile;
  // barriers: slab_full[2] | slab_empty[2] | tfull[2] | tempty[2] | b_full[8] | b_empty[8]
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  constexpr int kBFull = 8, kBEmpty = 16;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + (bar0 - base) + 8 * 24);
  float* bias_s = reinterpret_cast<float*>(smem + (bar0 - base) + 8 * 24 + 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(BAR(0 + i), 1);              // slab full (expect_tx)
      mbar_init(BAR(2 + i), 1);              // slab empty (tcgen05.commit)
      mbar_init(BAR(4 + i), 1);              // accumulators full
      mbar_init(BAR(6 + i), kSlabEpiWarps);  // accumulators drained
    }
    for (int i = 0; i < kSlab2WStages; ++i) {
      mbar_init(BAR(kBFull + i), 1);
      mbar_init(BAR(kBEmpty + i), 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&maps.a);
    tma_prefetch_desc(&maps.b);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 64; i += kSlabThreads) bias_s[i] = ep.bias ? ep.bias[i] : 0.f;
  // rows behind the TMA box: the bottom halo of the second frame (and the tail only padding rows read)
  for (int sb = 0; sb < 2; ++sb) {
    uint4* z = reinterpret_cast<uint4*>(smem + sb * g.slab_bytes + g.box_bytes);
    const int n16 = (int)((g.slab_bytes - g.box_bytes) >> 4);
    for (int i = threadIdx.x; i < n16; i += kSlabThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    // The slab of the NEXT tile is requested while the current tile is still being multiplied: once the third filter
    // row of tile t could be placed in the ring, row 0 of tile t has retired, hence tile t-1 is complete and its slab
    // buffer is free.  Requesting it only at the top of the next tile left ~85 KB of TMA latency exposed per tile.
    uint32_t si = 0, bi = 0;
    auto load_slab = [&](int64_t tile, uint32_t s) {
      const int sb = s & 1;
      mbar_wait(BAR(2 + sb), ((s >> 1) & 1u) ^ 1u);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(BAR(0 + sb), g.box_bytes);
        tma_load_4d(base + sb * g.slab_bytes, &maps.a, 0, -1, -1, (int)(tile * kSlab2Frames), BAR(0 + sb));
      }
      __syncwarp();
    };
    if ((int64_t)blockIdx.x < g.total_tiles) load_slab(blockIdx.x, 0);
    for (int64_t tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++si) {
      for (int it = 0; it < 6; ++it, ++bi) {  // three filter rows for each of the two block groups
        const int row = it < 3 ? it : it - 3;
        const int bs = bi % kSlab2WStages;
        mbar_wait(BAR(kBEmpty + bs), ((bi / kSlab2WStages) & 1u) ^ 1u);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(BAR(kBFull + bs), kSlab2WTile);
#pragma unroll
          for (int k = 0; k < 3; ++k)
            tma_load_2d(w_base + bs * kSlab2WTile + k * kSlab2WTap, &maps.b, (row * 3 + k) * BK, 0, BAR(kBFull + bs));
        }
        __syncwarp();
        if (it == kSlab2WStages && tile + gridDim.x < g.total_tiles) load_slab(tile + gridDim.x, si + 1);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The 6 (group, filter row) steps of a tile are straight-line code: ring slot, accumulator stage and every descriptor
    // offset are compile-time constants (6 % kSlab2WStages == 0, so a tile always starts at ring slot 0; group A owns
    // accumulator stage 0, group B stage 1).  With a rolled tap loop the ~170 instructions of per-tap bookkeeping in
    // this single thread took longer than the 8-12 MMAs of a tap (ncu: tensor pipe 40 % active, issue thread never
    // waiting on a barrier).
    static_assert(6 % kSlab2WStages == 0, "a tile must start at ring slot 0");
    constexpr uint32_t idesc = make_idesc(64);
    uint32_t tcnt = 0;  // tiles done by this CTA
    for (int64_t tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++tcnt) {
      const int sb = tcnt & 1;
      const uint32_t slab_lo = desc_lo(base + sb * g.slab_bytes);
      const uint32_t w_lo0 = desc_lo(w_base);
      const uint32_t row8 = (uint32_t)g.Wp * 8u;  // one image row in descriptor units
#pragma unroll
      for (int it = 0; it < 6; ++it) {
        const int grp = it / 3, fr = it - grp * 3;
        const int bs = it % kSlab2WStages;
        if (fr == 0) {
          mbar_wait(BAR(6 + grp), (tcnt & 1u) ^ 1u);
          if (grp == 0) mbar_wait(BAR(0 + sb), (tcnt >> 1) & 1u);
        }
        mbar_wait(BAR(kBFull + bs), (tcnt * (6 / kSlab2WStages) + it / kSlab2WStages) & 1u);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t d0 = tmem_acc + (uint32_t)grp * kSlab2AccCols;
          constexpr int kB = kSlab2Blocks - kSlab2GroupA;
#pragma unroll
          for (int fs = 0; fs < 3; ++fs) {
            const uint32_t w_lo = w_lo0 + (uint32_t)(bs * 3 + fs) * (kSlab2WTap >> 4);
            // descriptor low words count 16-byte units: one slab row = 8, one 128-row block = 1024
            const uint32_t a0 = slab_lo + (uint32_t)fr * row8 + (uint32_t)(fs * 8 + (grp == 0 ? 0 : kSlab2GroupA) * 1024);
#pragma unroll
            for (int m = 0; m < kSlab2GroupA; ++m) {
              if (grp == 1 && m >= kB) break;
              umma_f16_lo(d0 + m * 64, a0 + m * 1024u, w_lo, idesc, (fr | fs) != 0);
              umma_f16_lo(d0 + m * 64, a0 + m * 1024u + 2, w_lo + 2, idesc, 1);
              umma_f16_lo(d0 + m * 64, a0 + m * 1024u + 4, w_lo + 4, idesc, 1);
              umma_f16_lo(d0 + m * 64, a0 + m * 1024u + 6, w_lo + 6, idesc, 1);
            }
          }
          umma_commit(BAR(kBEmpty + bs));
          if (fr == 2) {
            umma_commit(BAR(4 + grp));
            if (grp == 1) umma_commit(BAR(2 + sb));
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue: warps 2..9 =================
    // A group has nblk * 2 units of 128 rows x 32 columns; unit u = (m, j) belongs to the four warps of epilogue group
    // (u & 1), one TMEM lane quarter each.  The residual of a thread's units is requested before the wait on the
    // accumulators and pinned behind it.
    const int q = warp & 3;
    const int eg = (warp - 2) >> 2;
    constexpr int kMaxUnits = kSlab2GroupA;  // (3 blocks x 2 column halves) / 2 epilogue groups
    const __nv_bfloat16* resp = ep.residual;
    __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(ep.C);
    const bool relu = ep.relu != 0;
    uint32_t tcnt = 0;
    for (int64_t tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++tcnt) {
      const int64_t n0 = tile * kSlab2Frames;
#pragma unroll
      for (int grp = 0; grp < 2; ++grp) {
        const int nblk = grp == 0 ? kSlab2GroupA : kSlab2Blocks - kSlab2GroupA;
        const int blk0 = grp == 0 ? 0 : kSlab2GroupA;
        const uint32_t acc = (uint32_t)grp, aph = tcnt & 1u;  // group A owns accumulator stage 0, group B stage 1
        int64_t off[kMaxUnits];
        bool ok[kMaxUnits];
        u32x8 rb[kMaxUnits][2] = {};
#pragma unroll
        for (int i = 0; i < kMaxUnits; ++i) {
          const int u = eg + 2 * i;
          ok[i] = false;
          off[i] = 0;
          if (u < nblk * 2) {
            const int m = u >> 1, j = u & 1;
            const int p = (blk0 + m) * 128 + q * 32 + lane;
            const int f = p / g.per_frame;
            const int rem = p - f * g.per_frame;
            const int y = rem / g.Wp;
            const int x = rem - y * g.Wp;
            ok[i] = f < kSlab2Frames && y < g.OH && x < g.OW && n0 + f < g.n_frames;
            off[i] = (((n0 + f) * g.OH + y) * g.OW + x) * ep.ldc + j * 32;
            if (ok[i] && resp) {
              rb[i][0] = ld_global_256(resp + off[i]);
              rb[i][1] = ld_global_256(resp + off[i] + 16);
            }
          }
        }
        mbar_wait(BAR(4 + acc), aph);
        tc_fence_after();
#pragma unroll
        for (int i = 0; i < kMaxUnits; ++i)
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int e = 0; e < 8; ++e) asm volatile("" : "+r"(rb[i][c].v[e]));
#pragma unroll
        for (int i = 0; i < kMaxUnits; ++i) {
          const int u = eg + 2 * i;
          if (u < nblk * 2) {  // warp-uniform
            const int m = u >> 1, j = u & 1;
            uint32_t v[32];
            tmem_ld32(tmem_acc + acc * kSlab2AccCols + (uint32_t)(m * 64 + j * 32) + ((uint32_t)(q * 32) << 16), v);
            tmem_ld_wait();
            if (ok[i]) {
              float f32[32];
              const float4* bp = reinterpret_cast<const float4*>(bias_s + j * 32);
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float4 b4 = bp[c];
                f32[4 * c + 0] = __uint_as_float(v[4 * c + 0]) + b4.x;
                f32[4 * c + 1] = __uint_as_float(v[4 * c + 1]) + b4.y;
                f32[4 * c + 2] = __uint_as_float(v[4 * c + 2]) + b4.z;
                f32[4 * c + 3] = __uint_as_float(v[4 * c + 3]) + b4.w;
              }
              if (resp) {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  const uint32_t w = rb[i][e >> 3].v[e & 7];
                  f32[2 * e + 0] += __uint_as_float(w << 16);
                  f32[2 * e + 1] += __uint_as_float(w & 0xFFFF0000u);
                }
              }
              __nv_bfloat16* cp = outp + off[i];
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                u32x8 o;
#pragma unroll
                for (int e = 0; e < 8; ++e)
                  o.v[e] = relu ? pack_relu_bf16x2(f32[16 * c + 2 * e], f32[16 * c + 2 * e + 1])
                                : pack_bf16x2(f32[16 * c + 2 * e], f32[16 * c + 2 * e + 1]);
                st_global_256(cp + 16 * c, o);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(6 + acc));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_acc, 512);
  }
}

}  // namespace tc
}  // namespace avvad
