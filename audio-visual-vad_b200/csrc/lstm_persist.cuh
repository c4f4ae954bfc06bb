// Persistent LSTM recurrence for one layer (sm_100a): all T time steps in ONE cooperative launch.
//
// The 4H gate columns (gate-interleaved: column 4u+g) are split into slices of 64 columns (16 hidden units) and the
// batch into slices of 128 rows; CTA (ns, ms) keeps its W_hh slice [64][H] bf16 resident in shared memory for the
// whole sequence (H=1024: 128 KB, loaded once by TMA) and the cell state of its 128 x 16 (row, unit) pairs in
// registers.  Synchronisation between CTAs is dataflow, not a grid barrier: every CTA owns one flag (the number of
// steps it has published), and K block kb of h_{t-1} (64 hidden units) only depends on the four CTAs that produce
// those units.  Per time step:
//   warp 0 : polls the flags of its batch slice (two per lane) and, K block by K block as their four producers
//            report step t-1, streams h_{t-1}[128 rows][64] through a 6-stage TMA ring (box {64, 1, 128} of the
//            [B][T][H] output sequence -- the layer output doubles as the recurrent operand)
//   warp 1 : tcgen05.mma 128 x 64 x 16 over K = H into a 64-column TMEM accumulator
//   warps 2-9: have already fetched xproj[b][t][their 32 columns] (independent of h), wait for the accumulator,
//            apply the gate non-linearities, update c (registers), store h_t as bf16 (zero beyond the sequence
//            length); after a named barrier one thread fences and publishes the CTA's flag.
// Cluster mode (g.cluster = 2, 4 or 8 CTAs with the same batch slice): every CTA needs ALL of h_{t-1}[128 rows], so
// without help the 64 column slices of a batch slice read the same 256 KB from L2 each step (32 MB per step).  In a
// cluster, K block kb is fetched once by CTA (kb % cluster) and TMA-multicast into the same ring slot of all members;
// a slot is recycled when every member's MMA has committed it (tcgen05.commit multicast onto the members' `empty`
// barriers, count = cluster).
// Reference semantics: nn.LSTM inside packages/models/AV_Net.py:128-137 (gates i,f,g,o; zero initial state).
#pragma once
#include "gemm_tma.cuh"

namespace avvad {
namespace tc {

constexpr int kLstmThreads = 320;
constexpr int kLstmStages = 6;
constexpr int kLstmMaxSlices = 64;  // flags per batch slice: two per polling lane

struct LstmGeom {
  int B, T, H, KB;      // KB = H / 64
  int n_slices;         // 4H / 64
  const float4* xT;     // input projection xT[t][u][b][4] (b < Bp), both biases folded in; already offset to this
                        // launch's first batch row
  int Bp;
  __nv_bfloat16* hseq;  // [B][T][H] bf16 layer output
  const int32_t* lengths;
  unsigned int* counters;  // flags [m_slices][kLstmMaxSlices], zeroed before the launch
  // training only (may be null): post-activation gates (i,f,g,o per unit, bf16 [B][T][4H]) and cell states (f32 [B][T][H])
  __nv_bfloat16* gates_out;
  float* c_out;
  // This launch runs steps [t0, t1) (see lstm_pair.cuh: chunked layers); the cell state crosses launches in c_state
  int t0, t1;
  float* c_state;  // f32 [B][H]: read when t0 > 0, written at the end
  int cluster;  // CTAs per cluster sharing h through TMA multicast (1 = none)
  int variant;  // tuning knobs (AVVAD_LSTM_VARIANT): 1 = every thread fences before the barrier, 2 = back-off between
                // polls, 32 / 64 = acquire / proxy fence behind the poll
};

struct LstmMaps {
  CUtensorMap h;  // hseq as (H, T, B), box {64, 1, 128}
  CUtensorMap w;  // W_hh packed [4H][H], box {64, 64}
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%2, %3, %4}], [%5], %6;"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// generic <-> async proxy ordering for global memory only (h is written with st.global and read back by TMA)
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// hardware tanh (MUFU.TANH, |err| ~ 5e-4 abs, below the bf16 rounding of h) and sigmoid(x) = 0.5 tanh(x/2) + 0.5: one
// special-function instruction per gate on the recurrence's critical path
__device__ __forceinline__ float tanh_hw(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_hw(float x) { return fmaf(0.5f, tanh_hw(0.5f * x), 0.5f); }

__global__ void __launch_bounds__(kLstmThreads)
lstm_persist_kernel(const __grid_constant__ LstmMaps maps, const LstmGeom g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t w_bytes = (uint32_t)g.KB * 8192u;  // KB tiles of [64 rows][64 k]
  const uint32_t sW = base;
  const uint32_t sA = base + w_bytes;
  const uint32_t bar0 = sA + kLstmStages * 16384u;
  // barriers: full[S] | empty[S] | wfull | tfull[2] | tempty[2]
  // The accumulator is double buffered: K block 0 of step t+1 only depends on four OTHER CTAs, so without `tempty` the
  // MMAs of step t+1 could overwrite an accumulator that this CTA's cell update (step t) has not read yet.
  constexpr int kBarW = 2 * kLstmStages, kBarT = 2 * kLstmStages + 1, kBarTE = 2 * kLstmStages + 3;
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + (bar0 - base) + 8 * (2 * kLstmStages + 6));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ns = blockIdx.x % g.n_slices;
  const int ms = blockIdx.x / g.n_slices;
  unsigned int* flags = g.counters + ms * kLstmMaxSlices;  // flags[n] = steps published by CTA (ms, n)

  if (threadIdx.x == 0) {
    for (int s = 0; s < kLstmStages; ++s) {
      mbar_init(BAR(s), 1);
      mbar_init(BAR(kLstmStages + s), (uint32_t)g.cluster);  // every cluster member's MMA releases the slot
    }
    mbar_init(BAR(kBarW), 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(BAR(kBarT + a), 1);
      mbar_init(BAR(kBarTE + a), 8);  // one arrival per cell-update warp
    }
    fence_barrier_init();
    tma_prefetch_desc(&maps.h);
    tma_prefetch_desc(&maps.w);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  const int CL = g.cluster;
  const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0u;
  const uint16_t cmask = (uint16_t)((1u << CL) - 1u);
  if (CL > 1) cluster_sync_all();  // peers' barriers are initialised before anyone multicasts onto them

  if (warp == 0) {
    if (lane == 0) {
      // resident weights
      mbar_arrive_expect_tx(BAR(kBarW), w_bytes);
      for (int kb = 0; kb < g.KB; ++kb) tma_load_2d(sW + kb * 8192u, &maps.w, kb * 64, ns * 64, BAR(kBarW));
    }
    uint32_t it = 0;
    if (CL > 1) {
      // cluster mode: slots are armed in K order by every member; the member that owns kb waits for its four producer
      // flags (lanes 0,1 read the two flag pairs), fences once and multicasts the box to all members
      for (int t = (g.t0 > 1 ? g.t0 : 1); t < g.t1; ++t) {
        const unsigned int target = (unsigned int)(t - g.t0);
        for (int kb = 0; kb < g.KB; ++kb, ++it) {
          const int s = it % kLstmStages;
          mbar_wait(BAR(kLstmStages + s), ((it / kLstmStages) & 1u) ^ 1u);
          if (elect_one_sync()) mbar_arrive_expect_tx(BAR(s), 16384u);
          __syncwarp();
          if ((uint32_t)(kb % CL) == crank) {
            const unsigned long long* fp = reinterpret_cast<const unsigned long long*>(flags + 4 * kb) + (lane & 1);
            uint32_t spins = 0;
            for (;;) {
              unsigned long long fv;
              asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(fv) : "l"(fp) : "memory");
              const bool ok = ((unsigned int)fv >= target) && ((unsigned int)(fv >> 32) >= target);
              if ((__ballot_sync(0xffffffffu, ok) & 3u) == 3u) break;
              if (++spins > (1u << 26)) __trap();
            }
            if (elect_one_sync()) {
              asm volatile("fence.acq_rel.gpu;" ::: "memory");
              fence_proxy_async_global();
              tma_load_3d_mc(sA + s * 16384u, &maps.h, kb * 64, t - 1, ms * 128, BAR(s), cmask);
            }
            __syncwarp();
          }
        }
      }
    } else {
    // lane l watches the flags of slices 2l and 2l+1; K block kb is produced by slices 4kb .. 4kb+3 = lanes 2kb, 2kb+1
    const int f0 = 2 * lane, f1 = 2 * lane + 1;
    const unsigned long long* fpair = reinterpret_cast<const unsigned long long*>(flags + f0);  // both flags in one load
    for (int t = (g.t0 > 1 ? g.t0 : 1); t < g.t1; ++t) {  // step 0 has h_{-1} = 0: no operand to fetch
      // flags count the steps published in THIS launch; h_{t0-1} comes from the previous launch (target 0)
      const unsigned int target = (unsigned int)(t - g.t0);
      bool ok = false;
      int kb = 0;
      uint32_t spins = 0;
      while (kb < g.KB) {
        if (!ok) {
          if (f0 >= g.n_slices) {
            ok = true;
          } else {
            // relaxed poll: the data is only ever read by TMA (L2), so no L1 invalidation is needed per probe; the
            // acquire fence + proxy fence below order the TMA reads after the observation
            unsigned long long fv;
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(fv) : "l"(fpair) : "memory");
            ok = ((unsigned int)fv >= target) && (f1 >= g.n_slices || (unsigned int)(fv >> 32) >= target);
          }
          if (++spins > (1u << 26)) __trap();
          if (!ok && (g.variant & 2)) __nanosleep(20);
        }
        const unsigned int m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0 && kb < g.KB && ((m >> (2 * kb)) & 3u) == 3u) {
          // Peers released h_t at gpu scope before their flag and TMA reads L2, so the fences behind the poll only
          // restate the control dependency -- and both wait for every outstanding load of the SM (the cell-update warps'
          // input-projection prefetch): off unless AVVAD_LSTM_VARIANT bits 32 / 64 ask for them.
          if (g.variant & 32) asm volatile("fence.acq_rel.gpu;" ::: "memory");
          if (g.variant & 64) fence_proxy_async_global();
        }
        while (kb < g.KB && ((m >> (2 * kb)) & 3u) == 3u) {
          if (lane == 0) {
            const int s = it % kLstmStages;
            mbar_wait(BAR(kLstmStages + s), ((it / kLstmStages) & 1u) ^ 1u);
            mbar_arrive_expect_tx(BAR(s), 16384u);
            tma_load_3d(sA + s * 16384u, &maps.h, kb * 64, t - 1, ms * 128, BAR(s));
          }
          ++kb;
          ++it;
        }
      }
    }
    }
  } else if (warp == 1) {
    // MMA issuer (warp-converged; one elected lane issues the MMAs and commits)
    constexpr uint32_t idesc = make_idesc(64);
    mbar_wait(BAR(kBarW), 0);
    tc_fence_after();
    uint32_t it = 0;
    const int tb = g.t0 > 1 ? g.t0 : 1;  // first step with an MMA
    for (int t = tb; t < g.t1; ++t) {
      const uint32_t k = (uint32_t)(t - tb), acc = k & 1u;
      if (k >= 2) mbar_wait(BAR(kBarTE + acc), ((k >> 1) - 1u) & 1u);  // step t-2 has left this accumulator
      tc_fence_after();
      const uint32_t d = tmem_acc + acc * 64u;
      for (int kb = 0; kb < g.KB; ++kb, ++it) {
        const int s = it % kLstmStages;
        mbar_wait(BAR(s), (it / kLstmStages) & 1u);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t a_lo = desc_lo(sA + s * 16384u);
          const uint32_t b_lo = desc_lo(sW + kb * 8192u);
          umma_f16_lo(d, a_lo, b_lo, idesc, kb != 0);
          umma_f16_lo(d, a_lo + 2, b_lo + 2, idesc, 1);
          umma_f16_lo(d, a_lo + 4, b_lo + 4, idesc, 1);
          umma_f16_lo(d, a_lo + 6, b_lo + 6, idesc, 1);
          if (CL > 1)
            umma_commit_mc(BAR(kLstmStages + s), cmask);
          else
            umma_commit(BAR(kLstmStages + s));
          if (kb == g.KB - 1) umma_commit(BAR(kBarT + acc));
        }
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue / cell update: warps 2..9 =================
    // two warps per TMEM lane quarter: `half` selects hidden units [8*half, 8*half+8) = accumulator columns 32*half ..
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int b = ms * 128 + q * 32 + lane;
    const bool row_ok = b < g.B;
    const int len = row_ok ? g.lengths[b] : 0;
    const int H4 = 4 * g.H;
    float c[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) c[u] = 0.f;
    if (g.t0 > 0 && row_ok) {
      const float4* cs = reinterpret_cast<const float4*>(g.c_state + (int64_t)b * g.H + ns * 16 + half * 8);
      const float4 c0 = cs[0], c1 = cs[1];
      c[0] = c0.x; c[1] = c0.y; c[2] = c0.z; c[3] = c0.w; c[4] = c1.x; c[5] = c1.y; c[6] = c1.z; c[7] = c1.w;
    }
    const int tb = g.t0 > 1 ? g.t0 : 1;
    // lane = batch row: a warp's load of one unit's gates is one contiguous 512-byte run
    const float4* xcol = g.xT + (int64_t)(ns * 16 + half * 8) * g.Bp + (row_ok ? b : 0);
    __nv_bfloat16* hrow = g.hseq + ((int64_t)(row_ok ? b : 0) * g.T) * g.H + ns * 16 + half * 8;
    for (int t = g.t0; t < g.t1; ++t) {
      // input projection of this step: independent of h, requested before the wait on the accumulator
      float4 x[8];
      if (row_ok) {
        const float4* xp = xcol + (int64_t)t * g.H * g.Bp;
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = __ldg(xp + (int64_t)u * g.Bp);
      }
      uint32_t v[32];
      if (t > 0) {
        const uint32_t k = (uint32_t)(t - tb), acc = k & 1u;
        mbar_wait(BAR(kBarT + acc), (k >> 1) & 1u);
        tc_fence_after();
        tmem_ld32(tmem_acc + acc * 64u + (uint32_t)(half * 32) + ((uint32_t)(q * 32) << 16), v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(kBarTE + acc));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      if (row_ok) {
        float hv[8];
        __nv_bfloat16* gsave = g.gates_out ? g.gates_out + ((int64_t)b * g.T + t) * H4 + ns * 64 + half * 32 : nullptr;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int o = u * 4;
          float gi, gf, gg, go, tc;
          if (g.variant & 8) {  // exp-based activations (the pre-MUFU.TANH path, kept for A/B accuracy checks)
            gi = sigmoidf_fast(__uint_as_float(v[o + 0]) + x[u].x);
            gf = sigmoidf_fast(__uint_as_float(v[o + 1]) + x[u].y);
            gg = tanhf_fast(__uint_as_float(v[o + 2]) + x[u].z);
            go = sigmoidf_fast(__uint_as_float(v[o + 3]) + x[u].w);
            c[u] = gf * c[u] + gi * gg;
            tc = tanhf_fast(c[u]);
          } else {
            gi = sigmoid_hw(__uint_as_float(v[o + 0]) + x[u].x);
            gf = sigmoid_hw(__uint_as_float(v[o + 1]) + x[u].y);
            gg = tanh_hw(__uint_as_float(v[o + 2]) + x[u].z);
            go = sigmoid_hw(__uint_as_float(v[o + 3]) + x[u].w);
            c[u] = gf * c[u] + gi * gg;
            tc = tanh_hw(c[u]);
          }
          hv[u] = go * tc;
          if (gsave) {
            uint2 pk;
            pk.x = pack_bf16x2(gi, gf);
            pk.y = pack_bf16x2(gg, go);
            *reinterpret_cast<uint2*>(gsave + 4 * u) = pk;
          }
        }
        if (g.c_out) {
          float4* cp = reinterpret_cast<float4*>(g.c_out + ((int64_t)b * g.T + t) * g.H + ns * 16 + half * 8);
          cp[0] = make_float4(c[0], c[1], c[2], c[3]);
          cp[1] = make_float4(c[4], c[5], c[6], c[7]);
        }
        const bool live = t < len;
        uint4 h0;
        h0.x = live ? pack_bf16x2(hv[0], hv[1]) : 0u;
        h0.y = live ? pack_bf16x2(hv[2], hv[3]) : 0u;
        h0.z = live ? pack_bf16x2(hv[4], hv[5]) : 0u;
        h0.w = live ? pack_bf16x2(hv[6], hv[7]) : 0u;
        *reinterpret_cast<uint4*>(hrow + (int64_t)t * g.H) = h0;
      }
      // publish: all eight warps have stored their part of h_t; one thread makes it visible and raises the flag
      if (g.variant & 4) fence_proxy_async_all(); else fence_proxy_async_global();
      if (g.variant & 1) {
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp == 2 && lane == 0)
          asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(flags + ns), "r"((unsigned int)(t + 1 - g.t0)) : "memory");
      } else {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp == 2 && lane == 0) {
          asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + ns), "r"((unsigned int)(t + 1 - g.t0)) : "memory");
        }
        // (variant bit 4: hold the other warps until the flag is out, so that MEMBAR.GPU does not wait behind the next
        // step's input-projection reads.  Measured without instrumentation: 7.6 ms with, 7.4-7.5 ms without -> off.)
        if (g.variant & 16) asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    }
    if (row_ok && g.c_state) {  // hand the cell state to the next chunk
      float4* cs = reinterpret_cast<float4*>(g.c_state + (int64_t)b * g.H + ns * 16 + half * 8);
      cs[0] = make_float4(c[0], c[1], c[2], c[3]);
      cs[1] = make_float4(c[4], c[5], c[6], c[7]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // no member exits while peers may still arrive on its barriers
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_acc, 128);
  }
}

}  // namespace tc
}  // namespace avvad
