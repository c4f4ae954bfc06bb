// Persistent LSTM recurrence on CTA PAIRS (sm_100a, tcgen05.mma.cta_group::2): all T time steps of one layer in ONE
// cooperative cluster launch, for batches of up to 256 rows per launch (with <= 128 rows the second CTA of every pair works
// on zero-filled rows: still faster than lstm_persist.cuh, 8.1-8.7 vs 9.4 us per step at B = 32..64, 8.3 vs 12.2 at B = 128).
//
// lstm_persist.cuh gives every (64 gate columns, 128 batch rows) block its own CTA: 128 CTAs, each of which pulls the
// whole h_{t-1} of its batch slice (256 KB) out of L2 every step (32 MB per step) and runs N = 64 UMMAs (48 cycles for
// 32 cycles of math: shared-memory-read bound).  Here a PAIR of CTAs (one TPC) owns 128 gate columns (32 hidden units)
// for ALL 256 batch rows as one M = 256, N = 128 tcgen05.mma.cta_group::2:
//   * CTA r of the pair stages the h rows of batch slice r (its 128 rows of A) and keeps ITS HALF of the pair's W_hh
//     slice (64 gate columns x H, 128 KB) resident; the tensor cores of both SMs read both halves (B is shared through
//     the pair), so a K = 16 step is 64 cycles of math on 48 cycles of operand reads: the MMA runs at its full rate;
//   * 64 CTAs instead of 128: the per-step all-gather of h through L2 halves (16 MB), and 84 SMs stay free;
//   * accumulators [128 rows][128 columns] per CTA, double buffered in TMEM (the MMAs of step t+1 never wait for the
//     cell update of step t-1 to leave TMEM, and cannot overtake it: `tempty`);
//   * 16 cell-update warps per CTA (one 32-row x 32-column block each), xproj of the step prefetched before the
//     accumulator wait, hardware tanh.
// Synchronisation between pairs is the same dataflow as in lstm_persist.cuh: every CTA publishes "steps done" in its own
// flag; K block kb of h_{t-1} (64 hidden units) of batch slice r depends on CTA r of pairs 2kb and 2kb+1 only.
//   warp 0 : polls the 32 flags of its batch slice (one per lane), streams ready K blocks through a 6-stage TMA ring
//            (cp.async.bulk.tensor...cta_group::2: the transaction bytes of BOTH CTAs complete on the leader's barrier)
//   warp 1 : (leader CTA only) tcgen05.mma.cta_group::2 128x... M = 256, N = 128, K = 16; commits are multicast to the
//            ring / accumulator barriers of both CTAs
//   warps 2-17 : cell update, h_t -> global (bf16), flag.
// Reference semantics: nn.LSTM inside packages/models/AV_Net.py:128-137 (gates i,f,g,o; zero initial state).
#pragma once
#include "lstm_persist.cuh"

namespace avvad {
namespace tc {

// cell-update warps per CTA: 16 (32 rows x 32 columns each, 16-byte h stores) or 8 (32 rows x 64 columns: every thread
// owns 16 hidden units = one full 32-byte sector of h_t)
__host__ __device__ constexpr int kPairThreadsOf(int epi_warps) { return 64 + 32 * epi_warps; }
// operand ring: whatever the resident weight slice leaves of the 227 KB (NP = 128: 6 x 16 KB, NP = 64: 9 x 16 KB)
__host__ __device__ constexpr int kPairStagesOf(int np) { return np == 128 ? 6 : 9; }
constexpr int kPairTraceSlots = 12;

struct PairGeom {
  int B, T, H, KB;      // KB = H / 64
  int n_pairs;          // 4H / NP (<= 64)
  const float4* xT;     // input projection xT[t][u][b][4] (b < Bp), both biases folded in; already offset to this
                        // launch's first batch row
  int Bp;
  __nv_bfloat16* hseq;  // [B][T][H] bf16 layer output
  const int32_t* lengths;
  unsigned int* counters;  // flags [2][kLstmMaxSlices], zeroed before the launch
  __nv_bfloat16* gates_out;  // training tape (may be null)
  float* c_out;
  // This launch runs steps [t0, t1) of the sequence (the driver splits a layer into chunks so that layer l+1 can follow
  // layer l one chunk behind on the idle SMs): h_{t0-1} is already in hseq, the cell state crosses launches in c_state
  int t0, t1;
  float* c_state;  // f32 [B][H]: read when t0 > 0, written at the end
  unsigned long long* trace;  // kTrace only: [CTA][T][kPairTraceSlots] globaltimer stamps
  // diagnostics (AVVAD_LSTM_VARIANT): 2 = acquire fence behind the poll, 4 = generic->async proxy fence behind the poll
  // (neither is needed: the producers release h_t at gpu scope before their flag and TMA reads L2; measured +0.6 us
  // per step because both wait for the SM's outstanding input-projection loads), 256 = no input-projection loads
  int variant;
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Both CTAs of the pair execute this; `bar_leader` is the cluster address of the LEADER's barrier
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                                uint32_t bar_leader) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar_leader)
      : "memory");
}
__device__ __forceinline__ void umma2_f16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                             uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accum), "r"(kDescHi)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// kind::f16 instruction descriptor for the pair: M = 256 (128 rows per CTA)
__host__ __device__ constexpr uint32_t make_idesc_pair(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// NP = gate columns per pair (128: 32 pairs = 64 CTAs; 64: 64 pairs = 128 CTAs, half the cell-update work per SM and a
// deeper operand ring, twice the h traffic through L2); kEW = cell-update warps per CTA.
template <bool kTrace, int kEW, int NP>
__global__ void __launch_bounds__(kPairThreadsOf(kEW), 1)
lstm_pair_kernel(const __grid_constant__ LstmMaps maps, const PairGeom g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  constexpr int S = kPairStagesOf(NP);
  constexpr uint32_t kWTile = (NP / 2) * 128u;  // one K block of this CTA's weight half: [NP/2 gate columns][64 k]
  constexpr int kUnitsPair = NP / 4;            // hidden units per pair
  constexpr int kPerKb = 64 / kUnitsPair;       // pairs that produce one K block of h (2 or 4)
  const uint32_t w_bytes = (uint32_t)g.KB * kWTile;
  const uint32_t sW = base;
  const uint32_t sA = base + w_bytes;
  const uint32_t bar0 = sA + S * 16384u;
  // barriers: full[S] | empty[S] | wfull | tfull[2] | tempty[2]
  constexpr int kBarW = 2 * S, kBarTF = 2 * S + 1, kBarTE = 2 * S + 3, kNumBars = 2 * S + 5;
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + (bar0 - base) + 8 * (kNumBars + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // = batch slice of this CTA; rank 0 issues the MMAs
  const int pair = blockIdx.x >> 1;
  unsigned int* flags = g.counters + rank * kLstmMaxSlices;  // flags[p] = steps published by CTA `rank` of pair p
  unsigned long long* trace = nullptr;
  if (kTrace) trace = g.trace + (size_t)blockIdx.x * g.T * kPairTraceSlots;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(BAR(s), 1);      // the leader's arrive.expect_tx (+ the transaction bytes of both CTAs)
      mbar_init(BAR(S + s), 1);  // tcgen05.commit (multicast to both CTAs)
    }
    mbar_init(BAR(kBarW), 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(BAR(kBarTF + a), 1);
      mbar_init(BAR(kBarTE + a), 2 * kEW);  // every cell-update warp of both CTAs (leader's copy is used)
    }
    fence_barrier_init();
    tma_prefetch_desc(&maps.h);
    tma_prefetch_desc(&maps.w);
    // this CTA's half of the pair's weight slice, resident for the whole sequence
    mbar_arrive_expect_tx(BAR(kBarW), w_bytes);
    for (int kb = 0; kb < g.KB; ++kb)
      tma_load_2d(sW + kb * kWTile, &maps.w, kb * 64, pair * NP + (int)rank * (NP / 2), BAR(kBarW));
  }
  if (warp == 1) {
    tmem_alloc2(smem_u32(const_cast<uint32_t*>(tmem_slot)), 2 * NP);
    tmem_relinquish2();
  }
  if (warp == 0) mbar_wait(BAR(kBarW), 0);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both halves of W resident, both CTAs' barriers initialised, TMEM allocated in both
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  const uint32_t bar0_leader = mapa_u32(bar0, 0);
  auto LBAR = [&](int i) { return bar0_leader + 8u * (uint32_t)i; };

  if (warp == 0) {
    // ================= producer: flags -> TMA ring =================
    // lane l watches kPerKb / 2 ... flags such that K block kb is covered by ballot bits 2kb, 2kb+1:
    //   NP = 128: one flag per lane (pairs 2kb, 2kb+1);  NP = 64: two flags per lane in one 8-byte load (pairs 4kb..4kb+3)
    uint32_t it = 0;
    for (int t = (g.t0 > 1 ? g.t0 : 1); t < g.t1; ++t) {  // step 0 has h_{-1} = 0: no operand to fetch
      // flags count the steps published in THIS launch; h_{t0-1} comes from the previous launch (target 0)
      const unsigned int target = (unsigned int)(t - g.t0);
      bool ok = false;
      int kb = 0;
      uint32_t spins = 0;
      while (kb < g.KB) {
        if (!ok) {
          // relaxed poll: the data is only ever read by TMA (L2)
          if (kPerKb == 2) {
            if (lane >= g.n_pairs) {
              ok = true;
            } else {
              unsigned int fv;
              asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(fv) : "l"(flags + lane) : "memory");
              ok = fv >= target;
            }
          } else {
            if (2 * lane >= g.n_pairs) {
              ok = true;
            } else {
              unsigned long long fv;
              asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(fv) : "l"(flags + 2 * lane) : "memory");
              ok = ((unsigned int)fv >= target) && ((unsigned int)(fv >> 32) >= target);
            }
          }
          if (++spins > (1u << 26)) __trap();
        }
        const unsigned int m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0 && ((m >> (2 * kb)) & 3u) == 3u) {
          if (kTrace && kb == 0) trace[t * kPairTraceSlots + 0] = globaltimer_ns();
          if (g.variant & 2) asm volatile("fence.acq_rel.gpu;" ::: "memory");
          if (g.variant & 4) fence_proxy_async_global();
        }
        while (kb < g.KB && ((m >> (2 * kb)) & 3u) == 3u) {
          if (lane == 0) {
            const int s = it % S;
            mbar_wait(BAR(S + s), ((it / S) & 1u) ^ 1u);
            if (rank == 0) mbar_arrive_expect_tx(BAR(s), 32768u);
            tma_load_3d_2sm(sA + s * 16384u, &maps.h, kb * 64, t - 1, (int)rank * 128, LBAR(s));
            if (kTrace && kb == 0) trace[t * kPairTraceSlots + 1] = globaltimer_ns();
          }
          ++kb;
          ++it;
        }
      }
      if (kTrace && lane == 0) trace[t * kPairTraceSlots + 2] = globaltimer_ns();
    }
  } else if (warp == 1) {
    // ================= MMA issuer: leader CTA only =================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_pair(NP);
      uint32_t it = 0;
      const int tb = g.t0 > 1 ? g.t0 : 1;  // first step with an MMA
      for (int t = tb; t < g.t1; ++t) {
        const uint32_t k = (uint32_t)(t - tb), a = k & 1u;
        if (k >= 2) mbar_wait(BAR(kBarTE + a), ((k >> 1) - 1u) & 1u);  // step t-2 has left this accumulator
        tc_fence_after();
        const uint32_t d = tmem_acc + a * (uint32_t)NP;
        for (int kb = 0; kb < g.KB; ++kb, ++it) {
          const int s = it % S;
          mbar_wait(BAR(s), (it / S) & 1u);
          tc_fence_after();
          if (elect_one_sync()) {
            if (kTrace && kb == 0) trace[t * kPairTraceSlots + 3] = globaltimer_ns();
            const uint32_t a_lo = desc_lo(sA + s * 16384u);
            const uint32_t b_lo = desc_lo(sW + kb * kWTile);
            umma2_f16_lo(d, a_lo, b_lo, idesc, kb != 0);
            umma2_f16_lo(d, a_lo + 2, b_lo + 2, idesc, 1);
            umma2_f16_lo(d, a_lo + 4, b_lo + 4, idesc, 1);
            umma2_f16_lo(d, a_lo + 6, b_lo + 6, idesc, 1);
            umma2_commit_mc(BAR(S + s), 3);
            if (kb == g.KB - 1) {
              umma2_commit_mc(BAR(kBarTF + a), 3);
              if (kTrace) trace[t * kPairTraceSlots + 4] = globaltimer_ns();
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================= cell update: warps 2.. =================
    constexpr int kCols = 4 * NP / kEW;  // accumulator columns per warp (32 or 64)
    constexpr int kU = kCols / 4;        // hidden units per thread (8 or 16)
    constexpr int kG = kU / 8;           // 32-column TMEM loads per thread
    static_assert(kCols == 32 || kCols == 64, "cell-update warps cover 32 or 64 accumulator columns");
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int chunk = (warp - 2) >> 2;
    const int b = (int)rank * 128 + q * 32 + lane;
    const bool row_ok = b < g.B;
    const int len = row_ok ? g.lengths[b] : 0;
    const int H4 = 4 * g.H;
    const bool tr = kTrace && warp == 2 && lane == 0;
    const int unit0 = pair * kUnitsPair + chunk * kU;  // first hidden unit of this thread
    float c[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) c[u] = 0.f;
    if (g.t0 > 0 && row_ok) {
      const float4* cs = reinterpret_cast<const float4*>(g.c_state + (int64_t)b * g.H + unit0);
#pragma unroll
      for (int u = 0; u < kU / 4; ++u) {
        const float4 v4 = cs[u];
        c[4 * u] = v4.x; c[4 * u + 1] = v4.y; c[4 * u + 2] = v4.z; c[4 * u + 3] = v4.w;
      }
    }
    const int tb = g.t0 > 1 ? g.t0 : 1;
    // lane = batch row: a warp's load of one unit's gates is one contiguous 512-byte run
    const float4* xcol = g.xT + (int64_t)unit0 * g.Bp + (row_ok ? b : 0);
    __nv_bfloat16* hrow = g.hseq + ((int64_t)(row_ok ? b : 0) * g.T) * g.H + unit0;
    const uint32_t t_row = tmem_acc + (uint32_t)(chunk * kCols) + ((uint32_t)(q * 32) << 16);
    for (int t = g.t0; t < g.t1; ++t) {
      // input projection of this step: independent of h, requested before the wait on the accumulator
      float4 x[kU];
      if (g.variant & 256) {  // diagnostic: no input-projection loads (wrong results; what do these loads cost?)
#pragma unroll
        for (int u = 0; u < kU; ++u) x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {  // rows past B read row 0 and store nothing
        const float4* xp = xcol + (int64_t)t * g.H * g.Bp;
#pragma unroll
        for (int u = 0; u < kU; ++u, xp += g.Bp) x[u] = __ldg(xp);
      }
      const uint32_t k = (uint32_t)(t - tb), a = k & 1u;
      if (t > 0) {
        mbar_wait(BAR(kBarTF + a), (k >> 1) & 1u);
        tc_fence_after();
        if (tr) trace[t * kPairTraceSlots + 5] = globaltimer_ns();
      }
      const bool live = t < len;
      uint32_t hp[kU / 2];
      __nv_bfloat16* gsave = (g.gates_out && row_ok) ? g.gates_out + ((int64_t)b * g.T + t) * H4 + 4 * unit0 : nullptr;
#pragma unroll
      for (int gi = 0; gi < kG; ++gi) {  // 32 accumulator columns = 8 hidden units at a time (bounds the live registers)
        uint32_t v[32];
        if (t > 0) {
          tmem_ld32(t_row + a * (uint32_t)NP + gi * 32, v);
          tmem_ld_wait();
          if (gi == kG - 1) {
            if (tr) trace[t * kPairTraceSlots + 6] = globaltimer_ns();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(LBAR(kBarTE + a));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {  // groups of four units = one 32-byte sector of the gate tape
          u32x8 gp;
          float hv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int u = 8 * gi + 4 * k + j;
            const int o = (4 * k + j) * 4;
            const float gi_ = sigmoid_hw(__uint_as_float(v[o + 0]) + x[u].x);
            const float gf = sigmoid_hw(__uint_as_float(v[o + 1]) + x[u].y);
            const float gg = tanh_hw(__uint_as_float(v[o + 2]) + x[u].z);
            const float go = sigmoid_hw(__uint_as_float(v[o + 3]) + x[u].w);
            c[u] = gf * c[u] + gi_ * gg;
            hv[j] = go * tanh_hw(c[u]);
            gp.v[2 * j] = pack_bf16x2(gi_, gf);
            gp.v[2 * j + 1] = pack_bf16x2(gg, go);
          }
          hp[4 * gi + 2 * k] = live ? pack_bf16x2(hv[0], hv[1]) : 0u;
          hp[4 * gi + 2 * k + 1] = live ? pack_bf16x2(hv[2], hv[3]) : 0u;
          if (gsave) st_global_256(gsave + 32 * gi + 16 * k, gp);  // training tape: post-activation gates, bf16
        }
      }
      if (row_ok) {
        // h_t: what the other CTAs wait for (16 units per thread = one full 32-byte sector)
        if (kU == 16) {
          u32x8 o;
#pragma unroll
          for (int e = 0; e < 8; ++e) o.v[e] = hp[e & (kU / 2 - 1)];
          st_global_256(hrow + (int64_t)t * g.H, o);
        } else {
          *reinterpret_cast<uint4*>(hrow + (int64_t)t * g.H) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        }
        if (g.c_out) {  // training tape: cell state, f32
          float* cp = g.c_out + ((int64_t)b * g.T + t) * g.H + unit0;
#pragma unroll
          for (int k = 0; k < kU / 8; ++k) {
            u32x8 o;
#pragma unroll
            for (int e = 0; e < 8; ++e) o.v[e] = __float_as_uint(c[8 * k + e]);
            st_global_256(cp + 8 * k, o);
          }
        }
      }
      // publish: all cell-update warps have stored their part of h_t; one thread makes it visible and raises the flag
      if (tr) trace[t * kPairTraceSlots + 7] = globaltimer_ns();
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEW) : "memory");
      if (warp == 2 && lane == 0) {
        if (kTrace) trace[t * kPairTraceSlots + 9] = globaltimer_ns();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + pair), "r"((unsigned int)(t + 1 - g.t0))
                     : "memory");
        if (kTrace) trace[t * kPairTraceSlots + 10] = globaltimer_ns();
      }
    }
    if (row_ok && g.c_state) {  // hand the cell state to the next chunk
      float4* cs = reinterpret_cast<float4*>(g.c_state + (int64_t)b * g.H + unit0);
#pragma unroll
      for (int u = 0; u < kU / 4; ++u) cs[u] = make_float4(c[4 * u], c[4 * u + 1], c[4 * u + 2], c[4 * u + 3]);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA exits (or frees TMEM) while its peer may still arrive on its barriers / use its operands
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc2(tmem_acc, 2 * NP);
  }
}

}  // namespace tc
}  // namespace avvad
