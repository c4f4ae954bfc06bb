// Persistent LSTM recurrence for one layer (sm_100a): all T time steps in ONE cooperative launch.
//
// The 4H gate columns (gate-interleaved: column 4u+g) are split into slices of 64 columns (16 hidden units) and the
// batch into slices of 128 rows; CTA (ns, ms) keeps its W_hh slice [64][H] bf16 resident in shared memory for the
// whole sequence (H=1024: 128 KB, loaded once by TMA) and the cell state of its 128 x 16 (row, unit) pairs in
// registers.  Per time step:
//   warp 0 : waits until every CTA of the same batch slice has published h_{t-1} (one global counter per batch
//            slice), then streams h_{t-1}[128 rows][H] through a 4-stage TMA ring (box {64, 1, 128} of the
//            [B][T][H] output sequence -- the layer output doubles as the recurrent operand)
//   warp 1 : tcgen05.mma 128 x 64 x 16 over K = H into a 64-column TMEM accumulator
//   warps 2-5: have already fetched xproj[b][t][64 columns] (independent of h), wait for the accumulator, apply the
//            gate non-linearities, update c (registers), store h_t as bf16 (zero beyond the sequence length), fence,
//            and one thread bumps the batch slice's counter.
// Reference semantics: nn.LSTM inside packages/models/AV_Net.py:128-137 (gates i,f,g,o; zero initial state).
#pragma once
#include "gemm_tma.cuh"

namespace avvad {
namespace tc {

constexpr int kLstmThreads = 192;
constexpr int kLstmStages = 4;

struct LstmGeom {
  int B, T, H, KB;      // KB = H / 64
  int n_slices;         // 4H / 64
  const float* xproj;   // [B][T][4H] fp32, gate-interleaved, both biases folded in
  __nv_bfloat16* hseq;  // [B][T][H] bf16 layer output
  const int32_t* lengths;
  unsigned int* counters;  // [m_slices], zeroed before the launch
  // training only (may be null): post-activation gates (i,f,g,o per unit, bf16 [B][T][4H]) and cell states (f32 [B][T][H])
  __nv_bfloat16* gates_out;
  float* c_out;
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
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

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
  // barriers: full[4] | empty[4] | wfull | tfull
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + (bar0 - base) + 8 * 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ns = blockIdx.x % g.n_slices;
  const int ms = blockIdx.x / g.n_slices;
  const unsigned int n_peers = (unsigned int)g.n_slices;
  unsigned int* counter = g.counters + ms;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kLstmStages; ++s) {
      mbar_init(BAR(s), 1);
      mbar_init(BAR(4 + s), 1);
    }
    mbar_init(BAR(8), 1);
    mbar_init(BAR(9), 1);
    fence_barrier_init();
    tma_prefetch_desc(&maps.h);
    tma_prefetch_desc(&maps.w);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // resident weights
      mbar_arrive_expect_tx(BAR(8), w_bytes);
      for (int kb = 0; kb < g.KB; ++kb) tma_load_2d(sW + kb * 8192u, &maps.w, kb * 64, ns * 64, BAR(8));
      uint32_t it = 0;
      for (int t = 1; t < g.T; ++t) {  // step 0 has h_{-1} = 0: no operand to fetch
        // wait until all CTAs of this batch slice have published h_{t-1}
        const unsigned int target = (unsigned int)t * n_peers;
        uint32_t spins = 0;
        while (ld_acquire_gpu(counter) < target) {
          __nanosleep(32);
          if (++spins > (1u << 26)) __trap();
        }
        fence_proxy_async_all();  // peers wrote h through the generic proxy; TMA reads through the async proxy
        for (int kb = 0; kb < g.KB; ++kb, ++it) {
          const int s = it % kLstmStages;
          mbar_wait(BAR(4 + s), ((it / kLstmStages) & 1u) ^ 1u);
          mbar_arrive_expect_tx(BAR(s), 16384u);
          tma_load_3d(sA + s * 16384u, &maps.h, kb * 64, t - 1, ms * 128, BAR(s));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(64);
      mbar_wait(BAR(8), 0);
      tc_fence_after();
      uint32_t it = 0;
      for (int t = 1; t < g.T; ++t) {
        for (int kb = 0; kb < g.KB; ++kb, ++it) {
          const int s = it % kLstmStages;
          mbar_wait(BAR(s), (it / kLstmStages) & 1u);
          tc_fence_after();
          const uint32_t a_lo = desc_lo(sA + s * 16384u);
          const uint32_t b_lo = desc_lo(sW + kb * 8192u);
          umma_f16_lo(tmem_acc, a_lo, b_lo, idesc, kb != 0);
          umma_f16_lo(tmem_acc, a_lo + 2, b_lo + 2, idesc, 1);
          umma_f16_lo(tmem_acc, a_lo + 4, b_lo + 4, idesc, 1);
          umma_f16_lo(tmem_acc, a_lo + 6, b_lo + 6, idesc, 1);
          umma_commit(BAR(4 + s));
        }
        umma_commit(BAR(9));
      }
    }
  } else {
    // ================= epilogue / cell update: warps 2..5 =================
    const int q = warp & 3;
    const int b = ms * 128 + q * 32 + lane;
    const bool row_ok = b < g.B;
    const int len = row_ok ? g.lengths[b] : 0;
    const int H4 = 4 * g.H;
    float c[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) c[u] = 0.f;
    const float* xrow = g.xproj + ((int64_t)(row_ok ? b : 0) * g.T) * H4 + ns * 64;
    __nv_bfloat16* hrow = g.hseq + ((int64_t)(row_ok ? b : 0) * g.T) * g.H + ns * 16;
    for (int t = 0; t < g.T; ++t) {
      float4 x[16];
      if (row_ok) {
        const float4* xp = reinterpret_cast<const float4*>(xrow + (int64_t)t * H4);
#pragma unroll
        for (int u = 0; u < 16; ++u) x[u] = __ldg(xp + u);
      }
      uint32_t v0[32], v1[32];
      if (t > 0) {
        mbar_wait(BAR(9), (uint32_t)(t - 1) & 1u);
        tc_fence_after();
        tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16), v0);
        tmem_ld32(tmem_acc + 32u + ((uint32_t)(q * 32) << 16), v1);
        tmem_ld_wait();
        tc_fence_before();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v0[i] = v1[i] = 0u;
      }
      if (row_ok) {
        float hv[16];
        __nv_bfloat16* gsave = g.gates_out ? g.gates_out + ((int64_t)b * g.T + t) * H4 + ns * 64 : nullptr;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const uint32_t* vv = (u < 8) ? v0 : v1;
          const int o = (u & 7) * 4;
          const float gi = sigmoidf_fast(__uint_as_float(vv[o + 0]) + x[u].x);
          const float gf = sigmoidf_fast(__uint_as_float(vv[o + 1]) + x[u].y);
          const float gg = tanhf_fast(__uint_as_float(vv[o + 2]) + x[u].z);
          const float go = sigmoidf_fast(__uint_as_float(vv[o + 3]) + x[u].w);
          c[u] = gf * c[u] + gi * gg;
          hv[u] = go * tanhf_fast(c[u]);
          if (gsave) {
            uint2 pk;
            pk.x = pack_bf16x2(gi, gf);
            pk.y = pack_bf16x2(gg, go);
            *reinterpret_cast<uint2*>(gsave + 4 * u) = pk;
          }
        }
        if (g.c_out) {
          float4* cp = reinterpret_cast<float4*>(g.c_out + ((int64_t)b * g.T + t) * g.H + ns * 16);
#pragma unroll
          for (int u4 = 0; u4 < 4; ++u4) cp[u4] = make_float4(c[4 * u4], c[4 * u4 + 1], c[4 * u4 + 2], c[4 * u4 + 3]);
        }
        const bool live = t < len;
        uint4 h0, h1;
        h0.x = live ? pack_bf16x2(hv[0], hv[1]) : 0u;
        h0.y = live ? pack_bf16x2(hv[2], hv[3]) : 0u;
        h0.z = live ? pack_bf16x2(hv[4], hv[5]) : 0u;
        h0.w = live ? pack_bf16x2(hv[6], hv[7]) : 0u;
        h1.x = live ? pack_bf16x2(hv[8], hv[9]) : 0u;
        h1.y = live ? pack_bf16x2(hv[10], hv[11]) : 0u;
        h1.z = live ? pack_bf16x2(hv[12], hv[13]) : 0u;
        h1.w = live ? pack_bf16x2(hv[14], hv[15]) : 0u;
        uint4* hp = reinterpret_cast<uint4*>(hrow + (int64_t)t * g.H);
        hp[0] = h0;
        hp[1] = h1;
      }
      // publish: make the stores visible to the other SMs' TMA reads, then count this CTA in
      fence_proxy_async_all();
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps
      if (warp == 2 && lane == 0) atomicAdd(counter, 1u);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_acc, 64);
  }
}

}  // namespace tc
}  // namespace avvad
