// bf16 tensor-core engine for sm_100a: C[M][N] = A[M][K] * W[N][K]^T with fp32 accumulation in TMEM.
//
// Persistent CTAs (one or two per SM) loop over 128 x BN output tiles with three warp roles:
//   warps 0-3 (128 threads) : operand producers.  They gather 128-byte K-slices (64 bf16) of A rows and
//                             W rows with 16-byte cp.async (zero-filling out-of-image taps / tail rows)
//                             into the canonical K-major SWIZZLE_128B shared-memory layout, then
//                             fence.proxy.async + mbarrier-arrive so the tensor core may read them.
//                             The A "row" is either a plain matrix row (GEMM) or an im2col row of an
//                             NHWC convolution (tap-major K order), so convolutions never materialise
//                             im2col in HBM.  The ring runs ahead across tile boundaries.
//   warp 4                  : allocates TMEM (2 accumulator stages of BN fp32 columns); its lane 0 issues
//                             tcgen05.mma (UMMA 128 x BN x 16, kind::f16, bf16 operands from shared-memory
//                             descriptors) and tcgen05.commit to free ring stages / publish an accumulator.
//   warps 5-8 (128 threads) : epilogue.  tcgen05.ld TMEM -> registers -> fused bias / residual / ReLU / LSTM
//                             cell -> global, overlapped with the next tile's main loop.
// Pipeline: kStages-deep ring of {A 16 KB, W BN*128 B} stages (full/empty mbarriers) + tmem full/empty.
#pragma once
#include "common.cuh"

namespace avvad {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;  // bf16 elements per K block = 128 bytes = one swizzle row
constexpr int kProducerThreads = 128;
constexpr int kThreads = 288;  // 4 producer warps + 1 MMA warp + 4 epilogue warps
constexpr int kLag = 2;  // cp.async groups in flight before a stage is published

enum AMode { A_PLAIN = 0, A_CONV = 1, A_CONV1 = 2 };
enum EpiMode { EPI_BF16 = 0, EPI_F32 = 1, EPI_LSTM = 2 };

struct AParams {
  const __nv_bfloat16* A;  // plain: [M][lda]; conv: NHWC input [n][H][W][Cin]
  const float* A32;        // A_CONV1: fp32 single-channel frames [n][H][W] (7x7 / stride 2 / pad 3 im2col, K 49 -> 64)
  int64_t lda;
  int H, W, Cin, OH, OW, R, S, stride, pad, cpb;  // conv only; cpb = Cin / 64
  int use_ca;                                     // conv only: cache A gathers in L1 (tap re-reads)
};

struct EpiParams {
  const float* bias;              // [N] or null
  const __nv_bfloat16* residual;  // [M][ldc] or null (EPI_BF16)
  void* C;                        // bf16 or f32 [M][ldc]
  int64_t ldc;
  int relu;
  // EPI_LSTM: accumulator column n = 4*u + gate (i,f,g,o); row = batch index b
  const float* xproj;     // [B][T][4H] gate-interleaved input projection (+ both biases)
  float* c_state;         // [B][H]
  __nv_bfloat16* h_next;  // [B][H]   A operand of the next step
  __nv_bfloat16* hseq;    // [B][T][H] layer output (zero for t >= len_b)
  const int32_t* lengths;
  int t, T, H4;
  int debug;  // experiments only (AVVAD_EPI_DEBUG): bit 0 = skip the output stores of the fast epilogue
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (reported as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16_ca(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (=1, unused for swizzled K-major) |
//   [32,46) SBO >> 4 (=1024 B between 8-row groups) | [46,48) version = 1 | [61,64) layout = 2 (SW128)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D=f32 (bit4), A=B=bf16 (bits 7,10), K-major A and B, N>>3 at [17,23),
// M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}
// Same instruction with the two descriptors given as (low word, shared high word): the high word of a K-major
// SWIZZLE_128B descriptor (SBO, version, layout) is constant, so stepping a descriptor is one 32-bit add on the
// low word ((address >> 4) | LBO << 16) -- keeps the single issuing thread ahead of 48-cycle MMAs.
constexpr uint32_t kDescHi = (64u) | (1u << 14) | (2u << 29);  // SBO=1024>>4 | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void umma_f16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                            uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accum), "r"(kDescHi)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets row (lane base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// One elected lane of a fully converged warp (deterministic for a given mask).  Issuing tcgen05.mma / commit / TMA from
// `if (elect_one_sync())` inside warp-uniform control flow lets the compiler keep descriptors in uniform registers and
// emit the instructions back to back; the older `if (lane == 0) { whole loop }` form made every issue pay ~10-20
// scalar instructions (R2UR, ELECT, BRA.U.ANY loops), which was the bound for N <= 128 tiles (48-64 cycle MMAs).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int BN, int AMODE = A_PLAIN>
struct Cfg {
  static constexpr int kStages = (BN == 256) ? 3 : ((BN == 128) ? 3 : 4);
  static constexpr uint32_t kABytes = BM * 128;
  static constexpr uint32_t kBBytes = BN * 128;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  // resident CTAs per SM: shared memory (227 KB) and TMEM (512 columns, 2*BN per CTA) permitting
  static constexpr int kCtasPerSm = (BN == 256) ? 1 : 2;
};

// ---- epilogue for one 32-column chunk held by one thread (row m, columns n0..n0+31) ----------------
__device__ __forceinline__ void epilogue_chunk(const EpiParams& ep, int mode, const uint32_t (&v)[32], int64_t m,
                                               int n0, int N) {
  if (mode == EPI_LSTM) {
    // 8 hidden units x (i,f,g,o)
    const int b = (int)m;
    const int u0 = n0 >> 2;
    const int H = ep.H4 >> 2;
    const float* xp = ep.xproj + ((int64_t)b * ep.T + ep.t) * ep.H4 + n0;
    float* cs = ep.c_state + (int64_t)b * H + u0;
    const bool live = ep.t < ep.lengths[b];
    float hv[8];
    float4 c_lo = *reinterpret_cast<const float4*>(cs);
    float4 c_hi = *reinterpret_cast<const float4*>(cs + 4);
    float cv[8] = {c_lo.x, c_lo.y, c_lo.z, c_lo.w, c_hi.x, c_hi.y, c_hi.z, c_hi.w};
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float4 x4 = *reinterpret_cast<const float4*>(xp + 4 * u);
      const float gi = __uint_as_float(v[4 * u + 0]) + x4.x;
      const float gf = __uint_as_float(v[4 * u + 1]) + x4.y;
      const float gg = __uint_as_float(v[4 * u + 2]) + x4.z;
      const float go = __uint_as_float(v[4 * u + 3]) + x4.w;
      const float c = sigmoidf_fast(gf) * cv[u] + sigmoidf_fast(gi) * tanhf_fast(gg);
      cv[u] = c;
      hv[u] = sigmoidf_fast(go) * tanhf_fast(c);
    }
    *reinterpret_cast<float4*>(cs) = make_float4(cv[0], cv[1], cv[2], cv[3]);
    *reinterpret_cast<float4*>(cs + 4) = make_float4(cv[4], cv[5], cv[6], cv[7]);
    uint4 hp;
    hp.x = pack_bf16x2(hv[0], hv[1]);
    hp.y = pack_bf16x2(hv[2], hv[3]);
    hp.z = pack_bf16x2(hv[4], hv[5]);
    hp.w = pack_bf16x2(hv[6], hv[7]);
    *reinterpret_cast<uint4*>(ep.h_next + (int64_t)b * H + u0) = hp;
    if (!live) hp = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(ep.hseq + ((int64_t)b * ep.T + ep.t) * H + u0) = hp;
    return;
  }
  // NOTE: every access to f[] below uses a compile-time index (fully unrolled loops with predicates); a runtime
  // index would push the array to local memory and turn the whole epilogue into L1 traffic.
  float f[32];
#pragma unroll
  for (int q = 0; q < 32; ++q) f[q] = __uint_as_float(v[q]);
  const bool full = (n0 + 32 <= N);
  if (ep.bias) {
    if (full) {
      const float4* bp = reinterpret_cast<const float4*>(ep.bias + n0);  // n0 % 32 == 0 -> 16-byte aligned
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 b4 = __ldg(bp + q);
        f[4 * q + 0] += b4.x; f[4 * q + 1] += b4.y; f[4 * q + 2] += b4.z; f[4 * q + 3] += b4.w;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 32; ++q)
        if (n0 + q < N) f[q] += __ldg(ep.bias + n0 + q);
    }
  }
  if (mode == EPI_BF16) {
    __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(ep.C) + m * ep.ldc + n0;
    if (full) {
      if (ep.residual) {
        const uint4* rp = reinterpret_cast<const uint4*>(ep.residual + m * ep.ldc + n0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 r = rp[q];
          float2 a = unpack_bf16x2(r.x), b2 = unpack_bf16x2(r.y), c2 = unpack_bf16x2(r.z), d2 = unpack_bf16x2(r.w);
          f[8 * q + 0] += a.x;  f[8 * q + 1] += a.y;  f[8 * q + 2] += b2.x; f[8 * q + 3] += b2.y;
          f[8 * q + 4] += c2.x; f[8 * q + 5] += c2.y; f[8 * q + 6] += d2.x; f[8 * q + 7] += d2.y;
        }
      }
      if (ep.relu) {
#pragma unroll
        for (int q = 0; q < 32; ++q) f[q] = fmaxf(f[q], 0.f);
      }
      uint4* cp = reinterpret_cast<uint4*>(crow);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 o;
        o.x = pack_bf16x2(f[8 * q + 0], f[8 * q + 1]);
        o.y = pack_bf16x2(f[8 * q + 2], f[8 * q + 3]);
        o.z = pack_bf16x2(f[8 * q + 4], f[8 * q + 5]);
        o.w = pack_bf16x2(f[8 * q + 6], f[8 * q + 7]);
        cp[q] = o;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        if (n0 + q < N) {
          float x = f[q];
          if (ep.residual) x += __bfloat162float(ep.residual[m * ep.ldc + n0 + q]);
          if (ep.relu) x = fmaxf(x, 0.f);
          crow[q] = __float2bfloat16_rn(x);
        }
      }
    }
  } else {  // EPI_F32
    float* crow = reinterpret_cast<float*>(ep.C) + m * ep.ldc + n0;
    if (ep.relu) {
#pragma unroll
      for (int q = 0; q < 32; ++q) f[q] = fmaxf(f[q], 0.f);
    }
    if (full && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
      float4* cp = reinterpret_cast<float4*>(crow);
#pragma unroll
      for (int q = 0; q < 8; ++q) cp[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
    } else {
#pragma unroll
      for (int q = 0; q < 32; ++q)
        if (n0 + q < N) crow[q] = f[q];
    }
  }
}

// ---- the kernel -------------------------------------------------------------------------------------
// Persistent and warp-specialised: every CTA loops over output tiles (tile = blockIdx.x + i*gridDim.x, n-tile
// fastest so concurrently running CTAs share A rows in L2).  The operand ring runs across tile boundaries and the
// accumulator is double-buffered in TMEM, so the epilogue of tile i overlaps the main loop of tile i+1.
template <int BN, int AMODE>
__global__ void __launch_bounds__(kThreads)
tc_gemm_kernel(AParams ap, const __nv_bfloat16* __restrict__ Wt, int64_t ldw, int64_t M, int N, int KB,
               int n_tiles, int64_t total_tiles, EpiParams ep, int epi_mode) {
  using C = Cfg<BN, AMODE>;
  constexpr int S = C::kStages;
  constexpr uint32_t kTmemCols = 2 * BN;  // two accumulator stages (BN in {64,128,256} -> power of two >= 32)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  // barriers: full[s] = bar0+8s | empty[s] = bar0+8(S+s) | tfull[a] = bar0+8(2S+a) | tempty[a] = bar0+8(2S+2+a)
  const uint32_t bar0 = base + S * C::kStageBytes;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + S * C::kStageBytes + 8 * (2 * S + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar0 + 8 * s, kProducerThreads);
      mbar_init(bar0 + 8 * (S + s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar0 + 8 * (2 * S + a), 1);
      mbar_init(bar0 + 8 * (2 * S + 2 + a), 4);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp < 4) {
    // ================= producers =================
    const int p = threadIdx.x;
    const int chunk = p & 7;
    const int r0 = p >> 3;  // this thread feeds rows r0 + 16 i
    const uint32_t sw_off = (uint32_t)(((r0 >> 3) << 10) + ((r0 & 7) << 7) + ((chunk ^ (r0 & 7)) << 4));
    uint32_t it = 0;  // global K-block counter of this CTA (ring position)
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_base = (int)(tile % n_tiles) * BN;
      const int64_t m_base = (tile / n_tiles) * BM;
      const __nv_bfloat16* wrow = Wt + (int64_t)(n_base + r0) * ldw + chunk * 8;

      if (AMODE == A_CONV1) {
        // Stem convolution (7x7 / stride 2 / pad 3 on an fp32 single-channel frame, K 49 -> 64, one K block per
        // tile): thread p builds im2col row p with read-only loads (L1 serves the overlap between neighbouring
        // rows) and stores its 8 swizzled 16-byte chunks.
        const int s = it % S;
        const uint32_t ph = (it / S) & 1u;
        mbar_wait(bar0 + 8 * (S + s), ph ^ 1u);
        const uint32_t stage = base + s * C::kStageBytes;
        const uint32_t sb = stage + C::kABytes + sw_off;
#pragma unroll
        for (int i = 0; i < BN / 16; ++i) {
          const bool ok = (n_base + r0 + 16 * i) < N;
          cp_async16(sb + i * 2048, ok ? (const void*)(wrow + (int64_t)i * 16 * ldw) : (const void*)Wt, ok ? 16u : 0u);
        }
        cp_async_commit();
        const int64_t m = m_base + p;
        uint32_t pk[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) pk[q] = 0u;
        if (m < M) {
          const int ohw = ap.OH * ap.OW;
          const int64_t n = m / ohw;
          const int rem = (int)(m - n * ohw);
          const int oh = rem / ap.OW, ow = rem - oh * ap.OW;
          const float* img = ap.A32 + n * ap.H * ap.W;
          const int ih0 = oh * 2 - 3, iw0 = ow * 2 - 3;
          float vals[50];
#pragma unroll
          for (int r = 0; r < 7; ++r) {
            const int ih = ih0 + r;
            const bool rok = (unsigned)ih < (unsigned)ap.H;
#pragma unroll
            for (int c = 0; c < 7; ++c) {
              const int iw = iw0 + c;
              vals[r * 7 + c] = (rok && (unsigned)iw < (unsigned)ap.W) ? __ldg(img + ih * ap.W + iw) : 0.f;
            }
          }
          vals[49] = 0.f;
#pragma unroll
          for (int q = 0; q < 25; ++q) pk[q] = pack_bf16x2(vals[2 * q], vals[2 * q + 1]);
        }
        const uint32_t rowoff = (uint32_t)(((p >> 3) << 10) + ((p & 7) << 7));
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t dst = stage + rowoff + (uint32_t)((c ^ (p & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * c]), "r"(pk[4 * c + 1]),
                       "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3])
                       : "memory");
        }
        cp_async_wait<0>();
        fence_proxy_async();
        mbar_arrive(bar0 + 8 * s);
        ++it;
        continue;
      }

      // per-row state for the 8 A rows this thread feeds
      int64_t a_off[8];
      int ihw0[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t m = m_base + r0 + 16 * i;
        if (AMODE == A_PLAIN) {
          a_off[i] = (m < M) ? m * ap.lda + chunk * 8 : -1;
          ihw0[i] = 0;
        } else {
          if (m < M) {
            const int ohw = ap.OH * ap.OW;
            const int64_t n = m / ohw;
            const int rem = (int)(m - n * ohw);
            const int oh = rem / ap.OW;
            const int ow = rem - oh * ap.OW;
            a_off[i] = n * ap.H * ap.W;  // pixel index of the frame origin
            const int ih0 = oh * ap.stride - ap.pad, iw0 = ow * ap.stride - ap.pad;
            ihw0[i] = (int)(((uint32_t)(ih0 + 1024) << 16) | (uint32_t)(iw0 + 1024));
          } else {
            a_off[i] = -1;
            ihw0[i] = 0;
          }
        }
      }

      for (int kb = 0; kb < KB; ++kb, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1u;
        mbar_wait(bar0 + 8 * (S + s), ph ^ 1u);
        const uint32_t sa = base + s * C::kStageBytes + sw_off;
        const uint32_t sb = sa + C::kABytes;
        if (AMODE == A_PLAIN) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const bool ok = a_off[i] >= 0;
            cp_async16(sa + i * 2048, ok ? (const void*)(ap.A + a_off[i] + (int64_t)kb * BK) : (const void*)ap.A,
                       ok ? 16u : 0u);
          }
        } else {
          const int tap = kb / ap.cpb;
          const int cb = kb - tap * ap.cpb;
          const int fr = tap / ap.S;
          const int fs = tap - fr * ap.S;
          const int coff = cb * BK + chunk * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int ih = (int)((uint32_t)ihw0[i] >> 16) - 1024 + fr;
            const int iw = (int)((uint32_t)ihw0[i] & 0xFFFFu) - 1024 + fs;
            const bool ok = (a_off[i] >= 0) && ((unsigned)ih < (unsigned)ap.H) && ((unsigned)iw < (unsigned)ap.W);
            const int64_t off = (a_off[i] + (int64_t)ih * ap.W + iw) * ap.Cin + coff;
            cp_async16(sa + i * 2048, ok ? (const void*)(ap.A + off) : (const void*)ap.A, ok ? 16u : 0u);
          }
        }
#pragma unroll
        for (int i = 0; i < BN / 16; ++i) {
          const bool ok = (n_base + r0 + 16 * i) < N;
          cp_async16(sb + i * 2048,
                     ok ? (const void*)(wrow + (int64_t)i * 16 * ldw + (int64_t)kb * BK) : (const void*)Wt,
                     ok ? 16u : 0u);
        }
        cp_async_commit();
        if (it >= (uint32_t)kLag) {
          cp_async_wait<kLag>();
          fence_proxy_async();
          mbar_arrive(bar0 + 8 * ((it - kLag) % S));
        }
      }
    }
    if (AMODE != A_CONV1 && it > 0) {
      // drain: publish the last min(kLag, it) stages
      if (it >= 2) {
        cp_async_wait<1>();
        fence_proxy_async();
        mbar_arrive(bar0 + 8 * ((it - 2) % S));
      }
      cp_async_wait<0>();
      fence_proxy_async();
      mbar_arrive(bar0 + 8 * ((it - 1) % S));
    }
  } else if (warp == 4) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      uint32_t it = 0, tl = 0;
      for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
        const uint32_t acc = tl & 1u, aph = (tl >> 1) & 1u;
        mbar_wait(bar0 + 8 * (2 * S + 2 + acc), aph ^ 1u);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_acc + acc * BN;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1u;
          mbar_wait(bar0 + 8 * s, ph);
          tc_fence_after();
          const uint32_t sa = base + s * C::kStageBytes;
          const uint32_t sb = sa + C::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16(d_tmem, make_sw128_desc(sa + k * 32), make_sw128_desc(sb + k * 32), idesc, (kb | k) != 0);
          umma_commit(bar0 + 8 * (S + s));
        }
        umma_commit(bar0 + 8 * (2 * S + acc));
      }
    }
  } else {
    // ================= epilogue warps 5..8 =================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    uint32_t tl = 0;
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
      const int n_base = (int)(tile % n_tiles) * BN;
      const int64_t m_base = (tile / n_tiles) * BM;
      const uint32_t acc = tl & 1u, aph = (tl >> 1) & 1u;
      mbar_wait(bar0 + 8 * (2 * S + acc), aph);
      tc_fence_after();
      const int64_t m = m_base + q * 32 + lane;
      const uint32_t t_row = tmem_acc + acc * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int j = 0; j < BN / 32; ++j) {
        uint32_t v[32];
        tmem_ld32(t_row + j * 32, v);
        tmem_ld_wait();
        const int n0 = n_base + j * 32;
        if (m < M && n0 < N) epilogue_chunk(ep, epi_mode, v, m, n0, N);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar0 + 8 * (2 * S + 2 + acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    tmem_dealloc(tmem_acc, kTmemCols);
  }
}

// ---- host-side launcher -----------------------------------------------------------------------------
int launch(int amode, const AParams& ap, const __nv_bfloat16* Wt, int64_t ldw, int64_t M, int N, int K,
           const EpiParams& ep, int epi_mode, int bn_hint, cudaStream_t st);
int prof_begin(cudaStream_t st, void** tok);
void prof_end(cudaStream_t st, void* tok, int cat, double flops);
int prof_group_begin(cudaStream_t st, void** tok);
void prof_group_end(cudaStream_t st, void* tok, int cat, double flops);
int gemm_dispatch(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* Wt, int64_t ldw, int64_t M, int N, int K,
                  const EpiParams& ep, int epi_mode, int bn_hint, cudaStream_t st);

}  // namespace tc
}  // namespace avvad
