// ResNet stem for single-channel 67x67 ROIs without an im2col build: conv 7x7 / stride 2 / pad 3 (3 identical input
// channels folded into one) + folded BN + ReLU + max-pool 3x3 / stride 2 / pad 1, optionally fused with the 30 -> 62.5
// fps index-exact gather + u8 -> standardised value conversion that precedes it in the pipeline.
//
// Idea: the A operand of the implicit GEMM is the zero-padded image itself.  The padded frame P (73x73) is stored in
// shared memory as two bf16 "planes" (even rows, odd rows) with a 160-byte pitch.  A GEMM row is the anchor
// (y, i) = 16 consecutive pixels P[2y+fr][8i .. 8i+15] (two 16-byte chunks), and because anchors of one image row
// are 16 bytes apart they form the 8-row core matrices of the un-swizzled K-major UMMA layout *in place*:
//     row r of a tile  <->  anchor (y = y0 + r/8, i = r%8):  address = plane + (y0 + r/8) * 160 + (r%8) * 16
//     descriptor: SBO = 160 (next image row), LBO = 16 (next chunk -- overlapping the next anchor, which is legal:
//     descriptors only generate addresses).  Filter row fr = 2a+b is a start-address shift of `a` plane rows.
// One anchor covers the four output columns ow = 4i+s (s = 0..3): the B operand holds four shifted copies of each
// filter row, W[fr][(s,ch)][e] = w[ch][fr][e-2s] (zero outside 0..6), so N = 4 x 64 = 256 and a thread of the
// epilogue owns four horizontally adjacent conv outputs.  With the s=3 value of its left neighbour lane (one
// shuffle) it reduces them to the two horizontally pooled columns pw = 2i, 2i+1 in registers.  The conv columns
// 32, 33 (pooled column 16) come from one extra "edge" tile whose anchors (pixel 60 of every image row) are staged
// as a separate transposed plane.  Horizontally pooled rows go to a 32-row shared-memory ring; the vertical 3-row
// maximum is taken from there and written to global memory as the pooled 17x17x64 bf16 map.
//
// Seven tcgen05.mma (128 x 256 x 16) per tile of 16 image rows + one "bias step" (A = ones, B = folded BN bias split
// into a bf16 high and low part), so the accumulator already holds conv + bias and the epilogue is max / ReLU / pack
// only; a batch of 2 frames is 5 such tiles + 1 edge tile.
//
//   warps 0-7  : builders  (global -> registers one frame ahead -> planes; double-buffered batches)
//   warp  8    : TMEM alloc; one elected lane issues the MMAs
//   warps 9-16 : epilogue  (TMEM -> horizontal pool -> ring; then vertical pool -> global); two warps per TMEM lane
//                quarter, 32 channels each (two 16-channel passes per tile)
#pragma once
#include "gemm_tc.cuh"

namespace avvad {
namespace tc {

constexpr int kS2Builders = 256;
constexpr int kS2BuilderWarps = kS2Builders / 32;
constexpr int kS2EpiWarps = 8;                         // two per TMEM lane quarter (16 warps measured slower: 3.1 vs 2.65 ms)
constexpr int kS2EpiThreads = kS2EpiWarps * 32;
constexpr int kS2ChPerWarp = 64 / (kS2EpiWarps / 4);   // channels of a row one epilogue warp owns (32)
constexpr int kS2Passes = kS2ChPerWarp / 16;           // 16-channel passes per tile and warp
constexpr int kS2Items = (67 * 37 + kS2Builders - 1) / kS2Builders;  // (row, pixel pair) items per builder thread and frame
constexpr int kS2Threads = kS2Builders + 32 + kS2EpiThreads;
constexpr int kS2FramesPerBatch = 2;
constexpr int kS2RowsPerFrame = 37;                    // plane rows per frame: y = 0..36 (34 outputs + 3 filter-row halo)
constexpr int kS2Pitch = 160;                          // bytes per plane row: 80 bf16 pixels (73 used)
constexpr int kS2PlaneRows = 83;                       // 5 tiles x 16 rows + 3 halo rows
constexpr uint32_t kS2PlaneBytes = kS2PlaneRows * kS2Pitch;      // 13,280
constexpr int kS2EdgeRows = 132;                       // 128 tile rows + 3 halo (+1)
constexpr uint32_t kS2EdgeChunkBytes = kS2EdgeRows * 16;         // one 8-pixel chunk column: 2,112
constexpr uint32_t kS2EdgePlaneBytes = 2 * kS2EdgeChunkBytes;    // chunks 0,1
constexpr uint32_t kS2BufBytes = 2 * kS2PlaneBytes + 2 * kS2EdgePlaneBytes;  // one batch buffer: 35,008
constexpr uint32_t kS2WStepBytes = 256 * 32;           // one filter row: N=256 x K=16 bf16
constexpr uint32_t kS2WBytes = 8 * kS2WStepBytes;      // 7 filter rows + the bias step: 65,536
constexpr uint32_t kS2OnesBytes = 128 * 32;            // 128 x 16 bf16 ones: A operand of the bias step
constexpr int kS2RingRows = 32;
constexpr uint32_t kS2RingRowBytes = 16 * 128;         // 16 pooled columns x 64 ch bf16
constexpr uint32_t kS2RingBytes = kS2RingRows * kS2RingRowBytes;  // 65,536
constexpr uint32_t kS2EdgeHpBytes = 80 * 128;          // pooled column 16 of every stream row of the batch
// layout: [buf0][buf1][W][ones][ring][edge_hp][lut 512][bias 256][barriers 128]
constexpr uint32_t kS2OffW = 2 * kS2BufBytes;
constexpr uint32_t kS2OffOnes = kS2OffW + kS2WBytes;
constexpr uint32_t kS2OffRing = kS2OffOnes + kS2OnesBytes;
constexpr uint32_t kS2OffEdgeHp = kS2OffRing + kS2RingBytes;
constexpr uint32_t kS2OffLut = kS2OffEdgeHp + kS2EdgeHpBytes;
constexpr uint32_t kS2OffBias = kS2OffLut + 512;
constexpr uint32_t kS2OffBar = kS2OffBias + 256;
// MODE 1: raw bytes of a u8 source frame as 16-byte chunks from the 16-byte boundary below its first byte (4,489 bytes
// + up to 15 of misalignment = 282 chunks), double buffered
constexpr uint32_t kS2StageBytes = 288 * 16;
constexpr uint32_t kS2OffStage = kS2OffBar + 128;
constexpr uint32_t kS2Smem = 1024 + kS2OffStage + 2 * kS2StageBytes;
static_assert(kS2Smem <= 227 * 1024, "stem kernel: shared memory");

struct StemS2Params {
  // input mode 0: fp32 frames [n_frames][67*67] (already standardised)
  const float* frames;
  // input mode 1: u8 source frames [B][f_max][67*67] at the source rate; output frame n = b * t_max + k reads
  // source frame upsample_src_index(k, n_src[b]) when k < n_out[b], else the collate zero frame
  const uint8_t* src;
  const int32_t* n_src;
  const int32_t* n_out;
  int f_max, t_max, num, den;
  float mean, denom;  // standardisation (v - mean) / denom; denom = std + eps
  int standardise;
  int64_t n_frames;
  int64_t first;             // mode 1: global index of local frame 0 (chunked calls)
  const __nv_bfloat16* w1b;  // folded conv1 weights [64][64], k = fr*7 + fs
  int64_t src_bytes;         // mode 1: size of the whole `src` buffer (the 16-byte chunk loads never leave it)
  const float* bias;         // folded BN bias [64] (NULL = zero)
  __nv_bfloat16* out;        // [n_frames][17][17][64]
  double* stats;             // STATS kernels: per-channel sum [64] and sum of squares [64] of the conv1 outputs
};

// un-swizzled K-major descriptor: low word = addr>>4 | (LBO>>4)<<16, high word = SBO>>4 | version 1 | layout 0
__device__ __forceinline__ uint32_t desc_lo_ns(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
}
__device__ __forceinline__ uint32_t desc_hi_ns(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14); }

__device__ __forceinline__ void umma_f16_ns(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum)
      : "memory");
}

__device__ __forceinline__ int s2_src_index(int k, int n_src, int num, int den) {
  // same closed form as upsample.cu: max{i : floor(i*num/den + 1/2) <= k}, clamped to the last source frame
  long long v = ((long long)den * (2LL * k + 1) - 1) / (2LL * num);
  return v < n_src - 1 ? (int)v : n_src - 1;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

// STATS = true: training-mode first pass.  Same operand staging and MMAs, but the epilogue only accumulates the
// per-channel sum and sum of squares of the 34x34 conv1 outputs of every frame (batch statistics of the BatchNorm that
// follows conv1); nothing is pooled or written.  The second pass is the normal kernel with the batch scale folded
// into the weights and the batch shift as bias.
template <int MODE, bool STATS = false>
__global__ void __launch_bounds__(kS2Threads, 1) stem_s2d_kernel(const StemS2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  float* bias_s = reinterpret_cast<float*>(smem + kS2OffBias);
  unsigned short* lut = reinterpret_cast<unsigned short*>(smem + kS2OffLut);
  const uint32_t bar0 = base + kS2OffBar;
  // barriers: planes_full[2] | planes_empty[2] | acc_full[2] | acc_empty[2]
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kS2OffBar + 96);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int64_t n_batches = (p.n_frames + kS2FramesPerBatch - 1) / kS2FramesPerBatch;

  // ---- one-off setup: zero the batch buffers (borders / halo rows stay zero for the whole kernel), shifted weight
  //      copies, bias, u8 -> bf16 lookup table, barriers, TMEM
  for (uint32_t i = tid; i < (2 * kS2BufBytes) / 16; i += kS2Threads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int idx = tid; idx < 7 * 256 * 16; idx += kS2Threads) {
    const int e = idx & 15, n = (idx >> 4) & 255, fr = idx >> 12;
    const int s = n >> 6, ch = n & 63;
    const int fs = e - 2 * s;
    __nv_bfloat16 v = __float2bfloat16_rn(0.f);
    if (fs >= 0 && fs < 7) v = p.w1b[ch * 64 + fr * 7 + fs];
    // step fr: core matrices [n/8][e/8] of 8 rows x 16 bytes; LBO (K) = 128, SBO (N) = 256
    const uint32_t off = kS2OffW + (uint32_t)fr * kS2WStepBytes + (uint32_t)(n >> 3) * 256u + (uint32_t)(e >> 3) * 128u +
                         (uint32_t)(n & 7) * 16u + (uint32_t)(e & 7) * 2u;
    *reinterpret_cast<__nv_bfloat16*>(smem + off) = v;
  }
  // bias step: B[n][0] = bf16(bias), B[n][1] = bf16(bias - hi) (together ~16 mantissa bits), other columns zero
  for (int idx = tid; idx < 256 * 16; idx += kS2Threads) {
    const int e = idx & 15, n = idx >> 4;
    const float bv = p.bias ? p.bias[n & 63] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(bv);
    __nv_bfloat16 v = __float2bfloat16_rn(0.f);
    if (e == 0) v = hi;
    if (e == 1) v = __float2bfloat16_rn(bv - __bfloat162float(hi));
    const uint32_t off = kS2OffW + 7u * kS2WStepBytes + (uint32_t)(n >> 3) * 256u + (uint32_t)(e >> 3) * 128u +
                         (uint32_t)(n & 7) * 16u + (uint32_t)(e & 7) * 2u;
    *reinterpret_cast<__nv_bfloat16*>(smem + off) = v;
  }
  for (uint32_t i = tid; i < kS2OnesBytes / 4; i += kS2Threads)
    reinterpret_cast<uint32_t*>(smem + kS2OffOnes)[i] = 0x3F803F80u;  // bf16 1.0 pairs
  if (tid < 64) bias_s[tid] = p.bias ? p.bias[tid] : 0.f;
  if (MODE == 1 && tid < 256) {
    float v = (float)tid;
    if (p.standardise) v = (v - p.mean) / p.denom;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    lut[tid] = *reinterpret_cast<const unsigned short*>(&h);
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(BAR(0 + i), kS2Builders);  // planes full: every builder thread arrives after its proxy fence
      mbar_init(BAR(2 + i), 1);            // planes empty: tcgen05.commit
      mbar_init(BAR(4 + i), 1);            // accumulator full: tcgen05.commit
      mbar_init(BAR(6 + i), kS2EpiWarps);  // accumulator empty: one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == kS2BuilderWarps) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
    tmem_relinquish();
  }
  fence_proxy_async();  // weights / zeros were written through the generic proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp < kS2BuilderWarps) {
    // ======================= builders =======================
    // item j of a thread = (source row r, pixel pair q): padded pixels 2q, 2q+1 <-> source columns 2q-3, 2q-2.
    // The pixels of a frame are fetched into registers one frame ahead, so the global latency overlaps the wait for
    // a free batch buffer and the stores of the previous frame.
    // MODE 0: px = packed bf16 pairs, fetched one frame ahead.
    // MODE 1: the raw bytes of the next frame are fetched as two aligned 16-byte chunks per thread (ck0, ck1) and NOT
    // touched until that frame's turn, so the loads really are in flight across the stores of the current frame (the
    // first version combined 20 byte loads per thread into px[] inside fetch(): ncu had 17 % of the kernel's samples
    // on that line, waiting for L2).  At its turn the chunks go to a staging buffer, a builder-only barrier makes the
    // frame visible, and the items take their bytes from shared memory.
    uint32_t px[MODE == 0 ? kS2Items : 1];
    uint4 ck0 = make_uint4(0u, 0u, 0u, 0u), ck1 = ck0;
    bool live_next = false;
    uint32_t mis_next = 0;
    auto load_chunk = [&](const uint8_t* ptr) -> uint4 {
      if (ptr >= p.src && ptr + 16 <= p.src + p.src_bytes) return __ldg(reinterpret_cast<const uint4*>(ptr));
      uint32_t w[4] = {0u, 0u, 0u, 0u};  // first / last chunk of the whole buffer: byte by byte
      for (int i = 0; i < 16; ++i)
        if (ptr + i >= p.src && ptr + i < p.src + p.src_bytes) w[i >> 2] |= (uint32_t)__ldg(ptr + i) << (8 * (i & 3));
      return make_uint4(w[0], w[1], w[2], w[3]);
    };
    auto fetch = [&](int64_t n) {
      if (MODE == 1) live_next = false;
      if (n >= p.n_frames) return;
      if (MODE == 0) {
        const float* f32 = p.frames + n * (67 * 67);
#pragma unroll
        for (int j = 0; j < kS2Items; ++j) {
          const int item = tid + j * kS2Builders;
          const int r = item / 37, q = item - r * 37;
          const int c0 = 2 * q - 3, c1 = c0 + 1;
          const bool ok0 = (unsigned)c0 < 67u && r < 67, ok1 = (unsigned)c1 < 67u && r < 67;
          const float v0 = ok0 ? __ldg(f32 + r * 67 + c0) : 0.f;
          const float v1 = ok1 ? __ldg(f32 + r * 67 + c1) : 0.f;
          px[j] = pack_bf16x2(v0, v1);  // bf16(0) == 0 keeps the border pixels zero
        }
      } else {
        const int64_t ng = p.first + n;
        const int b = (int)(ng / p.t_max), k = (int)(ng - (int64_t)b * p.t_max);
        const int F = p.n_src[b], T = p.n_out[b];
        if (!(k < T && F > 0)) return;  // collate zero frame: source value 0 everywhere
        const uint8_t* u8 = p.src + ((int64_t)b * p.f_max + s2_src_index(k, F, p.num, p.den)) * (67 * 67);
        const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(u8) & 15u);
        const uint8_t* a0 = u8 - mis;
        live_next = true;
        mis_next = mis;
        ck0 = load_chunk(a0 + 16 * tid);
        if (16u * (uint32_t)(kS2Builders + tid) < 67u * 67u + mis) ck1 = load_chunk(a0 + 16 * (kS2Builders + tid));
      }
    };
    uint32_t stage_sel = 0;
    const uint8_t* sb = nullptr;  // MODE 1: first byte of the current frame in its staging buffer
    bool live = false;            // MODE 1: the current frame has source pixels (else: collate zero frame)
    auto store = [&]() {          // MODE 1: the chunks fetched for this frame -> staging buffer
      if (MODE == 1) {
        uint8_t* stage = smem + kS2OffStage + stage_sel * kS2StageBytes;
        stage_sel ^= 1u;
        live = live_next;
        if (live) {
          reinterpret_cast<uint4*>(stage)[tid] = ck0;
          if (kS2Builders + tid < (int)(kS2StageBytes / 16)) reinterpret_cast<uint4*>(stage)[kS2Builders + tid] = ck1;
        }
        sb = stage + mis_next;
      }
    };
    auto build = [&](uint8_t* bufp, int fi) {
#pragma unroll
      for (int j = 0; j < kS2Items; ++j) {
        const int item = tid + j * kS2Builders;
        const int r = item / 37, q = item - r * 37;
        if (r >= 67) continue;
        uint32_t w;
        if (MODE == 0) {
          w = px[j];
        } else {
          const int c0 = 2 * q - 3, c1 = c0 + 1;
          const uint32_t lo = ((unsigned)c0 < 67u) ? (uint32_t)lut[live ? sb[r * 67 + c0] : 0] : 0u;
          const uint32_t hi = ((unsigned)c1 < 67u) ? (uint32_t)lut[live ? sb[r * 67 + c1] : 0] : 0u;
          w = lo | (hi << 16);
        }
        const int pr = r + 3;  // padded row
        const int yy = fi * kS2RowsPerFrame + (pr >> 1);
        uint8_t* plane = bufp + (uint32_t)(pr & 1) * kS2PlaneBytes;
        *reinterpret_cast<uint32_t*>(plane + yy * kS2Pitch + 4 * q) = w;
        if (q >= 30) {  // pixels 60..73: the edge anchor's two chunks, transposed (rows 16 bytes apart)
          uint8_t* edge = bufp + 2 * kS2PlaneBytes + (uint32_t)(pr & 1) * kS2EdgePlaneBytes;
          *reinterpret_cast<uint32_t*>(edge + (uint32_t)((q - 30) >> 2) * kS2EdgeChunkBytes + yy * 16 + 4 * ((q - 30) & 3)) = w;
        }
      }
    };
    uint32_t it = 0;
    fetch((int64_t)blockIdx.x * kS2FramesPerBatch);
    for (int64_t bi = blockIdx.x; bi < n_batches; bi += gridDim.x, ++it) {
      const uint32_t buf = it & 1u;
      mbar_wait(BAR(2 + buf), ((it >> 1) & 1u) ^ 1u);
      uint8_t* bufp = smem + buf * kS2BufBytes;
#pragma unroll
      for (int fi = 0; fi < kS2FramesPerBatch; ++fi) {
        const int64_t n = bi * kS2FramesPerBatch + fi;
        const bool have = n < p.n_frames;  // CTA-uniform
        if (have) store();
        if (MODE == 0 && have) build(bufp, fi);  // px holds THIS frame until the fetch below overwrites it
        // next frame of this CTA: the second of this batch, or the first of the next one
        const int64_t nn = (fi + 1 < kS2FramesPerBatch) ? n + 1 : (bi + gridDim.x) * kS2FramesPerBatch;
        if (fi + 1 < kS2FramesPerBatch || bi + gridDim.x < n_batches) fetch(nn);
        if (MODE == 1 && have) {
          // staging buffer complete; it is rewritten two frames from now, behind the next frame's barrier
          named_bar_sync(2, kS2Builders);
          build(bufp, fi);
        }
      }
      fence_proxy_async();
      mbar_arrive(BAR(0 + buf));
    }
  } else if (warp == kS2BuilderWarps) {
    // ======================= MMA issuer (warp-converged; one elected lane issues) =======================
    constexpr uint32_t idesc = make_idesc(256);
    const uint32_t a_hi_main = desc_hi_ns(kS2Pitch), a_hi_edge = desc_hi_ns(128), b_hi = desc_hi_ns(256);
    uint32_t it = 0, g = 0;  // batch / tile counters of this CTA
    for (int64_t bi = blockIdx.x; bi < n_batches; bi += gridDim.x, ++it) {
      const uint32_t buf = it & 1u;
      const int nf = (int)((p.n_frames - bi * kS2FramesPerBatch) < kS2FramesPerBatch
                               ? (p.n_frames - bi * kS2FramesPerBatch)
                               : kS2FramesPerBatch);
      const int n_main = (nf * kS2RowsPerFrame + 15) / 16;
      mbar_wait(BAR(0 + buf), (it >> 1) & 1u);
      tc_fence_after();
      const uint32_t bufa = base + buf * kS2BufBytes;
      for (int t = -1; t < n_main; ++t, ++g) {  // t = -1: edge tile
        const uint32_t acc = g & 1u;
        mbar_wait(BAR(6 + acc), ((g >> 1) & 1u) ^ 1u);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t d = tmem_acc + acc * 256u;
          // bias step: ones (128 x 16, core matrices [r/8][k/8]: SBO 256, LBO 128) x bias rows
          umma_f16_ns(d, desc_lo_ns(base + kS2OffOnes, 128u), b_hi, desc_lo_ns(base + kS2OffW + 7u * kS2WStepBytes, 128u),
                      b_hi, idesc, 0);
#pragma unroll
          for (int fr = 0; fr < 7; ++fr) {
            const int a = fr >> 1, b = fr & 1;
            uint32_t a_lo, a_hi;
            if (t < 0) {
              a_lo = desc_lo_ns(bufa + 2 * kS2PlaneBytes + (uint32_t)b * kS2EdgePlaneBytes + (uint32_t)a * 16u,
                                kS2EdgeChunkBytes);
              a_hi = a_hi_edge;
            } else {
              a_lo = desc_lo_ns(bufa + (uint32_t)b * kS2PlaneBytes + (uint32_t)(16 * t + a) * kS2Pitch, 16u);
              a_hi = a_hi_main;
            }
            const uint32_t b_lo = desc_lo_ns(base + kS2OffW + (uint32_t)fr * kS2WStepBytes, 128u);
            umma_f16_ns(d, a_lo, a_hi, b_lo, b_hi, idesc, 1);
          }
          umma_commit(BAR(4 + acc));
          if (t == n_main - 1) umma_commit(BAR(2 + buf));  // all MMAs reading this batch buffer have completed
        }
        __syncwarp();
      }
    }
  } else {
    // ======================= epilogue: the last 8 warps =======================
    const int ew = warp - (kS2BuilderWarps + 1);  // 0..kS2EpiWarps-1
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = ew >> 2;          // channel group: [kS2ChPerWarp*half, kS2ChPerWarp*(half+1))
    const int etid = tid - (kS2Builders + 32);  // 0..kS2EpiThreads-1
    const int L = q * 32 + lane;       // accumulator row
    uint8_t* ring = smem + kS2OffRing;
    uint8_t* edge_hp = smem + kS2OffEdgeHp;
    uint32_t g = 0;
    if (STATS) {
      // per-thread running sums of this thread's 32 channels over all its rows and tiles; reduced once at the end
      float sa[kS2ChPerWarp], sq[kS2ChPerWarp];
#pragma unroll
      for (int c = 0; c < kS2ChPerWarp; ++c) sa[c] = sq[c] = 0.f;
      for (int64_t bi = blockIdx.x; bi < n_batches; bi += gridDim.x) {
        const int nf = (int)((p.n_frames - bi * kS2FramesPerBatch) < kS2FramesPerBatch
                                 ? (p.n_frames - bi * kS2FramesPerBatch)
                                 : kS2FramesPerBatch);
        const int n_rows = nf * kS2RowsPerFrame;
        const int n_main = (n_rows + 15) / 16;
        for (int t = -1; t < n_main; ++t, ++g) {
          const uint32_t acc = g & 1u;
          mbar_wait(BAR(4 + acc), (g >> 1) & 1u);
          tc_fence_after();
          const uint32_t t_row = tmem_acc + acc * 256u + ((uint32_t)(q * 32) << 16);
          // stream row of this accumulator row and the shifts that are distinct conv outputs
          const int yy = (t < 0) ? L : 16 * t + (L >> 3);
          const int y = yy % kS2RowsPerFrame;
          const bool live = (yy < n_rows) && (y < 34) && (t >= 0 || L < 80);
          const int s_first = (t < 0) ? 2 : 0;  // edge tile: conv columns 30,31 (s = 0,1) belong to the main tiles
#pragma unroll 1
          for (int pass = 0; pass < kS2Passes; ++pass) {
            const int ch0 = half * kS2ChPerWarp + pass * 16;
#pragma unroll
            for (int sft = 0; sft < 4; ++sft) {
              uint32_t v[16];
              tmem_ld16(t_row + (uint32_t)sft * 64u + ch0, v);
              tmem_ld_wait();
              if (live && sft >= s_first) {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                  const float x = __uint_as_float(v[c]);
                  if (kS2Passes == 1 || pass == 0) {
                    sa[c] += x;
                    sq[c] = fmaf(x, x, sq[c]);
                  } else {
                    sa[(16 + c) % kS2ChPerWarp] += x;
                    sq[(16 + c) % kS2ChPerWarp] = fmaf(x, x, sq[(16 + c) % kS2ChPerWarp]);
                  }
                }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(6 + acc));
        }
      }
      // warp reduction over the 32 rows of the quarter, then one fp64 atomic per channel and warp
#pragma unroll
      for (int c = 0; c < kS2ChPerWarp; ++c) {
        float a = sa[c], b = sq[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (lane == 0) {
          atomicAdd(p.stats + half * kS2ChPerWarp + c, (double)a);
          atomicAdd(p.stats + 64 + half * kS2ChPerWarp + c, (double)b);
        }
      }
    } else
    for (int64_t bi = blockIdx.x; bi < n_batches; bi += gridDim.x) {
      const int nf = (int)((p.n_frames - bi * kS2FramesPerBatch) < kS2FramesPerBatch
                               ? (p.n_frames - bi * kS2FramesPerBatch)
                               : kS2FramesPerBatch);
      const int n_rows = nf * kS2RowsPerFrame;
      const int n_main = (n_rows + 15) / 16;
      for (int t = -1; t < n_main; ++t, ++g) {
        const uint32_t acc = g & 1u;
        mbar_wait(BAR(4 + acc), (g >> 1) & 1u);
        tc_fence_after();
        const uint32_t t_row = tmem_acc + acc * 256u + ((uint32_t)(q * 32) << 16);
        if (t < 0) {
          // edge tile: row L = stream row yy; shifts s = 1,2,3 are conv columns 31,32,33 -> pooled column 16
#pragma unroll 1
          for (int pass = 0; pass < kS2Passes; ++pass) {
            const int ch0 = half * kS2ChPerWarp + pass * 16;
            uint32_t v1[16], v2[16], v3[16];
            tmem_ld16(t_row + 64u + ch0, v1);
            tmem_ld16(t_row + 128u + ch0, v2);
            tmem_ld16(t_row + 192u + ch0, v3);
            tmem_ld_wait();
            if (L < 80) {
              uint32_t o[8];
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float x0 = fmaxf(fmaxf(__uint_as_float(v1[2 * c]), __uint_as_float(v2[2 * c])), __uint_as_float(v3[2 * c]));
                const float x1 = fmaxf(fmaxf(__uint_as_float(v1[2 * c + 1]), __uint_as_float(v2[2 * c + 1])),
                                       __uint_as_float(v3[2 * c + 1]));
                o[c] = pack_relu_bf16x2(x0, x1);
              }
              uint8_t* rowp = edge_hp + L * 128;
              const int k0 = ch0 >> 3;
              *reinterpret_cast<uint4*>(rowp + (((k0) ^ (L & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
              *reinterpret_cast<uint4*>(rowp + (((k0 + 1) ^ (L & 7)) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(6 + acc));
          named_bar_sync(1, kS2EpiThreads);  // edge_hp complete before any vertical pooling of this batch
          continue;
        }
        // main tile: row L = anchor (stream row yy = 16t + L/8, i = L%8): conv columns 4i+s, pooled 2i and 2i+1
        const int yy = 16 * t + (L >> 3);
        const int i = L & 7;
        uint8_t* rrow = ring + (uint32_t)(yy & (kS2RingRows - 1)) * kS2RingRowBytes;
#pragma unroll 1
        for (int pass = 0; pass < kS2Passes; ++pass) {
          const int ch0 = half * kS2ChPerWarp + pass * 16;
          uint32_t v0[16], v1[16], v2[16], v3[16];
          tmem_ld16(t_row + ch0, v0);
          tmem_ld16(t_row + 64u + ch0, v1);
          tmem_ld16(t_row + 128u + ch0, v2);
          tmem_ld16(t_row + 192u + ch0, v3);
          tmem_ld_wait();
          uint32_t oa[8], ob[8];
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            float pa[2], pb[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float s0 = __uint_as_float(v0[c + h]), s1 = __uint_as_float(v1[c + h]);
              const float s2 = __uint_as_float(v2[c + h]), s3 = __uint_as_float(v3[c + h]);
              float left = __shfl_up_sync(0xffffffffu, s3, 1);  // conv column 4i-1 (lane L-1 is anchor i-1 of the same row)
              left = (i == 0) ? s0 : left;
              pa[h] = fmaxf(fmaxf(left, s0), s1);  // accumulators already hold conv + bias; max commutes with ReLU / rounding
              pb[h] = fmaxf(fmaxf(s1, s2), s3);
            }
            oa[c >> 1] = pack_relu_bf16x2(pa[0], pa[1]);
            ob[c >> 1] = pack_relu_bf16x2(pb[0], pb[1]);
          }
          if (yy < n_rows) {
            const int k0 = ch0 >> 3;  // 16-byte chunk index of these channels within the 128-byte pixel
            uint8_t* pxa = rrow + (2 * i) * 128;
            uint8_t* pxb = pxa + 128;
            *reinterpret_cast<uint4*>(pxa + (((k0) ^ i) << 4)) = make_uint4(oa[0], oa[1], oa[2], oa[3]);
            *reinterpret_cast<uint4*>(pxa + (((k0 + 1) ^ i) << 4)) = make_uint4(oa[4], oa[5], oa[6], oa[7]);
            *reinterpret_cast<uint4*>(pxb + (((k0) ^ i) << 4)) = make_uint4(ob[0], ob[1], ob[2], ob[3]);
            *reinterpret_cast<uint4*>(pxb + (((k0 + 1) ^ i) << 4)) = make_uint4(ob[4], ob[5], ob[6], ob[7]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(6 + acc));
        named_bar_sync(1, kS2EpiThreads);  // this tile's ring rows are visible to all epilogue warps

        // vertical pool: every odd conv row y = 2ph+1 inside this tile completes pooled row ph (rows y-2, y-1, y).
        // thread = (16-byte channel chunk ch, row pair rr of the tile, column group pg); columns pg, pg+4, ...
        {
          const int ch = etid & 7, rr = (etid >> 3) & 7, pg = etid >> 6;  // pg in [0, kS2EpiThreads / 64)
          // candidate stream rows 16t + 2rr + {0,1}: the one whose in-frame row y is odd is the trigger
          int ys = 16 * t + 2 * rr;
          int fi = (ys >= kS2RowsPerFrame) ? 1 : 0;
          int y = ys - fi * kS2RowsPerFrame;
          if ((y & 1) == 0) {
            ++ys; ++y;
            if (y == kS2RowsPerFrame) { y = 0; ++fi; }
          }
          if ((y & 1) && y <= 33 && ys < n_rows && fi < nf) {
            const int ph = y >> 1;
            const uint8_t* r0 = ring + (uint32_t)(ys & (kS2RingRows - 1)) * kS2RingRowBytes;
            const uint8_t* r1 = ring + (uint32_t)((ys - 1) & (kS2RingRows - 1)) * kS2RingRowBytes;
            const uint8_t* r2 = ring + (uint32_t)((ys - 2) & (kS2RingRows - 1)) * kS2RingRowBytes;
            const bool top = (y >= 2);  // pooled row 0 has no conv row -1
            __nv_bfloat16* orow = p.out + ((bi * kS2FramesPerBatch + fi) * 17 + ph) * (17 * 64) + ch * 8;
#pragma unroll
            for (int pw = pg; pw < 17; pw += kS2EpiThreads / 64) {
              uint4 a, b, c;
              if (pw < 16) {
                const uint32_t off = (uint32_t)pw * 128u + (uint32_t)((ch ^ (pw >> 1)) << 4);
                a = *reinterpret_cast<const uint4*>(r0 + off);
                b = *reinterpret_cast<const uint4*>(r1 + off);
                c = top ? *reinterpret_cast<const uint4*>(r2 + off) : b;
              } else {
                a = *reinterpret_cast<const uint4*>(edge_hp + ys * 128 + ((ch ^ (ys & 7)) << 4));
                b = *reinterpret_cast<const uint4*>(edge_hp + (ys - 1) * 128 + ((ch ^ ((ys - 1) & 7)) << 4));
                c = top ? *reinterpret_cast<const uint4*>(edge_hp + (ys - 2) * 128 + ((ch ^ ((ys - 2) & 7)) << 4)) : b;
              }
              uint4 m;
              m.x = bf16x2_max(bf16x2_max(a.x, b.x), c.x);
              m.y = bf16x2_max(bf16x2_max(a.y, b.y), c.y);
              m.z = bf16x2_max(bf16x2_max(a.z, b.z), c.z);
              m.w = bf16x2_max(bf16x2_max(a.w, b.w), c.w);
              *reinterpret_cast<uint4*>(orow + pw * 64) = m;
            }
          }
        }
        named_bar_sync(1, kS2EpiThreads);  // ring rows of tile t-1 may be overwritten by tile t+1 from here on
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kS2BuilderWarps) {
    __syncwarp();
    tmem_dealloc(tmem_acc, 512);
  }
}

}  // namespace tc
}  // namespace avvad
