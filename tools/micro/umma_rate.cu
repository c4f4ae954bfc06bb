// Micro-benchmark: cycles per tcgen05.mma (SS mode, bf16, M=128) as a function of N (and M = 64 / 128), operands resident in smem.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I audio-visual-vad_b200/csrc tools/micro/umma_rate.cu -o gpurun_out/umma_rate
#include <cstdio>
#include "gemm_tc.cuh"
using namespace avvad::tc;

template <int BN, int M = 128>
__global__ void k(long long* out, int iters, int a_shift_rows) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  // A: 256 rows x 128 B, B: BN rows x 128 B, zero filled
  for (int i = threadIdx.x; i < (256 + BN) * 128 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  const uint32_t bar = base + (256 + BN) * 128;
  volatile uint32_t* slot = (volatile uint32_t*)(smem + (256 + BN) * 128 + 16);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32((void*)slot), 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = (make_idesc(BN) & ~(0x1Fu << 24)) | ((uint32_t)(M >> 4) << 24);
    const uint32_t sa = base + a_shift_rows * 128, sb = base + 256 * 128;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_f16(tm + (i & 1) * BN, make_sw128_desc(sa + kk * 32), make_sw128_desc(sb + kk * 32), idesc, 1);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int BN, int M = 128> void run(int grid, int shift) {
  long long* d; cudaMalloc(&d, sizeof(long long) * grid);
  size_t smem = (256 + BN) * 128 + 1024 + 64;
  cudaFuncSetAttribute(k<BN, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 2000;
  k<BN, M><<<grid, 128, smem>>>(d, iters, shift);
  k<BN, M><<<grid, 128, smem>>>(d, iters, shift);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[1024]; cudaMemcpy(h, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
  printf("M=%3d N=%3d grid=%3d shift=%2d: %.1f cycles per MxNx16 MMA (floor %d)  err=%s\n", M, BN, grid, shift, avg / (iters * 4.0), BN / 2, cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  for (int grid : {1, 148}) {
    run<64>(grid, 0); run<64>(grid, 19); run<64, 64>(grid, 0); run<64, 64>(grid, 18); run<128, 64>(grid, 0); run<128>(grid, 0); run<256>(grid, 0);
  }
  return 0;
}
