// Fused training-loop loss and metrics over padded batches.
//
// Reference semantics:
//   loss    : scripts/train_AV_net.py:298-301 -- SUM over utterances of packages/models/utils.py:113
//             binary_cross_entropy(pred[:len_b], target[:len_b], eps) = -mean(x log(sigmoid(r)+eps) + (1-x) log(1-sigmoid(r)+eps))
//             (the reference runs B Python iterations of ~8 small kernels each)
//   metrics : packages/models/utils.py:164-203 f1_loss per utterance on (sigmoid(r) > 0.5) vs target,
//             scripts/train_AV_net.py:311-329
// One CTA per utterance, fixed-order reductions (deterministic).  Optionally emits d(loss)/d(logits).
#include "common.cuh"

namespace avvad {

__device__ __forceinline__ float block_sum_256(float v, float* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w];
  }
  __syncthreads();
  return t;  // valid on thread 0
}

__global__ void __launch_bounds__(256)
bce_kernel(const float* __restrict__ logits, const float* __restrict__ target, const int32_t* __restrict__ lengths,
           int T, int Y, float eps, float* __restrict__ per_utt, float* __restrict__ dlogits) {
  __shared__ float sm[8];
  const int b = blockIdx.x;
  int len = lengths[b];
  len = len < 0 ? 0 : (len > T ? T : len);
  const int64_t base = (int64_t)b * T * Y;
  const int n = len * Y;
  const float inv = n > 0 ? 1.0f / (float)n : 0.f;
  float acc = 0.f;
  for (int i = threadIdx.x; i < T * Y; i += 256) {
    float g = 0.f;
    if (i < n) {
      const float r = logits[base + i];
      const float x = (float)(long long)target[base + i];  // the reference casts the target .long() before use
      const float s = 1.0f / (1.0f + expf(-r));
      acc += x * logf(s + eps) + (1.0f - x) * logf(1.0f - s + eps);
      const float ds = s * (1.0f - s);
      g = -(x * ds / (s + eps) - (1.0f - x) * ds / (1.0f - s + eps)) * inv;
    }
    if (dlogits) dlogits[base + i] = g;  // zero for padded steps
  }
  const float tot = block_sum_256(acc, sm);
  if (threadIdx.x == 0) per_utt[b] = n > 0 ? -tot * inv : 0.f;
}

__global__ void sum_utt_kernel(const float* __restrict__ per_utt, int B, float* __restrict__ loss) {
  // single thread, utterance order: the same order the reference's Python loop adds them in
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float t = 0.f;
    for (int b = 0; b < B; ++b) t += per_utt[b];
    loss[0] = t;
  }
}

// metrics[b] = (accuracy, precision, recall, f1) of utterance b over its first len_b frames (y_dim == 1)
__global__ void __launch_bounds__(256)
f1_kernel(const float* __restrict__ logits, const float* __restrict__ target, const int32_t* __restrict__ lengths,
          int T, float epsilon, float* __restrict__ metrics, int32_t* __restrict__ dec_out) {
  __shared__ float sm[8];
  const int b = blockIdx.x;
  int len = lengths[b];
  len = len < 0 ? 0 : (len > T ? T : len);
  float tp = 0.f, tn = 0.f, fp = 0.f, fn = 0.f;
  for (int t = threadIdx.x; t < T; t += 256) {
    const float r = logits[(int64_t)b * T + t];
    const int d = (1.0f / (1.0f + expf(-r))) > 0.5f ? 1 : 0;
    if (dec_out) dec_out[(int64_t)b * T + t] = d;
    if (t < len) {
      const int y = (int)(long long)target[(int64_t)b * T + t];
      tp += (float)(y * d);
      tn += (float)((1 - y) * (1 - d));
      fp += (float)((1 - y) * d);
      fn += (float)(y * (1 - d));
    }
  }
  const float s_tp = block_sum_256(tp, sm);
  const float s_tn = block_sum_256(tn, sm);
  const float s_fp = block_sum_256(fp, sm);
  const float s_fn = block_sum_256(fn, sm);
  if (threadIdx.x == 0) {
    const float acc = (s_tp + s_tn) / (s_tp + s_tn + s_fp + s_fn + epsilon);
    const float prec = s_tp / (s_tp + s_fp + epsilon);
    const float rec = s_tp / (s_tp + s_fn + epsilon);
    const float f1 = 2.f * (prec * rec) / (prec + rec + epsilon);
    metrics[b * 4 + 0] = acc;
    metrics[b * 4 + 1] = prec;
    metrics[b * 4 + 2] = rec;
    metrics[b * 4 + 3] = f1;
  }
}

}  // namespace avvad

using namespace avvad;

extern "C" int avvad_bce_loss(const float* logits, const float* target, const int32_t* lengths, int32_t B, int32_t T,
                              int32_t y_dim, float eps, float* loss, float* per_utt, float* dlogits, void* stream) {
  AVVAD_CHECK_ARG(logits && target && lengths && loss && per_utt, "null pointer");
  AVVAD_CHECK_ARG(B > 0 && T > 0 && y_dim > 0, "bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  bce_kernel<<<B, 256, 0, st>>>(logits, target, lengths, T, y_dim, eps, per_utt, dlogits);
  AVVAD_LAUNCHED();
  sum_utt_kernel<<<1, 32, 0, st>>>(per_utt, B, loss);
  AVVAD_LAUNCHED();
  return AVVAD_OK;
}

extern "C" int avvad_f1_metrics(const float* logits, const float* target, const int32_t* lengths, int32_t B, int32_t T,
                                float epsilon, float* metrics, int32_t* dec, void* stream) {
  AVVAD_CHECK_ARG(logits && target && lengths && metrics, "null pointer");
  AVVAD_CHECK_ARG(B > 0 && T > 0, "bad sizes");
  f1_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(logits, target, lengths, T, epsilon, metrics, dec);
  AVVAD_LAUNCHED();
  return AVVAD_OK;
}
