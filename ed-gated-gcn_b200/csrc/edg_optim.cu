// adam_multi : one Adam step (torch.optim.Adam semantics, no amsgrad, optional L2 weight decay) over up to 32 parameter
// tensors in ONE launch.  The reference trains with torch.optim.Adam (train.py:239-243); its fused multi-tensor kernel
// needs 39 us for the 0.58 M parameters of the gated block (about twenty small tensors: a couple of dozen blocks, each
// walking 64 K elements) -- on the critical path of a 0.9 ms step.  Here a block owns 1024 elements of one tensor
// (pointers and chunk prefix sums travel in the kernel parameters), so the step is one pass at memory speed.
// The step counters (one per tensor, as in torch: a tensor without gradient does not advance) live on the device -- the
// whole training step is replayed as a CUDA graph -- and adam_tick increments them.
#include <math.h>

#include "edg_common.cuh"

namespace edg {

constexpr int kAdamMax = 32;
constexpr int kAdamChunk = 1024;

struct AdamBatch {
  float* p[kAdamMax];
  const float* g[kAdamMax];
  float* m[kAdamMax];
  float* v[kAdamMax];
  float* step[kAdamMax];
  int64_t numel[kAdamMax];
  int chunk_start[kAdamMax + 1];
  int n;
};

__global__ void adam_tick_kernel(const __grid_constant__ AdamBatch b) {
  if ((int)threadIdx.x < b.n) *b.step[threadIdx.x] += 1.0f;
}

__global__ void __launch_bounds__(256)
adam_multi_kernel(const __grid_constant__ AdamBatch b, float lr, float beta1, float beta2,
                  float eps, float weight_decay, float grad_scale) {
  int t = 0;
  while (t + 1 < b.n && (int)blockIdx.x >= b.chunk_start[t + 1]) ++t;
  const int64_t base = (int64_t)((int)blockIdx.x - b.chunk_start[t]) * kAdamChunk;
  const float s = *b.step[t];
  const float bc1 = 1.0f - powf(beta1, s), bc2 = 1.0f - powf(beta2, s);
  const float step_size = lr / bc1, inv_sqrt_bc2 = 1.0f / sqrtf(bc2);
  float* __restrict__ p = b.p[t];
  const float* __restrict__ g = b.g[t];
  float* __restrict__ m = b.m[t];
  float* __restrict__ v = b.v[t];
  const int64_t n = b.numel[t];
#pragma unroll
  for (int k = 0; k < kAdamChunk / 256; ++k) {
    const int64_t i = base + threadIdx.x + 256 * k;
    if (i < n) {
      float gi = g[i] * grad_scale;
      const float pi = p[i];
      if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
      const float mi = fmaf(beta1, m[i], (1.0f - beta1) * gi);        // exp_avg.lerp_(grad, 1 - beta1)
      const float vi = fmaf(beta2, v[i], (1.0f - beta2) * gi * gi);   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      m[i] = mi;
      v[i] = vi;
      p[i] = pi - step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
    }
  }
}

}  // namespace edg

using namespace edg;

/* see include/edgcn.h */
extern "C" int edg_adam_multi(int32_t n, void* const* param, const void* const* grad, void* const* exp_avg,
                              void* const* exp_avg_sq, void* const* step, const int64_t* numel, float lr, float beta1,
                              float beta2, float eps, float weight_decay, float grad_scale, edg_stream stream) {
  if (n < 0 || n > kAdamMax) return EDG_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) return EDG_OK;
  AdamBatch b;
  memset(&b, 0, sizeof(b));
  b.n = n;
  int chunks = 0;
  for (int i = 0; i < n; ++i) {
    if (!param[i] || !grad[i] || !exp_avg[i] || !exp_avg_sq[i] || !step[i] || numel[i] < 0) return EDG_ERR_ARG;
    b.step[i] = (float*)step[i];
    b.p[i] = (float*)param[i]; b.g[i] = (const float*)grad[i]; b.m[i] = (float*)exp_avg[i]; b.v[i] = (float*)exp_avg_sq[i];
    b.numel[i] = numel[i];
    b.chunk_start[i] = chunks;
    const int64_t c = (numel[i] + kAdamChunk - 1) / kAdamChunk;
    if (c > (1 << 24)) return EDG_ERR_UNSUPPORTED;
    chunks += (int)c;
  }
  b.chunk_start[n] = chunks;
  adam_tick_kernel<<<1, kAdamMax, 0, s>>>(b);
  int rc = check_launch();
  if (rc) return rc;
  if (chunks == 0) return EDG_OK;
  adam_multi_kernel<<<chunks, 256, 0, s>>>(b, lr, beta1, beta2, eps, weight_decay, grad_scale);
  return check_launch();
}
