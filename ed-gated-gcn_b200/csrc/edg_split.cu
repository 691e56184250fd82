// fp32-parity GEMMs on the tensor cores: split operands.
//
// The reference computes the projection  text @ W  (models/gcn.py:34) and what autograd derives from it in fp32.
// The 5th-gen tensor cores have no fp32 input type, so an fp32 matrix x is carried as TWO fp16 matrices
//     hi = fp16(x s),  lo = fp16(x s - hi),        s = 2^k chosen from max|x| so that max|x s| lies in [2^13, 2^14)
// stored side by side in one row ([ hi | lo ], lo at column ld/2, both zero-padded to a multiple of 64 columns).
// hi + lo reproduces x s to 22 significant bits (absolute floor 2^-25, i.e. 2^-39 max|x|), and
//     a b = (a_hi b_hi + a_hi b_lo + a_lo b_hi) / (s_a s_b)
// drops only a_lo b_lo (<= 2^-22 |a||b|): three kind::f16 MMAs per k step with fp32 accumulation in TMEM
// (linear_ws_kernel<float, true>, launch_wgrad_split in edg_gemm_tc.cu).  The scales are powers of two and live on
// the device (`amax`, one float per split matrix), so nothing here synchronises with the host.
//
//   edg_split_f16     fp32 rows -> [hi | lo] fp16 rows + amax            (amax_kernel, split_f16_kernel)
//   edg_linear_split  C = act(A W^T + bias) from split A [M,K], W [Nout,K]
//   edg_wgrad_split   dW = A^T B (+ column sums) from split A [R,K1], B [R,K2]
#include <cuda_fp16.h>

#include "edg_common.cuh"

namespace edg {

// max |x| over a [rows, cols] fp32 matrix (pitch ld), combined with atomicMax on the bit pattern (non-negative floats
// order like unsigned integers); *amax must be zero on entry
__global__ void __launch_bounds__(256)
amax_kernel(const float* __restrict__ x, int64_t ld, int rows, int cols, float* __restrict__ amax) {
  const uint32_t c4 = (uint32_t)(cols + 3) >> 2;
  const uint32_t total = (uint32_t)rows * c4;                    // the launcher keeps rows * ceil(cols / 4) below 2^32
  const uint32_t stride = gridDim.x * blockDim.x;
  float m = 0.f;
  auto take = [&](const float4& v, uint32_t g) {
    const int c = 4 * (int)g;
    m = fmaxf(m, fabsf(v.x));
    if (c + 1 < cols) m = fmaxf(m, fabsf(v.y));
    if (c + 2 < cols) m = fmaxf(m, fabsf(v.z));
    if (c + 3 < cols) m = fmaxf(m, fabsf(v.w));
  };
  auto at = [&](uint32_t i, uint32_t& g) {                        // ld % 4 == 0: the 16-byte read stays inside the row
    const uint32_t r = i / c4;
    g = i - r * c4;
    return reinterpret_cast<const float4*>(x + (int64_t)r * ld + 4 * g);
  };
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < total && total - i > 3 * stride; i += 4 * stride) {  // four independent 16-byte loads in flight
    uint32_t g0, g1, g2, g3;
    const float4 v0 = __ldg(at(i, g0)), v1 = __ldg(at(i + stride, g1)), v2 = __ldg(at(i + 2 * stride, g2)),
                 v3 = __ldg(at(i + 3 * stride, g3));
    take(v0, g0); take(v1, g1); take(v2, g2); take(v3, g3);
  }
  for (; i < total; i += stride) {
    uint32_t g;
    const float4 v = __ldg(at(i, g));
    take(v, g);
  }
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float wm[8];
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmaxf(m, wm[w]);
    atomicMax(reinterpret_cast<unsigned int*>(amax), __float_as_uint(m));
  }
}

// one thread = 8 consecutive columns of one row: 2 x 16-byte loads, one 16-byte store into each half
__global__ void __launch_bounds__(256)
split_f16_kernel(const float* __restrict__ x, int64_t ld, int rows, int cols, const float* __restrict__ amax,
                 __half* __restrict__ out, int64_t ldo, int kp) {
  const uint32_t g8 = (uint32_t)kp >> 3;
  const uint32_t total = (uint32_t)rows * g8;                    // < 2^32 (launcher)
  const float s = split_scale(__ldg(amax));
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t rr = i / g8;
    const int r = (int)rr, c0 = 8 * (int)(i - rr * g8);
    float v[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 + 4 * h < cols) t = *reinterpret_cast<const float4*>(x + (int64_t)r * ld + c0 + 4 * h);
      v[4 * h] = t.x; v[4 * h + 1] = t.y; v[4 * h + 2] = t.z; v[4 * h + 3] = t.w;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float a = (c0 + 2 * p < cols) ? v[2 * p] * s : 0.f;
      float b = (c0 + 2 * p + 1 < cols) ? v[2 * p + 1] * s : 0.f;
      const __half ha = __float2half_rn(a), hb = __float2half_rn(b);
      const __half la = __float2half_rn(a - __half2float(ha)), lb = __float2half_rn(b - __half2float(hb));
      hi[p] = (uint32_t)__half_as_ushort(ha) | ((uint32_t)__half_as_ushort(hb) << 16);
      lo[p] = (uint32_t)__half_as_ushort(la) | ((uint32_t)__half_as_ushort(lb) << 16);
    }
    __half* o = out + (int64_t)r * ldo + c0;
    *reinterpret_cast<uint4*>(o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(o + kp) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// sum of the partial-sum slabs of the three split products, scaled back: dW by 1/(s_a s_b), the bias row / column by
// the inverse scale of the operand it sums
__global__ void __launch_bounds__(256)
split_reduce_bias_scaled_kernel(const float* __restrict__ partial, int slabs, int K1, int K2, int K1e, int K2e,
                                float* __restrict__ dW, int64_t lddw, float* __restrict__ dbias, int bias_of,
                                const float* __restrict__ amax_a, const float* __restrict__ amax_b) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t tot = (int64_t)K1e * K2e;
  if (idx >= tot) return;
  const int m = (int)(idx / K2e), n = (int)(idx - (int64_t)m * K2e);
  const float s = sum_slabs(partial + idx, slabs, tot);
  const float ia = split_inv_scale(__ldg(amax_a)), ib = split_inv_scale(__ldg(amax_b));
  if (m < K1 && n < K2) dW[(int64_t)m * lddw + n] = s * ia * ib;
  else if (bias_of == 2 && m == K1 && n < K2) dbias[n] = s * ib;
  else if (bias_of == 1 && n == K2 && m < K1) dbias[m] = s * ia;
}

void launch_split_reduce_bias_scaled(const float* partial, int slabs, int K1, int K2, int K1e, int K2e, float* dW,
                                     int64_t lddw, float* dbias, int bias_of, const float* amax_a, const float* amax_b,
                                     cudaStream_t s) {
  const int64_t tot = (int64_t)K1e * K2e;
  split_reduce_bias_scaled_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(partial, slabs, K1, K2, K1e, K2e, dW, lddw,
                                                                                dbias, bias_of, amax_a, amax_b);
}

}  // namespace edg

using namespace edg;

extern "C" int edg_set_sm_budget(int32_t n) { return set_sm_budget(n); }

extern "C" int64_t edg_split_pitch(int32_t cols) { return cols > 0 ? 2 * (((int64_t)cols + 63) / 64 * 64) : 0; }

extern "C" int edg_split_f16(const float* x, int64_t ldx, int32_t rows, int32_t cols, void* out, int64_t ldo,
                             float* amax, edg_stream stream) {
  if (rows < 0 || cols <= 0 || !amax) return EDG_ERR_ARG;
  if ((int64_t)rows * ((cols + 63) / 64 * 16) >= ((int64_t)1 << 31)) return EDG_ERR_UNSUPPORTED;  // 32-bit element counters
  if (ldo != edg_split_pitch(cols) || ldx < cols) return EDG_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(amax, 0, sizeof(float), s);
  if (rows == 0) return check_launch();
  if (!x || !out) return EDG_ERR_ARG;
  if (!aligned16(x) || !aligned16(out) || (ldx & 3)) return EDG_ERR_ALIGN;
  const int64_t n4 = (int64_t)rows * ((cols + 3) / 4);
  int grid = (int)((n4 + 255) / 256 < 8 * kNumSMs ? (n4 + 255) / 256 : 8 * kNumSMs);
  amax_kernel<<<grid, 256, 0, s>>>(x, ldx, rows, cols, amax);
  const int kp = (int)(ldo / 2);
  const int64_t n8 = (int64_t)rows * (kp / 8);
  grid = (int)((n8 + 255) / 256 < 16 * kNumSMs ? (n8 + 255) / 256 : 16 * kNumSMs);
  split_f16_kernel<<<grid, 256, 0, s>>>(x, ldx, rows, cols, amax, (__half*)out, ldo, kp);
  return check_launch();
}

extern "C" int edg_linear_split_ok(int32_t K, int32_t Nout) { return (K > 0 && Nout > 0 && linear_split_ok(K, Nout)) ? 1 : 0; }

extern "C" int edg_linear_split(const void* A2, int64_t lda, const float* amax_a, int32_t M, int32_t K, const void* W2,
                                int64_t ldw, const float* amax_w, int32_t Nout, const float* bias, int act, float* C,
                                int64_t ldc, edg_stream stream) {
  if (M < 0 || K <= 0 || Nout <= 0) return EDG_ERR_ARG;
  if (M == 0) return EDG_OK;
  if (!A2 || !W2 || !C || !amax_a || !amax_w) return EDG_ERR_ARG;
  if (act < EDG_ACT_NONE || act > EDG_ACT_RELU) return EDG_ERR_ARG;
  if (lda != edg_split_pitch(K) || ldw != lda || ldc < Nout) return EDG_ERR_ARG;
  if (!aligned16(A2) || !aligned16(W2) || !aligned16(C) || (ldc & 3)) return EDG_ERR_ALIGN;
  return launch_linear_split(A2, lda, amax_a, M, K, W2, ldw, amax_w, Nout, bias, act, C, ldc, (cudaStream_t)stream);
}

extern "C" size_t edg_wgrad_split_workspace(int32_t R, int32_t K1, int32_t K2) {
  if (R <= 0 || K1 <= 0 || K2 <= 0) return 16;
  return wgrad_split_workspace(R, K1, K2) + 256;
}

extern "C" int edg_wgrad_split(const void* A2, int64_t lda, const float* amax_a, int32_t K1, const void* B2, int64_t ldb,
                               const float* amax_b, int32_t K2, int32_t R, float* dW, int64_t lddw, float* dbias,
                               int bias_of, void* ws, size_t ws_bytes, edg_stream stream) {
  if (R < 0 || K1 <= 0 || K2 <= 0 || !dW || lddw < K2) return EDG_ERR_ARG;
  if (bias_of < 0 || bias_of > 2 || (bias_of && !dbias)) return EDG_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (R == 0) {
    cudaMemset2DAsync(dW, lddw * sizeof(float), 0, K2 * sizeof(float), K1, s);
    if (bias_of) cudaMemsetAsync(dbias, 0, (bias_of == 1 ? K1 : K2) * sizeof(float), s);
    return check_launch();
  }
  if (!A2 || !B2 || !ws || !amax_a || !amax_b) return EDG_ERR_ARG;
  if (lda != edg_split_pitch(K1) || ldb != edg_split_pitch(K2)) return EDG_ERR_ARG;
  if (!aligned16(A2) || !aligned16(B2) || !aligned16(ws)) return EDG_ERR_ALIGN;
  if (ws_bytes < edg_wgrad_split_workspace(R, K1, K2)) return EDG_ERR_WORKSPACE;
  return launch_wgrad_split(A2, lda, amax_a, K1, B2, ldb, amax_b, K2, R, dW, lddw, dbias, bias_of, (float*)ws, s);
}
