// Per-sentence algebra of the importance-score term (models/bert_amir5.py:645-646 in the collapsed form of
// SURVEY A9):  output_w = fc(cat[x_out, a]),  scores = <logits, output_w>  ==>
//     [v_b | va_b] = logits_b @ Wfc          ([C] x [C, 2D]),
//     c_b          = a_b . va_b + logits_b . bfc,            scores[b,t] = x_out[b,t,:] . v_b + c_b.
// Forward and backward of that [B,*]-sized map as two fp32 kernels (they replace three skinny cuBLAS GEMMs and
// a dozen elementwise launches of the torch formulation).  A block owns 32 sentences, a warp 4 of them; Wfc is
// read through L1 (34 x 600 floats = 82 KB at C2).  Reductions run in a fixed order (deterministic).
#include "edg_common.cuh"

namespace edg {

constexpr int kHeadGraphs = 32;     // sentences per block
constexpr int kHeadThreads = 256;   // 8 warps x 4 sentences
constexpr int kHeadMaxC = 64;

// ---- forward ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kHeadThreads)
fc_head_fwd_kernel(const float* __restrict__ lg, int64_t ldl, const float* __restrict__ W, int64_t ldw,
                   const float* __restrict__ fcb, const float* __restrict__ a, int64_t lda, int B, int D, int C,
                   float* __restrict__ v, float* __restrict__ c) {
  __shared__ __align__(16) float lg_t[kHeadMaxC][kHeadGraphs];          // [class][sentence]
  const int b0 = blockIdx.x * kHeadGraphs;
  const int nb = min(kHeadGraphs, B - b0);
  for (int i = threadIdx.x; i < C * kHeadGraphs; i += kHeadThreads) {
    const int g = i / C, cc = i - g * C;                                 // consecutive threads read consecutive classes
    lg_t[cc][g] = g < nb ? lg[(int64_t)(b0 + g) * ldl + cc] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g0 = warp * 4;
  if (g0 >= nb) return;
  float cacc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j = lane; j < 2 * D; j += 32) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int cc = 0; cc < C; ++cc) {
      const float w = __ldg(W + (int64_t)cc * ldw + j);
      const float4 l = *reinterpret_cast<const float4*>(&lg_t[cc][g0]);
      s[0] = fmaf(l.x, w, s[0]); s[1] = fmaf(l.y, w, s[1]); s[2] = fmaf(l.z, w, s[2]); s[3] = fmaf(l.w, w, s[3]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (g0 + q >= nb) break;
      const int64_t b = b0 + g0 + q;
      if (j < D) v[b * D + j] = s[q];
      else cacc[q] = fmaf(__ldg(a + b * lda + (j - D)), s[q], cacc[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) cacc[q] = warp_sum(cacc[q]);
  if (lane < 4 && g0 + lane < nb) {
    float t = lane == 0 ? cacc[0] : lane == 1 ? cacc[1] : lane == 2 ? cacc[2] : cacc[3];
    for (int cc = 0; cc < C; ++cc) t = fmaf(lg_t[cc][g0 + lane], __ldg(fcb + cc), t);
    c[b0 + g0 + lane] = t;
  }
}

// ---- backward ---------------------------------------------------------------------------------
// With u_b = [dv_b | dc_b * a_b] (length 2D; dv, dc optionally scaled by the device scalar *scale):
//   d logits[b,cc] = u_b . Wfc[cc,:] + dc_b * bfc[cc]
//   d a[b,j]       = dc_b * va[b,j],   va = logits_b @ Wfc[:, D:]
//   d Wfc[cc,j]    = sum_b logits[b,cc] * u_b[j],   d bfc[cc] = sum_b dc_b * logits[b,cc]
// The parameter gradients leave as per-block partials [nblocks][C][2D+1] (column 2D = bias), reduced by
// fc_head_reduce_kernel in block order.
template <int CP>
__global__ void __launch_bounds__(kHeadThreads)
fc_head_bwd_kernel(const float* __restrict__ lg, int64_t ldl, const float* __restrict__ W, int64_t ldw,
                   const float* __restrict__ fcb, const float* __restrict__ a, int64_t lda,
                   const float* __restrict__ dv, const float* __restrict__ dc, const float* __restrict__ scale,
                   int B, int D, int C, float* __restrict__ dlg, int64_t lddl, float* __restrict__ da,
                   float* __restrict__ partial) {
  __shared__ __align__(16) float lg_t[CP][kHeadGraphs];                  // [class][sentence]
  __shared__ __align__(16) float lg_s[kHeadGraphs][CP];                  // [sentence][class]
  __shared__ float dc_s[kHeadGraphs];
  const int b0 = blockIdx.x * kHeadGraphs;
  const int nb = min(kHeadGraphs, B - b0);
  const float sc = scale ? __ldg(scale) : 1.f;
  for (int i = threadIdx.x; i < CP * kHeadGraphs; i += kHeadThreads) {
    const int g = i / CP, cc = i - g * CP;
    const float x = (g < nb && cc < C) ? lg[(int64_t)(b0 + g) * ldl + cc] : 0.f;
    lg_t[cc][g] = x;
    lg_s[g][cc] = x;
  }
  if (threadIdx.x < kHeadGraphs) dc_s[threadIdx.x] = threadIdx.x < nb ? sc * dc[b0 + threadIdx.x] : 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g0 = warp * 4;
  const int W2 = 2 * D;

  // ---- d a: four sentences of the warp share every Wfc load (same loop as the forward kernel)
  if (g0 < nb) {
    for (int j = lane; j < D; j += 32) {
      float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
      for (int cc = 0; cc < C; ++cc) {
        const float w = __ldg(W + (int64_t)cc * ldw + D + j);
        const float4 l = *reinterpret_cast<const float4*>(&lg_t[cc][g0]);
        s[0] = fmaf(l.x, w, s[0]); s[1] = fmaf(l.y, w, s[1]); s[2] = fmaf(l.z, w, s[2]); s[3] = fmaf(l.w, w, s[3]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (g0 + q < nb) da[(int64_t)(b0 + g0 + q) * D + j] = dc_s[g0 + q] * s[q];
    }
    // ---- d logits: one sentence at a time, lanes stride the 2D columns, CP accumulators per lane
    for (int q = 0; q < 4 && g0 + q < nb; ++q) {
      const int64_t b = b0 + g0 + q;
      const float dcb = dc_s[g0 + q];
      float acc[CP];
#pragma unroll
      for (int cc = 0; cc < CP; ++cc) acc[cc] = 0.f;
      for (int j = lane; j < W2; j += 32) {
        const float u = j < D ? sc * __ldg(dv + b * D + j) : dcb * __ldg(a + b * lda + (j - D));
#pragma unroll
        for (int cc = 0; cc < CP; ++cc)
          if (cc < C) acc[cc] = fmaf(u, __ldg(W + (int64_t)cc * ldw + j), acc[cc]);
      }
      float mine = 0.f;                                                   // lane cc keeps class cc (and cc + 32)
      float mine2 = 0.f;
#pragma unroll
      for (int cc = 0; cc < CP; ++cc) {
        const float t = warp_sum(acc[cc]);
        if (cc < 32) { if (lane == cc) mine = t; }
        else if (lane == cc - 32) mine2 = t;
      }
      if (lane < C) dlg[b * lddl + lane] = fmaf(dcb, __ldg(fcb + lane), mine);
      if (CP > 32 && lane + 32 < C) dlg[b * lddl + lane + 32] = fmaf(dcb, __ldg(fcb + lane + 32), mine2);
    }
  }

  // ---- parameter-gradient partials of this block: thread = column j of [dv | dc*a | dc]
  float* P = partial + (int64_t)blockIdx.x * C * (W2 + 1);
  for (int j = threadIdx.x; j <= W2; j += kHeadThreads) {
    float acc[CP];
#pragma unroll
    for (int cc = 0; cc < CP; ++cc) acc[cc] = 0.f;
    for (int g = 0; g < nb; ++g) {
      const int64_t b = b0 + g;
      const float u = j < D ? sc * __ldg(dv + b * D + j) : (j < W2 ? dc_s[g] * __ldg(a + b * lda + (j - D)) : dc_s[g]);
#pragma unroll
      for (int c4 = 0; c4 < CP; c4 += 4) {
        const float4 l = *reinterpret_cast<const float4*>(&lg_s[g][c4]);
        acc[c4] = fmaf(l.x, u, acc[c4]); acc[c4 + 1] = fmaf(l.y, u, acc[c4 + 1]);
        acc[c4 + 2] = fmaf(l.z, u, acc[c4 + 2]); acc[c4 + 3] = fmaf(l.w, u, acc[c4 + 3]);
      }
    }
#pragma unroll
    for (int cc = 0; cc < CP; ++cc)
      if (cc < C) P[(int64_t)cc * (W2 + 1) + j] = acc[cc];
  }
}

// d Wfc[cc, j] = sum over blocks of partial[.][cc][j] (j < 2D), d bfc[cc] = column 2D; fixed order.
__global__ void __launch_bounds__(256)
fc_head_reduce_kernel(const float* __restrict__ partial, int nblocks, int C, int W2, float* __restrict__ dW,
                      int64_t lddw, float* __restrict__ db) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per = (int64_t)C * (W2 + 1);
  if (idx >= per) return;
  const int cc = (int)(idx / (W2 + 1)), j = (int)(idx - (int64_t)cc * (W2 + 1));
  float s = 0.f;
  for (int k = 0; k < nblocks; ++k) s += partial[k * per + idx];
  if (j < W2) dW[(int64_t)cc * lddw + j] = s;
  else db[cc] = s;
}

}  // namespace edg

using namespace edg;

extern "C" int edg_fc_head_fwd(const float* logits, int64_t ldl, const float* fc_w, int64_t ldw, const float* fc_b,
                               const float* a, int64_t lda, int32_t B, int32_t D, int32_t C, float* v, float* c,
                               edg_stream stream) {
  if (B < 0 || D <= 0 || C <= 0) return EDG_ERR_ARG;
  if (C > kHeadMaxC) return EDG_ERR_UNSUPPORTED;
  if (B == 0) return EDG_OK;
  if (!logits || !fc_w || !fc_b || !a || !v || !c) return EDG_ERR_ARG;
  if (ldl < C || ldw < 2 * D || lda < D) return EDG_ERR_ARG;
  const int blocks = (B + kHeadGraphs - 1) / kHeadGraphs;
  fc_head_fwd_kernel<<<blocks, kHeadThreads, 0, (cudaStream_t)stream>>>(logits, ldl, fc_w, ldw, fc_b, a, lda, B, D, C, v, c);
  return check_launch();
}

extern "C" size_t edg_fc_head_bwd_workspace(int32_t B, int32_t D, int32_t C) {
  if (B <= 0 || D <= 0 || C <= 0) return 16;
  const size_t blocks = (size_t)(B + kHeadGraphs - 1) / kHeadGraphs;
  return blocks * (size_t)C * (2 * (size_t)D + 1) * sizeof(float);
}

extern "C" int edg_fc_head_bwd(const float* logits, int64_t ldl, const float* fc_w, int64_t ldw, const float* fc_b,
                               const float* a, int64_t lda, const float* dv, const float* dc, const float* scale,
                               int32_t B, int32_t D, int32_t C, float* d_logits, int64_t lddl, float* d_a,
                               float* d_fc_w, int64_t lddw, float* d_fc_b, void* ws, size_t ws_bytes,
                               edg_stream stream) {
  if (B < 0 || D <= 0 || C <= 0) return EDG_ERR_ARG;
  if (C > kHeadMaxC) return EDG_ERR_UNSUPPORTED;
  if (!d_fc_w || !d_fc_b || lddw < 2 * D) return EDG_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (B == 0) {
    cudaMemset2DAsync(d_fc_w, lddw * sizeof(float), 0, 2 * (size_t)D * sizeof(float), C, s);
    cudaMemsetAsync(d_fc_b, 0, C * sizeof(float), s);
    return check_launch();
  }
  if (!logits || !fc_w || !fc_b || !a || !dv || !dc || !d_logits || !d_a || !ws) return EDG_ERR_ARG;
  if (ldl < C || ldw < 2 * D || lda < D || lddl < C) return EDG_ERR_ARG;
  if (ws_bytes < edg_fc_head_bwd_workspace(B, D, C)) return EDG_ERR_WORKSPACE;
  const int blocks = (B + kHeadGraphs - 1) / kHeadGraphs;
  float* partial = reinterpret_cast<float*>(ws);
  if (C <= 8)
    fc_head_bwd_kernel<8><<<blocks, kHeadThreads, 0, s>>>(logits, ldl, fc_w, ldw, fc_b, a, lda, dv, dc, scale, B, D, C,
                                                          d_logits, lddl, d_a, partial);
  else if (C <= 40)
    fc_head_bwd_kernel<40><<<blocks, kHeadThreads, 0, s>>>(logits, ldl, fc_w, ldw, fc_b, a, lda, dv, dc, scale, B, D, C,
                                                           d_logits, lddl, d_a, partial);
  else
    fc_head_bwd_kernel<64><<<blocks, kHeadThreads, 0, s>>>(logits, ldl, fc_w, ldw, fc_b, a, lda, dv, dc, scale, B, D, C,
                                                           d_logits, lddl, d_a, partial);
  int rc = check_launch();
  if (rc) return rc;
  const int64_t per = (int64_t)C * (2 * D + 1);
  fc_head_reduce_kernel<<<(unsigned)((per + 255) / 256), 256, 0, s>>>(partial, blocks, C, 2 * D, d_fc_w, lddw, d_fc_b);
  return check_launch();
}
