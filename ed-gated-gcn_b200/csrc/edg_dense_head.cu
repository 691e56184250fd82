// The classifier head and its loss on [B,*]-sized data (models/bert_amir5.py:643 `self.dense(cat[aspect, pooled])`,
// train.py:121 `criterion(outputs, targets)` = nn.CrossEntropyLoss):
//     logits_b = W [a_b | p_b] + bias                      W fp32 [C, 2D]
//     loss     = mean_b ( logsumexp(logits_b) - logits_b[target_b] )
// and what autograd derives from them.  In the torch formulation this [4096 x 600] x [600 x 34] problem is a cat, three
// sm80 cuBLAS GEMMs, a split-K reduce, four loss kernels and a handful of elementwise launches (~130 us of the 0.79 ms
// step at config 2, all of it between the forward and the backward pass); here it is four launches.  Same scheme as
// edg_head.cu: a block owns 32 sentences, W and the block's rows are staged in shared memory with wide loads, every
// reduction runs in a fixed order.  (Included after edg_head.cu: shares its staging helpers and its reduce kernel.)
#include "edg_common.cuh"

namespace edg {

// rows [b0, b0+nb) of x[:, d0 : d0+kHeadPitch) into Xs [32][kHeadPitch], zero beyond nb / lim
__device__ __forceinline__ void stage_rows_chunk(float* __restrict__ Xs, const float* __restrict__ x, int64_t ldx, int b0,
                                                 int nb, int d0, int lim, bool vec4) {
  if (vec4) {
    constexpr int P4 = kHeadPitch / 4;
#pragma unroll 4
    for (int i = threadIdx.x; i < kHeadGraphs * P4; i += kHeadThreads) {
      const int g = i / P4, jj = (i - g * P4) << 2;
      float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g < nb && jj < lim) u = __ldg(reinterpret_cast<const float4*>(x + (int64_t)(b0 + g) * ldx + d0 + jj));
      *reinterpret_cast<float4*>(Xs + g * kHeadPitch + jj) = u;
    }
  } else {
#pragma unroll 4
    for (int i = threadIdx.x; i < kHeadGraphs * kHeadPitch; i += kHeadThreads) {
      const int g = i / kHeadPitch, jj = i - g * kHeadPitch;
      Xs[i] = (g < nb && jj < lim) ? __ldg(x + (int64_t)(b0 + g) * ldx + d0 + jj) : 0.f;
    }
  }
}

// ---- forward: logits = [a | p] W^T + bias ------------------------------------------------------------
// Block (x, h): 32 sentences x column half h (a against W[:, :D] or p against W[:, D:]); a warp owns four sentences and
// takes them two at a time so that every staged W element feeds two rows.  The two halves of a sentence add their
// parts into the zero-initialised logits with one atomicAdd each (half 0 carries the bias): two commuting additions
// onto zero, so the result does not depend on their order.
template <int CP>
__global__ void __launch_bounds__(kHeadThreads, 2)
dense_head_fwd_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ p, int64_t ldp,
                      const float* __restrict__ W, int64_t ldw, const float* __restrict__ bias, int B, int D, int C,
                      int vec4, float* __restrict__ logits, int64_t ldl) {
  extern __shared__ __align__(16) float head_smem[];
  __shared__ float lg_s[kHeadGraphs][CP];
  float* Ws = head_smem;                                                 // [CP][kHeadPitch]
  float* Xs = head_smem + CP * kHeadPitch;                               // [32][kHeadPitch]
  const int b0 = blockIdx.x * kHeadGraphs;
  const int nb = min(kHeadGraphs, B - b0);
  const int half = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g0 = warp * 4;
  for (int i = threadIdx.x; i < kHeadGraphs * CP; i += kHeadThreads) (&lg_s[0][0])[i] = 0.f;
  const float* x = half == 0 ? a : p;
  const int64_t ldx = half == 0 ? lda : ldp;
  for (int d0 = 0; d0 < D; d0 += kHeadPitch) {
    __syncthreads();                                                     // previous pass consumed (and lg_s zeroed)
    const int lim = min(kHeadPitch, D - d0);
    stage_w_chunk<CP>(Ws, W, ldw, C, half * D + d0, (half + 1) * D, vec4 != 0);
    stage_rows_chunk(Xs, x, ldx, b0, nb, d0, lim, vec4 != 0);
    __syncthreads();
    if (g0 >= nb) continue;
#pragma unroll 1
    for (int q = 0; q < 4; q += 2) {
      if (g0 + q >= nb) break;
      const float* u0 = Xs + (g0 + q) * kHeadPitch;
      const float* u1 = u0 + kHeadPitch;                                  // rows >= nb are zero
      constexpr int CG = (CP % 18 == 0) ? 18 : (CP >= 16 ? 16 : CP);
#pragma unroll 1
      for (int c0 = 0; c0 < CP; c0 += CG) {
        float acc0[CG], acc1[CG];
#pragma unroll
        for (int cc = 0; cc < CG; ++cc) { acc0[cc] = 0.f; acc1[cc] = 0.f; }
#pragma unroll 2
        for (int jj = lane; jj < kHeadPitch; jj += 32) {                  // columns >= lim hold zeros
          const float x0 = u0[jj], x1 = u1[jj];
          const float* wcol = Ws + c0 * kHeadPitch + jj;
#pragma unroll
          for (int cc = 0; cc < CG; ++cc) {
            const float w = wcol[cc * kHeadPitch];
            acc0[cc] = fmaf(x0, w, acc0[cc]);
            acc1[cc] = fmaf(x1, w, acc1[cc]);
          }
        }
#pragma unroll
        for (int cc = 0; cc < CG; ++cc) {
          const float t0 = warp_sum(acc0[cc]), t1 = warp_sum(acc1[cc]);
          if (lane == cc) { lg_s[g0 + q][c0 + cc] += t0; lg_s[g0 + q + 1][c0 + cc] += t1; }   // the warp owns these rows
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb * C; i += kHeadThreads) {
    const int g = i / C, cc = i - g * C;
    atomicAdd(logits + (int64_t)(b0 + g) * ldl + cc, lg_s[g][cc] + ((half == 0 && bias) ? __ldg(bias + cc) : 0.f));
  }
}

// ---- backward ------------------------------------------------------------------------------------------
//   d a = g W[:, :D],  d p = g W[:, D:]      (g = d logits, [B,C]);   d W = g^T [a | p],  d bias = colsum(g)
// Block (x, h): 32 sentences x column half h.  Parameter gradients leave as per-block partials [nblocks][C][2D+1]
// (column 2D = bias, written by half 1) and are summed in block order by fc_head_reduce_kernel.
template <int CP>
__global__ void __launch_bounds__(kHeadThreads, 2)
dense_head_bwd_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ a, int64_t lda,
                      const float* __restrict__ p, int64_t ldp, const float* __restrict__ W, int64_t ldw, int B, int D,
                      int C, int vec4, int parts, const float* __restrict__ g2, int64_t ldg2,
                      const float* __restrict__ da_add, const float* __restrict__ dp_add, float* __restrict__ da,
                      float* __restrict__ dp, float* __restrict__ partial) {
  // g2 (optional): a second d logits term added to g on load (the loss' share + the scores term's share); da_add /
  // dp_add (optional, [B, D] contiguous): added to the input gradients on store (what else flows into a and pooled)
  // parts: bit 0 = input gradients (needs W), bit 1 = parameter-gradient partials (needs the rows): two launches on two
  // streams keep the parameter gradients, which nothing downstream waits for, off the critical path
  extern __shared__ __align__(16) float head_smem[];
  __shared__ __align__(16) float g_t[CP][kHeadGraphs];                   // [class][sentence]
  __shared__ __align__(16) float g_s[kHeadGraphs][CP];                   // [sentence][class]
  float* Ws = head_smem;
  float* Xs = head_smem + CP * kHeadPitch;
  const int b0 = blockIdx.x * kHeadGraphs;
  const int nb = min(kHeadGraphs, B - b0);
  const int half = blockIdx.y;
  const int W2 = 2 * D;
  for (int i = threadIdx.x; i < CP * kHeadGraphs; i += kHeadThreads) {
    const int gi = i / CP, cc = i - gi * CP;
    float x = (gi < nb && cc < C) ? g[(int64_t)(b0 + gi) * ldg + cc] : 0.f;
    if (g2 && gi < nb && cc < C) x += g2[(int64_t)(b0 + gi) * ldg2 + cc];
    g_t[cc][gi] = x;
    g_s[gi][cc] = x;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g0 = warp * 4;
  const float* x = half == 0 ? a : p;
  const int64_t ldx = half == 0 ? lda : ldp;
  float* dx = half == 0 ? da : dp;
  const float* dx_add = half == 0 ? da_add : dp_add;
  float* P = partial + (int64_t)blockIdx.x * C * (W2 + 1);
  for (int d0 = 0; d0 < D; d0 += kHeadPitch) {
    __syncthreads();
    const int lim = min(kHeadPitch, D - d0);
    if (parts & 1) stage_w_chunk<CP>(Ws, W, ldw, C, half * D + d0, (half + 1) * D, vec4 != 0);
    if (parts & 2) stage_rows_chunk(Xs, x, ldx, b0, nb, d0, lim, vec4 != 0);
    __syncthreads();
    if (g0 < nb && dx && (parts & 1)) {
      for (int jj = lane; jj < lim; jj += 32) {
        float s[4];
        dot4<CP>(Ws + jj, g_t, g0, s);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (g0 + q < nb) {
            const int64_t o = (int64_t)(b0 + g0 + q) * D + d0 + jj;
            dx[o] = dx_add ? s[q] + __ldg(dx_add + o) : s[q];
          }
      }
    }
    for (int jj = threadIdx.x; (parts & 2) && jj < lim; jj += kHeadThreads) {   // thread = column of the pass
      float acc[CP];
#pragma unroll
      for (int cc = 0; cc < CP; ++cc) acc[cc] = 0.f;
      const float* ucol = Xs + jj;
#pragma unroll 2
      for (int gi = 0; gi < kHeadGraphs; ++gi) {                          // rows >= nb are zero
        const float u = ucol[gi * kHeadPitch];
#pragma unroll
        for (int c4 = 0; c4 < CP; c4 += 4) {
          const float4 l = *reinterpret_cast<const float4*>(&g_s[gi][c4]);
          acc[c4] = fmaf(l.x, u, acc[c4]); acc[c4 + 1] = fmaf(l.y, u, acc[c4 + 1]);
          acc[c4 + 2] = fmaf(l.z, u, acc[c4 + 2]); acc[c4 + 3] = fmaf(l.w, u, acc[c4 + 3]);
        }
      }
      float* pc = P + half * D + d0 + jj;
#pragma unroll
      for (int cc = 0; cc < CP; ++cc)
        if (cc < C) pc[(int64_t)cc * (W2 + 1)] = acc[cc];
    }
  }
  if ((parts & 2) && half == 1 && threadIdx.x < C) {
    float t = 0.f;
    for (int gi = 0; gi < nb; ++gi) t += g_s[gi][threadIdx.x];
    P[(int64_t)threadIdx.x * (W2 + 1) + W2] = t;
  }
}

// ---- cross entropy (mean over the rows whose target is not `ignore`) ----------------------------------
// A warp per row (lanes over the classes).  Every block leaves its (loss sum, row count) in `scratch`; the block that
// takes the last ticket adds the partials in block order (deterministic) and writes out[0] = mean loss, out[1] = count.
constexpr int kCeWarps = 8;
__global__ void __launch_bounds__(kCeWarps * 32)
ce_fwd_kernel(const float* __restrict__ lg, int64_t ldl, const int64_t* __restrict__ tgt, int B, int C, int64_t ignore,
              float* __restrict__ out, int* __restrict__ bad, float* __restrict__ scratch, unsigned* __restrict__ ticket) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kCeWarps + warp;
  float loss = 0.f, cnt = 0.f;
  if (b < B) {
    const int64_t t = tgt[b];
    if (t != ignore && (t < 0 || t >= C)) {
      if (lane == 0) atomicOr(bad, 1);
    } else if (t != ignore) {
      const float* row = lg + (int64_t)b * ldl;
      float m = -INFINITY;
      for (int c = lane; c < C; c += 32) m = fmaxf(m, __ldg(row + c));
      m = warp_max(m);
      float se = 0.f;
      for (int c = lane; c < C; c += 32) se += expf(__ldg(row + c) - m);
      se = warp_sum(se);
      loss = (m + logf(se)) - __ldg(row + t);
      cnt = 1.f;
    }
  }
  __shared__ float ws[kCeWarps], wc[kCeWarps];
  __shared__ bool last;
  if (lane == 0) { ws[warp] = loss; wc[warp] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f, c = 0.f;
    for (int w = 0; w < kCeWarps; ++w) { s += ws[w]; c += wc[w]; }
    scratch[2 * blockIdx.x] = s;
    scratch[2 * blockIdx.x + 1] = c;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // the last block: fixed-order sum of the per-block partials (lane-strided, then a shuffle tree)
  if (warp == 0) {
    float s = 0.f, c = 0.f;
    for (unsigned k = lane; k < gridDim.x; k += 32) { s += __ldcg(scratch + 2 * k); c += __ldcg(scratch + 2 * k + 1); }
    s = warp_sum(s);
    c = warp_sum(c);
    if (lane == 0) { out[0] = c > 0.f ? s / c : 0.f; out[1] = c; *ticket = 0u; }
  }
}

// d logits[b, c] = (softmax(logits_b)[c] - [c == target_b]) * gscale / count   (zero rows for ignored targets)
__global__ void __launch_bounds__(kCeWarps * 32)
ce_bwd_kernel(const float* __restrict__ lg, int64_t ldl, const int64_t* __restrict__ tgt, int B, int C, int64_t ignore,
              const float* __restrict__ gscale, const float* __restrict__ fwd_out, float* __restrict__ dlg, int64_t lddl) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kCeWarps + warp;
  if (b >= B) return;
  const int64_t t = tgt[b];
  float* o = dlg + (int64_t)b * lddl;
  if (t == ignore || t < 0 || t >= C) {
    for (int c = lane; c < C; c += 32) o[c] = 0.f;
    return;
  }
  const float cnt = __ldg(fwd_out + 1);
  const float k = (gscale ? __ldg(gscale) : 1.f) / (cnt > 0.f ? cnt : 1.f);
  const float* row = lg + (int64_t)b * ldl;
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, __ldg(row + c));
  m = warp_max(m);
  float se = 0.f;
  for (int c = lane; c < C; c += 32) se += expf(__ldg(row + c) - m);
  se = warp_sum(se);
  const float inv = 1.f / se;
  for (int c = lane; c < C; c += 32) o[c] = (expf(__ldg(row + c) - m) * inv - (c == t ? 1.f : 0.f)) * k;
}

// ---- loss = w0 t0 + w1 t1 + w2 t2 on device scalars (train.py:115-118: CE + gate_w xy + kl_w kl) and its backward ----
__global__ void loss_combine_kernel(const float* __restrict__ t0, const float* __restrict__ t1, const float* __restrict__ t2,
                                    float w0, float w1, float w2, float* __restrict__ out) {
  if (threadIdx.x == 0) out[0] = (t0 ? w0 * t0[0] : 0.f) + (t1 ? w1 * t1[0] : 0.f) + (t2 ? w2 * t2[0] : 0.f);
}
__global__ void loss_combine_bwd_kernel(const float* __restrict__ g, float w0, float w1, float w2, float* __restrict__ out) {
  if (threadIdx.x == 0) {
    const float gg = g ? g[0] : 1.f;
    out[0] = gg * w0; out[1] = gg * w1; out[2] = gg * w2;
  }
}

}  // namespace edg

using namespace edg;

extern "C" int edg_loss_combine(const float* t0, const float* t1, const float* t2, float w0, float w1, float w2, float* out,
                                edg_stream stream) {
  if (!out) return EDG_ERR_ARG;
  loss_combine_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(t0, t1, t2, w0, w1, w2, out);
  return check_launch();
}

extern "C" int edg_loss_combine_bwd(const float* g, float w0, float w1, float w2, float* out3, edg_stream stream) {
  if (!out3) return EDG_ERR_ARG;
  loss_combine_bwd_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(g, w0, w1, w2, out3);
  return check_launch();
}

extern "C" int edg_dense_head_fwd(const float* a, int64_t lda, const float* p, int64_t ldp, const float* W, int64_t ldw,
                                  const float* bias, int32_t B, int32_t D, int32_t C, float* logits, int64_t ldl,
                                  edg_stream stream) {
  if (B < 0 || D <= 0 || C <= 0) return EDG_ERR_ARG;
  if (C > kHeadMaxC) return EDG_ERR_UNSUPPORTED;
  if (B == 0) return EDG_OK;
  if (!a || !p || !W || !logits) return EDG_ERR_ARG;
  if (lda < D || ldp < D || ldw < 2 * D || ldl < C) return EDG_ERR_ARG;
  const int vec4 = ((D & 3) == 0 && (ldw & 3) == 0 && (lda & 3) == 0 && (ldp & 3) == 0 && aligned16(W) && aligned16(a) &&
                    aligned16(p)) ? 1 : 0;
  const int blocks = (B + kHeadGraphs - 1) / kHeadGraphs;
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemset2DAsync(logits, ldl * sizeof(float), 0, C * sizeof(float), B, s);      // the two halves add into it
  auto launch = [&](auto kern, int CP) -> int {
    const size_t smem = (size_t)(CP + kHeadGraphs) * kHeadPitch * sizeof(float);
    if (int rc_ = ensure_dyn_smem((const void*)kern, smem)) return rc_;
    kern<<<dim3(blocks, 2), kHeadThreads, smem, s>>>(a, lda, p, ldp, W, ldw, bias, B, D, C, vec4, logits, ldl);
    return check_launch();
  };
  if (C <= 8) return launch(dense_head_fwd_kernel<8>, 8);
  if (C <= 36) return launch(dense_head_fwd_kernel<36>, 36);
  return launch(dense_head_fwd_kernel<64>, 64);
}

extern "C" size_t edg_dense_head_bwd_workspace(int32_t B, int32_t D, int32_t C) {
  if (B <= 0 || D <= 0 || C <= 0) return 16;
  const size_t blocks = (size_t)(B + kHeadGraphs - 1) / kHeadGraphs;
  return blocks * (size_t)C * (2 * (size_t)D + 1) * sizeof(float);
}

extern "C" int edg_dense_head_bwd(const float* g, int64_t ldg, const float* a, int64_t lda, const float* p, int64_t ldp,
                                  const float* W, int64_t ldw, int32_t B, int32_t D, int32_t C, int parts,
                                  const float* g2, int64_t ldg2, const float* da_add, const float* dp_add, float* da,
                                  float* dp, float* dW, int64_t lddw, float* dbias, void* ws, size_t ws_bytes,
                                  edg_stream stream) {
  if (B < 0 || D <= 0 || C <= 0 || parts < 1 || parts > 3) return EDG_ERR_ARG;
  if ((parts & 2) && (!dW || !dbias || lddw < 2 * D)) return EDG_ERR_ARG;
  if (C > kHeadMaxC) return EDG_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  if (B == 0) {
    if (parts & 2) {
      cudaMemset2DAsync(dW, lddw * sizeof(float), 0, 2 * (size_t)D * sizeof(float), C, s);
      cudaMemsetAsync(dbias, 0, C * sizeof(float), s);
    }
    return check_launch();
  }
  if (!g || !a || !p || !W || !ws) return EDG_ERR_ARG;
  if (ldg < C || lda < D || ldp < D || ldw < 2 * D || (g2 && ldg2 < C)) return EDG_ERR_ARG;
  if (ws_bytes < edg_dense_head_bwd_workspace(B, D, C)) return EDG_ERR_WORKSPACE;
  const int vec4 = ((D & 3) == 0 && (ldw & 3) == 0 && (lda & 3) == 0 && (ldp & 3) == 0 && aligned16(W) && aligned16(a) &&
                    aligned16(p)) ? 1 : 0;
  const int blocks = (B + kHeadGraphs - 1) / kHeadGraphs;
  float* partial = reinterpret_cast<float*>(ws);
  auto launch = [&](auto kern, int CP) -> int {
    const size_t smem = (size_t)(CP + kHeadGraphs) * kHeadPitch * sizeof(float);
    if (int rc_ = ensure_dyn_smem((const void*)kern, smem)) return rc_;
    kern<<<dim3(blocks, 2), kHeadThreads, smem, s>>>(g, ldg, a, lda, p, ldp, W, ldw, B, D, C, vec4, parts, g2, ldg2, da_add, dp_add,
                                                     da, dp, partial);
    return check_launch();
  };
  int rc;
  if (C <= 8) rc = launch(dense_head_bwd_kernel<8>, 8);
  else if (C <= 36) rc = launch(dense_head_bwd_kernel<36>, 36);
  else rc = launch(dense_head_bwd_kernel<64>, 64);
  if (rc) return rc;
  if (!(parts & 2)) return EDG_OK;
  const int64_t per = (int64_t)C * (2 * D + 1);
  fc_head_reduce_kernel<<<(unsigned)((per + 255) / 256), 256, 0, s>>>(partial, blocks, C, 2 * D, dW, lddw, dbias, nullptr, 0, B,
                                                                      nullptr, 0, 2, 0u);
  return check_launch();
}

// out: device float[2] = { mean loss over the counted rows, number of counted rows }; bad: device int[1], set to 1 when a
// target lies outside [0, C) and is not `ignore_index` (the caller decides when to look: no host sync here);
// ws: edg_cross_entropy_workspace(B) bytes, its LAST 4 bytes a ticket counter that must be zero on the first call (the
// kernel leaves it zero)
extern "C" size_t edg_cross_entropy_workspace(int32_t B) {
  const size_t blocks = B > 0 ? ((size_t)B + kCeWarps - 1) / kCeWarps : 1;
  return (2 * blocks + 1) * sizeof(float);
}

extern "C" int edg_cross_entropy_fwd(const float* logits, int64_t ldl, const int64_t* target, int32_t B, int32_t C,
                                     int64_t ignore_index, float* out, int32_t* bad, void* ws, size_t ws_bytes,
                                     edg_stream stream) {
  if (B < 0 || C <= 0 || !out || !bad || !ws) return EDG_ERR_ARG;
  if (B > 0 && (!logits || !target || ldl < C)) return EDG_ERR_ARG;
  if (ws_bytes < edg_cross_entropy_workspace(B)) return EDG_ERR_WORKSPACE;
  const int blocks = B > 0 ? (B + kCeWarps - 1) / kCeWarps : 1;
  float* scratch = (float*)ws;
  unsigned* ticket = (unsigned*)(scratch + 2 * (size_t)blocks);
  ce_fwd_kernel<<<blocks, kCeWarps * 32, 0, (cudaStream_t)stream>>>(logits, ldl, target, B, C, ignore_index, out, bad, scratch,
                                                                    ticket);
  return check_launch();
}

extern "C" int edg_cross_entropy_bwd(const float* logits, int64_t ldl, const int64_t* target, int32_t B, int32_t C,
                                     int64_t ignore_index, const float* gscale, const float* fwd_out, float* dlogits,
                                     int64_t lddl, edg_stream stream) {
  if (B < 0 || C <= 0) return EDG_ERR_ARG;
  if (B == 0) return EDG_OK;
  if (!logits || !target || !fwd_out || !dlogits || ldl < C || lddl < C) return EDG_ERR_ARG;
  ce_bwd_kernel<<<(B + kCeWarps - 1) / kCeWarps, kCeWarps * 32, 0, (cudaStream_t)stream>>>(logits, ldl, target, B, C,
                                                                                         ignore_index, gscale, fwd_out, dlogits,
                                                                                         lddl);
  return check_launch();
}
