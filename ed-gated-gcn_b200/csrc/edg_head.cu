// Per-sentence algebra of the importance-score term (models/bert_amir5.py:645-646 in the collapsed form of
// SURVEY A9):  output_w = fc(cat[x_out, a]),  scores = <logits, output_w>  ==>
//     [v_b | va_b] = logits_b @ Wfc          ([C] x [C, 2D]),
//     c_b          = a_b . va_b + logits_b . bfc,            scores[b,t] = x_out[b,t,:] . v_b + c_b.
// Forward and backward of that [B,*]-sized map as two fp32 kernels (they replace three skinny cuBLAS GEMMs and
// a dozen elementwise launches of the torch formulation).  A block owns 32 sentences, a warp 4 of them; Wfc (34 x 600 floats = 82 KB at C2)
// and the block's u rows are staged in shared memory with wide coalesced loads (a first version read Wfc through
// L1 with dependent loads and was bound by L2 latency: 250 us instead of ~10).  Reductions run in a fixed order (deterministic).
#include "edg_common.cuh"

namespace edg {

constexpr int kHeadGraphs = 32;     // sentences per block
constexpr int kHeadThreads = 256;   // 8 warps x 4 sentences
constexpr int kHeadMaxC = 64;
constexpr int kHeadPitch = 320;     // columns staged per pass (compile-time: shared-memory offsets become immediates)

// Block (x, h) owns sentences [32x, 32x+32) and column half h of [v | va] (h = 0: the v columns, h = 1: the va
// columns), walked in passes of kHeadPitch columns.  CP = class count padded to a multiple of 4 (rows >= C of the
// staged Wfc and of logits are zero, so the unrolled loops need no predicates).

// cooperative, coalesced copy of Wfc[:, col0 : col0 + kHeadPitch) into shared memory (zero rows >= C, zero columns
// >= col_end); 128-bit loads when the pitch / alignment allow (vec4), several loads in flight per thread
template <int CP>
__device__ __forceinline__ void stage_w_chunk(float* __restrict__ Ws, const float* __restrict__ W, int64_t ldw, int C,
                                              int col0, int col_end, bool vec4) {
  if (vec4) {
    constexpr int P4 = kHeadPitch / 4;
#pragma unroll 4
    for (int i = threadIdx.x; i < CP * P4; i += kHeadThreads) {
      const int cc = i / P4, jj = (i - cc * P4) << 2;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cc < C && col0 + jj < col_end) w = __ldg(reinterpret_cast<const float4*>(W + (int64_t)cc * ldw + col0 + jj));
      *reinterpret_cast<float4*>(Ws + cc * kHeadPitch + jj) = w;
    }
  } else {
#pragma unroll 4
    for (int i = threadIdx.x; i < CP * kHeadPitch; i += kHeadThreads) {
      const int cc = i / kHeadPitch, jj = i - cc * kHeadPitch;
      Ws[i] = (cc < C && col0 + jj < col_end) ? __ldg(W + (int64_t)cc * ldw + col0 + jj) : 0.f;
    }
  }
}

// s[q] = sum_cc lg_t[cc][g0 + q] * Ws[cc][jj] for the warp's four sentences (straight-line, immediates only)
template <int CP>
__device__ __forceinline__ void dot4(const float* __restrict__ wcol, const float (*lg_t)[kHeadGraphs], int g0, float (&s)[4]) {
  s[0] = s[1] = s[2] = s[3] = 0.f;
  constexpr int U = (CP % 9 == 0) ? 9 : 8;          // partial unroll: enough loads in flight without spilling
#pragma unroll 1
  for (int c0 = 0; c0 < CP; c0 += U) {
#pragma unroll
    for (int cc = 0; cc < U; ++cc) {
      const float w = wcol[(c0 + cc) * kHeadPitch];
      const float4 l = *reinterpret_cast<const float4*>(&lg_t[c0 + cc][g0]);
      s[0] = fmaf(l.x, w, s[0]); s[1] = fmaf(l.y, w, s[1]); s[2] = fmaf(l.z, w, s[2]); s[3] = fmaf(l.w, w, s[3]);
    }
  }
}

// ---- forward ----------------------------------------------------------------------------------
template <int CP>
__global__ void __launch_bounds__(kHeadThreads, 2)
fc_head_fwd_kernel(const float* __restrict__ lg, int64_t ldl, const float* __restrict__ W, int64_t ldw,
                   const float* __restrict__ fcb, const float* __restrict__ a, int64_t lda, int B, int D, int C,
                   int vec4, float* __restrict__ v, float* __restrict__ c) {
  extern __shared__ __align__(16) float head_smem[];
  __shared__ __align__(16) float lg_t[CP][kHeadGraphs];                  // [class][sentence]
  float* Ws = head_smem;                                                 // [CP][kHeadPitch]
  const int b0 = blockIdx.x * kHeadGraphs;
  const int nb = min(kHeadGraphs, B - b0);
  const int half = blockIdx.y;
  for (int i = threadIdx.x; i < CP * kHeadGraphs; i += kHeadThreads) {
    const int g = i / CP, cc = i - g * CP;
    lg_t[cc][g] = (g < nb && cc < C) ? lg[(int64_t)(b0 + g) * ldl + cc] : 0.f;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g0 = warp * 4;
  float cacc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int d0 = 0; d0 < D; d0 += kHeadPitch) {                           // columns half*D + [d0, d0 + pitch)
    __syncthreads();                                                     // previous pass consumed (and lg_t written)
    stage_w_chunk<CP>(Ws, W, ldw, C, half * D + d0, (half + 1) * D, vec4 != 0);
    __syncthreads();
    if (g0 < nb) {
      const int lim = min(kHeadPitch, D - d0);
      for (int jj = lane; jj < lim; jj += 32) {
        float s[4];
        dot4<CP>(Ws + jj, lg_t, g0, s);
        const int d = d0 + jj;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (g0 + q >= nb) break;
          const int64_t b = b0 + g0 + q;
          if (half == 0) v[b * D + d] = s[q];
          else cacc[q] = fmaf(__ldg(a + b * lda + d), s[q], cacc[q]);
        }
      }
    }
  }
  if (half == 0 || g0 >= nb) return;
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
//   d logits[b,cc] = u_b . Wfc[cc,:] + dc_b * bfc[cc]        (each column half writes its part: dlg_part[h][B][C])
//   d a[b,j]       = dc_b * va[b,j],   va = logits_b @ Wfc[:, D:]                                  (half 1)
//   d Wfc[cc,j]    = sum_b logits[b,cc] * u_b[j],   d bfc[cc] = sum_b dc_b * logits[b,cc]
// Per pass both Wfc[:, pass] and the block's u[:, pass] are staged in shared memory; every phase then reads
// shared memory only.  The parameter gradients leave as per-block partials [nblocks][C][2D+1] (column 2D =
// bias), reduced by fc_head_reduce_kernel in block order (which also adds the two d logits parts).
template <int CP>
__global__ void __launch_bounds__(kHeadThreads, 2)
fc_head_bwd_kernel(const float* __restrict__ lg, int64_t ldl, const float* __restrict__ W, int64_t ldw,
                   const float* __restrict__ fcb, const float* __restrict__ a, int64_t lda,
                   const float* __restrict__ dv, const float* __restrict__ dc, const float* __restrict__ scale,
                   int B, int D, int C, int vec4, int parts, float* __restrict__ dlg_part, float* __restrict__ da,
                   float* __restrict__ partial) {
  // parts: bit 0 = input gradients (d logits, d a), bit 1 = parameter-gradient partials (two launches on two
  // streams keep the parameter gradients, which nothing downstream waits for, off the critical path)
  extern __shared__ __align__(16) float head_smem[];
  __shared__ __align__(16) float lg_t[CP][kHeadGraphs];                  // [class][sentence]
  __shared__ __align__(16) float lg_s[kHeadGraphs][CP];                  // [sentence][class]
  __shared__ float dlg_s[kHeadGraphs][CP];                               // d logits accumulated over the passes
  __shared__ float dc_s[kHeadGraphs];
  float* Ws = head_smem;                                                 // [CP][kHeadPitch]
  float* Us = head_smem + CP * kHeadPitch;                               // [32][kHeadPitch]
  const int b0 = blockIdx.x * kHeadGraphs;
  const int nb = min(kHeadGraphs, B - b0);
  const int half = blockIdx.y;
  const float sc = scale ? __ldg(scale) : 1.f;
  const int W2 = 2 * D;
  for (int i = threadIdx.x; i < CP * kHeadGraphs; i += kHeadThreads) {
    const int g = i / CP, cc = i - g * CP;
    const float x = (g < nb && cc < C) ? lg[(int64_t)(b0 + g) * ldl + cc] : 0.f;
    lg_t[cc][g] = x;
    lg_s[g][cc] = x;
    dlg_s[g][cc] = 0.f;
  }
  if (threadIdx.x < kHeadGraphs) dc_s[threadIdx.x] = threadIdx.x < nb ? sc * dc[b0 + threadIdx.x] : 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g0 = warp * 4;
  float* P = partial + (int64_t)blockIdx.x * C * (W2 + 1);

  for (int d0 = 0; d0 < D; d0 += kHeadPitch) {
    __syncthreads();                                                     // previous pass consumed, dc_s / lg_* written
    if (parts & 1) stage_w_chunk<CP>(Ws, W, ldw, C, half * D + d0, (half + 1) * D, vec4 != 0);
    const int lim = min(kHeadPitch, D - d0);
    // u rows of the block for these columns: sc * dv (half 0) or dc * a (half 1); zero beyond `lim` / `nb`
    if (vec4) {
      constexpr int P4 = kHeadPitch / 4;
#pragma unroll 4
      for (int i = threadIdx.x; i < kHeadGraphs * P4; i += kHeadThreads) {
        const int g = i / P4, jj = (i - g * P4) << 2;
        float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < nb && jj < lim) {
          const int64_t b = b0 + g;
          const float m = half == 0 ? sc : dc_s[g];
          u = __ldg(reinterpret_cast<const float4*>(half == 0 ? dv + b * D + d0 + jj : a + b * lda + d0 + jj));
          u.x *= m; u.y *= m; u.z *= m; u.w *= m;
        }
        *reinterpret_cast<float4*>(Us + g * kHeadPitch + jj) = u;
      }
    } else {
#pragma unroll 4
      for (int i = threadIdx.x; i < kHeadGraphs * kHeadPitch; i += kHeadThreads) {
        const int g = i / kHeadPitch, jj = i - g * kHeadPitch;
        float u = 0.f;
        if (g < nb && jj < lim) {
          const int64_t b = b0 + g;
          u = half == 0 ? sc * __ldg(dv + b * D + d0 + jj) : dc_s[g] * __ldg(a + b * lda + d0 + jj);
        }
        Us[i] = u;
      }
    }
    __syncthreads();
    if (g0 < nb && (parts & 1)) {
      // ---- d a (half 1): four sentences of the warp share every Wfc element
      if (half == 1) {
        for (int jj = lane; jj < lim; jj += 32) {
          float s[4];
          dot4<CP>(Ws + jj, lg_t, g0, s);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (g0 + q < nb) da[(int64_t)(b0 + g0 + q) * D + d0 + jj] = dc_s[g0 + q] * s[q];
        }
      }
      // ---- d logits: two sentences at a time share every Wfc element; lanes stride the columns
#pragma unroll 1
      for (int q = 0; q < 4; q += 2) {
        if (g0 + q >= nb) break;
        const float* u0 = Us + (g0 + q) * kHeadPitch;
        const float* u1 = u0 + kHeadPitch;                                // rows >= nb are zero
        // CG classes per sweep keep 2 x CG accumulators in registers (two blocks per SM need <= 128 registers)
        constexpr int CG = (CP % 18 == 0) ? 18 : (CP >= 16 ? 16 : CP);
#pragma unroll 1
        for (int c0 = 0; c0 < CP; c0 += CG) {
          float acc0[CG], acc1[CG];
#pragma unroll
          for (int cc = 0; cc < CG; ++cc) { acc0[cc] = 0.f; acc1[cc] = 0.f; }
#pragma unroll 2
          for (int jj = lane; jj < kHeadPitch; jj += 32) {                // columns >= lim hold zeros
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
            if (lane == cc) { dlg_s[g0 + q][c0 + cc] += t0; dlg_s[g0 + q + 1][c0 + cc] += t1; }   // the warp owns these rows
          }
        }
      }
    }
    // ---- parameter-gradient partials of this block: thread = column of the pass
    for (int jj = threadIdx.x; (parts & 2) && jj < lim; jj += kHeadThreads) {
      float acc[CP];
#pragma unroll
      for (int cc = 0; cc < CP; ++cc) acc[cc] = 0.f;
      const float* ucol = Us + jj;
#pragma unroll 2
      for (int g = 0; g < kHeadGraphs; ++g) {                             // rows >= nb are zero
        const float u = ucol[g * kHeadPitch];
#pragma unroll
        for (int c4 = 0; c4 < CP; c4 += 4) {
          const float4 l = *reinterpret_cast<const float4*>(&lg_s[g][c4]);
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
  __syncthreads();
  // this half's part of d logits (+ the bias term once); bias-gradient partial (column 2D of the slab)
  float* dl = dlg_part + (int64_t)half * B * C;
  for (int i = threadIdx.x; (parts & 1) && i < nb * C; i += kHeadThreads) {
    const int g = i / C, cc = i - g * C;
    dl[(int64_t)(b0 + g) * C + cc] = half == 0 ? fmaf(dc_s[g], __ldg(fcb + cc), dlg_s[g][cc]) : dlg_s[g][cc];
  }
  if ((parts & 2) && half == 1 && threadIdx.x < C) {
    float t = 0.f;
    for (int g = 0; g < nb; ++g) t = fmaf(dc_s[g], lg_s[g][threadIdx.x], t);
    P[(int64_t)threadIdx.x * (W2 + 1) + W2] = t;
  }
}

// d Wfc[cc, j] = sum over blocks of partial[.][cc][j] (j < 2D), d bfc[cc] = column 2D; fixed order.  The tail of
// the grid adds the two column-half parts of d logits.
__global__ void __launch_bounds__(256)
fc_head_reduce_kernel(const float* __restrict__ partial, int nblocks, int C, int W2, float* __restrict__ dW,
                      int64_t lddw, float* __restrict__ db, const float* __restrict__ dlg_part, int64_t BC,
                      int B, float* __restrict__ dlg, int64_t lddl, int parts, unsigned first_block) {
  const int64_t idx = (int64_t)(blockIdx.x + first_block) * blockDim.x + threadIdx.x;
  const int64_t per = (int64_t)C * (W2 + 1);
  if (idx < per) {
    if (!(parts & 2)) return;
    const int cc = (int)(idx / (W2 + 1)), j = (int)(idx - (int64_t)cc * (W2 + 1));
    float s = 0.f;
#pragma unroll 8
    for (int k = 0; k < nblocks; ++k) s += partial[k * per + idx];
    if (j < W2) dW[(int64_t)cc * lddw + j] = s;
    else db[cc] = s;
  } else if (idx - per < BC && (parts & 1)) {
    const int64_t i = idx - per;
    const int64_t b = i / C;
    dlg[b * lddl + (i - b * C)] = dlg_part[i] + dlg_part[BC + i];
  }
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
  const dim3 grid((B + kHeadGraphs - 1) / kHeadGraphs, 2);
  const int vec4 = ((D & 3) == 0 && (ldw & 3) == 0 && aligned16(fc_w)) ? 1 : 0;
  cudaStream_t s = (cudaStream_t)stream;
  auto launch = [&](auto kern, int CP) -> int {
    const size_t smem = (size_t)CP * kHeadPitch * sizeof(float);
    if (int rc_ = ensure_dyn_smem((const void*)kern, smem)) return rc_;
    kern<<<grid, kHeadThreads, smem, s>>>(logits, ldl, fc_w, ldw, fc_b, a, lda, B, D, C, vec4, v, c);
    return check_launch();
  };
  if (C <= 8) return launch(fc_head_fwd_kernel<8>, 8);
  if (C <= 36) return launch(fc_head_fwd_kernel<36>, 36);
  return launch(fc_head_fwd_kernel<64>, 64);
}

// workspace: per-block parameter-gradient slabs + the two column-half parts of d logits (calls with parts = 1 and
// parts = 2 may run concurrently on two streams: they touch disjoint regions, pass each its own workspace or the
// same one)
extern "C" size_t edg_fc_head_bwd_workspace(int32_t B, int32_t D, int32_t C) {
  if (B <= 0 || D <= 0 || C <= 0) return 16;
  const size_t blocks = (size_t)(B + kHeadGraphs - 1) / kHeadGraphs;
  return (blocks * (size_t)C * (2 * (size_t)D + 1) + 2 * (size_t)B * C) * sizeof(float);
}

extern "C" int edg_fc_head_bwd(const float* logits, int64_t ldl, const float* fc_w, int64_t ldw, const float* fc_b,
                               const float* a, int64_t lda, const float* dv, const float* dc, const float* scale,
                               int32_t B, int32_t D, int32_t C, int parts, float* d_logits, int64_t lddl, float* d_a,
                               float* d_fc_w, int64_t lddw, float* d_fc_b, void* ws, size_t ws_bytes,
                               edg_stream stream) {
  if (B < 0 || D <= 0 || C <= 0 || parts < 1 || parts > 3) return EDG_ERR_ARG;
  if (C > kHeadMaxC) return EDG_ERR_UNSUPPORTED;
  const bool want_in = parts & 1, want_par = parts & 2;
  if (want_par && (!d_fc_w || !d_fc_b || lddw < 2 * D)) return EDG_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (B == 0) {
    if (want_par) {
      cudaMemset2DAsync(d_fc_w, lddw * sizeof(float), 0, 2 * (size_t)D * sizeof(float), C, s);
      cudaMemsetAsync(d_fc_b, 0, C * sizeof(float), s);
    }
    return check_launch();
  }
  if (!logits || !fc_w || !fc_b || !a || !dv || !dc || !ws) return EDG_ERR_ARG;
  if (want_in && (!d_logits || !d_a || lddl < C)) return EDG_ERR_ARG;
  if (ldl < C || ldw < 2 * D || lda < D) return EDG_ERR_ARG;
  if (ws_bytes < edg_fc_head_bwd_workspace(B, D, C)) return EDG_ERR_WORKSPACE;
  const int blocks = (B + kHeadGraphs - 1) / kHeadGraphs;
  float* partial = reinterpret_cast<float*>(ws);
  float* dlg_part = partial + (size_t)blocks * C * (2 * (size_t)D + 1);
  const int vec4 = ((D & 3) == 0 && (ldw & 3) == 0 && (lda & 3) == 0 && aligned16(fc_w) && aligned16(a) && aligned16(dv)) ? 1 : 0;
  auto launch = [&](auto kern, int CP) -> int {
    const size_t smem = (size_t)(CP + kHeadGraphs) * kHeadPitch * sizeof(float);
    if (int rc_ = ensure_dyn_smem((const void*)kern, smem)) return rc_;
    kern<<<dim3(blocks, 2), kHeadThreads, smem, s>>>(logits, ldl, fc_w, ldw, fc_b, a, lda, dv, dc, scale, B, D, C, vec4, parts,
                                                      dlg_part, d_a, partial);
    return check_launch();
  };
  int rc;
  if (C <= 8) rc = launch(fc_head_bwd_kernel<8>, 8);
  else if (C <= 36) rc = launch(fc_head_bwd_kernel<36>, 36);
  else rc = launch(fc_head_bwd_kernel<64>, 64);
  if (rc) return rc;
  const int64_t per = (int64_t)C * (2 * D + 1), BC = (int64_t)B * C;
  // parameter-gradient slabs first, the d logits halves after them: launch only the part of the grid that has work
  const int64_t lo = want_par ? 0 : per, hi = want_in ? per + BC : per;
  const unsigned first = (unsigned)(lo / 256), nblk = (unsigned)((hi + 255) / 256) - first;
  fc_head_reduce_kernel<<<nblk, 256, 0, s>>>(partial, blocks, C, 2 * D, d_fc_w, lddw, d_fc_b, dlg_part, BC, B, d_logits, lddl,
                                             parts, first);
  return check_launch();
}
