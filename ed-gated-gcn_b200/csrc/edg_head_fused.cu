// head_du : the gradient that enters the GCN chain from the top of the gated block (bert_amir5.py:639-648), produced
// directly in the form the fused layer kernel consumes and WITHOUT reading h_L or materialising dh_L:
//
//     dh_L[t,:] = ds_t (g o v_b)  +  [t == arg(b,:)] (g o gp_b)         ds_t = g_kl u_t (+ g_scores_t)
//     du_L      = A^T dh_L                                                (adjoint aggregation, gcn.py:35,41)
//     db_L      = colsum(dh_L) = sum_b [ (sum_t ds_t) (g o v_b) + g o gp_b ]
//     dgate_L   = v_b o sum_t ds_t h_t + gp_b o h_arg = g_kl (v_b o sf_b) + gp_b o hmax_b
//
// u_t = d kl / d scores_t and sf_b = sum_t u_t h_t come from the forward sweep (edg_scores_kl_fwd), hmax_b = the column
// maxima of h_L from the fused layer kernel, gp_b = d loss / d pooled_b.  dh_L is rank one per sentence plus one entry per
// (sentence, column), so  du_L[i,:] = (sum_{j in N(i)} ds_j / (deg_j+1)) (g o v_b) + [arg(b,:) in N(i)] (g o gp_b) / (deg_arg+1):
// a row-scalar aggregation, one multiply-add per element and a byte-SIMD membership test of the arg-max row against
// the row's neighbour list -- one write of the row matrix, nothing read but [N]- and [B,D]-sized arrays.
// Same sentence-aligned tiles (edg_tile_plan) and 16-byte-per-row neighbour words as edg_gcn_layer.
#include "edg_common.cuh"

namespace edg {

constexpr int kHdThreads = 256;
constexpr int kHdMaxD = 320;

struct HeadDuParams {
  const int32_t* tile_info; const int32_t* n_tiles;
  const int32_t* row_ptr; const int32_t* col; const int32_t* sent_ptr;
  const float* u_unit; const float* g_kl; const float* g_scores;   // [N], device scalar or null, [N] or null
  const float* gate; const float* v; const float* gp; const int32_t* arg;   // [B,D] each (gp / arg may be null)
  const float* sf_unit; const float* hmax;                                   // [B,D]
  __nv_bfloat16* du; int64_t lddu;
  float* dgate;                // [B,D]
  float* db_part;              // [gridDim.x][D] partial bias gradient
  int D, B;
};

__global__ void __launch_bounds__(kHdThreads, 8)
head_du_kernel(const HeadDuParams P) {
  __shared__ __align__(16) float q_s[kFMaxSent][kHdMaxD];       // g o v
  __shared__ __align__(16) float pg_s[kFMaxSent][kHdMaxD];      // g o gp / (deg_arg + 1)
  __shared__ __align__(16) uint8_t arg_s[kFMaxSent][kHdMaxD];   // tile-local arg-max row (0xff = none)
  __shared__ __align__(16) uint4 meta_s[kFRows];                // neighbour ids | deg, sentence | first CSR entry
  __shared__ uint16_t rp_s[kFRows + 8];
  __shared__ uint8_t cl_s[kFMaxNnz];
  __shared__ float ds_s[kFRows], dsa_s[kFRows], dsum_s[kFMaxSent];
  __shared__ uint8_t sfirst_s[kFMaxSent + 8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = P.D;
  const int n_tiles = __ldg(P.n_tiles);
  const float gkl = P.g_kl ? __ldg(P.g_kl) : 0.f;
  const int nch = (D + 7) >> 3;                       // 8-column chunks per row
  float dbacc[2] = {0.f, 0.f};                        // columns tid and tid + 256
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int s0 = __ldg(P.tile_info + 8 * t), s1 = __ldg(P.tile_info + 8 * t + 1);
    const int r0 = __ldg(P.tile_info + 8 * t + 2), r1 = __ldg(P.tile_info + 8 * t + 3);
    const int e0 = __ldg(P.tile_info + 8 * t + 4), e1 = __ldg(P.tile_info + 8 * t + 5);
    const int n = r1 - r0, ns = s1 - s0, nnz = min(e1 - e0, kFMaxNnz);
    __syncthreads();                                  // the previous tile's readers are done
    // ---- raw CSR slice, ds, sentence starts (coalesced; one latency round)
    for (int i = tid; i <= n; i += kHdThreads) rp_s[i] = (uint16_t)(__ldg(P.row_ptr + r0 + i) - e0);
    for (int i = tid; i < nnz; i += kHdThreads) cl_s[i] = (uint8_t)(__ldg(P.col + e0 + i) - r0);
    for (int i = tid; i < n; i += kHdThreads) {
      float d = gkl * __ldg(P.u_unit + r0 + i);
      if (P.g_scores) d += __ldg(P.g_scores + r0 + i);
      ds_s[i] = d;
    }
    if (tid <= ns) sfirst_s[tid] = (uint8_t)(__ldg(P.sent_ptr + s0 + tid) - r0);
    __syncthreads();
    // ---- per row: neighbour word (self first, unused = 0xff) and ds_i / (deg_i + 1)
    for (int i = tid; i < n; i += kHdThreads) {
      const int e_beg = rp_s[i], deg = rp_s[i + 1] - e_beg;
      uint64_t ids = (0xffffffffffffff00ull) | (uint64_t)i;
      int cnt = 0;
      for (int qq = 0; qq < 8 && qq < deg; ++qq) {
        const uint32_t j = cl_s[e_beg + qq];
        if ((int)j != i && cnt < 7) { ++cnt; ids = (ids & ~(0xffull << (8 * cnt))) | ((uint64_t)j << (8 * cnt)); }
      }
      int s = 0;
      for (int qq = 1; qq < ns; ++qq) s += (sfirst_s[qq] <= i) ? 1 : 0;
      meta_s[i] = make_uint4((uint32_t)ids, (uint32_t)(ids >> 32), (uint32_t)deg | ((uint32_t)s << 8), (uint32_t)e_beg);
      dsa_s[i] = ds_s[i] * __frcp_rn((float)(deg + 1));
    }
    // sum_t ds_t per sentence (fixed order)
    if (warp < ns) {
      const int s = warp;
      float a = 0.f;
      for (int i = sfirst_s[s] + lane; i < sfirst_s[s + 1]; i += 32) a += ds_s[i];
      a = warp_sum(a);
      if (lane == 0) dsum_s[s] = a;
    }
    __syncthreads();
    // ---- per-sentence vectors -> shared memory; dgate and the bias-gradient partials on the way
    for (int idx = tid; idx < ns * D; idx += kHdThreads) {
      const int s = idx / D, d = idx - s * D;
      const int64_t o = (int64_t)(s0 + s) * D + d;
      const float g = __ldg(P.gate + o), vv = __ldg(P.v + o);
      const float gp = P.gp ? __ldg(P.gp + o) : 0.f;
      const int a = (P.gp && P.arg) ? __ldg(P.arg + o) - r0 : -1;
      const bool has = a >= 0 && a < n;
      const float qv = g * vv, gg = g * gp;
      q_s[s][d] = qv;
      float pg = 0.f;
      if (has) pg = gg * __frcp_rn((float)(rp_s[a + 1] - rp_s[a] + 1));
      pg_s[s][d] = pg;
      arg_s[s][d] = has ? (uint8_t)a : (uint8_t)0xff;
      if (P.dgate) P.dgate[o] = fmaf(gkl * vv, __ldg(P.sf_unit + o), gp * __ldg(P.hmax + o));
    }
    // bias gradient: thread = column, sentences in order (deterministic)
    __syncthreads();
    for (int c = 0; c < 2; ++c) {
      const int d = tid + c * kHdThreads;
      if (d < D) {
        float a = dbacc[c];
        for (int s = 0; s < ns; ++s) {
          const uint8_t ar = arg_s[s][d];
          float gg = 0.f;
          if (ar != 0xff) gg = pg_s[s][d] * (float)(rp_s[ar + 1] - rp_s[ar] + 1);
          a += fmaf(dsum_s[s], q_s[s][d], gg);
        }
        dbacc[c] = a;
      }
    }
    // ---- aggregated row scalars: dsagg_i = sum_{j in N(i)} ds_j / (deg_j + 1)   (overwrites ds_s)
    for (int i = tid; i < n; i += kHdThreads) {
      const uint4 m = meta_s[i];
      const int deg = m.z & 0xff;
      float a = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) { const uint32_t j = (m.x >> (8 * u)) & 0xff; if (j != 0xff) a += dsa_s[j]; }
#pragma unroll
      for (int u = 0; u < 4; ++u) { const uint32_t j = (m.y >> (8 * u)) & 0xff; if (j != 0xff) a += dsa_s[j]; }
      if (deg > 8) {
        int cnt = 0;
        for (int e = 0; e < deg; ++e) {
          const int j = cl_s[m.w + e];
          if (j == i) continue;
          if (cnt++ < 7) continue;
          a += dsa_s[j];
        }
      }
      ds_s[i] = a;
    }
    __syncthreads();
    // ---- rows: thread = (row, 8-column chunk)
    for (int task = tid; task < n * nch; task += kHdThreads) {
      const int i = task / nch, k8 = task - i * nch;
      const uint4 m = meta_s[i];
      const int deg = m.z & 0xff, s = (m.z >> 8) & 0xff;
      const float dsi = ds_s[i];
      const float4 q0 = *reinterpret_cast<const float4*>(&q_s[s][8 * k8]), q1 = *reinterpret_cast<const float4*>(&q_s[s][8 * k8 + 4]);
      const float4 p0 = *reinterpret_cast<const float4*>(&pg_s[s][8 * k8]), p1 = *reinterpret_cast<const float4*>(&pg_s[s][8 * k8 + 4]);
      const uint2 ab = *reinterpret_cast<const uint2*>(&arg_s[s][8 * k8]);
      const float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
      const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t a = ((k < 4 ? ab.x : ab.y) >> (8 * (k & 3))) & 0xffu;
        const uint32_t rep = a * 0x01010101u;
        bool hit = (__vcmpeq4(m.x, rep) | __vcmpeq4(m.y, rep)) != 0u;       // arg-max row among the first 8 entries of N(i)
        if (deg > 8 && !hit && a != 0xffu) {
          int cnt = 0;
          for (int e = 0; e < deg; ++e) {
            const uint32_t j = cl_s[m.w + e];
            if ((int)j == i) continue;
            if (cnt++ < 7) continue;
            hit |= (j == a);
          }
        }
        o[k] = fmaf(dsi, qv[k], (hit && a != 0xffu) ? pv[k] : 0.f);
        if (8 * k8 + k >= D) o[k] = 0.f;
      }
      Vec16<__nv_bfloat16>::store(P.du + (int64_t)(r0 + i) * P.lddu + 8 * k8, o);
    }
  }
  if (P.db_part) {
    for (int c = 0; c < 2; ++c) {
      const int d = tid + c * kHdThreads;
      if (d < D) P.db_part[(int64_t)blockIdx.x * D + d] = dbacc[c];
    }
  }
}

}  // namespace edg

using namespace edg;

constexpr int kHdCtasPerSm = 8;      // latency-bound phases (five dependent rounds per tile): occupancy hides them
extern "C" size_t edg_head_du_workspace(int32_t D) { return (size_t)kHdCtasPerSm * kNumSMs * (size_t)D * sizeof(float); }

/* see include/edgcn.h */
extern "C" int edg_head_du(const float* u_unit, const float* g_kl, const float* g_scores, const float* gate, const float* v,
                           const float* g_pooled, const int32_t* arg, const float* sf_unit, const float* hmax, int32_t N,
                           int32_t B, int32_t D, const int32_t* row_ptr, const int32_t* col, const int32_t* sent_ptr,
                           const int32_t* tile_info, const int32_t* n_tiles, void* du, int64_t lddu, float* dgate,
                           float* dbias, void* ws, size_t ws_bytes, edg_stream stream) {
  if (N < 0 || B < 0 || D <= 0) return EDG_ERR_ARG;
  if (N == 0 || B == 0) return EDG_OK;
  if (!u_unit || !gate || !v || !row_ptr || !col || !sent_ptr || !tile_info || !n_tiles || !du) return EDG_ERR_ARG;
  if (g_pooled && !arg) return EDG_ERR_ARG;
  if (dgate && (!sf_unit || !hmax)) return EDG_ERR_ARG;
  if (D > kHdMaxD || D > 2 * kHdThreads) return EDG_ERR_UNSUPPORTED;
  if (lddu < ((D + 7) / 8) * 8 || (lddu & 7) || !aligned16(du)) return EDG_ERR_ALIGN;
  const int grid = kHdCtasPerSm * kNumSMs;
  if (dbias && ws_bytes < (size_t)grid * D * sizeof(float)) return EDG_ERR_WORKSPACE;
  HeadDuParams P;
  P.tile_info = tile_info; P.n_tiles = n_tiles; P.row_ptr = row_ptr; P.col = col; P.sent_ptr = sent_ptr;
  P.u_unit = u_unit; P.g_kl = g_kl; P.g_scores = g_scores; P.gate = gate; P.v = v; P.gp = g_pooled; P.arg = arg;
  P.sf_unit = sf_unit; P.hmax = hmax; P.du = (__nv_bfloat16*)du; P.lddu = lddu; P.dgate = dgate;
  P.db_part = dbias ? (float*)ws : nullptr; P.D = D; P.B = B;
  cudaStream_t s = (cudaStream_t)stream;
  head_du_kernel<<<grid, kHdThreads, 0, s>>>(P);
  int rc = check_launch();
  if (rc) return rc;
  if (dbias) {
    colsum_part_reduce_kernel<<<(D + 255) / 256, 256, 0, s>>>((const float*)ws, grid, D, dbias, 0);
    rc = check_launch();
  }
  return rc;
}
