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

constexpr int kHdThreads = 512;
constexpr int kHdMaxD = 320;
constexpr int kHdCtasPerSm = 2;

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

// Shared memory (dynamic): the tile of du rows in bf16 with the GLOBAL row pitch (so that it leaves as one contiguous
// block of 16-byte stores), the per-sentence vectors, and the tile's CSR slice.
struct HeadDuSmem {
  float q[kFMaxSent][kHdMaxD];          // gate o v
  float pg[kFMaxSent][kHdMaxD];         // gate o g_pooled / (deg_arg + 1)
  uint8_t arg[kFMaxSent][kHdMaxD];      // tile-local arg-max row (0xff = none)
  uint16_t rp[kFRows + 8];
  uint8_t cl[kFMaxNnz];
  float ds[kFRows], dsa[kFRows], dsagg[kFRows], dsum[kFMaxSent];
  uint8_t srow[kFRows];                 // sentence (tile-local) of each row
  uint8_t sfirst[kFMaxSent + 8];
};

// One tile = a run of whole sentences (edg_tile_plan).  Phases (block barriers in between):
//   0  CSR slice, ds_t, sentence starts                      (coalesced global loads)
//   1  per row: ds_t / (deg_t + 1), sentence id
//   2  per row: dsagg_i = sum_{j in N(i)} ds_j / (deg_j + 1);  per (sentence, column): q = g o v, pg, arg, dgate
//   3  base rows  tile[i][:] = dsagg_i * q_s(i)[:]            thread = (16-byte chunk, row slice);  bias partials
//   4  patch      tile[j][c] += pg[s][c] for j in N(arg[s][c])  thread = (sentence, column): ~3 rows each, no conflicts
//   5  tile -> global, 16 bytes per thread, contiguous
__global__ void __launch_bounds__(kHdThreads, kHdCtasPerSm)
head_du_kernel(const HeadDuParams P) {
  extern __shared__ __align__(16) uint8_t hd_smem[];
  HeadDuSmem& S = *reinterpret_cast<HeadDuSmem*>(hd_smem);
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(hd_smem + ((sizeof(HeadDuSmem) + 15) & ~(size_t)15));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = P.D;
  const int ldt = (int)P.lddu;                        // tile pitch = global pitch (elements)
  const int n_tiles = __ldg(P.n_tiles);
  const float gkl = P.g_kl ? __ldg(P.g_kl) : 0.f;
  const int nch = ldt >> 3;                           // 16-byte chunks per row (padding columns included: written as zeros)
  const int nsl = kHdThreads / nch;
  const int sl = tid / nch, k8 = tid - sl * nch;
  float dbacc = 0.f;                                  // column tid (D <= kHdThreads)
  int t = blockIdx.x;
  int4 ti0 = make_int4(0, 0, 0, 0); int2 ti1 = make_int2(0, 0);
  if (t < n_tiles) { ti0 = __ldg(reinterpret_cast<const int4*>(P.tile_info + 8 * t)); ti1 = __ldg(reinterpret_cast<const int2*>(P.tile_info + 8 * t + 4)); }
  for (; t < n_tiles; t += gridDim.x) {
    const int s0 = ti0.x, s1 = ti0.y, r0 = ti0.z, r1 = ti0.w, e0 = ti1.x, e1 = ti1.y;
    const int n = r1 - r0, ns = s1 - s0, nnz = min(e1 - e0, kFMaxNnz);
    const int tn = t + gridDim.x;                     // the next tile's header is in flight during this tile
    if (tn < n_tiles) { ti0 = __ldg(reinterpret_cast<const int4*>(P.tile_info + 8 * tn)); ti1 = __ldg(reinterpret_cast<const int2*>(P.tile_info + 8 * tn + 4)); }
    // ---- 0
    for (int i = tid; i <= n; i += kHdThreads) S.rp[i] = (uint16_t)(__ldg(P.row_ptr + r0 + i) - e0);
    for (int i = tid; i < nnz; i += kHdThreads) S.cl[i] = (uint8_t)(__ldg(P.col + e0 + i) - r0);
    for (int i = tid; i < n; i += kHdThreads) {
      float d = gkl * __ldg(P.u_unit + r0 + i);
      if (P.g_scores) d += __ldg(P.g_scores + r0 + i);
      S.ds[i] = d;
    }
    if (tid <= ns) S.sfirst[tid] = (uint8_t)(__ldg(P.sent_ptr + s0 + tid) - r0);
    __syncthreads();
    // ---- 1
    for (int i = tid; i < n; i += kHdThreads) {
      const int deg = S.rp[i + 1] - S.rp[i];
      S.dsa[i] = S.ds[i] * __frcp_rn((float)(deg + 1));
      int s = 0;
      for (int qq = 1; qq < ns; ++qq) s += (S.sfirst[qq] <= i) ? 1 : 0;
      S.srow[i] = (uint8_t)s;
    }
    __syncthreads();
    // ---- 2
    if (tid < n) {
      const int i = tid;
      float a = 0.f;
      for (int e = S.rp[i]; e < S.rp[i + 1]; ++e) a += S.dsa[S.cl[e]];
      S.dsagg[i] = a;
    } else if (warp >= kFRows / 32 && warp - kFRows / 32 < ns) {      // sum_t ds_t per sentence (fixed order)
      const int s = warp - kFRows / 32;
      float a = 0.f;
      for (int i = S.sfirst[s] + lane; i < S.sfirst[s + 1]; i += 32) a += S.ds[i];
      a = warp_sum(a);
      if (lane == 0) S.dsum[s] = a;
    }
    for (int idx = tid; idx < ns * D; idx += kHdThreads) {
      const int s = idx / D, d = idx - s * D;
      const int64_t o = (int64_t)(s0 + s) * D + d;
      const float g = __ldg(P.gate + o), vv = __ldg(P.v + o);
      const float gp = P.gp ? __ldg(P.gp + o) : 0.f;
      const int a = (P.gp && P.arg) ? __ldg(P.arg + o) - r0 : -1;
      const bool has = a >= 0 && a < n;
      S.q[s][d] = g * vv;
      S.pg[s][d] = has ? g * gp : 0.f;                 // un-scaled here (the bias gradient wants it); phase 4 scales
      S.arg[s][d] = has ? (uint8_t)a : (uint8_t)0xff;
      if (P.dgate) P.dgate[o] = fmaf(gkl * vv, __ldg(P.sf_unit + o), gp * __ldg(P.hmax + o));
    }
    if (D < ldt) {                                      // padding columns of q: the base rows write zeros there
      for (int idx = tid; idx < ns * (ldt - D); idx += kHdThreads) { const int s = idx / (ldt - D); S.q[s][D + idx - s * (ldt - D)] = 0.f; }
    }
    __syncthreads();
    // ---- 3
    if (sl < nsl) {
      int cur_s = -1;
      float qv[8];
      for (int i = sl; i < n; i += nsl) {
        const int s = S.srow[i];
        if (s != cur_s) {
          cur_s = s;
          const float4 q0 = *reinterpret_cast<const float4*>(&S.q[s][8 * k8]), q1 = *reinterpret_cast<const float4*>(&S.q[s][8 * k8 + 4]);
          qv[0] = q0.x; qv[1] = q0.y; qv[2] = q0.z; qv[3] = q0.w; qv[4] = q1.x; qv[5] = q1.y; qv[6] = q1.z; qv[7] = q1.w;
        }
        const float dsi = S.dsagg[i];
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = dsi * qv[k];
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { __nv_bfloat162 h = __floats2bfloat162_rn(o[2 * k], o[2 * k + 1]); w[k] = *reinterpret_cast<uint32_t*>(&h); }
        *reinterpret_cast<uint4*>(tile + (size_t)i * ldt + 8 * k8) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    if (tid < D) {                                      // bias gradient: thread = column, sentences in order (deterministic)
      float a = dbacc;
      for (int s = 0; s < ns; ++s) a += fmaf(S.dsum[s], S.q[s][tid], S.pg[s][tid]);
      dbacc = a;
    }
    __syncthreads();
    // ---- 4
    for (int idx = tid; idx < ns * D; idx += kHdThreads) {
      const int s = idx / D, d = idx - s * D;
      const int a = S.arg[s][d];
      if (a != 0xff) {
        const int eb = S.rp[a], ee = S.rp[a + 1];
        const float pgv = S.pg[s][d] * __frcp_rn((float)(ee - eb + 1));
        for (int e = eb; e < ee; ++e) {
          __nv_bfloat16* p = tile + (size_t)S.cl[e] * ldt + d;
          *p = __float2bfloat16_rn(__bfloat162float(*p) + pgv);
        }
      }
    }
    __syncthreads();
    // ---- 5
    {
      const uint4* src = reinterpret_cast<const uint4*>(tile);
      uint4* dst = reinterpret_cast<uint4*>(P.du + (int64_t)r0 * P.lddu);
      const int total = n * nch;
      for (int i = tid; i < total; i += kHdThreads) dst[i] = src[i];
    }
    // (the next tile's phases 0-2 touch neither `tile` nor anything phase 5 reads; its phase 3 comes after three barriers)
  }
  if (P.db_part && tid < D) P.db_part[(int64_t)blockIdx.x * D + tid] = dbacc;
}

}  // namespace edg

using namespace edg;

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
  if (D > kHdMaxD || D > kHdThreads || lddu > kHdMaxD) return EDG_ERR_UNSUPPORTED;
  if (lddu < ((D + 7) / 8) * 8 || (lddu & 7) || !aligned16(du)) return EDG_ERR_ALIGN;
  const int grid = kHdCtasPerSm * kNumSMs;
  if (dbias && ws_bytes < (size_t)grid * D * sizeof(float)) return EDG_ERR_WORKSPACE;
  HeadDuParams P;
  P.tile_info = tile_info; P.n_tiles = n_tiles; P.row_ptr = row_ptr; P.col = col; P.sent_ptr = sent_ptr;
  P.u_unit = u_unit; P.g_kl = g_kl; P.g_scores = g_scores; P.gate = gate; P.v = v; P.gp = g_pooled; P.arg = arg;
  P.sf_unit = sf_unit; P.hmax = hmax; P.du = (__nv_bfloat16*)du; P.lddu = lddu; P.dgate = dgate;
  P.db_part = dbias ? (float*)ws : nullptr; P.D = D; P.B = B;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = ((sizeof(HeadDuSmem) + 15) & ~(size_t)15) + (size_t)kFRows * lddu * sizeof(__nv_bfloat16);
  if (int rc_ = ensure_dyn_smem((const void*)head_du_kernel, smem)) return rc_;
  head_du_kernel<<<grid, kHdThreads, smem, s>>>(P);
  int rc = check_launch();
  if (rc) return rc;
  if (dbias) {
    colsum_part_reduce_kernel<<<(D + 31) / 32, 256, 0, s>>>((const float*)ws, grid, D, dbias, 0);
    rc = check_launch();
  }
  return rc;
}
