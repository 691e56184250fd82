// Sentence-window (shared-memory staged) versions of the row-streaming kernels of the gated block:
// gated max-pool views, importance scores + softmax product, and the head backward.  Same arithmetic as
// the per-sentence kernels in edg_block.cu (which remain for sentences too long for a window); the rows
// come from ONE bulk async copy per window instead of a dependent global load per row.
#include <math.h>

#include "edg_staged.cuh"

namespace edg {

template <int I64> __device__ __forceinline__ float sdist_at(const void* dist, int64_t i) {
  return I64 ? (float)__ldg(reinterpret_cast<const long long*>(dist) + i) : (float)__ldg(reinterpret_cast<const int32_t*>(dist) + i);
}

// ---- gated max-pool views (bert_amir5.py:627-636, :640) -------------------------------------------
// thread = (16-byte column chunk x, sentence lane y): the sentences of the window are dealt to the y lanes and
// each thread walks its sentence's rows in shared memory (measured faster than splitting one sentence's rows
// over the lanes and combining through shared memory: no block barriers on the critical path).
template <typename T, int V>
__global__ void __launch_bounds__(256)
pool_staged_kernel(const T* __restrict__ h, int64_t ldh, const int32_t* __restrict__ sent_ptr,
                   const int32_t* __restrict__ row_sent, int N, int B, int D, int tile_rows, int cap_rows,
                   const float* __restrict__ gates, float* __restrict__ pooled, int32_t* __restrict__ arg) {
  constexpr int E = Vec16<T>::kElems;
  extern __shared__ __align__(128) uint8_t win[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int sh[4];
  const RowWindow w = stage_window<T>(h, ldh, N, B, tile_rows, sent_ptr, row_sent, win, &bar, sh);
  if (w.r1 <= w.r0) return;
  const int pitch = (int)(ldh * (int64_t)sizeof(T));
  const int c = threadIdx.x * E;
  const int64_t BD = (int64_t)B * D;
  const uint32_t xs = stg_smem_u32(win) + threadIdx.x * 16;
  bool waited = false;
  for (int s = w.s0 + threadIdx.y; s < w.s1; s += blockDim.y) {
    const int beg = __ldg(sent_ptr + s), end = __ldg(sent_ptr + s + 1);
    float g[V][E], best[V][E];
    int32_t where[V][E];
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
      for (int k = 0; k < E; ++k) {
        g[v][k] = (c + k < D) ? __ldg(gates + v * BD + (int64_t)s * D + c + k) : 0.f;
        best[v][k] = -INFINITY;
        where[v][k] = beg < end ? beg : -1;
      }
    if (!waited) { wait_window(&bar); waited = true; }
    for (int t = beg; t < end; ++t) {
      float f[E];
      SVec16<T>::load(xs + (uint32_t)((t - w.r0) * pitch), f);
#pragma unroll
      for (int v = 0; v < V; ++v)
#pragma unroll
        for (int k = 0; k < E; ++k) {
          const float a = f[k] * g[v][k];
          if (a > best[v][k]) { best[v][k] = a; where[v][k] = t; }      // strict: first row wins ties
        }
    }
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
      for (int k = 0; k < E; ++k)
        if (c + k < D) {
          pooled[v * BD + (int64_t)s * D + c + k] = (beg < end) ? best[v][k] : 0.f;
          arg[v * BD + (int64_t)s * D + c + k] = where[v][k];
        }
  }
  if (!waited) wait_window(&bar);        // never leave a bulk copy in flight into a dead block
}

// ---- importance scores, softmax product, and d kl / d (v, c) units (bert_amir5.py:645-648) ----------
// blockDim = (32, 8).  All 8 warps work on ONE sentence at a time: warp y takes rows y, y+8, ... (lanes = 16-byte
// chunks, dot product by warp shuffles); partial dv sums of the 8 warps are combined through shared memory.
constexpr int kSMaxQ = 4;

template <typename T, int I64>
__global__ void __launch_bounds__(256)
scores_staged_kernel(const T* __restrict__ h, int64_t ldh, const int32_t* __restrict__ sent_ptr,
                     const int32_t* __restrict__ row_sent, int N, int B, int D, int chunks, int tile_rows, int cap_rows,
                     const float* __restrict__ gate, const float* __restrict__ vvec, const float* __restrict__ cvec,
                     const void* __restrict__ dist, float* __restrict__ scores, float* __restrict__ kl_b,
                     float* __restrict__ dv_unit, float* __restrict__ dc_unit) {
  constexpr int E = Vec16<T>::kElems;
  extern __shared__ __align__(128) uint8_t win[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int sh[4];
  const RowWindow w = stage_window<T>(h, ldh, N, B, tile_rows, sent_ptr, row_sent, win, &bar, sh);
  if (w.r1 <= w.r0) return;
  const int pitch = (int)(ldh * (int64_t)sizeof(T));
  float* sc_s = reinterpret_cast<float*>(win + (size_t)cap_rows * pitch);      // [cap_rows] scores, then u_t
  float* part = sc_s + cap_rows;                                               // [8][chunks*E] partial dv sums
  const int lane = threadIdx.x, warp = threadIdx.y, tid = warp * 32 + lane;
  const int width = chunks * E;
  const uint32_t xs = stg_smem_u32(win);
  wait_window(&bar);
  for (int s = w.s0; s < w.s1; ++s) {
    const int beg = __ldg(sent_ptr + s), end = __ldg(sent_ptr + s + 1);
    float gq[kSMaxQ][E], wq[kSMaxQ][E];
#pragma unroll
    for (int q = 0; q < kSMaxQ; ++q) {
      const int c = (lane + 32 * q) * E;
#pragma unroll
      for (int k = 0; k < E; ++k) {
        const bool ok = (lane + 32 * q < chunks) && (c + k < D);
        gq[q][k] = ok ? __ldg(gate + (int64_t)s * D + c + k) : 0.f;
        wq[q][k] = ok ? gq[q][k] * __ldg(vvec + (int64_t)s * D + c + k) : 0.f;
      }
    }
    const float cb = cvec ? __ldg(cvec + s) : 0.f;
    // sweep 1: scores (warp per row)
    for (int t = beg + warp; t < end; t += 8) {
      float acc[kSMaxQ];
#pragma unroll
      for (int q = 0; q < kSMaxQ; ++q) {
        acc[q] = 0.f;
        if (lane + 32 * q < chunks) {
          float f[E];
          SVec16<T>::load(xs + (uint32_t)((t - w.r0) * pitch + (lane + 32 * q) * 16), f);
#pragma unroll
          for (int k = 0; k < E; ++k) acc[q] = fmaf(f[k], wq[q][k], acc[q]);
        }
      }
      float tot = (acc[0] + acc[1]) + (acc[2] + acc[3]);
      tot = warp_sum(tot) + cb;
      if (lane == 0) { sc_s[t - w.r0] = tot; scores[t] = tot; }
    }
    __syncthreads();
    // softmax statistics of scores and float(dist) over the sentence (every warp, redundantly: n is small)
    float ms = -INFINITY, mq = -INFINITY;
    for (int t = beg + lane; t < end; t += 32) { ms = fmaxf(ms, sc_s[t - w.r0]); mq = fmaxf(mq, sdist_at<I64>(dist, t)); }
    ms = warp_max(ms); mq = warp_max(mq);
    float zs = 0.f, zq = 0.f;
    for (int t = beg + lane; t < end; t += 32) { zs += expf(sc_s[t - w.r0] - ms); zq += expf(sdist_at<I64>(dist, t) - mq); }
    zs = warp_sum(zs); zq = warp_sum(zq);
    float kacc = 0.f;
    for (int t = beg + lane; t < end; t += 32)
      kacc += (expf(sc_s[t - w.r0] - ms) / zs) * (expf(sdist_at<I64>(dist, t) - mq) / zq);
    const float klb = warp_sum(kacc);
    if (tid == 0) kl_b[s] = klb;
    if (dv_unit == nullptr) { __syncthreads(); continue; }
    __syncthreads();                       // everyone has read the scores: overwrite them with u_t
    const float invB = 1.0f / (float)B;
    for (int t = beg + tid; t < end; t += 256)
      sc_s[t - w.r0] = (expf(sc_s[t - w.r0] - ms) / zs) * ((expf(sdist_at<I64>(dist, t) - mq) / zq) - klb) * invB;
    __syncthreads();
    // sweep 2: dv_unit[d] = gate[d] * sum_t u_t h[t,d]; warp y sums its rows, then the 8 partials are combined
    float dvu[kSMaxQ][E];
#pragma unroll
    for (int q = 0; q < kSMaxQ; ++q)
#pragma unroll
      for (int k = 0; k < E; ++k) dvu[q][k] = 0.f;
    for (int t = beg + warp; t < end; t += 8) {
      const float u = sc_s[t - w.r0];
#pragma unroll
      for (int q = 0; q < kSMaxQ; ++q)
        if (lane + 32 * q < chunks) {
          float f[E];
          SVec16<T>::load(xs + (uint32_t)((t - w.r0) * pitch + (lane + 32 * q) * 16), f);
#pragma unroll
          for (int k = 0; k < E; ++k) dvu[q][k] = fmaf(u, f[k], dvu[q][k]);
        }
    }
#pragma unroll
    for (int q = 0; q < kSMaxQ; ++q)
      if (lane + 32 * q < chunks) {
#pragma unroll
        for (int k = 0; k < E; ++k) part[warp * width + k * chunks + (lane + 32 * q)] = dvu[q][k] * gq[q][k];
      }
    __syncthreads();
    for (int j = tid; j < width; j += 256)
      if (j < D) {
        float tot = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) tot += part[y * width + (j % E) * chunks + j / E];
        dv_unit[(int64_t)s * D + j] = tot;
      }
    if (warp == 0 && dc_unit) {
      float dcu = 0.f;
      for (int t = beg + lane; t < end; t += 32) dcu += sc_s[t - w.r0];
      dcu = warp_sum(dcu);
      if (lane == 0) dc_unit[s] = dcu;
    }
    __syncthreads();
  }
}

// ---- head backward (scores/kl, final max-pool, optional direct x_out gradient) -----------------------
template <typename T, int I64>
__global__ void __launch_bounds__(256)
head_bwd_staged_kernel(const T* __restrict__ h, int64_t ldh, const int32_t* __restrict__ sent_ptr,
                       const int32_t* __restrict__ row_sent, int N, int B, int D, int tile_rows, int cap_rows,
                       const float* __restrict__ gate, const float* __restrict__ vvec, const void* __restrict__ dist,
                       const float* __restrict__ scores, const float* __restrict__ kl_b, const float* __restrict__ g_kl,
                       const float* __restrict__ g_scores, const float* __restrict__ g_pooled, const int32_t* __restrict__ arg,
                       const T* __restrict__ g_xout, int64_t ldgx, T* __restrict__ dh, int64_t lddh,
                       float* __restrict__ dgate, float* __restrict__ dv, float* __restrict__ dc) {
  constexpr int E = Vec16<T>::kElems;
  extern __shared__ __align__(128) uint8_t win[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int sh[4];
  const RowWindow w = stage_window<T>(h, ldh, N, B, tile_rows, sent_ptr, row_sent, win, &bar, sh);
  if (w.r1 <= w.r0) return;
  const int pitch = (int)(ldh * (int64_t)sizeof(T));
  float* ds_s = reinterpret_cast<float*>(win + (size_t)cap_rows * pitch);      // [cap_rows] d loss / d scores
  const int nthreads = blockDim.x * blockDim.y;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int lane = tid & 31, wid = tid >> 5, nwarps = (nthreads + 31) >> 5;
  const bool have_s = (scores != nullptr) && (g_kl != nullptr || g_scores != nullptr);
  // phase 1 (overlaps the bulk copy): ds for every row of the window, one warp per sentence
  for (int s = w.s0 + wid; s < w.s1; s += nwarps) {
    const int beg = __ldg(sent_ptr + s), end = __ldg(sent_ptr + s + 1);
    float tot = 0.f;
    if (have_s) {
      float ms = 0.f, zs = 1.f, mq = 0.f, zq = 1.f, klb = 0.f, gk = 0.f;
      if (g_kl) {
        gk = __ldg(g_kl) / (float)B;
        klb = __ldg(kl_b + s);
        ms = -INFINITY; mq = -INFINITY;
        for (int t = beg + lane; t < end; t += 32) { ms = fmaxf(ms, __ldg(scores + t)); mq = fmaxf(mq, sdist_at<I64>(dist, t)); }
        ms = warp_max(ms); mq = warp_max(mq);
        zs = 0.f; zq = 0.f;
        for (int t = beg + lane; t < end; t += 32) { zs += expf(__ldg(scores + t) - ms); zq += expf(sdist_at<I64>(dist, t) - mq); }
        zs = warp_sum(zs); zq = warp_sum(zq);
      }
      for (int t = beg + lane; t < end; t += 32) {
        float v = 0.f;
        if (g_kl) v = gk * (expf(__ldg(scores + t) - ms) / zs) * ((expf(sdist_at<I64>(dist, t) - mq) / zq) - klb);
        if (g_scores) v += __ldg(g_scores + t);
        ds_s[t - w.r0] = v;
        tot += v;
      }
      tot = warp_sum(tot);
    } else {
      for (int t = beg + lane; t < end; t += 32) ds_s[t - w.r0] = 0.f;
    }
    if (lane == 0 && dc) dc[s] = tot;
  }
  __syncthreads();
  wait_window(&bar);
  // phase 2: thread = (16-byte column chunk x, sentence lane y); the sentence's vectors stay in registers
  const int c = threadIdx.x * E;
  const uint32_t xs = stg_smem_u32(win) + threadIdx.x * 16;
  for (int s = w.s0 + threadIdx.y; s < w.s1; s += blockDim.y) {
    const int beg = __ldg(sent_ptr + s), end = __ldg(sent_ptr + s + 1);
    float g[E], vv[E], gp[E], ag[E], av[E];
    int32_t where[E];
#pragma unroll
    for (int k = 0; k < E; ++k) {
      const bool ok = c + k < D;
      const int64_t o = (int64_t)s * D + c + k;
      g[k] = ok ? __ldg(gate + o) : 0.f;
      vv[k] = (ok && vvec) ? __ldg(vvec + o) : 0.f;
      gp[k] = (ok && g_pooled) ? __ldg(g_pooled + o) : 0.f;
      where[k] = (ok && g_pooled) ? __ldg(arg + o) : -1;
      ag[k] = 0.f; av[k] = 0.f;
    }
    for (int t = beg; t < end; ++t) {
      float f[E], gx[E], o[E];
      SVec16<T>::load(xs + (uint32_t)((t - w.r0) * pitch), f);
      if (g_xout) Vec16<T>::load(g_xout + (int64_t)t * ldgx + c, gx);
      const float dst = ds_s[t - w.r0];
#pragma unroll
      for (int k = 0; k < E; ++k) {
        float coef = dst * vv[k];
        if (where[k] == t) coef += gp[k];
        if (g_xout) coef += gx[k];
        o[k] = g[k] * coef;
        ag[k] = fmaf(f[k], coef, ag[k]);
        av[k] = fmaf(dst * f[k], g[k], av[k]);
      }
      if (dh) Vec16<T>::store(dh + (int64_t)t * lddh + c, o);
    }
#pragma unroll
    for (int k = 0; k < E; ++k)
      if (c + k < D) {
        const int64_t o = (int64_t)s * D + c + k;
        if (dgate) dgate[o] = ag[k];
        if (dv) dv[o] = av[k];
      }
  }
}

// ---- host launchers (called from the C-ABI functions in edg_block.cu) --------------------------------
template <typename K> static int opt_in_smem(K kernel, size_t smem, size_t* seen) {
  if (smem > *seen) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return check_launch();
    *seen = smem;
  }
  return EDG_OK;
}

static inline dim3 chunk_block(int chunks) {
  int y = 256 / chunks;
  if (y > 8) y = 8;
  if (y < 1) y = 1;
  return dim3(chunks, y);
}

// returns 1 when the staged path does not apply (caller falls back to the per-sentence kernels)
template <typename T>
int pool_fwd_staged(const void* h, int64_t ldh, const int32_t* sent_ptr, const int32_t* row_sent, int N, int B, int D,
                    int max_len, const float* gates, int V, float* pooled, int32_t* arg, cudaStream_t s) {
  constexpr int E = Vec16<T>::kElems;
  const int chunks = (D + E - 1) / E;
  if (!row_sent || max_len <= 0 || chunks > 256) return 1;
  const dim3 blk = chunk_block(chunks);
  const WindowPlan p = plan_window((size_t)ldh * sizeof(T), max_len);
  if (p.tile_rows < 8) return 1;
  const unsigned blocks = (unsigned)((N + p.tile_rows - 1) / p.tile_rows);
  const int64_t BD = (int64_t)B * D;
  static size_t seen[4] = {0, 0, 0, 0};
  for (int v0 = 0; v0 < V; v0 += 2) {
    const int nv = (V - v0) < 2 ? (V - v0) : 2;
    const size_t smem = p.smem_rows;
    const float* g = gates + v0 * BD;
    float* pp = pooled + v0 * BD;
    int32_t* aa = arg + v0 * BD;
    int rc = EDG_OK;
#define EDG_POOL_CASE(NV)                                                                                   \
    rc = opt_in_smem(pool_staged_kernel<T, NV>, smem, &seen[NV - 1]);                                       \
    if (rc) return rc;                                                                                      \
    pool_staged_kernel<T, NV><<<blocks, blk, smem, s>>>((const T*)h, ldh, sent_ptr, row_sent, N, B, D, p.tile_rows,    \
                                                       p.cap_rows, g, pp, aa);
    switch (nv) {
      case 1: EDG_POOL_CASE(1) break;
      default: EDG_POOL_CASE(2) break;
    }
#undef EDG_POOL_CASE
  }
  return check_launch();
}

template <typename T>
int scores_kl_staged(const void* h, int64_t ldh, const int32_t* sent_ptr, const int32_t* row_sent, int N, int B, int D,
                     int max_len, const float* gate, const float* v, const float* c, const void* dist, int dist_i64,
                     float* scores, float* kl_b, float* dv_unit, float* dc_unit, cudaStream_t s) {
  constexpr int E = Vec16<T>::kElems;
  const int chunks = (D + E - 1) / E;
  if (!row_sent || max_len <= 0 || chunks > 32 * kSMaxQ) return 1;
  const size_t part_bytes = (size_t)8 * chunks * E * sizeof(float);
  const WindowPlan p = plan_window((size_t)ldh * sizeof(T), max_len, 4, part_bytes);
  if (p.tile_rows < 8) return 1;
  const size_t smem = p.smem_rows + (size_t)p.cap_rows * sizeof(float) + part_bytes;
  const unsigned blocks = (unsigned)((N + p.tile_rows - 1) / p.tile_rows);
  static size_t seen[2] = {0, 0};
  int rc;
  if (dist_i64) {
    rc = opt_in_smem(scores_staged_kernel<T, 1>, smem, &seen[1]);
    if (rc) return rc;
    scores_staged_kernel<T, 1><<<blocks, dim3(32, 8), smem, s>>>((const T*)h, ldh, sent_ptr, row_sent, N, B, D, chunks, p.tile_rows,
                                                                p.cap_rows, gate, v, c, dist, scores, kl_b, dv_unit, dc_unit);
  } else {
    rc = opt_in_smem(scores_staged_kernel<T, 0>, smem, &seen[0]);
    if (rc) return rc;
    scores_staged_kernel<T, 0><<<blocks, dim3(32, 8), smem, s>>>((const T*)h, ldh, sent_ptr, row_sent, N, B, D, chunks, p.tile_rows,
                                                                p.cap_rows, gate, v, c, dist, scores, kl_b, dv_unit, dc_unit);
  }
  return check_launch();
}

template <typename T>
int head_bwd_staged(const void* h, int64_t ldh, const int32_t* sent_ptr, const int32_t* row_sent, int N, int B, int D,
                    int max_len, const float* gate, const float* v, const void* dist, int dist_i64, const float* scores,
                    const float* kl_b, const float* g_kl, const float* g_scores, const float* g_pooled, const int32_t* arg,
                    const void* g_xout, int64_t ldgx, void* dh, int64_t lddh, float* dgate, float* dv, float* dc,
                    cudaStream_t s) {
  constexpr int E = Vec16<T>::kElems;
  const int chunks = (D + E - 1) / E;
  if (!row_sent || max_len <= 0 || chunks > 256) return 1;
  const dim3 blk = chunk_block(chunks);
  const WindowPlan p = plan_window((size_t)ldh * sizeof(T), max_len, 4);
  if (p.tile_rows < 8) return 1;
  const size_t smem = p.smem_rows + (size_t)p.cap_rows * sizeof(float);
  const unsigned blocks = (unsigned)((N + p.tile_rows - 1) / p.tile_rows);
  static size_t seen[2] = {0, 0};
  int rc;
  if (dist_i64) {
    rc = opt_in_smem(head_bwd_staged_kernel<T, 1>, smem, &seen[1]);
    if (rc) return rc;
    head_bwd_staged_kernel<T, 1><<<blocks, chunk_block(chunks), smem, s>>>(
        (const T*)h, ldh, sent_ptr, row_sent, N, B, D, p.tile_rows, p.cap_rows, gate, v, dist, scores, kl_b, g_kl, g_scores,
        g_pooled, arg, (const T*)g_xout, ldgx, (T*)dh, lddh, dgate, dv, dc);
  } else {
    rc = opt_in_smem(head_bwd_staged_kernel<T, 0>, smem, &seen[0]);
    if (rc) return rc;
    head_bwd_staged_kernel<T, 0><<<blocks, chunk_block(chunks), smem, s>>>(
        (const T*)h, ldh, sent_ptr, row_sent, N, B, D, p.tile_rows, p.cap_rows, gate, v, dist, scores, kl_b, g_kl, g_scores,
        g_pooled, arg, (const T*)g_xout, ldgx, (T*)dh, lddh, dgate, dv, dc);
  }
  return check_launch();
}

}  // namespace edg
