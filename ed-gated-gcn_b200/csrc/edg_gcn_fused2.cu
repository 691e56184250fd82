// gcn_layer2 : the fused graph-convolution layer (see edg_gcn_fused.cu for the algebra and the tile plan) with the
// TREE AGGREGATION ON THE TENSOR CORES as well.  Per tile of whole sentences (<= 127 rows):
//
//     U  = X W^T                       tcgen05.mma, A = activation tile (TMA, SW128 K-major), B = stationary weight slice
//     S  = bf16(U)  (adjoint: D^-1 U, + views patch)   accumulator -> registers -> shared memory, written directly in
//                                      the canonical MN-major SWIZZLE_64B layout of a UMMA B operand
//     Y  = Adj S                       second tcgen05.mma, A = the tile's 0/1 adjacency (self loops included) as a
//                                      [128 x 128] bf16 matrix built in shared memory from the CSR slice, K = rows of the tile
//     y  = D^-1 Y + b  (forward)       accumulator -> registers -> bf16 -> global; max-pool from a staged copy
//
// 0/1 entries are exact in bf16 and the accumulation is fp32, so Y is the exact sum of the staged bf16 rows (the
// CUDA-core version added them in fp32 too).  What this buys: the epilogue no longer gathers -- ~100 instructions per
// 16-byte output chunk became ~40 -- and the kernel was bound by the issue slots of its 8 epilogue warps.
// The bias gradient (column sums of the un-aggregated rows, adjoint mode) is one more row of the same product: row 127
// of Adj holds (deg_j + 1), so accumulator lane 127 is sum_j (deg_j + 1) S_j.
//
//   warp 0      TMA producer (activation K-block ring; once, the CTA's stationary weight slice)
//   warp 1      tcgen05.mma issuer of U (3 TMEM accumulator stages of 160 columns)
//   warp 2      tcgen05.mma issuer of Y = Adj S (into the SAME TMEM columns as the tile's U, which S has left by then)
//   warp 3      builds Adj (+ per-row 1/(deg+1), sentence ids) for the NEXT tile from the global CSR
//   warps 4-11  epilogue
//
// Unity build: included after edg_gcn_fused.cu.
#include "edg_common.cuh"

namespace edg {

constexpr int kF2UBytes = 5 * 8192;        // S staging: 5 column atoms (32 columns each) x 128 rows x 64 bytes
constexpr int kF2AdjBytes = 2 * 16384;     // Adj: 2 K blocks x [128 rows x 128 bytes]
constexpr int kF2MetaBytes = 512 + 128 + 16 + 16;   // inv_deg f32[128] | sentence u8[128] | sent_first u8[16] | hdr int[4]
constexpr int kF2VecBytes = 160 * 4;       // bias (forward) / column-sum accumulators (adjoint)

struct GcnLayer2Params {
  CUtensorMap map_a;                    // [N, K] bf16 rows, box [32 rows x 64 cols], SW128
  CUtensorMap map_w;                    // [Nout, K] bf16, box [16 x 64], SW128
  const int32_t* tile_info; const int32_t* n_tiles;
  const int32_t* row_ptr; const int32_t* col; const int32_t* sent_ptr;
  const uint4* row_meta;                // [N][2] neighbour words (edg_row_meta)
  const float* bias;
  __nv_bfloat16* y; int64_t ldy;
  float* hmax; int32_t* harg; int64_t ldpool;
  const float* patch_val; const int32_t* patch_arg; int64_t ldpatch;
  float* colsum_part;
  long long* trace;                     // bring-up (EDG_FUSED_DEBUG & 32): clock64 timelines of CTA 0 (epilogue | MMA2 | builder)
  int K, Nout, num_kb, mode, n_split, stages, rows_cap, debug;
  int bn[2];
  uint32_t idesc[2], idesc2[2][2];      // idesc2[half][column group]
  int gh;                               // 16-column granules in column group 0 (the aggregation product runs per group)
  int opitch, osw_shift, osw_mask;      // pool staging: row pitch in bytes and the chunk swizzle  c ^ ((row >> shift) & mask)
  uint32_t off_a, off_u, off_adj, off_tab, off_meta, off_vec, off_bar;
};

__device__ __forceinline__ void f2_zero16(uint32_t addr) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
}

#define F2_TRACE(region, var) do { if (P.trace && blockIdx.x == 0 && var < 160) P.trace[(region) * 160 + var++] = clock64(); } while (0)

template <int kF2EpiWarps>
__global__ void __launch_bounds__((4 + kF2EpiWarps) * 32, 1)
gcn_layer2_kernel(const __grid_constant__ GcnLayer2Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int kEpiThreads = kF2EpiWarps * 32;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = (P.n_split == 2) ? (int)(blockIdx.x & 1) : 0;
  const int t_first = (int)blockIdx.x / P.n_split, t_step = (int)gridDim.x / P.n_split;
  const int n0 = half * P.bn[0];
  const int bn = P.bn[half];
  const int num_kb = P.num_kb;
  const uint32_t w_kb_bytes = (uint32_t)bn * 128u;
  auto Wblk = [&](int kb) { return base + (uint32_t)kb * w_kb_bytes; };
  auto Ablk = [&](int s) { return base + P.off_a + (uint32_t)s * 16384u; };
  const uint32_t u_s = base + P.off_u, adj_s = base + P.off_adj;
  const uint32_t bars = base + P.off_bar;
  auto full = [&](int s) { return bars + 8 * s; };
  auto empty = [&](int s) { return bars + 8 * (kFMaxStages + s); };
  auto tfull = [&](int s) { return bars + 8 * (2 * kFMaxStages + s); };
  auto tempty = [&](int s) { return bars + 8 * (2 * kFMaxStages + kFAccStages + s); };
  const uint32_t bx = bars + 8 * (2 * kFMaxStages + 2 * kFAccStages);
  const uint32_t wfull = bx, afull = bx + 24, aempty = bx + 32;
  auto ufull = [&](int g) { return g ? bx + 80 : bx + 8; };
  auto yfull = [&](int g) { return g ? bx + 88 : bx + 16; };
  auto mfull = [&](int b) { return bx + 40 + 8 * b; };
  auto mempty = [&](int b) { return bx + 56 + 8 * b; };
  const uint32_t tmem_slot = bx + 72;
  auto meta = [&](int b) { return sm + P.off_meta + b * kF2MetaBytes; };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&P.map_a);
    tma_prefetch_desc(&P.map_w);
    for (int s = 0; s < P.stages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int s = 0; s < kFAccStages; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), kF2EpiWarps); }
    mbar_init(wfull, 1); mbar_init(afull, 1); mbar_init(aempty, 1);
    for (int g = 0; g < 2; ++g) { mbar_init(ufull(g), kF2EpiWarps); mbar_init(yfull(g), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(mfull(b), 1); mbar_init(mempty(b), kF2EpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const int n_tiles = __ldg(P.n_tiles);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(wfull, (uint32_t)num_kb * w_kb_bytes);
      for (int kb = 0; kb < num_kb; ++kb)
        for (int r = 0; r < bn; r += 16) tma_load_2d(Wblk(kb) + (uint32_t)r * 128u, &P.map_w, wfull, kb * 64, n0 + r);
      int stage = 0; uint32_t phase = 0;
      int2 rr = make_int2(0, 0);
      if (t_first < n_tiles) rr = __ldg(reinterpret_cast<const int2*>(P.tile_info + 8 * t_first + 2));
      for (int t = t_first; t < n_tiles; t += t_step) {
        const int r0 = rr.x, r1 = rr.y;
        if (t + t_step < n_tiles) rr = __ldg(reinterpret_cast<const int2*>(P.tile_info + 8 * (t + t_step) + 2));
        const int nb32 = (r1 - r0 + 31) >> 5;
        for (int kb = 0; kb < num_kb; ++kb) {
          f2_wait(empty(stage), phase ^ 1);
          mbar_expect_tx(full(stage), (uint32_t)nb32 * 4096u);
          for (int b = 0; b < nb32; ++b)
            tma_load_2d(Ablk(stage) + (uint32_t)b * 4096u, &P.map_a, full(stage), kb * 64, r0 + 32 * b);
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: U = X W^T
    if (lane == 0) {
      const uint32_t idesc = P.idesc[half];
      f2_wait(wfull, 0);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int t = t_first; t < n_tiles; t += t_step, ++it) {
        const int as = it % kFAccStages;
        const uint32_t aphase = (uint32_t)(it / kFAccStages) & 1u;
        f2_wait(tempty(as), aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * kFAccStride);
        for (int kb = 0; kb < num_kb; ++kb) {
          f2_wait(full(stage), phase);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (kb * 64 + k * 16 < P.K) {
              const uint64_t ad = make_desc_sw128(Ablk(stage) + k * 32, 16, 1024);
              const uint64_t bd = make_desc_sw128(Wblk(kb) + k * 32, 16, 1024);
              umma_bf16(d_tmem, ad, bd, idesc, (kb | k) != 0);
            }
          }
          umma_commit(empty(stage));
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull(as));
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ MMA issuer: Y = Adj S
    if (lane == 0) {
      int it = 0, tr = 0;
      for (int t = t_first; t < n_tiles; t += t_step, ++it) {
        const int as = it % kFAccStages;
        const uint32_t ph = (uint32_t)it & 1u;
        F2_TRACE(1, tr);
        f2_wait(afull, ph);
        F2_TRACE(1, tr);
        const int n = reinterpret_cast<const volatile int32_t*>(meta(it & 1) + 656)[1];
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * kFAccStride);
        const int nk = (n + 15) >> 4;
        const int ngran_ = bn >> 4;
        for (int grp = 0; grp < 2; ++grp) {                  // column groups: the product of group 0 runs under phase A of group 1
          const int g0 = grp ? P.gh : 0, g1 = grp ? ngran_ : min(P.gh, ngran_);
          if (g1 <= g0) break;
          f2_wait(ufull(grp), ph);
          if (grp == 0) F2_TRACE(1, tr);
          tc_fence_after();
          for (int k = 0; k < nk; ++k) {
            const uint64_t ad = make_desc_sw128(adj_s + (uint32_t)(k >> 2) * 16384u + (uint32_t)(k & 3) * 32u, 16, 1024);
            const uint64_t bd = make_desc(u_s + (uint32_t)(g0 >> 1) * 8192u + (uint32_t)k * 1024u, 8192, 512, 4);
            umma_bf16(d_tmem + (uint32_t)(16 * g0), ad, bd, P.idesc2[half][grp], k != 0);
          }
          umma_commit(yfull(grp));
        }
        umma_commit(aempty);
        F2_TRACE(1, tr);
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ Adj builder (one tile ahead of the epilogue)
    // Reads two 16-byte words per row (edg_row_meta: sentence-local ids of self + up to 15 neighbours | degree, sentence),
    // prefetched one tile ahead in registers, the tile headers two tiles ahead: no dependent global round trip is waited
    // for inside a tile.  Adj is zeroed once; afterwards each tile CLEARS the entries of the tile before (their words are
    // still in registers) and sets its own -- a few 2-byte stores per row instead of 32 KB of zeros.
    auto ld_info = [&](int t) { return (t < n_tiles) ? __ldg(reinterpret_cast<const int4*>(P.tile_info + 8 * t)) : make_int4(0, 0, 0, 0); };
    auto ld_nb = [&](const int4& ti, int k) {
      const int i = lane + 32 * k;
      return (i < ti.w - ti.z) ? __ldg(P.row_meta + 2 * (int64_t)(ti.z + i)) : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    };
    auto ld_ds = [&](const int4& ti, int k) {
      const int i = lane + 32 * k;
      return (i < ti.w - ti.z) ? __ldg(reinterpret_cast<const uint2*>(P.row_meta + 2 * (int64_t)(ti.z + i) + 1)) : make_uint2(0u, 0u);
    };
    auto adj_addr = [&](int i, int j) {
      return adj_s + (uint32_t)(j >> 6) * 16384u + (uint32_t)i * 128u + (uint32_t)(((((j & 63) >> 3) ^ (i & 7)) << 4) | ((j & 7) << 1));
    };
    auto put = [&](uint32_t a, uint16_t val) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(val) : "memory"); };
    // entries of one row from its word: slot q holds a sentence-local id
    auto row_entries = [&](int i, const uint4& nb, int cnt, uint16_t val) {
      const int sbase = i - (int)(nb.x & 0xffu);               // tile-local first row of the sentence
      uint32_t w0 = nb.x, w1 = nb.y, w2 = nb.z, w3 = nb.w;     // shifted through: slot q is always the low byte of w0
#pragma unroll 1
      for (int q = 0; q < cnt; ++q) {
        put(adj_addr(i, sbase + (int)(w0 & 0xffu)), val);
        w0 = __funnelshift_r(w0, w1, 8); w1 = __funnelshift_r(w1, w2, 8); w2 = __funnelshift_r(w2, w3, 8); w3 >>= 8;
      }
    };
    for (int o = lane * 16; o < kF2AdjBytes; o += 512) f2_zero16(adj_s + (uint32_t)o);
    int4 ti_cur = ld_info(t_first), ti_nxt = ld_info(t_first + t_step);
    uint4 nb_cur[4], nb_prev[4];
    uint2 ds_cur[4];
    int cnt_prev[4], n_prev = 0;
    bool hub_prev = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) { nb_cur[k] = ld_nb(ti_cur, k); ds_cur[k] = ld_ds(ti_cur, k); nb_prev[k] = nb_cur[k]; cnt_prev[k] = 0; }
    int sf_cur = (lane <= ti_cur.y - ti_cur.x) ? __ldg(P.sent_ptr + ti_cur.x + lane) : 0;
    int it = 0, tr = 0;
    for (int t = t_first; t < n_tiles; t += t_step, ++it) {
      const int b = it & 1;
      const int s0 = ti_cur.x, r0 = ti_cur.z, n = ti_cur.w - ti_cur.z, ns = ti_cur.y - ti_cur.x;
      // ---- prefetch: words of the next tile, header of the one after
      const int4 ti_n2 = ld_info(t + 2 * t_step);
      uint4 nb_nxt[4];
      uint2 ds_nxt[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { nb_nxt[k] = ld_nb(ti_nxt, k); ds_nxt[k] = ld_ds(ti_nxt, k); }
      const int sf_nxt = (t + t_step < n_tiles && lane <= ti_nxt.y - ti_nxt.x) ? __ldg(P.sent_ptr + ti_nxt.x + lane) : 0;
      // ---- per-row meta of the tile (double-buffered: the epilogue may still read the tile before last)
      if (lane == 0) F2_TRACE(2, tr);
      f2_wait(mempty(b), (((uint32_t)it >> 1) & 1u) ^ 1u);
      if (lane == 0) F2_TRACE(2, tr);
      uint8_t* mb = meta(b);
      float* invd = reinterpret_cast<float*>(mb);
      uint8_t* srow = mb + 512;
      uint8_t* sfirst = mb + 640;
      int32_t* hdr = reinterpret_cast<int32_t*>(mb + 656);
      if (lane <= ns) sfirst[lane] = (uint8_t)(sf_cur - r0);
      bool hub = false;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = lane + 32 * k;
        const int deg = (int)ds_cur[k].x;
        invd[i] = (i < n) ? __frcp_rn((float)(deg + 1)) : 0.f;
        srow[i] = (i < n) ? (uint8_t)((int)ds_cur[k].y - s0) : (uint8_t)0;
        hub |= (i < n) && deg > 16;
      }
      hub = __any_sync(0xffffffffu, hub);
      if (lane == 0) { hdr[0] = r0; hdr[1] = n; hdr[2] = ns; hdr[3] = s0; }
      __syncwarp();
      if (lane == 0) mbar_arrive(mfull(b));
      // ---- Adj of this tile
      f2_wait(aempty, ((uint32_t)it & 1u) ^ 1u);
      if (lane == 0) F2_TRACE(2, tr);
      if (hub_prev) {                                          // the tile before held a row with more than 16 entries: start clean
        for (int o = lane * 16; o < kF2AdjBytes; o += 512) f2_zero16(adj_s + (uint32_t)o);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = lane + 32 * k;
          if (cnt_prev[k] > 0) row_entries(i, nb_prev[k], cnt_prev[k], 0);
          if (P.colsum_part && i >= n && i < n_prev) put(adj_addr(127, i), 0);
        }
      }
      __syncwarp();
      if (lane == 0) F2_TRACE(2, tr);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = lane + 32 * k;
        const int deg = (int)ds_cur[k].x;
        cnt_prev[k] = 0;
        if (i < n) {
          if (deg <= 16) {
            cnt_prev[k] = deg;
            row_entries(i, nb_cur[k], deg, 0x3F80);
          } else {                                             // hub: every entry from the global CSR
            const int eb = __ldg(P.row_ptr + r0 + i), ee = __ldg(P.row_ptr + r0 + i + 1);
            for (int e = eb; e < ee; ++e) {
              const int j = __ldg(P.col + e) - r0;
              if (j >= 0 && j < n) put(adj_addr(i, j), 0x3F80);
            }
          }
          if (P.colsum_part) {                 // row 127 = (deg_j + 1): accumulator lane 127 becomes the column sum
            const __nv_bfloat16 wv = __float2bfloat16_rn((float)(deg + 1));
            put(adj_addr(127, i), *reinterpret_cast<const uint16_t*>(&wv));
          }
        }
        nb_prev[k] = nb_cur[k];
      }
      hub_prev = hub; n_prev = n;
      if (lane == 0) F2_TRACE(2, tr);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(afull);
      ti_cur = ti_nxt; ti_nxt = ti_n2; sf_cur = sf_nxt;
#pragma unroll
      for (int k = 0; k < 4; ++k) { nb_cur[k] = nb_nxt[k]; ds_cur[k] = ds_nxt[k]; }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - 4;
    const int q = warp & 3;
    const int part = ew >> 2;
    constexpr int nparts = kF2EpiWarps / 4;
    const int et = threadIdx.x - 128;                       // 0 .. kEpiThreads-1
    uint32_t* tab = reinterpret_cast<uint32_t*>(sm + P.off_tab);
    float* vec = reinterpret_cast<float*>(sm + P.off_vec);  // bias (forward) or column-sum accumulators (adjoint)
    const uint32_t vec_s = base + P.off_vec;
    const int tabw = P.bn[0];
    const bool pool = P.hmax != nullptr;
    const bool patch = P.patch_val != nullptr;
    const bool fwd = P.mode == 0;
    const int ngran = bn >> 4;
    // phase D mapping: thread = (8-column chunk k8, row slice sl)
    const int nch = bn >> 3;
    const int nsl = kEpiThreads / nch;
    const int sl = et / nch, k8 = et - sl * nch;
    const bool d_active = sl < nsl;
    for (int c = et; c < 160; c += kEpiThreads)
      vec[c] = (fwd && P.bias && c < bn && n0 + c < P.Nout) ? __ldg(P.bias + n0 + c) : 0.f;
    if (pool) for (int i = et; i < kFMaxSent * tabw; i += kEpiThreads) tab[i] = 0u;
    epi_bar(kEpiThreads);
    const int row = q * 32 + lane;
    const uint32_t urow = u_s + (uint32_t)row * 64u;
    const uint32_t usw = (uint32_t)((row >> 1) & 3);
    // patch entries (adjoint): global arg-max row and value per (sentence of the tile, this thread's column)
    float pv[kFMaxSent];
    int pg[kFMaxSent];
    auto load_patch = [&](const int32_t* h) {
      const int s0_ = h[3], ns_ = h[2];
      const int gc = n0 + et;
#pragma unroll
      for (int s = 0; s < kFMaxSent; ++s) {
        pv[s] = 0.f; pg[s] = -1;
        if (et < bn && gc < P.Nout && s < ns_) {
          const int64_t o = (int64_t)(s0_ + s) * P.ldpatch + gc;
          pg[s] = __ldg(P.patch_arg + o);
          pv[s] = __ldg(P.patch_val + o);
        }
      }
    };
    if (patch && t_first < n_tiles) {
      f2_wait(mfull(0), 0);
      load_patch(reinterpret_cast<const int32_t*>(meta(0) + 656));
    }
    int it = 0, tr = 0;
    for (int t = t_first; t < n_tiles; t += t_step, ++it) {
      const int as = it % kFAccStages;
      const uint32_t aphase = (uint32_t)(it / kFAccStages) & 1u;
      const int b = it & 1;
      const uint8_t* mb = meta(b);
      const float* invd = reinterpret_cast<const float*>(mb);
      const uint8_t* srow = mb + 512;
      const uint8_t* sfirst = mb + 640;
      const int32_t* hdr = reinterpret_cast<const int32_t*>(mb + 656);
      if (et == 0) F2_TRACE(0, tr);
      f2_wait(mfull(b), ((uint32_t)it >> 1) & 1u);
      if (et == 0) F2_TRACE(0, tr);
      const int r0 = hdr[0], n = hdr[1], ns = hdr[2], s0 = hdr[3];
      const int n16 = (n + 15) & ~15;
      // ---- patch entries of the tile (adjoint): thread = column, one entry per sentence (loaded one tile ahead)
      uint32_t pa_lo = 0xffffffffu, pa_hi = 0xffffffffu;
      if (patch) {
#pragma unroll
        for (int s = 0; s < kFMaxSent; ++s) {
          const int a = pg[s] - r0;
          const uint32_t lr = (pg[s] >= 0 && a >= 0 && a < n) ? (uint32_t)a : 0xffu;
          if (s < 4) pa_lo = (pa_lo & ~(0xffu << (8 * s))) | (lr << (8 * s));
          else pa_hi = (pa_hi & ~(0xffu << (8 * (s - 4)))) | (lr << (8 * (s - 4)));
        }
      }
      const float inv = invd[row];                           // 0 for rows >= n
      f2_wait(tfull(as), aphase);
      if (et == 0) F2_TRACE(0, tr);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * kFAccStride);
      // ---- phase A: U -> (adjoint: x 1/(deg+1)) -> bf16 -> S, thread = row (= K index of the second product).  Two column
      // groups: the aggregation product of group 0 is issued as soon as its columns are staged and runs under group 1.
      const int gsplit = min(P.gh, ngran);
      auto phase_a = [&](int g_beg, int g_end) {
        if (q * 32 < n16) {
          const float scale = fwd ? 1.f : inv;
          uint32_t ra[16];
#pragma unroll 1
          for (int g = g_beg + part; g < g_end; g += nparts) {   // (one copy of the body: the kernel's code footprint matters)
            tmem_ld_32x32_x16(tbase + g * 16, ra);
            tmem_wait_ld();
            if (row < n16) {
              uint32_t w[8];
#pragma unroll
              for (int k = 0; k < 8; ++k)
                w[k] = fwd ? pack_bf16x2(__uint_as_float(ra[2 * k]), __uint_as_float(ra[2 * k + 1]))
                           : pack_bf16x2(__uint_as_float(ra[2 * k]) * scale, __uint_as_float(ra[2 * k + 1]) * scale);
              const uint32_t a = urow + (uint32_t)(g >> 1) * 8192u;
              const uint32_t c0 = (uint32_t)(g & 1) * 2u;
              st_shared_v4(a + (((c0) ^ usw) << 4), w[0], w[1], w[2], w[3]);
              st_shared_v4(a + (((c0 + 1u) ^ usw) << 4), w[4], w[5], w[6], w[7]);
            }
          }
        }
      };
      auto publish = [&](int grp) {                          // S (generic-proxy stores) -> visible to the tensor core
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ufull(grp));
      };
      phase_a(0, gsplit);
      if (!patch) publish(0);
      if (gsplit < ngran) phase_a(gsplit, ngran);
      // ---- patch (adjoint): S[arg][c] += val / (deg_arg + 1); thread = column: distinct addresses
      if (patch) {
        epi_bar(kEpiThreads);
        if (et < bn) {
#pragma unroll
          for (int s = 0; s < kFMaxSent; ++s) {
            const uint32_t lr = ((s < 4 ? pa_lo : pa_hi) >> (8 * (s & 3))) & 0xffu;
            if (lr != 0xffu) {
              const uint32_t a = u_s + (uint32_t)(et >> 5) * 8192u + lr * 64u +
                                 ((((uint32_t)(et & 31) >> 3) ^ ((lr >> 1) & 3u)) << 4) + (uint32_t)(et & 7) * 2u;
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(sm + (a - base));
              *dst = __float2bfloat16_rn(__bfloat162float(*dst) + pv[s] * invd[lr]);
            }
          }
        }
        publish(0);
      }
      if (gsplit < ngran) publish(1);
      if (patch && t + t_step < n_tiles) {                   // the next tile's patch entries: in flight during phases C and D
        f2_wait(mfull(b ^ 1), ((uint32_t)(it + 1) >> 1) & 1u);
        load_patch(reinterpret_cast<const int32_t*>(meta(b ^ 1) + 656));
      }
      if (et == 0) F2_TRACE(0, tr);
      // ---- phase C: Y -> (forward: x 1/(deg+1) + bias) -> bf16 -> staging tile, in place of S (row per lane, same layout)
      auto phase_c = [&](int g_beg, int g_end) {
        if (q * 32 < n || (q == 3 && P.colsum_part)) {
          const float sc = fwd ? inv : 1.f;
          __nv_bfloat16* yrow = P.y + (int64_t)(r0 + row) * P.ldy + n0;
          uint32_t r[16];
#pragma unroll 1
          for (int g = g_beg + part; g < g_end; g += nparts) {
            tmem_ld_32x32_x16(tbase + g * 16, r);
            tmem_wait_ld();
            if (row < n) {
              float bb[16];
              if (fwd) {
#pragma unroll
                for (int k = 0; k < 4; ++k) lds_v4(vec_s + (uint32_t)g * 64u + 16u * k, *reinterpret_cast<float(*)[4]>(&bb[4 * k]));
              }
              uint32_t w[8];
#pragma unroll
              for (int k = 0; k < 8; ++k)
                w[k] = fwd ? pack_bf16x2(fmaf(__uint_as_float(r[2 * k]), sc, bb[2 * k]), fmaf(__uint_as_float(r[2 * k + 1]), sc, bb[2 * k + 1]))
                           : pack_bf16x2(__uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1]));
              const int gc = n0 + 16 * g;                                // 32 bytes of the row per lane: whole sectors
              if (gc + 8 <= (int)P.ldy) *reinterpret_cast<uint4*>(yrow + 16 * g) = make_uint4(w[0], w[1], w[2], w[3]);
              if (gc + 16 <= (int)P.ldy) *reinterpret_cast<uint4*>(yrow + 16 * g + 8) = make_uint4(w[4], w[5], w[6], w[7]);
              if (pool) {                                                // staged copy for the max-pool, in place of S: group
                const uint32_t a = urow + (uint32_t)(g >> 1) * 8192u;    // g's columns overwrite only what its own product consumed
                const uint32_t c0 = (uint32_t)(g & 1) * 2u;
                st_shared_v4(a + (((c0) ^ usw) << 4), w[0], w[1], w[2], w[3]);
                st_shared_v4(a + (((c0 + 1u) ^ usw) << 4), w[4], w[5], w[6], w[7]);
              }
            } else if (row == 127 && P.colsum_part) {        // lane 127 of the accumulator: this tile's column sums
#pragma unroll
              for (int k = 0; k < 16; ++k) vec[16 * g + k] += __uint_as_float(r[k]);
            }
          }
        }
      };
      f2_wait(yfull(0), (uint32_t)it & 1u);
      if (et == 0) F2_TRACE(0, tr);
      tc_fence_after();
      phase_c(0, gsplit);
      if (gsplit < ngran) {
        f2_wait(yfull(1), (uint32_t)it & 1u);
        tc_fence_after();
        phase_c(gsplit, ngran);
      }
      tc_fence_before();                                     // accumulator drained: the MMA warp may reuse the stage
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(as));
      if (et == 0) F2_TRACE(0, tr);
      if (pool) epi_bar(kEpiThreads);
      if (et == 0) F2_TRACE(0, tr);
      // ---- phase D (max-pool only): per-sentence column maxima of the staged rows and their FIRST rows (torch.max
      // semantics), thread = (16-byte chunk, row slice).  Four rows per step: independent loads (8 warps cannot hide a
      // dependent chain per row)
      if (pool && d_active) {
        int cur_s = -1;
        uint32_t vmax[4] = {0u, 0u, 0u, 0u}, varg[4] = {0u, 0u, 0u, 0u};     // packed bf16x2 maxima / u16x2 rows
        auto flush = [&]() {
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const int cbase = cur_s * tabw + 8 * k8 + 2 * w;
            atomicMax(tab + cbase, pool_key(vmax[w] << 16, 0xffffu - (varg[w] & 0xffffu)));
            atomicMax(tab + cbase + 1, pool_key(vmax[w], 0xffffu - (varg[w] >> 16)));
          }
        };
        const uint32_t dcol = u_s + (uint32_t)(k8 >> 2) * 8192u;
        const uint32_t dch = (uint32_t)(k8 & 3);
#pragma unroll 1
        for (int i0 = sl; i0 < n; i0 += 4 * nsl) {
          uint32_t pw[4][4];
          int srw[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * nsl;
            if (i < n) {
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(pw[u][0]), "=r"(pw[u][1]), "=r"(pw[u][2]), "=r"(pw[u][3])
                           : "r"(dcol + (uint32_t)i * 64u + ((dch ^ (uint32_t)((i >> 1) & 3)) << 4)));
              srw[u] = srow[i];
            }
          }
          {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int i = i0 + u * nsl;
              if (i < n) {
                const int s = srw[u];
                const uint32_t ipk = (uint32_t)i * 0x00010001u;
                if (s != cur_s) {
                  if (cur_s >= 0) flush();
                  cur_s = s;
#pragma unroll
                  for (int w = 0; w < 4; ++w) { vmax[w] = pw[u][w]; varg[w] = ipk; }
                } else {
#pragma unroll
                  for (int w = 0; w < 4; ++w) {
                    const __nv_bfloat162 nv = *reinterpret_cast<const __nv_bfloat162*>(&pw[u][w]);
                    const __nv_bfloat162 ov = *reinterpret_cast<const __nv_bfloat162*>(&vmax[w]);
                    const uint32_t gt = __hgt2_mask(nv, ov);
                    const __nv_bfloat162 mx = __hmax2(ov, nv);
                    vmax[w] = *reinterpret_cast<const uint32_t*>(&mx);
                    varg[w] = (varg[w] & ~gt) | (ipk & gt);
                  }
                }
              }
            }
          }
        }
        if (cur_s >= 0) flush();
      }
      if (et == 0) F2_TRACE(0, tr);
      if (pool) epi_bar(kEpiThreads);                        // the staging tile is free for the next tile's S
      if (et == 0) F2_TRACE(0, tr);
      if (pool) {
        // ---- pool table -> global (and reset): thread = column; the table's next writers come after another barrier
        if (et < bn) {
          const int gc = n0 + et;
          uint32_t key[kFMaxSent];
#pragma unroll
          for (int s = 0; s < kFMaxSent; ++s)
            if (s < ns) { key[s] = tab[s * tabw + et]; tab[s * tabw + et] = 0u; }
          if (gc < P.Nout) {
#pragma unroll
            for (int s = 0; s < kFMaxSent; ++s)
              if (s < ns) {
                const int64_t o = (int64_t)(s0 + s) * P.ldpool + gc;
                const bool any = sfirst[s + 1] > sfirst[s];
                P.hmax[o] = any ? pool_key_value(key[s]) : 0.f;
                P.harg[o] = any ? r0 + (int)(0xffffu - (key[s] & 0xffffu)) : -1;
              }
          }
        }
      }
      if (et == 0) F2_TRACE(0, tr);
      __syncwarp();
      if (lane == 0) mbar_arrive(mempty(b));
    }
    // ---- partial column sums of this CTA (adjoint)
    if (P.colsum_part) {
      epi_bar(kEpiThreads);
      for (int c = et; c < bn; c += kEpiThreads) {
        const int gc = n0 + c;
        if (gc < P.Nout) P.colsum_part[(int64_t)t_first * P.Nout + gc] = vec[c];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Two 16-byte words per row.  Word 0: sentence-local ids (u8) of the row itself (slot 0) and of its first fifteen other
// neighbours in CSR order (0xff = unused).  Word 1: number of CSR entries of the row (= rowsum(adj), self loop included),
// global sentence index, 0, 0.
__global__ void __launch_bounds__(256)
row_meta_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col, const int32_t* __restrict__ row_sent,
                const int32_t* __restrict__ sent_ptr, int N, uint4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int s = __ldg(row_sent + i);
  const int ss = __ldg(sent_ptr + s);
  const int eb = __ldg(row_ptr + i), ee = __ldg(row_ptr + i + 1);
  uint32_t w[4] = {0xffffff00u | (uint32_t)((i - ss) & 0xff), 0xffffffffu, 0xffffffffu, 0xffffffffu};
  int cnt = 0;
  for (int e = eb; e < ee && cnt < 15; ++e) {
    const int j = __ldg(col + e);
    if (j != i) {
      ++cnt;
      w[cnt >> 2] = (w[cnt >> 2] & ~(0xffu << (8 * (cnt & 3)))) | ((uint32_t)((j - ss) & 0xff) << (8 * (cnt & 3)));
    }
  }
  out[2 * (int64_t)i] = make_uint4(w[0], w[1], w[2], w[3]);
  out[2 * (int64_t)i + 1] = make_uint4((uint32_t)(ee - eb), (uint32_t)s, 0u, 0u);
}

struct Fused2Plan { int stages, rows_cap, bn0, bn1, n_split, opitch, osw_shift, osw_mask; size_t smem;
                    uint32_t off_a, off_u, off_adj, off_tab, off_meta, off_vec, off_bar; bool ok; };

static Fused2Plan plan_fused2(int K, int Nout) {
  Fused2Plan p;
  memset(&p, 0, sizeof(p));
  if (K <= 0 || Nout <= 0 || K > 64 * 5 || Nout > 2 * kFAccStride) return p;
  const int num_kb = (K + 63) / 64;
  const int n16 = (Nout + 15) / 16 * 16;
  if (n16 <= kFAccStride) { p.n_split = 1; p.bn0 = n16; p.bn1 = 0; }
  else { p.n_split = 2; p.bn0 = ((n16 / 2 + 15) / 16) * 16; p.bn1 = n16 - p.bn0; }
  p.rows_cap = 127;                                   // row 127 of Adj carries the column-sum weights
  // pool staging tile: [128 rows x bn0 bf16], unpadded, 16-byte chunks XOR-swizzled so that both the row-per-lane stores
  // and the chunk-per-lane loads are conflict-free: chunks mod 8 = 0 -> c ^ (row & 7); 4 -> c ^ ((row >> 1) & 3); 2, 6 -> c ^ ((row >> 2) & 1)
  p.opitch = p.bn0 * 2;
  const int m8 = (p.bn0 / 8) & 7;
  if (m8 == 0) { p.osw_shift = 0; p.osw_mask = 7; }
  else if (m8 == 4) { p.osw_shift = 1; p.osw_mask = 3; }
  else { p.osw_shift = 2; p.osw_mask = 1; }
  const size_t budget = 227 * 1024 - 1024;
  const size_t w_bytes = ((size_t)num_kb * p.bn0 * 128 + 1023) & ~(size_t)1023;
  const size_t tab = (size_t)kFMaxSent * p.bn0 * 4;
  const size_t fixed = w_bytes + kF2UBytes + kF2AdjBytes + tab + 2 * kF2MetaBytes + kF2VecBytes + 256;
  if (fixed + 2 * 16384 > budget) return p;
  int stages = (int)((budget - fixed) / 16384);
  if (stages > kFMaxStages) stages = kFMaxStages;
  int ws = fused_env("EDG_FUSED_STAGES", 0);
  if (ws >= 2 && ws < stages) stages = ws;
  p.stages = stages;
  uint32_t off = (uint32_t)w_bytes;
  p.off_a = off; off += (uint32_t)stages * 16384u;
  p.off_u = off; off += kF2UBytes;
  p.off_adj = off; off += kF2AdjBytes;
  p.off_tab = off; off += (uint32_t)tab;
  p.off_meta = off; off += 2 * kF2MetaBytes;
  p.off_vec = off; off += kF2VecBytes;
  off = (off + 7u) & ~7u;
  p.off_bar = off; off += 256;
  p.smem = (size_t)off + 1024;
  p.ok = p.smem <= 227 * 1024;
  return p;
}

static int fused2_tile_rows(int K, int Nout) {
  const Fused2Plan p = plan_fused2(K, Nout);
  return p.ok ? p.rows_cap : 0;
}

static int launch_gcn_layer2(const void* x, int64_t ldx, int32_t N, int32_t K, const void* w, int64_t ldw, int32_t Nout,
                             const float* bias, int mode, const int32_t* row_ptr, const int32_t* col, const int32_t* sent_ptr,
                             const int32_t* tile_info, const int32_t* n_tiles, int32_t tile_rows, void* y, int64_t ldy, float* hmax,
                             int32_t* harg, int64_t ldpool, const float* patch_val, const int32_t* patch_arg, int64_t ldpatch,
                             float* colsum, int colsum_accumulate, void* ws, size_t ws_bytes, const void* row_meta, cudaStream_t s) {
  if (!row_meta) return EDG_ERR_ARG;
  const Fused2Plan p = plan_fused2(K, Nout);
  if (!p.ok || tile_rows > p.rows_cap) return EDG_ERR_UNSUPPORTED;
  const int groups = kNumSMs / p.n_split;
  if (colsum && ws_bytes < (size_t)groups * Nout * sizeof(float)) return EDG_ERR_WORKSPACE;
  GcnLayer2Params P;
  memset(&P, 0, sizeof(P));
  int rc = make_map_bf16(&P.map_a, x, N, K, ldx, 64, 32);
  if (rc) return rc;
  rc = make_map_bf16(&P.map_w, w, Nout, K, ldw, 64, 16);
  if (rc) return rc;
  P.tile_info = tile_info; P.n_tiles = n_tiles; P.row_ptr = row_ptr; P.col = col; P.sent_ptr = sent_ptr;
  P.row_meta = (const uint4*)row_meta;
  P.bias = bias; P.y = (__nv_bfloat16*)y; P.ldy = ldy; P.hmax = hmax; P.harg = harg; P.ldpool = ldpool;
  P.patch_val = patch_val; P.patch_arg = patch_arg; P.ldpatch = ldpatch;
  P.colsum_part = colsum ? (float*)ws : nullptr;
  P.trace = (!colsum && ws && ws_bytes >= 3 * 160 * 8 && (fused_env("EDG_FUSED_DEBUG", 0) & 32)) ? (long long*)ws : nullptr;
  P.K = K; P.Nout = Nout; P.num_kb = (K + 63) / 64; P.mode = mode; P.n_split = p.n_split;
  P.stages = p.stages; P.rows_cap = p.rows_cap; P.debug = fused_env("EDG_FUSED_DEBUG", 0);
  P.bn[0] = p.bn0; P.bn[1] = p.bn1;
  P.idesc[0] = make_idesc_bf16(kFRows, p.bn0, 0, 0);
  P.idesc[1] = make_idesc_bf16(kFRows, p.bn1 > 0 ? p.bn1 : 16, 0, 0);
  P.gh = 6;                                           // column group 0 = three 32-column atoms
  for (int h = 0; h < 2; ++h) {
    const int bnh = h ? p.bn1 : p.bn0;
    const int c0 = bnh < 16 * P.gh ? bnh : 16 * P.gh, c1 = bnh - c0;
    P.idesc2[h][0] = make_idesc_bf16(kFRows, c0 > 0 ? c0 : 16, 0, 1);
    P.idesc2[h][1] = make_idesc_bf16(kFRows, c1 > 0 ? c1 : 16, 0, 1);
  }
  P.opitch = p.opitch; P.osw_shift = p.osw_shift; P.osw_mask = p.osw_mask;
  P.off_a = p.off_a; P.off_u = p.off_u; P.off_adj = p.off_adj; P.off_tab = p.off_tab; P.off_meta = p.off_meta;
  P.off_vec = p.off_vec; P.off_bar = p.off_bar;
  const int epi = fused_env("EDG_FUSED_EPI_WARPS", 8) == 16 ? 16 : 8;
  if (epi == 16) {
    if (int rc_ = ensure_dyn_smem((const void*)gcn_layer2_kernel<16>, p.smem)) return rc_;
    gcn_layer2_kernel<16><<<groups * p.n_split, (4 + 16) * 32, p.smem, s>>>(P);
  } else {
    if (int rc_ = ensure_dyn_smem((const void*)gcn_layer2_kernel<8>, p.smem)) return rc_;
    gcn_layer2_kernel<8><<<groups * p.n_split, (4 + 8) * 32, p.smem, s>>>(P);
  }
  rc = check_launch();
  if (rc) return rc;
  if (colsum) {
    colsum_part_reduce_kernel<<<(Nout + 31) / 32, 256, 0, s>>>((const float*)ws, groups, Nout, colsum, colsum_accumulate);
    rc = check_launch();
  }
  return rc;
}

}  // namespace edg

/* see include/edgcn.h */
extern "C" int edg_row_meta(const int32_t* row_ptr, const int32_t* col, const int32_t* row_sent, const int32_t* sent_ptr,
                            int32_t N, void* row_meta, edg_stream stream) {
  if (N < 0) return EDG_ERR_ARG;
  if (N == 0) return EDG_OK;
  if (!row_ptr || !col || !row_sent || !sent_ptr || !row_meta || !edg::aligned16(row_meta)) return EDG_ERR_ARG;
  edg::row_meta_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(row_ptr, col, row_sent, sent_ptr, N, (uint4*)row_meta);
  return edg::check_launch();
}
