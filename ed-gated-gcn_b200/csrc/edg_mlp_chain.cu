// mlp_chain : the gate MLPs (models/bert_amir5.py:562-571, called at :621-622) as ONE launch per direction.
//
// A gate is a chain of 2-3 (Linear(D,D), Sigmoid) pairs applied to the [B, D] trigger vectors; the reference
// runs it as 5-7 framework kernels per gate and direction, the per-Linear kernels of this library still needed
// 4 + 9 launches of ~10 us each for ~1 us of tensor-core work.  Here CTA (m, g) owns 128 sentences of gate g
// and walks the whole chain with the activation tile RESIDENT in shared memory:
//
//   A_0 (TMA, 128B swizzle)  ->  [ tcgen05.mma  A_s x W_s^T -> TMEM ]  ->  epilogue: tcgen05.ld, bias + sigmoid
//   (forward) or multiply by y(1-y) of the saved activation (backward), store the global copy the other
//   direction / the weight-gradient GEMM needs, and write the bf16 result back into the SAME swizzled
//   shared-memory tile as A_{s+1}  (generic-proxy stores + fence.proxy.async, then an mbarrier hands the tile
//   to the MMA thread).  Only the [D x D] weights stream (3-stage TMA ring of [2*n_half x 64] K blocks).
//   Warp 0 = TMA producer, warp 1 = MMA issuer, 16 epilogue warps (the epilogue is the critical path: a first
//   version with 8 warps spent 12.8 cycles per issued instruction and left the tensor pipe waiting).
//
// D <= 320 (5 K blocks, two N halves of <= 160 TMEM columns); larger widths use the per-Linear kernels.
// Unity build: included after edg_gemm_tc.cu (PTX wrappers, descriptors, tile constants).
#include "edg_common.cuh"

namespace edg {

constexpr int kChainMaxStages = 3;
constexpr int kChainMaxGroups = 4;
constexpr int kChainWStages = 3;
constexpr int kChainMaxKb = 5;
constexpr int kChainBiasPitch = 64 * kChainMaxKb;   // 320
constexpr int kChainEpiWarps = 16;                  // 4 per TMEM lane quarter: the epilogue is the critical path
constexpr int kChainThreads = 64 + 32 * kChainEpiWarps;

struct ChainStageDev {
  const float* bias;            // forward: [D] or null
  const __nv_bfloat16* y;       // backward: saved activation [M, ldy] whose sigmoid' multiplies the result (null = none)
  int64_t ldy;
  void* out;                    // global copy of the stage output (null = none)
  int64_t ldo;
  int out_f32;
};
struct ChainParams {
  CUtensorMap map_a[kChainMaxGroups];
  CUtensorMap map_w[kChainMaxGroups][kChainMaxStages];
  ChainStageDev st[kChainMaxGroups][kChainMaxStages];
};

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(kChainThreads, 1)
mlp_chain_kernel(const __grid_constant__ ChainParams P, int M, int D, int n_stages, int mode, int n_half, int num_kb,
                 uint32_t idesc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int g = blockIdx.y;
  const int m0 = blockIdx.x * kBlockM;
  const uint32_t w_stage_bytes = (uint32_t)n_half * 256u;            // [2*n_half rows x 64 k] bf16
  const uint32_t w_base = base + (uint32_t)num_kb * kABytes;
  const uint32_t bias_off = w_base + kChainWStages * w_stage_bytes;
  float* bias_s = reinterpret_cast<float*>(smem_raw + (bias_off - smem_u32(smem_raw)));
  const uint32_t bars = bias_off + kChainMaxStages * kChainBiasPitch * 4;
  const uint32_t afull = bars, tfull = bars + 8, aready = bars + 16, tmem_slot = bars + 24;
  auto wfull = [&](int s) { return bars + 32 + 8 * s; };
  auto wempty = [&](int s) { return bars + 32 + 8 * (kChainWStages + s); };
  auto Ablk = [&](int kb) { return base + (uint32_t)kb * kABytes; };
  auto Wst = [&](int s) { return w_base + (uint32_t)s * w_stage_bytes; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < n_stages * kChainBiasPitch; i += kChainThreads) {
    const int s = i / kChainBiasPitch, col = i - s * kChainBiasPitch;
    const float* b = P.st[g][s].bias;
    bias_s[i] = (b && col < D) ? b[col] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.map_a[g]);
    for (int s = 0; s < n_stages; ++s) tma_prefetch_desc(&P.map_w[g][s]);
    mbar_init(afull, 1); mbar_init(tfull, 1); mbar_init(aready, kChainEpiWarps);
    for (int s = 0; s < kChainWStages; ++s) { mbar_init(wfull(s), 1); mbar_init(wempty(s), 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(afull, (uint32_t)num_kb * kABytes);
      for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(Ablk(kb), &P.map_a[g], afull, kb * kBlockK, m0);
      const int box_rows = n_half / 2;
      int stage = 0; uint32_t phase = 0;
      for (int s = 0; s < n_stages; ++s) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(wempty(stage), phase ^ 1);
          mbar_expect_tx(wfull(stage), w_stage_bytes);
          for (int j = 0; j < 4; ++j)
            tma_load_2d(Wst(stage) + (uint32_t)(j * box_rows) * 128u, &P.map_w[g][s], wfull(stage), kb * kBlockK, j * box_rows);
          if (++stage == kChainWStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(afull, 0);
      tc_fence_after();
      int stage = 0; uint32_t phase = 0;
      for (int s = 0; s < n_stages; ++s) {
        if (s > 0) {                                   // the epilogue has drained TMEM and rewritten the A tile
          mbar_wait(aready, (uint32_t)((s - 1) & 1));
          tc_fence_after();
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(wfull(stage), phase);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t ad = make_desc_sw128(Ablk(kb) + k * 32, 16, 1024);
            const uint64_t bd0 = make_desc_sw128(Wst(stage) + k * 32, 16, 1024);
            const uint64_t bd1 = make_desc_sw128(Wst(stage) + (uint32_t)n_half * 128u + k * 32, 16, 1024);
            umma_bf16(tmem_base, ad, bd0, idesc, (kb | k) != 0);
            umma_bf16(tmem_base + n_half, ad, bd1, idesc, (kb | k) != 0);
          }
          umma_commit(wempty(stage));
          if (++stage == kChainWStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull);
      }
    }
  } else {
    // 16 epilogue warps: warp w reads TMEM lane quarter (w & 3); the four warps of a quarter take every fourth
    // 32-column chunk.  The saved activation y of the next chunk (backward) is requested before the wait on the
    // accumulator, so its DRAM/L2 latency overlaps the MMAs.
    const int q = warp & 3;
    const int sub = (warp - 2) >> 2;
    const int r_tile = q * 32 + lane;
    const int row = m0 + r_tile;
    const bool row_ok = row < M;
    const int n_chunks = 2 * num_kb;
    const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int s = 0; s < n_stages; ++s) {
      const ChainStageDev& S = P.st[g][s];
      const bool last = s == n_stages - 1;
      const bool has_y = mode == 1 && S.y != nullptr;
      const __nv_bfloat16* yrow = has_y ? S.y + (int64_t)(row_ok ? row : 0) * S.ldy : nullptr;
      uint4 yq[4];
      auto load_y = [&](int c0) {
#pragma unroll
        for (int v8 = 0; v8 < 4; ++v8) {
          const int col = c0 + v8 * 8;
          yq[v8] = (row_ok && col + 8 <= S.ldy) ? __ldg(reinterpret_cast<const uint4*>(yrow + col)) : make_uint4(0u, 0u, 0u, 0u);
        }
      };
      if (has_y && sub < n_chunks) load_y(sub * 32);
      mbar_wait(tfull, (uint32_t)(s & 1));
      tc_fence_after();
      for (int ci = sub; ci < n_chunks; ci += 4) {
        const int c0 = ci * 32;
        const bool full = c0 + 32 <= D;                      // chunk-uniform: only the last chunk masks columns
        uint32_t r[32];
        tmem_ld_32x32_nowait(tbase + c0, r);
        tmem_wait_ld();
        float val[32];
        if (mode == 0) {
          const float* bs = bias_s + s * kChainBiasPitch + c0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float sg = __fdividef(1.0f, 1.0f + __expf(-(__uint_as_float(r[j]) + bs[j])));
            val[j] = (full || c0 + j < D) ? sg : 0.f;
          }
        } else if (has_y) {
#pragma unroll
          for (int v8 = 0; v8 < 4; ++v8) {
            const uint32_t w[4] = {yq[v8].x, yq[v8].y, yq[v8].z, yq[v8].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {                    // bf16 -> fp32 is a 16-bit shift; padding columns of y are 0
              const float y0 = __uint_as_float(w[i] << 16), y1 = __uint_as_float(w[i] & 0xffff0000u);
              const int j = v8 * 8 + 2 * i;
              const float a0 = __uint_as_float(r[j]) * y0 * (1.f - y0), a1 = __uint_as_float(r[j + 1]) * y1 * (1.f - y1);
              val[j] = (full || c0 + j < D) ? a0 : 0.f;
              val[j + 1] = (full || c0 + j + 1 < D) ? a1 : 0.f;
            }
          }
          if (ci + 4 < n_chunks) load_y((ci + 4) * 32);      // next chunk's y while this one is stored
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) val[j] = (full || c0 + j < D) ? __uint_as_float(r[j]) : 0.f;
        }
        // global copy (saved activation / gate / gradient)
        if (row_ok && S.out != nullptr) {
          if (S.out_f32) {
            float* o = reinterpret_cast<float*>(S.out) + (int64_t)row * S.ldo;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (full || c0 + j + 4 <= D) *reinterpret_cast<float4*>(o + c0 + j) = make_float4(val[j], val[j + 1], val[j + 2], val[j + 3]);
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(S.out) + (int64_t)row * S.ldo;
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              if (full || c0 + j + 8 <= S.ldo)
                *reinterpret_cast<uint4*>(o + c0 + j) =
                    make_uint4(pack_bf16x2(val[j], val[j + 1]), pack_bf16x2(val[j + 2], val[j + 3]),
                               pack_bf16x2(val[j + 4], val[j + 5]), pack_bf16x2(val[j + 6], val[j + 7]));
          }
        }
        // next stage's A operand: K block c0/64, row r_tile, 16-byte chunks XOR-swizzled by (row & 7)
        if (!last) {
          const uint32_t rowaddr = Ablk(c0 >> 6) + (uint32_t)r_tile * 128u;
          const int chunk0 = (c0 & 63) >> 3;
#pragma unroll
          for (int v8 = 0; v8 < 4; ++v8) {
            const uint32_t addr = rowaddr + ((uint32_t)((chunk0 + v8) ^ (r_tile & 7)) << 4);
            st_shared_v4(addr, pack_bf16x2(val[v8 * 8], val[v8 * 8 + 1]), pack_bf16x2(val[v8 * 8 + 2], val[v8 * 8 + 3]),
                         pack_bf16x2(val[v8 * 8 + 4], val[v8 * 8 + 5]), pack_bf16x2(val[v8 * 8 + 6], val[v8 * 8 + 7]));
          }
        }
      }
      if (!last) {
        fence_proxy_async();          // generic-proxy stores -> visible to the tensor core (async proxy)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(aready);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace edg

using namespace edg;

extern "C" int edg_mlp_chain(int mode, int32_t n_groups, int32_t n_stages, const void* const* a0, int64_t lda,
                             const edg_chain_stage* stages, int32_t M, int32_t D, edg_stream stream) {
  if (mode < 0 || mode > 1 || n_groups <= 0 || n_stages <= 0 || M < 0 || D <= 0 || !a0 || !stages) return EDG_ERR_ARG;
  if (n_groups > kChainMaxGroups || n_stages > kChainMaxStages || D > 64 * kChainMaxKb || D < 16) return EDG_ERR_UNSUPPORTED;
  if (M == 0) return EDG_OK;
  if (lda < D || !row_pitch_ok(EDG_BF16, lda)) return EDG_ERR_ALIGN;
  const int num_kb = (D + kBlockK - 1) / kBlockK;
  const int n_half = 16 * ((D + 31) / 32);
  static_assert(sizeof(ChainParams) <= 4000, "kernel parameter space");
  ChainParams P;
  memset(&P, 0, sizeof(P));
  for (int g = 0; g < n_groups; ++g) {
    if (!a0[g] || !aligned16(a0[g])) return EDG_ERR_ALIGN;
    int rc = make_map_bf16(&P.map_a[g], a0[g], M, D, lda, kBlockK, kBlockM);
    if (rc) return rc;
    for (int s = 0; s < n_stages; ++s) {
      const edg_chain_stage& h = stages[g * n_stages + s];
      if (!h.w || !aligned16(h.w) || h.ldw < D || !row_pitch_ok(EDG_BF16, h.ldw)) return EDG_ERR_ALIGN;
      rc = make_map_bf16(&P.map_w[g][s], h.w, D, D, h.ldw, kBlockK, n_half / 2);
      if (rc) return rc;
      ChainStageDev& d = P.st[g][s];
      d.bias = mode == 0 ? h.bias : nullptr;
      d.y = mode == 1 ? reinterpret_cast<const __nv_bfloat16*>(h.y) : nullptr;
      d.ldy = h.ldy;
      d.out = h.out;
      d.ldo = h.ldo;
      d.out_f32 = h.out_dtype == EDG_F32;
      if (d.y && (!aligned16(d.y) || !row_pitch_ok(EDG_BF16, h.ldy) || h.ldy < D)) return EDG_ERR_ALIGN;
      if (d.out) {
        if (h.out_dtype != EDG_F32 && h.out_dtype != EDG_BF16) return EDG_ERR_DTYPE;
        if (!aligned16(d.out) || !row_pitch_ok(h.out_dtype, h.ldo) || h.ldo < D) return EDG_ERR_ALIGN;
        if (h.out_dtype == EDG_F32 && (D & 3)) return EDG_ERR_UNSUPPORTED;
      }
    }
  }
  const size_t smem = 1024 + (size_t)num_kb * kABytes + (size_t)kChainWStages * n_half * 256 +
                      (size_t)kChainMaxStages * kChainBiasPitch * 4 + 128;
  if (int rc_ = ensure_dyn_smem((const void*)mlp_chain_kernel, smem)) return rc_;
  const int m_tiles = (M + kBlockM - 1) / kBlockM;
  const uint32_t idesc = make_idesc_bf16(kBlockM, n_half, 0, 0);
  mlp_chain_kernel<<<dim3(m_tiles, n_groups), kChainThreads, smem, (cudaStream_t)stream>>>(P, M, D, n_stages, mode, n_half,
                                                                                         num_kb, idesc);
  return check_launch();
}
