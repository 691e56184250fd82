// K2: degree-normalised neighbour aggregation over the packed CSR of many sentences
// (the `adj @ hidden / (rowsum + 1)` half of models/gcn.py:35,41).
//
// HBM-bound (about 1.5 flop/byte).  Work is flattened to (row, 16-byte chunk): consecutive
// threads read consecutive 128-bit chunks of one hidden row, then of the next row, so every
// request is a fully used 32-byte sector stream.  A row's neighbours live in the same
// sentence (a few KB away), so the 2-3 extra row reads per output row hit L1/L2; compulsory
// DRAM traffic is one read + one write of the [N,D] matrix plus 16 B/row of CSR.
#include "edg_common.cuh"

namespace edg {

template <typename TO, int E>
__device__ __forceinline__ void store_chunk(TO* p, const float (&f)[E]);
template <> __device__ __forceinline__ void store_chunk<float, 4>(float* p, const float (&f)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
}
template <> __device__ __forceinline__ void store_chunk<float, 8>(float* p, const float (&f)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}
template <> __device__ __forceinline__ void store_chunk<__nv_bfloat16, 8>(__nv_bfloat16* p, const float (&f)[8]) {
  Vec16<__nv_bfloat16>::store(p, f);
}
template <> __device__ __forceinline__ void store_chunk<__nv_bfloat16, 4>(__nv_bfloat16* p, const float (&f)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}

template <typename TI, typename TO, int MODE>
__global__ void __launch_bounds__(256)
aggregate_kernel(const TI* __restrict__ x, int64_t ldx, TO* __restrict__ y, int64_t ldy, int N, int chunks,
                 const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col) {
  constexpr int E = Vec16<TI>::kElems;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int row = (int)(idx / chunks);
  if (row >= N) return;
  const int c = (int)(idx - (int64_t)row * chunks) * E;
  const int beg = __ldg(row_ptr + row), end = __ldg(row_ptr + row + 1);
  float acc[E];
#pragma unroll
  for (int k = 0; k < E; ++k) acc[k] = 0.f;
  for (int e = beg; e < end; e += 4) {
    int nb[4];
    float w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      nb[u] = (e + u < end) ? __ldg(col + e + u) : -1;
      w[u] = 1.f;
    }
    if (MODE == 1) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (nb[u] >= 0) w[u] = 1.0f / (float)(__ldg(row_ptr + nb[u] + 1) - __ldg(row_ptr + nb[u]) + 1);
    }
    float v[4][E];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (nb[u] >= 0) Vec16<TI>::load(x + (int64_t)nb[u] * ldx + c, v[u]);
      else {
#pragma unroll
        for (int k = 0; k < E; ++k) v[u][k] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int k = 0; k < E; ++k) acc[k] = (MODE == 1) ? fmaf(v[u][k], w[u], acc[k]) : acc[k] + v[u][k];
    }
  }
  if (MODE == 0) {
    const float den = (float)(end - beg + 1);     // rowsum(adj) + 1, gcn.py:35
#pragma unroll
    for (int k = 0; k < E; ++k) acc[k] = acc[k] / den;
  }
  store_chunk<TO, E>(y + (int64_t)row * ldy + c, acc);
}

template <typename TI, typename TO>
static int launch_aggregate(const void* x, int64_t ldx, void* y, int64_t ldy, int N, int D,
                            const int32_t* row_ptr, const int32_t* col, int mode, cudaStream_t s) {
  constexpr int E = Vec16<TI>::kElems;
  const int chunks = (D + E - 1) / E;
  if (ldx < (int64_t)chunks * E || ldy < (int64_t)chunks * E) return EDG_ERR_ALIGN;
  const int64_t total = (int64_t)N * chunks;
  const int64_t blocks = (total + 255) / 256;
  if (blocks > 0x7fffffffll) return EDG_ERR_UNSUPPORTED;
  if (mode == 0)
    aggregate_kernel<TI, TO, 0><<<(unsigned)blocks, 256, 0, s>>>((const TI*)x, ldx, (TO*)y, ldy, N, chunks, row_ptr, col);
  else
    aggregate_kernel<TI, TO, 1><<<(unsigned)blocks, 256, 0, s>>>((const TI*)x, ldx, (TO*)y, ldy, N, chunks, row_ptr, col);
  return check_launch();
}

}  // namespace edg

using namespace edg;

extern "C" int edg_aggregate(const void* x, int x_dtype, int64_t ldx, void* y, int y_dtype, int64_t ldy,
                             int32_t N, int32_t D, const int32_t* row_ptr, const int32_t* col, int mode,
                             edg_stream stream) {
  if (N < 0 || D <= 0 || (mode != 0 && mode != 1)) return EDG_ERR_ARG;
  if (N == 0) return EDG_OK;
  if (!x || !y || !row_ptr || !col) return EDG_ERR_ARG;
  if (!aligned16(x) || !aligned16(y) || !row_pitch_ok(x_dtype, ldx) || !row_pitch_ok(y_dtype, ldy)) return EDG_ERR_ALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  if (x_dtype == EDG_BF16 && y_dtype == EDG_BF16) return launch_aggregate<__nv_bfloat16, __nv_bfloat16>(x, ldx, y, ldy, N, D, row_ptr, col, mode, s);
  if (x_dtype == EDG_F32 && y_dtype == EDG_F32) return launch_aggregate<float, float>(x, ldx, y, ldy, N, D, row_ptr, col, mode, s);
  if (x_dtype == EDG_F32 && y_dtype == EDG_BF16) return launch_aggregate<float, __nv_bfloat16>(x, ldx, y, ldy, N, D, row_ptr, col, mode, s);
  if (x_dtype == EDG_BF16 && y_dtype == EDG_F32) return launch_aggregate<__nv_bfloat16, float>(x, ldx, y, ldy, N, D, row_ptr, col, mode, s);
  return EDG_ERR_DTYPE;
}
