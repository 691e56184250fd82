// Unity build of libedgcn.so (one translation unit: no relocatable device code needed).
#include "edg_api.cu"
#include "edg_graph.cu"
#include "edg_aggregate.cu"
#include "edg_gemm_simt.cu"
#include "edg_gemm_tc.cu"
#include "edg_block_staged.cu"
#include "edg_block.cu"
