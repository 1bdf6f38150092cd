// tree_internal.h - the device tree object shared by tree.cu (tree maintenance, CUDA-core scans) and nn_tensor.cu
// (tensor-core batched scan).  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nwb_internal.h"

// Tensor-core scan copy of the nodes: tiles of TC_N nodes x TC_K bf16 slots, each tile stored in the tcgen05 K-major
// no-swizzle "core matrix" order [row block 16][16-byte chunk 10][row 8][8 bf16], so one 1-D bulk copy lands a tile in
// shared memory exactly as the MMA shared-memory descriptor reads it.
namespace nwb {
constexpr int TC_M = 128;                         // queries per group = TMEM lanes
constexpr int TC_N = 128;                         // nodes per tile = TMEM columns of one accumulator
constexpr int TC_K = 80;                          // bf16 slots per row (5 MMA K-steps of 16)
constexpr int TC_KB = TC_K / 8;                   // 16-byte chunks per row
constexpr int TC_TILE_BYTES = TC_N * TC_K * 2;    // 20480
// slot map of a row (node side | query side):
//   [ 0,21) n_hi | -2 q_hi      [21,42) n_mid | -2 q_hi      [42,63) n_hi | -2 q_mid
//   [63,66) |n|^2 split in three bf16 pieces | 1.0            [66,80) zero
// fp32 scan copy: blocks of 128 nodes, [block][21 joints][128 nodes] - 84 B/node as before, but a tile of the scans is a
// run of CONSECUTIVE blocks, i.e. one contiguous bulk copy (10 752 B per block) instead of 21 row copies 4 MB apart
constexpr int SOA_BLK = 128;
constexpr int SOA_BLK_FLOATS = SOA_BLK * NWB_NJ;
__host__ __device__ inline size_t soa_index(int64_t node, int j) {
  return ((size_t)(node >> 7) * NWB_NJ + j) * SOA_BLK + (size_t)(node & (SOA_BLK - 1));
}
__host__ __device__ inline size_t tc_slot_offset(int64_t node, int slot) {
  const int64_t tile = node / TC_N;
  const int r = (int)(node - tile * TC_N);
  return (size_t)tile * TC_TILE_BYTES + (size_t)(r >> 3) * (TC_KB * 128) + (size_t)(slot >> 3) * 128 + (size_t)(r & 7) * 16 +
         (size_t)(slot & 7) * 2;
}
}  // namespace nwb

struct nwb_tree {
  nwb_ctx* ctx = nullptr;
  int64_t size = 0, cap = 0;
  float* soa = nullptr;          // fp32 scan copy, blocked: soa_index(node, joint)
  uint16_t* h16 = nullptr;       // fp16 filter copy of the single-query stream over large trees, same blocked order (42 B/node)
  double* aos = nullptr;
  int32_t* pred = nullptr;
  unsigned char* tc = nullptr;   // tensor-core scan copy: ceil(cap / 128) tiles of 20480 B (see above)
  unsigned* max_nn = nullptr;    // float bits of max |node|^2 over the tree (error bound of the tensor-core scan)
  void* tc_scratch = nullptr;    // chunk minima + slice minima of the two-phase tensor-core scan (grow-only)
  size_t tc_scratch_bytes = 0;
  // scratch
  double* part_d = nullptr;
  int32_t* part_i = nullptr;
  size_t part_cap = 0;
  unsigned* tickets = nullptr;   // one counter per query group (zero between calls)
  size_t ticket_cap = 0;
  void* stage = nullptr;      // device staging for host-pointer calls
  size_t stage_bytes = 0;
};

// nn_tensor.cu: batched nearest neighbour on the tensor cores (tcgen05); idx/dist on the device
cudaError_t nn_tc_launch(nwb_tree* t, const double* q_dev, int64_t nq, int32_t* idx_dev, double* dist_dev, float* debug_tile,
                         int* n_launches);
