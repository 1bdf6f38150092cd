// nn_tensor.cu - batched nearest neighbour on the 5th-generation tensor cores (tcgen05 / TMEM) for sm_100a.
//
// RRT_Planner::findNearestNeighbour (rrt_connect_planner/src/rrt_planner.cpp:336-419) for MANY queries against one
// tree is a contraction: |q - n|^2 = |q|^2 + |n|^2 - 2 q.n, i.e. a (Q x 21) . (21 x N) product plus norms.  The
// CUDA-core scan of tree.cu does that on the FP32 pipe (14.5 TFLOP/s of contraction at 1024 x 10^6); here the product
// runs as tcgen05.mma with the accumulator in tensor memory:
//
//   * operands are bf16, and a bf16 product alone would be far too coarse, so every value x is split x = hi + mid
//     (bf16 each, 16 significant bits together) and the three leading cross terms q_hi n_hi + q_hi n_mid + q_mid n_hi
//     are laid side by side along K (63 slots); three more slots carry |n|^2 (split in three bf16 pieces) against 1.0,
//     and the query side is pre-scaled by -2, so ONE accumulator element is v = |n|^2 - 2 q.n up to a bounded error E
//     (the |q|^2 term is constant per query and irrelevant for the argmin).  K = 80 (5 MMA steps of 16).
//   * A = 128 queries (one TMEM lane each), B = tiles of 128 nodes streamed with 1-D bulk copies (UBLKCP) through a
//     shared-memory ring; the node copy is stored in HBM already in the tcgen05 K-major core-matrix order
//     (tree_internal.h), so no tensor map and no swizzle are needed.  D = 128 x 128 fp32, double-buffered in TMEM: the
//     MMA of tile t+1 runs while the four warps read tile t back with tcgen05.ld.
//   * The epilogue is a filter, not the answer: every thread owns one query, keeps its running minimum of v and lists
//     the nodes within 2E of it.  Only those candidates are re-evaluated in fp64 with the reference's exact rounding
//     sequence (separately rounded multiply/add, sqrt, strict <, lowest index on ties), so the result is bit-identical
//     to the CUDA-core scan and to the oracle.  A tile whose candidates overflow the list is evaluated exactly, pair by
//     pair, by the CTA itself - no CPU path, no approximation.
//
// Grid: (node slices, query groups).  Partials (dist, idx) per (query, slice) are folded by the last CTA of a group.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "tree_internal.h"

namespace nwb {

constexpr int TC_THREADS = 512;                  // 16 warps: 4 lane quadrants x 4 column chunks of a tile
constexpr int TC_STAGES = 6;
constexpr int TC_CAND = 2048;

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TCWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TCDONE_%=;\n"
      "bra TCWAIT_%=;\n"
      "TCDONE_%=:\n"
      "}\n" ::"r"(tc_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(tc_smem_u32(bar))
               : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle: 8-row x 16-byte core matrices; LBO = distance between the two
// 16-byte K chunks of one MMA step, SBO = distance between 8-row groups (both in 16-byte units), descriptor version 1
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)((TC_KB * 128u) >> 4) << 32) |
         (1ull << 46);
}
// instruction descriptor, kind::f16: D = F32, A = B = BF16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kTcInstrDesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kTcInstrDesc), "r"(accumulate)
      : "memory");
}
// one lane of a fully converged warp (the others skip the guarded statement)
__device__ __forceinline__ bool tc_elect_one() {
  uint32_t is_elected;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(is_elected));
  return is_elected != 0;
}
// same MMA with N = 256: one instruction covers a PAIR of node tiles (consecutive tiles continue the 8-row group stride)
constexpr uint32_t kTcInstrDesc256 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
__device__ __forceinline__ void tc_mma256(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kTcInstrDesc256), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// 32 lanes x 32 consecutive columns: thread i of the warp receives row (lane base + i)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Error bound of one accumulator element against the true |n|^2 - 2 q.n: dropped cross terms of the
// bf16 split (3.1 * 2^-18 per product, doubled by the -2) + 66 accumulation steps of at most 2^-22 of the running
// magnitude each (twice the fp32 unit round-off: the tensor core may truncate) + 25 % slack.
__device__ __forceinline__ float tc_error_bound(float qnorm, float nmax) {
  const float qn = qnorm * nmax;
  return 1.25f * (2.4e-5f * qn + 1.6e-5f * (fmaf(nmax, nmax, 2.02f * qn)) + 1e-7f);
}

struct TcSmem {
  static constexpr size_t kA = 0;
  static constexpr size_t kB = kA + TC_TILE_BYTES;
  static constexpr size_t kQd = kB + (size_t)TC_STAGES * TC_TILE_BYTES;               // double [128][21]
  static constexpr size_t kCandD = kQd + sizeof(double) * TC_M * NWB_NJ;              // double [TC_CAND]
  static constexpr size_t kBestKey = kCandD + sizeof(double) * TC_CAND;               // u64 [128]
  static constexpr size_t kOldKey = kBestKey + 8 * TC_M;                              // u64 [128]
  static constexpr size_t kBars = kOldKey + 8 * TC_M;                                 // u64 [TC_STAGES + 2]
  static constexpr size_t kCandNode = kBars + 8 * (TC_STAGES + 2);                    // int [TC_CAND]
  static constexpr size_t kCandQ = kCandNode + 4 * TC_CAND;                           // int [TC_CAND]
  static constexpr size_t kCandV = kCandQ + 4 * TC_CAND;                              // float [TC_CAND]
  static constexpr size_t kMPub = kCandV + 4 * TC_CAND;                               // float [4][128]
  static constexpr size_t kE2 = kMPub + 4 * 4 * TC_M;                                 // float [128]
  static constexpr size_t kBestIdx = kE2 + 4 * TC_M;                                  // int [128]
  static constexpr size_t kIdxTmp = kBestIdx + 4 * TC_M;                              // int [128]
  static constexpr size_t kMisc = kIdxTmp + 4 * TC_M;                                 // int [8]: cand_n, tmem base, last flag
  static constexpr size_t kBytes = kMisc + 32;
};

__device__ __forceinline__ double tc_exact_dist(const double* __restrict__ aos, const double* qrow, int64_t node) {
  const double* row = aos + node * NWB_NJ;
  double r[NWB_NJ];
#pragma unroll
  for (int j = 0; j < NWB_NJ; ++j) r[j] = row[j];
  double sum = 0.0;
#pragma unroll
  for (int j = 0; j < NWB_NJ; ++j) {
    const double d = __dsub_rn(qrow[j], r[j]);
    sum = __dadd_rn(sum, __dmul_rn(d, d));               // separately rounded multiply and add (RP:361-366)
  }
  return sqrt(sum);
}

// Thread layout: 16 warps.  Warp w reads TMEM lanes 32 (w & 3) .. +31 (the lane quadrant the hardware lets it touch)
// and the 32 columns (w >> 2) of every accumulator tile, so a query has four threads, one per column chunk, each with
// its own running minimum (a candidate test against a minimum that is too LARGE only lists more nodes, never fewer).
// Four warps per scheduler hide the TMEM-load and dependent-issue latencies that one warp per scheduler exposed
// (profiles/r02_nn_tc_first_*: 9.5 cycles per issued instruction with 4 warps).
__global__ void __launch_bounds__(TC_THREADS, 1)
nn_tc_kernel(const unsigned char* __restrict__ tc, const double* __restrict__ aos, const unsigned* __restrict__ max_nn_bits,
             int64_t N, const double* __restrict__ q, int64_t nq, double* __restrict__ part_d, int32_t* __restrict__ part_i,
             unsigned* __restrict__ tickets, int32_t* __restrict__ idx_out, double* __restrict__ dist_out,
             float* __restrict__ debug_tile) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* a_tile = smem + TcSmem::kA;
  unsigned char* b_ring = smem + TcSmem::kB;
  double* qd = reinterpret_cast<double*>(smem + TcSmem::kQd);
  double* cand_d = reinterpret_cast<double*>(smem + TcSmem::kCandD);
  unsigned long long* best_key = reinterpret_cast<unsigned long long*>(smem + TcSmem::kBestKey);
  unsigned long long* old_key = reinterpret_cast<unsigned long long*>(smem + TcSmem::kOldKey);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TcSmem::kBars);        // [0, STAGES): tile landed; [STAGES, +2): MMA done
  int* cand_node = reinterpret_cast<int*>(smem + TcSmem::kCandNode);
  int* cand_q = reinterpret_cast<int*>(smem + TcSmem::kCandQ);
  float* cand_v = reinterpret_cast<float*>(smem + TcSmem::kCandV);
  float* m_pub = reinterpret_cast<float*>(smem + TcSmem::kMPub);             // [4 chunks][128 queries]
  float* e2s = reinterpret_cast<float*>(smem + TcSmem::kE2);
  int* best_idx = reinterpret_cast<int*>(smem + TcSmem::kBestIdx);
  int* idx_tmp = reinterpret_cast<int*>(smem + TcSmem::kIdxTmp);
  int* misc = reinterpret_cast<int*>(smem + TcSmem::kMisc);
  int& cand_n = misc[0];
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&misc[1]);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ql = (warp & 3) * 32 + lane;                     // the query (= TMEM lane) this thread reads
  const int chunk = warp >> 2;                               // its 32 columns of every tile
  const bool owner = tid < TC_M;                             // warps 0-3: tid == ql, they also own the per-query state
  const int slice = blockIdx.x, n_slices = gridDim.x;
  const int64_t q0 = (int64_t)blockIdx.y * TC_M;
  const bool q_live = q0 + ql < nq;

  // ---- set-up: barriers, tensor memory, the query tile ---------------------------------------------------------
  if (tid == 0) {
    for (int b = 0; b < TC_STAGES; ++b) tc_mbar_init(&bars[b], 1);
    tc_mbar_init(&bars[TC_STAGES], 1);
    tc_mbar_init(&bars[TC_STAGES + 1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    cand_n = 0;
  }
  if (warp == 0) {                                           // 2 accumulators x 128 columns
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // node tiles of this slice: contiguous range
  const int64_t tiles_total = (N + TC_N - 1) / TC_N;
  const int64_t per = (tiles_total + n_slices - 1) / n_slices;
  const int64_t tile_lo = min(tiles_total, (int64_t)slice * per), tile_hi = min(tiles_total, tile_lo + per);
  const int T = (int)(tile_hi - tile_lo);
  if (tid == 0) {                                            // start the node stream before anything else needs it
    for (int s = 0; s < TC_STAGES && s < T; ++s) {
      tc_mbar_expect_tx(&bars[s], TC_TILE_BYTES);
      tc_bulk_g2s(b_ring + (size_t)s * TC_TILE_BYTES, tc + (size_t)(tile_lo + s) * TC_TILE_BYTES, TC_TILE_BYTES, &bars[s]);
    }
  }
  // query row -> fp64 copy for the exact pass, bf16 split for the MMA (slot map: tree_internal.h)
  if (owner) {
    float qnorm2 = 0.f;
    __nv_bfloat16 row[TC_K];
#pragma unroll
    for (int j = 0; j < NWB_NJ; ++j) {
      const double xd = q_live ? q[(q0 + tid) * NWB_NJ + j] : 0.0;
      qd[tid * NWB_NJ + j] = xd;
      const float x = (float)xd;
      qnorm2 = fmaf(x, x, qnorm2);
      const __nv_bfloat16 hi = __float2bfloat16_rn(x);
      const __nv_bfloat16 mid = __float2bfloat16_rn(x - __bfloat162float(hi));
      const __nv_bfloat16 hi2 = __float2bfloat16_rn(-2.f * __bfloat162float(hi));      // exact: power of two
      const __nv_bfloat16 mid2 = __float2bfloat16_rn(-2.f * __bfloat162float(mid));
      row[j] = hi2;
      row[NWB_NJ + j] = hi2;
      row[2 * NWB_NJ + j] = mid2;
    }
    row[63] = row[64] = row[65] = __float2bfloat16_rn(q_live ? 1.f : 0.f);
#pragma unroll
    for (int k = 66; k < TC_K; ++k) row[k] = __float2bfloat16_rn(0.f);
    unsigned char* base = a_tile + (size_t)(tid >> 3) * (TC_KB * 128) + (size_t)(tid & 7) * 16;
#pragma unroll
    for (int kb = 0; kb < TC_KB; ++kb) *reinterpret_cast<uint4*>(base + kb * 128) = *reinterpret_cast<const uint4*>(&row[kb * 8]);
    const float nmax = sqrtf(__uint_as_float(*max_nn_bits));
    e2s[tid] = 2.f * tc_error_bound(sqrtf(qnorm2) * (1.f + 1e-6f), nmax);
    best_key[tid] = 0x7FF0000000000000ull;                   // +inf
    best_idx[tid] = -1;
  }
  m_pub[chunk * TC_M + ql] = 3.0e38f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the query tile is read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t a_addr = tc_smem_u32(a_tile);
  const float e2 = e2s[ql];

  auto mma_issue = [&](int t) {                              // thread 0: D[t % 2] = A . B(tile t)^T
    const int stage = t % TC_STAGES;
    tc_mbar_wait(&bars[stage], (uint32_t)((t / TC_STAGES) & 1));
    tc_fence_after();
    const uint32_t b_addr = tc_smem_u32(b_ring + (size_t)stage * TC_TILE_BYTES);
    const uint32_t d_addr = tmem_base + (uint32_t)((t & 1) * TC_N);
#pragma unroll
    for (int ks = 0; ks < TC_K / 16; ++ks)                    // one step = 16 slots = two 16-byte chunks = 256 B further
      tc_mma(d_addr, tc_smem_desc(a_addr + ks * 256), tc_smem_desc(b_addr + ks * 256), ks > 0 ? 1u : 0u);
    tc_commit(&bars[TC_STAGES + (t & 1)]);
  };

  float m_run = 3.0e38f;
  if (tid == 0 && T > 0) mma_issue(0);

  // Fold (key = bits of an exact fp64 distance, node) contributions into (best_key, best_idx) per query,
  // lexicographically: smaller distance first, lower index on equal distance.  Collective: every thread calls it
  // with the same number of rounds; `contribute(round, ql_out, key_out, node_out)` returns false for "nothing".
  auto fold_begin = [&]() {
    if (owner) {
      old_key[tid] = best_key[tid];
      idx_tmp[tid] = 0x7FFFFFFF;
    }
    __syncthreads();
  };
  auto fold_end = [&]() {
    __syncthreads();
    if (owner) {
      if (best_key[tid] < old_key[tid]) best_idx[tid] = idx_tmp[tid];
      else if (idx_tmp[tid] != 0x7FFFFFFF && (best_idx[tid] < 0 || idx_tmp[tid] < best_idx[tid])) best_idx[tid] = idx_tmp[tid];
    }
    __syncthreads();
  };

  // exact fp64 evaluation of the listed candidates that are still under their query's bound
  auto flush = [&]() {
    const int n_c = min(cand_n, TC_CAND);
    fold_begin();
    for (int c = tid; c < n_c; c += TC_THREADS) {
      const int cq = cand_q[c];
      const float mq = fminf(fminf(m_pub[cq], m_pub[TC_M + cq]), fminf(m_pub[2 * TC_M + cq], m_pub[3 * TC_M + cq]));
      double d = __longlong_as_double(0x7FF0000000000000ll);
      if (cand_v[c] <= mq + e2s[cq]) {
        d = tc_exact_dist(aos, qd + cq * NWB_NJ, cand_node[c]);
        atomicMin(&best_key[cq], (unsigned long long)__double_as_longlong(d));     // d >= 0: bit order = value order
      }
      cand_d[c] = d;
    }
    __syncthreads();
    for (int c = tid; c < n_c; c += TC_THREADS) {
      const int cq = cand_q[c];
      if ((unsigned long long)__double_as_longlong(cand_d[c]) == best_key[cq]) atomicMin(&idx_tmp[cq], cand_node[c]);
    }
    fold_end();
    if (tid == 0) cand_n = 0;
    __syncthreads();
  };

  for (int t = 0; t < T; ++t) {
    if (tid == 0 && t + 1 < T) mma_issue(t + 1);            // runs on the tensor core while tile t is read back
    tc_mbar_wait(&bars[TC_STAGES + (t & 1)], (uint32_t)((t >> 1) & 1));
    tc_fence_after();
    const int64_t n0 = (tile_lo + t) * TC_N + chunk * 32;    // first node of this thread's columns
    const bool full_chunk = n0 + 32 <= N;
    bool want_flush = false, overflow = false;
    float v[32];
    tc_ld32(tmem_base + (uint32_t)((t & 1) * TC_N + chunk * 32) + ((uint32_t)((warp & 3) * 32) << 16), v);
    if (debug_tile && slice == 0 && blockIdx.y == 0 && t == 0) {      // diagnostics: raw accumulator of the first tile
#pragma unroll
      for (int u = 0; u < 32; ++u) debug_tile[ql * TC_N + chunk * 32 + u] = v[u];
    }
    // minimum of the chunk as a tree (independent operations; a serial chain costs 31 dependent latencies); the
    // second level r8[k] = min of columns {k, k+8, k+16, k+24} is kept for the candidate search
    float r16[16], r8[8];
#pragma unroll
    for (int u = 0; u < 16; ++u) r16[u] = fminf(v[u], v[u + 16]);
#pragma unroll
    for (int u = 0; u < 8; ++u) r8[u] = fminf(r16[u], r16[u + 8]);
    const float cm = fminf(fminf(fminf(r8[0], r8[1]), fminf(r8[2], r8[3])), fminf(fminf(r8[4], r8[5]), fminf(r8[6], r8[7])));
    if (cm <= m_run + e2) {                                   // rare: a new running minimum or a near tie in this chunk
      // The whole warp takes this path when ONE lane does and the CTA waits for it at the tile barrier, so it is kept
      // short: lower the running minimum first, then list what lies within the band of it - 8 group tests, and the 4
      // members only of a group that passes.  (A threshold above the final minimum + band lists a superset.)
      if (full_chunk) m_run = fminf(m_run, cm);
      else {
#pragma unroll
        for (int u = 0; u < 32; ++u)
          if (n0 + u < N) m_run = fminf(m_run, v[u]);
      }
      const float thr = m_run + e2;
      auto push = [&](int u, float x) {
        if (x <= thr && q_live && (full_chunk || n0 + u < N)) {
          const int slot = atomicAdd(&cand_n, 1);
          if (slot < TC_CAND) {
            cand_node[slot] = (int)(n0 + u);
            cand_q[slot] = ql;
            cand_v[slot] = x;
          } else {
            overflow = true;
          }
          want_flush = want_flush || slot >= TC_CAND / 2;
        }
      };
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (r8[k] <= thr) {
          push(k, v[k]);
          push(k + 8, v[k + 8]);
          push(k + 16, v[k + 16]);
          push(k + 24, v[k + 24]);
        }
      m_pub[chunk * TC_M + ql] = m_run;
    }
    __syncwarp();
    tc_fence_before();
    const int act = __syncthreads_or((want_flush ? 1 : 0) | (overflow ? 2 : 0));
    if (tid == 0 && t + TC_STAGES < T) {                     // the MMA of tile t is complete: its stage is free again
      const int stage = t % TC_STAGES;
      tc_mbar_expect_tx(&bars[stage], TC_TILE_BYTES);
      tc_bulk_g2s(b_ring + (size_t)stage * TC_TILE_BYTES, tc + (size_t)(tile_lo + t + TC_STAGES) * TC_TILE_BYTES, TC_TILE_BYTES,
                  &bars[stage]);
    }
    if (act & 2) {
      // the list could not hold this tile's candidates: every thread evaluates its 32 (query, node) pairs exactly
      unsigned long long bk = 0x7FF0000000000000ull;
      int bi = 0x7FFFFFFF;
      if (q_live) {
        for (int u = 0; u < 32; ++u) {
          const int64_t node = n0 + u;
          if (node >= N) break;
          const unsigned long long k = (unsigned long long)__double_as_longlong(tc_exact_dist(aos, qd + ql * NWB_NJ, node));
          if (k < bk) { bk = k; bi = (int)node; }             // ascending node order: the first minimum has the lowest index
        }
      }
      fold_begin();
      if (bi != 0x7FFFFFFF) atomicMin(&best_key[ql], bk);
      __syncthreads();
      if (bi != 0x7FFFFFFF && bk == best_key[ql]) atomicMin(&idx_tmp[ql], bi);
      fold_end();
    }
    if (act) flush();
  }
  flush();

  // ---- partial of this slice, then the last CTA of the query group folds the slices --------------------------------
  if (owner && q_live) {
    part_d[(q0 + tid) * n_slices + slice] = __longlong_as_double((long long)best_key[tid]);
    part_i[(q0 + tid) * n_slices + slice] = best_idx[tid];
  }
  tc_fence_before();
  __threadfence();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
  if (tid == 0) misc[2] = atomicAdd(&tickets[blockIdx.y], 1u) == (unsigned)n_slices - 1 ? 1 : 0;
  __syncthreads();
  if (!misc[2]) return;
  __threadfence();
  if (owner && q_live) {
    double d = 10000.0;                                       // RP:347: best_norm starts at 10000.0, strict <
    int i = -1;
    for (int s = 0; s < n_slices; ++s) {
      const double ds = part_d[(q0 + tid) * n_slices + s];
      const int is = part_i[(q0 + tid) * n_slices + s];
      if (is >= 0 && (ds < d || (ds == d && i >= 0 && is < i))) { d = ds; i = is; }
    }
    idx_out[q0 + tid] = i;
    if (dist_out) dist_out[q0 + tid] = d;
  }
  if (tid == 0) tickets[blockIdx.y] = 0u;                    // ready for the next call
}


// =====================================================================================================================
// Two-phase form (default).  The one-pass kernel above interleaves the filter with its exact candidates: a lane that
// meets a new running minimum drags its warp - and, through the per-tile barrier, the CTA - into the candidate path,
// which with 128 queries per CTA happened in nearly every tile (profiles/r02_nn_tc_onepass_*: 46 % of all warp time
// waiting at that barrier).  Here the tensor-core pass does NOTHING but reduce:
//
//   phase A  nn_tc_min_kernel: a TMA thread, an MMA thread and 16 reduction warps, coupled only by mbarriers (stage
//            landed / stage free / accumulator full / accumulator free) - no CTA-wide barrier in the loop.  A reduction
//            thread reads 64 accumulator columns (one query x one half tile of 64 nodes), takes their minimum with a
//            tree of FMNMX and stores that ONE float: chunk_min[half-tile row][query].  4 B per (query, 64 nodes).
//   phase B  nn_tc_exact_kernel: per query, threshold = (smallest chunk minimum) + 2E; every chunk whose minimum is
//            under it is evaluated exactly - 64 fp64 distances by one warp, the reference's rounding sequence - and
//            folded lexicographically (distance, index).  The true nearest node n* has v(n*) <= v_min + 2E, so its chunk
//            is always among them; so is every node that ties with it.  A flood of near ties only makes phase B longer.
// =====================================================================================================================
constexpr int TC2_EPI_WARPS = 16;
constexpr int TC2_THREADS = (TC2_EPI_WARPS + 2) * 32;          // + TMA warp + MMA warp (one elected lane each)
constexpr int TC2_DSTAGES = 3;                                 // ring of double stages: two consecutive node tiles (40 KB) each
struct Tc2Smem {
  static constexpr size_t kA = 0;
  static constexpr size_t kB = kA + TC_TILE_BYTES;
  static constexpr size_t kBars = kB + (size_t)TC2_DSTAGES * 2 * TC_TILE_BYTES;   // full[S] empty[S] done[2] accfree[2]
  static constexpr size_t kMisc = kBars + 8 * (2 * TC2_DSTAGES + 4);
  static constexpr size_t kBytes = kMisc + 16;
};
__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}

// Node tiles are processed in PAIRS: one tcgen05.mma with N = 256 covers two consecutive tiles (their 8-row groups
// continue at the same stride in shared memory), five K-steps per pair.  Two accumulator sets of 256 columns: the MMA
// warp fills one while the 16 reduction warps read the other, so all 512 TMEM columns are in use.
//
// What bounds it (measured, profiles/r02_nn_tc_*): every accumulator element has to come out of tensor memory once -
// 128 lanes x 256 columns x 4 B = 128 KB per pair - and the 16 warps drain it at ~120 B per clock and SM, i.e. ~1100
// cycles per pair against 640 cycles of MMA (K = 80 is short).  A variant with TWO query groups per CTA (half the L2
// stream per MMA) measured the same ~1200 cycles per 128 KB of accumulator and was dropped: the pass is TMEM-read bound,
// 1024 x 10^6 = 4.2 GB of accumulator through 148 x ~120 B/clk = ~115 us.  (sm_103 has tcgen05.ld.red, a TMEM load with
// a fused min; sm_100a does not.)
__global__ void __launch_bounds__(TC2_THREADS, 1)
nn_tc_min_kernel(const unsigned char* __restrict__ tc, int64_t N, const double* __restrict__ q, int64_t nq, int64_t qp,
                 float* __restrict__ chunk_min /*[tiles * 2][qp]: one row per half tile of 64 nodes*/, float* __restrict__ slice_min /*[slices * 4][qp]*/,
                 float* __restrict__ debug_tile) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* a_tile = smem + Tc2Smem::kA;
  unsigned char* b_ring = smem + Tc2Smem::kB;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Tc2Smem::kBars);
  uint64_t* empty = full + TC2_DSTAGES;
  uint64_t* done = empty + TC2_DSTAGES;
  uint64_t* accfree = done + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Tc2Smem::kMisc);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slice = blockIdx.x, n_slices = gridDim.x;
  const int64_t q0 = (int64_t)blockIdx.y * TC_M;

  if (tid == 0) {
    for (int b = 0; b < TC2_DSTAGES; ++b) { tc_mbar_init(&full[b], 1); tc_mbar_init(&empty[b], 1); }
    for (int b = 0; b < 2; ++b) { tc_mbar_init(&done[b], 1); tc_mbar_init(&accfree[b], TC2_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // this slice's node tiles: a contiguous range of whole PAIRS (the copy holds one spare tile past the end of the tree)
  const int64_t pairs_total = ((N + TC_N - 1) / TC_N + 1) / 2;
  const int64_t per = (pairs_total + n_slices - 1) / n_slices;
  const int64_t pair_lo = min(pairs_total, (int64_t)slice * per), pair_hi = min(pairs_total, pair_lo + per);
  const int P = (int)(pair_hi - pair_lo);
  if (tid < TC_M) {                                            // query row -> bf16 split (slot map: tree_internal.h)
    const bool live = q0 + tid < nq;
    __nv_bfloat16 row[TC_K];
#pragma unroll
    for (int j = 0; j < NWB_NJ; ++j) {
      const float x = live ? (float)q[(q0 + tid) * NWB_NJ + j] : 0.f;
      const __nv_bfloat16 hi = __float2bfloat16_rn(x);
      const __nv_bfloat16 mid = __float2bfloat16_rn(x - __bfloat162float(hi));
      const __nv_bfloat16 hi2 = __float2bfloat16_rn(-2.f * __bfloat162float(hi));
      const __nv_bfloat16 mid2 = __float2bfloat16_rn(-2.f * __bfloat162float(mid));
      row[j] = hi2;
      row[NWB_NJ + j] = hi2;
      row[2 * NWB_NJ + j] = mid2;
    }
    row[63] = row[64] = row[65] = __float2bfloat16_rn(live ? 1.f : 0.f);
#pragma unroll
    for (int k = 66; k < TC_K; ++k) row[k] = __float2bfloat16_rn(0.f);
    unsigned char* base = a_tile + (size_t)(tid >> 3) * (TC_KB * 128) + (size_t)(tid & 7) * 16;
#pragma unroll
    for (int kb = 0; kb < TC_KB; ++kb) *reinterpret_cast<uint4*>(base + kb * 128) = *reinterpret_cast<const uint4*>(&row[kb * 8]);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == TC2_EPI_WARPS) {                                 // ---- TMA thread: keeps the ring full --------------------
    if (lane == 0)
      for (int p = 0; p < P; ++p) {
        const int stage = p % TC2_DSTAGES;
        if (p >= TC2_DSTAGES) tc_mbar_wait(&empty[stage], (uint32_t)(((p / TC2_DSTAGES) - 1) & 1));
        tc_mbar_expect_tx(&full[stage], 2 * TC_TILE_BYTES);
        tc_bulk_g2s(b_ring + (size_t)stage * 2 * TC_TILE_BYTES, tc + (size_t)(pair_lo + p) * 2 * TC_TILE_BYTES, 2 * TC_TILE_BYTES,
                    &full[stage]);
      }
  } else if (warp == TC2_EPI_WARPS + 1) {                      // ---- MMA warp ---------------------------------------------
    // The whole warp runs the loop (waits included) and ONE elected lane issues: with warp-uniform control flow the
    // descriptors live in uniform registers.  (Guarding the loop itself with lane == 0 made the compiler rebuild every
    // operand through R2UR / ELECT sequences: ~160 cycles per issued MMA, four times the MMA itself.)
    const uint64_t a_desc = tc_smem_desc(tc_smem_u32(a_tile));
    for (int p = 0; p < P; ++p) {
      const int stage = p % TC2_DSTAGES, set = p & 1;
      tc_mbar_wait(&full[stage], (uint32_t)((p / TC2_DSTAGES) & 1));
      if (p >= 2) tc_mbar_wait(&accfree[set], (uint32_t)(((p >> 1) - 1) & 1));
      tc_fence_after();
      const uint64_t b_desc = tc_smem_desc(tc_smem_u32(b_ring + (size_t)stage * 2 * TC_TILE_BYTES));
      const uint32_t d0 = tmem_base + (uint32_t)(set * 2 * TC_N);
      if (tc_elect_one()) {
#pragma unroll
        for (int ks = 0; ks < TC_K / 16; ++ks)                 // a K-step is 256 B further = 16 descriptor units
          tc_mma256(d0, a_desc + ks * 16, b_desc + ks * 16, ks > 0 ? 1u : 0u);
        tc_commit(&empty[stage]);                              // both arrive when the MMAs above have completed
        tc_commit(&done[set]);
      }
      __syncwarp();
    }
  } else {                                                     // ---- 16 reduction warps ----------------------------------
    // warp -> TMEM lane quadrant (its 32 queries) x one HALF TILE (64 accumulator columns = 64 nodes) of the pair;
    // it stores ONE float per query and half tile: row (tile * 2 + half) of chunk_min
    const int quad = warp & 3, part = warp >> 2, ql = quad * 32 + lane;      // part = tile-in-pair * 2 + half
    float m_run = 3.0e38f;
    float* out = chunk_min + ((size_t)pair_lo * 4 + part) * qp + q0 + ql;
    for (int p = 0; p < P; ++p) {
      const int set = p & 1;
      tc_mbar_wait(&done[set], (uint32_t)((p >> 1) & 1));
      tc_fence_after();
      float v[2][32];
      const uint32_t taddr = tmem_base + (uint32_t)(set * 2 * TC_N + part * 64) + ((uint32_t)(quad * 32) << 16);
      tc_ld32(taddr, v[0]);
      tc_ld32(taddr + 32, v[1]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) tc_mbar_arrive(&accfree[set]);            // the accumulators may be overwritten: values are in registers
      const int64_t n0 = (pair_lo + p) * 2 * TC_N + part * 64;
      if (n0 + 64 > N) {                                       // tail of the tree: columns past the end never win
#pragma unroll
        for (int u = 0; u < 64; ++u)
          if (n0 + u >= N) v[u >> 5][u & 31] = 3.0e38f;
      }
      if (debug_tile && slice == 0 && blockIdx.y == 0 && p == 0 && part < 2) {
#pragma unroll
        for (int u = 0; u < 64; ++u) debug_tile[ql * TC_N + part * 64 + u] = v[u >> 5][u & 31];
      }
      // minimum of the 64 values as a tree of 3-input minima (FMNMX3: the min runs on the half-rate ALU pipe, which a
      // 2-input tree kept 60 % busy - profiles/r02_nn_tc_min_metrics.txt): 64 -> 22 -> 8 -> 3 -> 1
      float a[22];
#pragma unroll
      for (int u = 0; u < 21; ++u) a[u] = fminf(fminf(v[0][u], v[1][u]), (u < 11 ? v[0][21 + u] : v[1][10 + u]));
      a[21] = v[1][31];
      float b8[8];
#pragma unroll
      for (int u = 0; u < 7; ++u) b8[u] = fminf(fminf(a[3 * u], a[3 * u + 1]), a[3 * u + 2]);
      b8[7] = a[21];
      const float cm = fminf(fminf(fminf(b8[0], b8[1]), b8[2]), fminf(fminf(b8[3], b8[4]), fminf(fminf(b8[5], b8[6]), b8[7])));
      m_run = fminf(m_run, cm);
      out[(size_t)p * 4 * qp] = cm;                            // 32 consecutive floats per warp: one 128-byte store
    }
    slice_min[((size_t)slice * 4 + part) * qp + q0 + ql] = m_run;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// Phase B.  One thread per query scans that query's chunk minima of a range of chunk rows, TC2_WINDOW rows at a time with
// all loads of a window in flight (the scan is a 128 MB stream at 1024 x 10^6: memory-level parallelism is what it
// needs); chunks under the threshold are listed, and between two windows the block evaluates the listed chunks
// exactly, one warp per chunk.  The list holds a whole window (128 x TC2_WINDOW entries), so it cannot overflow.
constexpr int TC2_WINDOW = 16;
// threshold per query = smallest accumulator value over the whole tree + 2E (computed once, not once per block of phase B)
__global__ void __launch_bounds__(TC_M)
nn_tc_threshold_kernel(const float* __restrict__ slice_min, int n_slice_rows, int64_t qp, const unsigned* __restrict__ max_nn_bits,
                       const double* __restrict__ q, int64_t nq, float* __restrict__ thr_out) {
  const int64_t k = (int64_t)blockIdx.x * TC_M + threadIdx.x;
  const bool live = k < nq;
  float qnorm2 = 0.f;
#pragma unroll
  for (int j = 0; j < NWB_NJ; ++j) {
    const float x = live ? (float)q[k * NWB_NJ + j] : 0.f;
    qnorm2 = fmaf(x, x, qnorm2);
  }
  float vmin = 3.0e38f;
#pragma unroll 8
  for (int r = 0; r < n_slice_rows; ++r) vmin = fminf(vmin, slice_min[(size_t)r * qp + k]);
  thr_out[k] = vmin + 2.f * tc_error_bound(sqrtf(qnorm2) * (1.f + 1e-6f), sqrtf(__uint_as_float(*max_nn_bits)));
}

__global__ void __launch_bounds__(TC_M)
nn_tc_exact_kernel(const float* __restrict__ chunk_min, const float* __restrict__ thr_in, int n_slice_rows, int64_t chunk_rows,
                   int64_t qp, const double* __restrict__ aos, const unsigned* __restrict__ max_nn_bits, int64_t N,
                   const double* __restrict__ q, int64_t nq, double* __restrict__ part_d, int32_t* __restrict__ part_i,
                   unsigned* __restrict__ tickets, int32_t* __restrict__ idx_out, double* __restrict__ dist_out) {
  __shared__ unsigned long long best_key[TC_M], old_key[TC_M], cand_key[TC_M * TC2_WINDOW];
  __shared__ int best_idx[TC_M], idx_tmp[TC_M], cand_idx[TC_M * TC2_WINDOW];
  __shared__ int cand_row[TC_M * TC2_WINDOW];
  __shared__ unsigned char cand_q[TC_M * TC2_WINDOW];
  __shared__ int cand_n;
  __shared__ bool last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int range = blockIdx.x, n_ranges = gridDim.x;
  const int64_t q0 = (int64_t)blockIdx.y * TC_M;
  const bool live = q0 + tid < nq;
  const float thr = thr_in[q0 + tid];
  (void)n_slice_rows;
  (void)max_nn_bits;
  best_key[tid] = 0x7FF0000000000000ull;
  best_idx[tid] = -1;
  if (tid == 0) cand_n = 0;
  __syncthreads();
  const int64_t per = (chunk_rows + n_ranges - 1) / n_ranges;
  const int64_t r_lo = min(chunk_rows, (int64_t)range * per), r_hi = min(chunk_rows, r_lo + per);
  const float* col = chunk_min + q0 + tid;
  for (int64_t w0 = r_lo; w0 < r_hi; w0 += TC2_WINDOW) {
    float v[TC2_WINDOW];
#pragma unroll
    for (int k = 0; k < TC2_WINDOW; ++k) v[k] = w0 + k < r_hi ? __ldcs(col + (size_t)(w0 + k) * qp) : 3.0e38f;   // streamed once
    bool any = false;
#pragma unroll
    for (int k = 0; k < TC2_WINDOW; ++k)
      if (v[k] <= thr && live) {
        const int slot = atomicAdd(&cand_n, 1);
        cand_row[slot] = (int)(w0 + k);
        cand_q[slot] = (unsigned char)tid;
        any = true;
      }
    if (!__syncthreads_or(any ? 1 : 0)) continue;              // nothing listed in this window (the common case)
    const int n_c = cand_n;
    old_key[tid] = best_key[tid];
    idx_tmp[tid] = 0x7FFFFFFF;
    __syncthreads();
    for (int c = warp; c < n_c; c += TC_M / 32) {              // one warp per listed half tile: 64 exact distances
      const int cq = cand_q[c];
      unsigned long long key = 0x7FF0000000000000ull;
      int idx = 0x7FFFFFFF;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t node = (int64_t)cand_row[c] * 64 + h * 32 + lane;
        if (node < N) {
          const unsigned long long k = (unsigned long long)__double_as_longlong(tc_exact_dist(aos, q + (q0 + cq) * NWB_NJ, node));
          if (k < key) { key = k; idx = (int)node; }           // ascending node order per lane: first minimum = lowest index
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {                       // lexicographic (distance, index) minimum of the warp
        const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
        if (k2 < key || (k2 == key && i2 < idx)) { key = k2; idx = i2; }
      }
      if (lane == 0) {
        cand_key[c] = key;
        cand_idx[c] = idx;
        atomicMin(&best_key[cq], key);
      }
    }
    __syncthreads();
    for (int c = tid; c < n_c; c += TC_M)
      if (cand_key[c] == best_key[cand_q[c]]) atomicMin(&idx_tmp[cand_q[c]], cand_idx[c]);
    __syncthreads();
    if (best_key[tid] < old_key[tid]) best_idx[tid] = idx_tmp[tid];
    else if (idx_tmp[tid] != 0x7FFFFFFF && (best_idx[tid] < 0 || idx_tmp[tid] < best_idx[tid])) best_idx[tid] = idx_tmp[tid];
    if (tid == 0) cand_n = 0;
    __syncthreads();
  }
  if (live) {
    part_d[(q0 + tid) * n_ranges + range] = __longlong_as_double((long long)best_key[tid]);
    part_i[(q0 + tid) * n_ranges + range] = best_idx[tid];
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) last = atomicAdd(&tickets[blockIdx.y], 1u) == (unsigned)n_ranges - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (live) {
    double d = 10000.0;                                        // RP:347: best_norm starts at 10000.0, strict <
    int i = -1;
    for (int s0 = 0; s0 < n_ranges; s0 += 8) {                 // 8 partials in flight: the fold is a chain of L2 round trips
      double ds[8];
      int is[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const bool in = s0 + k < n_ranges;
        ds[k] = in ? part_d[(q0 + tid) * n_ranges + s0 + k] : 0.0;
        is[k] = in ? part_i[(q0 + tid) * n_ranges + s0 + k] : -1;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (is[k] >= 0 && (ds[k] < d || (ds[k] == d && i >= 0 && is[k] < i))) { d = ds[k]; i = is[k]; }
    }
    idx_out[q0 + tid] = i;
    if (dist_out) dist_out[q0 + tid] = d;
  }
  if (tid == 0) tickets[blockIdx.y] = 0u;
}

}  // namespace nwb

using namespace nwb;

static cudaError_t nn_tc_onepass_launch(nwb_tree* t, const double* q_dev, int64_t nq, int32_t* idx_dev, double* dist_dev,
                                        float* debug_tile, int* n_launches);

static cudaError_t tc_reserve_partials(nwb_tree* t, size_t need, size_t groups, cudaStream_t s) {
  cudaError_t e;
  if (need > t->part_cap) {
    if (t->part_d) cudaFree(t->part_d);
    if (t->part_i) cudaFree(t->part_i);
    t->part_d = nullptr;
    t->part_i = nullptr;
    t->part_cap = 0;
    if ((e = cudaMalloc(&t->part_d, sizeof(double) * need)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&t->part_i, sizeof(int32_t) * need)) != cudaSuccess) return e;
    t->part_cap = need;
  }
  if (groups > t->ticket_cap) {
    if (t->tickets) cudaFree(t->tickets);
    t->tickets = nullptr;
    t->ticket_cap = 0;
    const size_t cap = std::max<size_t>(groups, 1024);
    if ((e = cudaMalloc(&t->tickets, sizeof(unsigned) * cap)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(t->tickets, 0, sizeof(unsigned) * cap, s)) != cudaSuccess) return e;
    t->ticket_cap = cap;
  }
  return cudaSuccess;
}

// Two-phase batched scan (see above).  Query groups are worked through in waves so that the chunk-minimum buffer
// (4 B per query and 64 nodes) stays under 1 GiB whatever the batch.
cudaError_t nn_tc_launch(nwb_tree* t, const double* q_dev, int64_t nq, int32_t* idx_dev, double* dist_dev, float* debug_tile,
                         int* n_launches) {
  static const bool onepass = getenv("NWB_NN_TC_ONEPASS") != nullptr;
  if (onepass) return nn_tc_onepass_launch(t, q_dev, nq, idx_dev, dist_dev, debug_tile, n_launches);
  const int64_t groups = (nq + TC_M - 1) / TC_M;
  const int64_t pairs = std::max<int64_t>(1, ((t->size + TC_N - 1) / TC_N + 1) / 2);   // node tiles go through the MMA in pairs
  const int64_t tiles = pairs;                                                            // (scheduling unit of phase A)
  const int64_t chunk_rows = pairs * 4;                                                   // one row per half tile (64 nodes)
  const int sm = std::max(1, t->ctx->sm_count);
  const size_t bytes_per_group = (size_t)chunk_rows * TC_M * sizeof(float);
  const int64_t wave = std::max<int64_t>(1, std::min<int64_t>(groups, (int64_t)(((size_t)1 << 30) / bytes_per_group)));
  const int64_t wave_use = std::min<int64_t>(wave, 65535);
  const int64_t slices = std::max<int64_t>(1, std::min<int64_t>(tiles, sm / std::min<int64_t>(wave_use, sm)));
  const int64_t ranges = std::max<int64_t>(1, std::min<int64_t>((chunk_rows + TC2_WINDOW - 1) / TC2_WINDOW, (4 * sm + wave_use - 1) / wave_use));
  const int64_t qp = wave_use * TC_M;
  cudaStream_t s = t->ctx->stream;
  cudaError_t e;
  const size_t need_chunk = (size_t)chunk_rows * qp * sizeof(float), need_slice = (size_t)(slices * 4 + 1) * qp * sizeof(float);
  if (need_chunk + need_slice > t->tc_scratch_bytes) {
    if (t->tc_scratch) cudaFree(t->tc_scratch);
    t->tc_scratch = nullptr;
    t->tc_scratch_bytes = 0;
    if ((e = cudaMalloc(&t->tc_scratch, need_chunk + need_slice)) != cudaSuccess) return e;
    t->tc_scratch_bytes = need_chunk + need_slice;
  }
  float* chunk_min = reinterpret_cast<float*>(t->tc_scratch);
  float* slice_min = reinterpret_cast<float*>(static_cast<char*>(t->tc_scratch) + need_chunk);
  if ((e = tc_reserve_partials(t, (size_t)wave_use * TC_M * ranges, (size_t)wave_use, s)) != cudaSuccess) return e;
  static std::atomic<bool> configured[64];
  const int dev = t->ctx->device >= 0 && t->ctx->device < 64 ? t->ctx->device : 0;
  if (!configured[dev].load(std::memory_order_relaxed)) {
    if ((e = cudaFuncSetAttribute(nn_tc_min_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tc2Smem::kBytes)) != cudaSuccess) return e;
    configured[dev].store(true, std::memory_order_relaxed);
  }
  int launches = 0;
  for (int64_t g0 = 0; g0 < groups; g0 += wave_use) {
    const int64_t ng = std::min<int64_t>(wave_use, groups - g0);
    const int64_t qoff = g0 * TC_M;
    nn_tc_min_kernel<<<dim3((unsigned)slices, (unsigned)ng), TC2_THREADS, Tc2Smem::kBytes, s>>>(
        t->tc, t->size, q_dev + qoff * NWB_NJ, nq - qoff, qp, chunk_min, slice_min, debug_tile);
    float* thr = slice_min + (size_t)slices * 4 * qp;
    nn_tc_threshold_kernel<<<(unsigned)ng, TC_M, 0, s>>>(slice_min, (int)(slices * 4), qp, t->max_nn, q_dev + qoff * NWB_NJ, nq - qoff, thr);
    nn_tc_exact_kernel<<<dim3((unsigned)ranges, (unsigned)ng), TC_M, 0, s>>>(
        chunk_min, thr, (int)(slices * 4), (((t->size + TC_N - 1) / TC_N + 1) / 2) * 4, qp, t->aos, t->max_nn, t->size, q_dev + qoff * NWB_NJ, nq - qoff,
        t->part_d, t->part_i, t->tickets, idx_dev + qoff, dist_dev ? dist_dev + qoff : nullptr);
    launches += 3;
    debug_tile = nullptr;
  }
  if (n_launches) *n_launches = launches;
  return cudaGetLastError();
}

static cudaError_t nn_tc_onepass_launch(nwb_tree* t, const double* q_dev, int64_t nq, int32_t* idx_dev, double* dist_dev,
                                        float* debug_tile, int* n_launches) {
  const int64_t groups = (nq + TC_M - 1) / TC_M;
  const int64_t tiles = (t->size + TC_N - 1) / TC_N;
  const int sm = std::max(1, t->ctx->sm_count);
  // one CTA per SM (the shared-memory ring allows no more): slices x groups fills the machine once
  int64_t slices = std::max<int64_t>(1, sm / std::min<int64_t>(groups, sm));
  slices = std::max<int64_t>(1, std::min(slices, tiles));
  cudaStream_t s = t->ctx->stream;
  cudaError_t e;
  if ((e = tc_reserve_partials(t, (size_t)groups * TC_M * slices, (size_t)groups, s)) != cudaSuccess) return e;
  static std::atomic<bool> configured[64];
  const int dev = t->ctx->device >= 0 && t->ctx->device < 64 ? t->ctx->device : 0;
  if (!configured[dev].load(std::memory_order_relaxed)) {
    if ((e = cudaFuncSetAttribute(nn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcSmem::kBytes)) != cudaSuccess) return e;
    configured[dev].store(true, std::memory_order_relaxed);
  }
  int launches = 0;
  for (int64_t g0 = 0; g0 < groups; g0 += 65535) {
    const int64_t ng = std::min<int64_t>(65535, groups - g0);
    const int64_t qoff = g0 * TC_M;
    nn_tc_kernel<<<dim3((unsigned)slices, (unsigned)ng), TC_THREADS, TcSmem::kBytes, s>>>(
        t->tc, t->aos, t->max_nn, t->size, q_dev + qoff * NWB_NJ, nq - qoff, t->part_d + (size_t)qoff * slices,
        t->part_i + (size_t)qoff * slices, t->tickets + g0, idx_dev + qoff, dist_dev ? dist_dev + qoff : nullptr, debug_tile);
    ++launches;
    debug_tile = nullptr;
  }
  if (n_launches) *n_launches = launches;
  return cudaGetLastError();
}
