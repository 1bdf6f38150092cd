// tree.cu - device-resident RRT tree + brute-force nearest-neighbour scan for sm_100a.
//
// Replaces struct Node / struct Tree (rrt_connect_planner/include/rrt_connect_planner/rrt_planner.h:30-46:
// an AoS std::vector<Node>, each node a heap std::vector<double>) and
// RRT_Planner::findNearestNeighbour (rrt_connect_planner/src/rrt_planner.cpp:336-419), which takes the
// whole Tree BY VALUE (rrt_planner.h:160) and scans it with one sqrt per node.
//
// Layout in HBM (per tree, capacity C):
//   soa  float  [21][C]   scan copy: one coalesced row per joint            (84 B / node, the bytes the
//                                                                             roofline is quoted on)
//   aos  double [C][21]   master copy, exact re-evaluation + read-back      (168 B / node, touched only
//                                                                             for near-minimum candidates)
//   pred int32  [C]       predecessor_index
// The scan is HBM-bound: every thread keeps 21 independent row loads in flight, computes the fp32
// squared distance, and only when that comes within a rigorous fp32 error bound of its running
// minimum re-evaluates the distance exactly as the reference does (double, separately rounded
// multiply and add, then sqrt, strict <) from the master copy.  Lexicographic (distance, index)
// reductions (warp shuffle -> block -> last-block-done grid fold, one launch) then reproduce "first strict minimum wins".
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <new>

#include "tree_internal.h"

namespace nwb {

constexpr int NN_THREADS = 256;

__global__ void tree_append_kernel(const double* __restrict__ src, const int32_t* __restrict__ pred_src, int64_t n,
                                   float* __restrict__ soa, uint16_t* __restrict__ h16, double* __restrict__ aos,
                                   int32_t* __restrict__ pred, int64_t cap, int64_t size) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * NWB_NJ) return;
  int64_t i = e / NWB_NJ;
  int j = (int)(e - i * NWB_NJ);
  double v = src[e];
  aos[(size + i) * NWB_NJ + j] = v;
  soa[soa_index(size + i, j)] = (float)v;
  h16[soa_index(size + i, j)] = __half_as_ushort(__double2half(v));     // one rounding: |error| <= half an fp16 ulp of |v|
  if (j == 0) pred[size + i] = pred_src ? pred_src[i] : -1;
}

// Tensor-core scan copy of the appended nodes (layout and slot map: tree_internal.h): one thread per node.
__global__ void tree_append_tc_kernel(const double* __restrict__ src, int64_t n, unsigned char* __restrict__ tc, int64_t size,
                                      unsigned* __restrict__ max_nn) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  __nv_bfloat16 row[TC_K];
  double nn = 0.0;
#pragma unroll
  for (int j = 0; j < NWB_NJ; ++j) {
    const double xd = src[i * NWB_NJ + j];
    nn = fma(xd, xd, nn);
    const float x = (float)xd;
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const __nv_bfloat16 mid = __float2bfloat16_rn(x - __bfloat162float(hi));
    row[j] = hi;
    row[NWB_NJ + j] = mid;
    row[2 * NWB_NJ + j] = hi;
  }
  const float nnf = (float)nn;
  const __nv_bfloat16 n0 = __float2bfloat16_rn(nnf);
  const float r1 = nnf - __bfloat162float(n0);
  const __nv_bfloat16 n1 = __float2bfloat16_rn(r1);
  const __nv_bfloat16 n2 = __float2bfloat16_rn(r1 - __bfloat162float(n1));
  row[63] = n0;
  row[64] = n1;
  row[65] = n2;
#pragma unroll
  for (int k = 66; k < TC_K; ++k) row[k] = __float2bfloat16_rn(0.f);
  unsigned char* base = tc + tc_slot_offset(size + i, 0);
#pragma unroll
  for (int kb = 0; kb < TC_KB; ++kb) *reinterpret_cast<uint4*>(base + kb * 128) = *reinterpret_cast<const uint4*>(&row[kb * 8]);
  // max |node|^2 (non-negative floats order like their bit patterns); rounded up so the bound stays a bound
  atomicMax(max_nn, __float_as_uint(nnf * (1.0f + 1e-6f)));
}

__device__ __forceinline__ bool lex_less(double da, int ia, double db, int ib) {
  // smaller distance first; on equal distance the lower index (a found node beats "none" = -1 only
  // through a strictly smaller distance, like the reference's best_norm = 10000.0 start)
  return (da < db) || (da == db && (unsigned)ia < (unsigned)ib);
}

// fp32 error bound of the scan: if node m has the smallest fp32 squared distance s_m, the exactly
// nearest node w satisfies s_w <= s_m + 2 e(s_m) with e(s) = 1.2e-5 sqrt(s) + 3e-6 s + 1e-10: inputs
// rounded to fp32 (|q| <= pi per joint: 2 * 2^-24 * pi * sqrt(21) * d) plus 21 fma roundings (relative
// 21 * 2^-24 of s).  Only nodes under that bound are re-evaluated exactly.
__device__ __forceinline__ float nn_candidate_bound(float m) {
  return (m + 2.4e-5f * sqrtf(m) + 6e-6f * m + 2e-10f) * (1.0f + 1e-6f);
}

// ---- TMA bulk-copy helpers (cp.async.bulk 1-D, completion on an mbarrier) -------------------------
__device__ __forceinline__ uint32_t nn_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void nn_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nn_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void nn_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nn_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void nn_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "NNWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra NNDONE_%=;\n"
      "bra NNWAIT_%=;\n"
      "NNDONE_%=:\n"
      "}\n" ::"r"(nn_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void nn_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(nn_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(nn_smem_u32(bar))
               : "memory");
}

constexpr int NN_TILE = 512;      // nodes per tile: 21 rows x 2 KB = 43 KB per stage
constexpr int NN_STAGES = 4;      // 172 KB of shared memory: one persistent CTA per SM
constexpr int NN_CAND_MAX = 512;  // exact-pass candidates a block lists (a handful per query is typical)

// one pass over the nodes for NN_QT queries; writes one (dist, idx) partial per (query, block).
// Persistent CTAs (one per SM) stream the 21 SoA rows of their tiles through a 4-stage shared-memory
// ring with TMA bulk copies (UBLKCP, mbarrier complete_tx), so ~130 KB of loads are in flight per SM
// while the threads compute on the stage that has landed.  Per tile the block first agrees on the
// running fp32 minimum per query (warp shuffle + shared atomicMin on the float bits), then only the
// nodes within the error bound of that minimum take the exact fp64 path.
// nodes per thread and tile geometry: one query streams 2 nodes per thread through 4 stages of 512 nodes; a query
// group takes 4 nodes per thread (a broadcast load of the query values then feeds twice the arithmetic) through
// 2 stages of 1024 nodes - the same 172 KB of shared memory either way.
template <int QT> struct NnCfg {
  static constexpr int NPT = QT > 1 ? 4 : 2;
  static constexpr int TILE = NN_THREADS * NPT;
  static constexpr int STAGES = (NN_TILE * NN_STAGES) / TILE;
};

template <int QT>
__global__ void __launch_bounds__(NN_THREADS, 1)
nn_scan_kernel(const float* __restrict__ soa, const double* __restrict__ aos, int64_t cap, int64_t N,
               const double* __restrict__ q, int64_t nq, double* __restrict__ part_d, int32_t* __restrict__ part_i,
               unsigned* __restrict__ tickets, int32_t* __restrict__ idx_out, double* __restrict__ dist_out) {
  constexpr int NPT = NnCfg<QT>::NPT, TILE = NnCfg<QT>::TILE, STAGES = NnCfg<QT>::STAGES;
  extern __shared__ __align__(128) unsigned char nn_smem[];
  float* ring = reinterpret_cast<float*>(nn_smem);                           // [STAGES][21][TILE]
  uint64_t* bars = reinterpret_cast<uint64_t*>(nn_smem + sizeof(float) * STAGES * NWB_NJ * TILE);
  __shared__ double qd[QT][NWB_NJ];
  __shared__ float qf[QT][NWB_NJ];
  __shared__ __align__(16) float2 qn2[NWB_NJ][QT];     // (-q, -q): second operand of the packed subtract
  __shared__ unsigned run_min[QT];                    // running fp32 minimum of the block (float bits, s >= 0)
  __shared__ double red_d[QT][NN_THREADS / 32];
  __shared__ int red_i[QT][NN_THREADS / 32];
  __shared__ int overflow;                            // some thread saw more than two candidates for a query
  const int tid = threadIdx.x;
  const int64_t q0 = (int64_t)blockIdx.y * QT;
  for (int e = tid; e < QT * NWB_NJ; e += NN_THREADS) {
    int k = e / NWB_NJ, j = e % NWB_NJ;
    double v = (q0 + k < nq) ? q[(q0 + k) * NWB_NJ + j] : 0.0;
    qd[k][j] = v;
    qf[k][j] = (float)v;
    qn2[j][k] = make_float2(-(float)v, -(float)v);
  }
  if (tid < QT) run_min[tid] = 0x7f7fffffu;           // FLT_MAX
  if (tid == 0) {
    overflow = 0;
    for (int sg = 0; sg < STAGES; ++sg) nn_mbar_init(&bars[sg], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t tiles = (N + TILE - 1) / TILE;
  auto issue = [&](int64_t tile, int stage) {          // thread 0: the tile's blocks are consecutive: ONE bulk copy
    const int64_t n0 = tile * TILE;
    const int64_t blocks = min((int64_t)(TILE / SOA_BLK), (N - n0 + SOA_BLK - 1) / SOA_BLK);
    const uint32_t bytes = (uint32_t)(blocks * SOA_BLK_FLOATS * sizeof(float));
    nn_mbar_expect_tx(&bars[stage], bytes);
    nn_bulk_g2s(ring + (size_t)stage * NWB_NJ * TILE, soa + (size_t)(n0 / SOA_BLK) * SOA_BLK_FLOATS, bytes, &bars[stage]);
  };
  if (tid == 0)
    for (int sg = 0; sg < STAGES; ++sg) {
      const int64_t t = (int64_t)blockIdx.x + (int64_t)sg * gridDim.x;
      if (t < tiles) issue(t, sg);
    }

  // Each thread keeps its two best fp32 candidates per query; the exact fp64 evaluation is deferred
  // to the end of the scan so that no tile waits on the long-latency exact path.
  float cs[QT][2];
  int ci[QT][2];
  bool ovf = false;
#pragma unroll
  for (int k = 0; k < QT; ++k) {
    cs[k][0] = cs[k][1] = 3.0e38f;
    ci[k][0] = ci[k][1] = -1;
  }
  uint32_t phase = 0;
  int stage = 0;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    nn_mbar_wait(&bars[stage], phase);
    const float* x = ring + (size_t)stage * NWB_NJ * TILE;
    const int64_t n0 = tile * TILE;
    // A thread owns the node pair (2 tid, 2 tid + 1) of the tile and works on it with packed fp32 pairs
    // (add.f32x2 / fma.f32x2 -> FADD2 / FFMA2: two nodes per dispatch slot): per joint one 8-byte load of
    // the pair, then per query one packed subtract of the broadcast (-q, -q) and one packed multiply-add.
    float2 s2[QT][NPT / 2];
#pragma unroll
    for (int k = 0; k < QT; ++k)
#pragma unroll
      for (int h = 0; h < NPT / 2; ++h) s2[k][h] = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < NWB_NJ; ++j) {
      float2 x2[NPT / 2];
      // shared-memory image of a tile = its blocks one after the other: [block][21][128]
      constexpr int PER_BLK = SOA_BLK / NPT;             // threads per block of 128 nodes
      const float* xb = x + ((tid / PER_BLK) * NWB_NJ + j) * SOA_BLK + (tid % PER_BLK) * NPT;
      if constexpr (NPT == 4) {
        const float4 v = *reinterpret_cast<const float4*>(xb);
        x2[0] = make_float2(v.x, v.y);
        x2[1] = make_float2(v.z, v.w);
      } else {
        x2[0] = *reinterpret_cast<const float2*>(xb);
      }
#pragma unroll
      for (int k = 0; k < QT; ++k) {
        const float2 c = qn2[j][k];
#pragma unroll
        for (int h = 0; h < NPT / 2; ++h) {
          const float2 d = __fadd2_rn(x2[h], c);
          s2[k][h] = __ffma2_rn(d, d, s2[k][h]);
        }
      }
    }
    float s[QT][NPT];
#pragma unroll
    for (int k = 0; k < QT; ++k)
#pragma unroll
      for (int h = 0; h < NPT / 2; ++h) {
        s[k][2 * h] = s2[k][h].x;
        s[k][2 * h + 1] = s2[k][h].y;
      }
    float lo[QT];                                       // the smallest of the thread's distances per query
#pragma unroll
    for (int k = 0; k < QT; ++k) {
      float m = 3.0e38f;
#pragma unroll
      for (int u = 0; u < NPT; ++u) {
        if (!(n0 + NPT * tid + u < N)) s[k][u] = 3.0e38f;
        m = fminf(m, s[k][u]);
      }
      lo[k] = m;
      // non-negative floats order like their bit patterns: one REDUX.MIN instead of a shuffle ladder
      const unsigned wmin = __reduce_min_sync(0xffffffffu, __float_as_uint(m));
      if ((tid & 31) == 0) atomicMin(&run_min[k], wmin);
    }
    __syncthreads();                                    // the stage is consumed and run_min is current
    if (tid == 0) {                                     // refill this stage with the tile STAGES rounds ahead
      const int64_t nt = tile + (int64_t)STAGES * gridDim.x;
      if (nt < tiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(nt, stage);
      }
    }
#pragma unroll
    for (int k = 0; k < QT; ++k) {
      const float bound = nn_candidate_bound(__uint_as_float(run_min[k]));
      if (lo[k] <= bound) {                             // rare: at least one of the two nodes is a candidate
#pragma unroll
        for (int u = 0; u < NPT; ++u) {
          const float v = s[k][u];
          if (v <= bound) {                             // candidate (nodes arrive in ascending index per thread)
            const int node = (int)(n0 + NPT * tid + u);
            // a slot that is pushed out may still hold a node under the bound (the bound only shrinks, so one
            // that is above it now stays out for good): that is a third live candidate -> exact rescan below
            if (ci[k][1] >= 0 && cs[k][1] <= bound) ovf = true;
            if (v < cs[k][0]) { cs[k][1] = cs[k][0]; ci[k][1] = ci[k][0]; cs[k][0] = v; ci[k][0] = node; }
            else if (v < cs[k][1]) { cs[k][1] = v; ci[k][1] = node; }
            else ovf = true;                            // neither slot gives way: also a third candidate
          }
        }
      }
    }
    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
  }
  if (ovf) overflow = 1;
  __syncthreads();

  // ---- exact pass (the reference's arithmetic, RP:361-377) ------------------------------------------------
  // Candidates under the final bound go to a shared list, then EVERY candidate gets its own thread: one round
  // of independent global loads for the whole block instead of a chain of dependent round trips per thread.
  __shared__ int cand_n;
  __shared__ int cand_k[NN_CAND_MAX], cand_node[NN_CAND_MAX];
  __shared__ double cand_d[NN_CAND_MAX];
  if (tid == 0) cand_n = 0;
  __syncthreads();
  auto exact_dist = [&](int k, int64_t node) {
    const double* row = aos + node * NWB_NJ;
    double r[NWB_NJ];
#pragma unroll
    for (int j = 0; j < NWB_NJ; ++j) r[j] = row[j];     // all 21 loads in flight
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < NWB_NJ; ++j) {
      double d = __dsub_rn(qd[k][j], r[j]);
      sum = __dadd_rn(sum, __dmul_rn(d, d));            // separately rounded multiply and add
    }
    return sqrt(sum);
  };
  auto push = [&](int k, int node) {
    const int slot = atomicAdd(&cand_n, 1);
    if (slot < NN_CAND_MAX) { cand_k[slot] = k; cand_node[slot] = node; }
  };
  if (!overflow) {
#pragma unroll
    for (int k = 0; k < QT; ++k) {
      const float bound = nn_candidate_bound(__uint_as_float(run_min[k]));
#pragma unroll
      for (int u = 0; u < 2; ++u)
        if (ci[k][u] >= 0 && cs[k][u] <= bound) push(k, ci[k][u]);
    }
  }
  __syncthreads();
  const bool list_ok = !overflow && cand_n <= NN_CAND_MAX;
  if (list_ok)
    for (int c = tid; c < cand_n; c += NN_THREADS) cand_d[c] = exact_dist(cand_k[c], cand_node[c]);
  __syncthreads();
  if (list_ok) {
    if (tid < QT && q0 + tid < nq) {                    // first strict minimum in index order, 10000.0 / -1 if none (RP:347)
      double d = 10000.0;
      int i = -1;
      for (int c = 0; c < cand_n; ++c)
        if (cand_k[c] == tid && lex_less(cand_d[c], cand_node[c], d, i) && cand_d[c] < 10000.0) { d = cand_d[c]; i = cand_node[c]; }
      part_d[(q0 + tid) * gridDim.x + blockIdx.x] = d;
      part_i[(q0 + tid) * gridDim.x + blockIdx.x] = i;
    }
  } else {
    // pathological input (three or more near-ties in one thread, or a flood of candidates): rescan this block's
    // tiles from global memory and take the exact path for every node under the final bound
    double best[QT];
    int besti[QT];
#pragma unroll
    for (int k = 0; k < QT; ++k) {
      best[k] = 10000.0;
      besti[k] = -1;
    }
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x)
      for (int u = 0; u < NPT; ++u) {
        const int64_t node = tile * TILE + NPT * tid + u;
        if (node >= N) continue;
        for (int k = 0; k < QT; ++k) {
          float sv = 0.f;
          for (int j = 0; j < NWB_NJ; ++j) {
            float d = qf[k][j] - soa[soa_index(node, j)];
            sv = fmaf(d, d, sv);
          }
          if (sv <= nn_candidate_bound(__uint_as_float(run_min[k]))) {
            const double dn = exact_dist(k, node);
            if (dn < best[k] || (dn == best[k] && (unsigned)node < (unsigned)besti[k])) { best[k] = dn; besti[k] = (int)node; }
          }
        }
      }
    const int lane = tid & 31, warp = tid >> 5;
    for (int k = 0; k < QT; ++k) {
      double d = best[k];
      int i = besti[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        double d2 = __shfl_xor_sync(0xffffffffu, d, o);
        int i2 = __shfl_xor_sync(0xffffffffu, i, o);
        if (lex_less(d2, i2, d, i)) { d = d2; i = i2; }
      }
      if (lane == 0) { red_d[k][warp] = d; red_i[k][warp] = i; }
    }
    __syncthreads();
    if (tid < QT) {
      double d = red_d[tid][0];
      int i = red_i[tid][0];
      for (int w = 1; w < NN_THREADS / 32; ++w)
        if (lex_less(red_d[tid][w], red_i[tid][w], d, i)) { d = red_d[tid][w]; i = red_i[tid][w]; }
      if (q0 + tid < nq) {
        part_d[(q0 + tid) * gridDim.x + blockIdx.x] = d;
        part_i[(q0 + tid) * gridDim.x + blockIdx.x] = i;
      }
    }
  }
  // grid reduction without a second launch: the last block to finish this query group folds the partials
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (tid == 0) last = atomicAdd(&tickets[blockIdx.y], 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int k = 0; k < QT; ++k) {
    if (q0 + k >= nq) break;
    double d = 10000.0;
    int i = -1;
    bool first = true;
    for (int b = tid; b < (int)gridDim.x; b += NN_THREADS) {
      const double db = part_d[(q0 + k) * gridDim.x + b];
      const int ib = part_i[(q0 + k) * gridDim.x + b];
      if (first || lex_less(db, ib, d, i)) { d = db; i = ib; first = false; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double d2 = __shfl_xor_sync(0xffffffffu, d, o);
      int i2 = __shfl_xor_sync(0xffffffffu, i, o);
      if (lex_less(d2, i2, d, i)) { d = d2; i = i2; }
    }
    if ((tid & 31) == 0) { red_d[0][tid >> 5] = d; red_i[0][tid >> 5] = i; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < NN_THREADS / 32; ++w)
        if (lex_less(red_d[0][w], red_i[0][w], d, i)) { d = red_d[0][w]; i = red_i[0][w]; }
      idx_out[q0 + k] = i;
      if (dist_out) dist_out[q0 + k] = d;
    }
    __syncthreads();
  }
  if (tid == 0) tickets[blockIdx.y] = 0u;              // ready for the next call
}


// ---- single query, small trees: exact fp64 straight from the master copy -------------------------------------------
// Below ~2 10^5 nodes the scan is latency-, not bandwidth-bound, so the filter + exact second trip of the fp32 scan
// costs more than it saves: one block per 256 nodes pulls its rows (43 KB, contiguous in the AoS master copy) into
// shared memory with ONE bulk copy and every thread evaluates its node exactly as RP:361-377 does.  No candidates, no
// second dependent memory trip; grid reduction by a last-block-done ticket.
constexpr int NN1_THREADS = 256;
constexpr int64_t kNnSmallMax = 65536;     // nodes up to which the exact small-tree kernel serves a single query

__device__ __forceinline__ void nn_lex_warp_min(double& d, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double d2 = __shfl_xor_sync(0xffffffffu, d, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, i, o);
    if (lex_less(d2, i2, d, i)) { d = d2; i = i2; }
  }
}

__global__ void __launch_bounds__(NN1_THREADS)
nn_small_kernel(const double* __restrict__ aos, int64_t N, const double* __restrict__ q, double* __restrict__ part_d,
                int32_t* __restrict__ part_i, unsigned* __restrict__ ticket, int32_t* __restrict__ idx_out,
                double* __restrict__ dist_out) {
  __shared__ __align__(128) double rows[NN1_THREADS * NWB_NJ];
  __shared__ double qd[NWB_NJ];
  __shared__ uint64_t bar;
  __shared__ double red_d[NN1_THREADS / 32];
  __shared__ int red_i[NN1_THREADS / 32];
  __shared__ bool last;
  const int tid = threadIdx.x;
  const int64_t n0 = (int64_t)blockIdx.x * NN1_THREADS;
  const int cnt = (int)min((int64_t)NN1_THREADS, N - n0);
  if (tid == 0) {
    nn_mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bytes = (uint32_t)(cnt * NWB_NJ * sizeof(double));      // 168 B per node: a multiple of 8, ...
    const uint32_t b16 = bytes & ~15u;                                      // ... the bulk copy moves whole 16-byte units
    if (b16) {
      nn_mbar_expect_tx(&bar, b16);
      nn_bulk_g2s(rows, aos + n0 * NWB_NJ, b16, &bar);
    } else {
      nn_mbar_expect_tx(&bar, 0);
    }
    if (bytes != b16) rows[b16 / 8] = aos[n0 * NWB_NJ + b16 / 8];           // odd node count: the last 8 bytes by hand
  }
  if (tid < NWB_NJ) qd[tid] = q[tid];
  __syncthreads();
  nn_mbar_wait(&bar, 0);
  double d = 10000.0;                                  // RP:347: best_norm starts at 10000.0, strict <
  int i = -1;
  if (tid < cnt) {
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < NWB_NJ; ++j) {
      const double e = __dsub_rn(qd[j], rows[tid * NWB_NJ + j]);
      sum = __dadd_rn(sum, __dmul_rn(e, e));           // separately rounded multiply and add
    }
    const double dn = sqrt(sum);
    if (dn < 10000.0) { d = dn; i = (int)(n0 + tid); }
  }
  nn_lex_warp_min(d, i);
  if ((tid & 31) == 0) { red_d[tid >> 5] = d; red_i[tid >> 5] = i; }
  __syncthreads();
  if (tid < 32) {
    d = tid < NN1_THREADS / 32 ? red_d[tid] : 10000.0;
    i = tid < NN1_THREADS / 32 ? red_i[tid] : -1;
    nn_lex_warp_min(d, i);
    if (tid == 0) {
      if (gridDim.x == 1) {
        idx_out[0] = i;
        if (dist_out) dist_out[0] = d;
        last = false;
      } else {
        part_d[blockIdx.x] = d;
        part_i[blockIdx.x] = i;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
      }
    }
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  d = 10000.0;
  i = -1;
  for (int b = tid; b < (int)gridDim.x; b += NN1_THREADS) {
    const double db = part_d[b];
    const int ib = part_i[b];
    if (ib >= 0 && lex_less(db, ib, d, i)) { d = db; i = ib; }
  }
  nn_lex_warp_min(d, i);
  if ((tid & 31) == 0) { red_d[tid >> 5] = d; red_i[tid >> 5] = i; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < NN1_THREADS / 32; ++w)
      if (red_i[w] >= 0 && lex_less(red_d[w], red_i[w], d, i)) { d = red_d[w]; i = red_i[w]; }
    idx_out[0] = i;
    if (dist_out) dist_out[0] = d;
    *ticket = 0u;
  }
}

// ---- single query, large trees: fp32 SoA stream + exact re-evaluation of the listed candidates ----------------------
// One persistent CTA per SM streams its CONTIGUOUS node range (balanced to 4 nodes) through a 4-stage ring.  The 21 row
// copies of a tile are issued by 21 lanes of warp 0 side by side (one thread issuing them one after the other was on the
// critical path: ~0.5 us per tile).  A node whose fp32 squared distance lies within the rigorous error bound of the
// block's running minimum is listed in shared memory the moment it is seen (and its master-copy row is prefetched into
// L2); there are no per-thread candidate slots that could silently overflow.  After the stream only the listed nodes
// still under the final bound are evaluated in fp64.  A list overflow (a flood of near ties) falls back to an exact
// rescan of the block's range.
constexpr int NN2_THREADS = 256;
constexpr int NN2_TILE = 512;
constexpr int NN2_STAGES = 4;
constexpr int NN2_STAGES_H = 8;      // the fp16 ring holds twice the tiles in the same 172 KB
constexpr int NN2_CAND = 1024;

// HALF: the stream reads the fp16 filter copy (42 B/node instead of 84).  A node stored in fp16 is off by at most half
// an fp16 ulp per joint, so its distance to the query moves by at most eps = sqrt(21) * (ulp16(max |node|) / 2)
// (triangle inequality); the tree's max |node|^2 is tracked on the device.  With m the running minimum of the APPROXIMATE
// distances, the exact nearest node w has d(w) <= m + eps, so every node that can be nearest or tie has an approximate
// distance <= m + 2 eps: those are listed and re-evaluated in fp64 from the master copy, exactly as in the fp32 stream.
// For joint angles (|q| < 4) eps = 4.5e-3; around the nearest of 10^6 nodes in 21 dimensions that shell holds a
// handful of nodes.
__device__ __forceinline__ float nn_candidate_bound16(float m2, float eps) {
  const float m = sqrtf(m2) * (1.0f + 1e-5f) + 2.0f * eps;
  return m * m * (1.0f + 1e-5f) + 1e-12f;
}

template <bool HALF>
__global__ void __launch_bounds__(NN2_THREADS, 1)
nn_stream_kernel(const float* __restrict__ soa, const uint16_t* __restrict__ h16, const unsigned* __restrict__ max_nn,
                 const double* __restrict__ aos, int64_t cap, int64_t N,
                 const double* __restrict__ q, double* __restrict__ part_d, int32_t* __restrict__ part_i,
                 unsigned* __restrict__ ticket, int32_t* __restrict__ idx_out, double* __restrict__ dist_out) {
  constexpr int STAGES = HALF ? NN2_STAGES_H : NN2_STAGES;
  constexpr size_t ELEM = HALF ? sizeof(uint16_t) : sizeof(float);
  extern __shared__ __align__(128) unsigned char nn_smem[];
  unsigned char* ring = nn_smem;                                             // [STAGES][21][TILE] floats or halves
  uint64_t* bars = reinterpret_cast<uint64_t*>(nn_smem + ELEM * STAGES * NWB_NJ * NN2_TILE);
  __shared__ float eps16;
  __shared__ double qd[NWB_NJ];
  __shared__ __align__(8) float2 qn2[NWB_NJ];
  __shared__ unsigned run_min;
  __shared__ int cand_n;
  __shared__ int cand_node[NN2_CAND];
  __shared__ float cand_s[NN2_CAND];
  __shared__ double red_d[NN2_THREADS / 32];
  __shared__ int red_i[NN2_THREADS / 32];
  __shared__ bool last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // contiguous node range of this CTA, boundaries on whole blocks of 128 nodes (the unit of the blocked copy)
  const int64_t nblk = (N + SOA_BLK - 1) / SOA_BLK;
  const int64_t lo = (nblk * blockIdx.x / gridDim.x) * SOA_BLK, hi = min(N, (nblk * (blockIdx.x + 1) / gridDim.x) * SOA_BLK);
  const int tiles = (int)((hi - lo + NN2_TILE - 1) / NN2_TILE);
  if (tid == 0) {
    for (int sg = 0; sg < STAGES; ++sg) nn_mbar_init(&bars[sg], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    run_min = 0x7f7fffffu;
    cand_n = 0;
    if constexpr (HALF) {
      const float max_abs = sqrtf(__uint_as_float(*max_nn)) * (1.0f + 1e-6f);   // >= every |coordinate| of the tree
      const int ex = max_abs > 6.1035156e-5f ? ilogbf(max_abs) : -14;           // below 2^-14 fp16 is subnormal: ulp 2^-24
      const float half_ulp = max_abs < 6.0e4f ? ldexpf(1.0f, ex - 11) : 3.0e38f; // beyond the fp16 range: everything is a candidate
      float qmax = 0.f;                                                          // the query is used in fp32: its own rounding
      for (int j = 0; j < NWB_NJ; ++j) qmax = fmaxf(qmax, fabsf((float)q[j]));
      eps16 = 4.5825757f * (half_ulp * 1.001f + qmax * 6.0e-8f) + 1e-12f;        // sqrt(21) x the per-joint errors
    }
  }
  __syncthreads();
  auto issue = [&](int tile, int stage) {              // one bulk copy per tile: its blocks of 128 nodes are consecutive
    const int64_t t0 = lo + (int64_t)tile * NN2_TILE;
    const int64_t blocks = min((int64_t)(NN2_TILE / SOA_BLK), (hi - t0 + SOA_BLK - 1) / SOA_BLK);
    const uint32_t bytes = (uint32_t)(blocks * SOA_BLK_FLOATS * ELEM);
    if (lane == 0) {
      nn_mbar_expect_tx(&bars[stage], bytes);
      const unsigned char* src = HALF ? reinterpret_cast<const unsigned char*>(h16) : reinterpret_cast<const unsigned char*>(soa);
      nn_bulk_g2s(ring + (size_t)stage * NWB_NJ * NN2_TILE * ELEM, src + (size_t)(t0 / SOA_BLK) * SOA_BLK_FLOATS * ELEM, bytes, &bars[stage]);
    }
  };
  if (warp == 0)
    for (int sg = 0; sg < STAGES && sg < tiles; ++sg) issue(sg, sg);
  if (tid >= 32 && tid < 32 + NWB_NJ) {                // the query arrives while the first tiles are in flight
    const double v = q[tid - 32];
    qd[tid - 32] = v;
    qn2[tid - 32] = make_float2(-(float)v, -(float)v);
  }
  __syncthreads();

  uint32_t phase = 0;
  int stage = 0;
  bool flooded = false;
  for (int tile = 0; tile < tiles; ++tile) {
    nn_mbar_wait(&bars[stage], phase);
    const int64_t t0 = lo + (int64_t)tile * NN2_TILE;
    float2 s2 = make_float2(0.f, 0.f);
    if constexpr (HALF) {
      const __half* x = reinterpret_cast<const __half*>(ring + (size_t)stage * NWB_NJ * NN2_TILE * ELEM);
#pragma unroll
      for (int j = 0; j < NWB_NJ; ++j) {
        const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&x[((tid >> 6) * NWB_NJ + j) * SOA_BLK + 2 * (tid & 63)]));
        const float2 d = __fadd2_rn(v, qn2[j]);
        s2 = __ffma2_rn(d, d, s2);
      }
    } else {
      const float* x = reinterpret_cast<const float*>(ring + (size_t)stage * NWB_NJ * NN2_TILE * ELEM);
#pragma unroll
      for (int j = 0; j < NWB_NJ; ++j) {
        const float2 d = __fadd2_rn(*reinterpret_cast<const float2*>(&x[((tid >> 6) * NWB_NJ + j) * SOA_BLK + 2 * (tid & 63)]), qn2[j]);
        s2 = __ffma2_rn(d, d, s2);
      }
    }
    const int64_t node0 = t0 + 2 * tid;
    const float sa = node0 < hi ? s2.x : 3.0e38f, sb = node0 + 1 < hi ? s2.y : 3.0e38f;
    const unsigned wmin = __reduce_min_sync(0xffffffffu, __float_as_uint(fminf(sa, sb)));
    if (lane == 0) atomicMin(&run_min, wmin);
    __syncthreads();                                    // the stage is consumed and run_min is current
    if (warp == 0 && tile + STAGES < tiles) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(tile + STAGES, stage);
    }
    const float bound = HALF ? nn_candidate_bound16(__uint_as_float(run_min), eps16) : nn_candidate_bound(__uint_as_float(run_min));
    if (fminf(sa, sb) <= bound) {                       // rare
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float v = u ? sb : sa;
        if (v <= bound) {
          const int slot = atomicAdd(&cand_n, 1);
          if (slot < NN2_CAND) {
            cand_node[slot] = (int)(node0 + u);
            cand_s[slot] = v;
            const double* row = aos + (node0 + u) * NWB_NJ;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 16));
          } else {
            flooded = true;
          }
        }
      }
    }
    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
  }
  const int any_flood = __syncthreads_or(flooded ? 1 : 0);

  auto exact_dist = [&](int64_t node) {
    const double* row = aos + node * NWB_NJ;
    double r[NWB_NJ];
#pragma unroll
    for (int j = 0; j < NWB_NJ; ++j) r[j] = row[j];
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < NWB_NJ; ++j) {
      const double d = __dsub_rn(qd[j], r[j]);
      sum = __dadd_rn(sum, __dmul_rn(d, d));
    }
    return sqrt(sum);
  };
  double d = 10000.0;
  int i = -1;
  const float fbound = HALF ? nn_candidate_bound16(__uint_as_float(run_min), eps16) : nn_candidate_bound(__uint_as_float(run_min));
  if (!any_flood) {
    const int n_c = cand_n;
    for (int c = tid; c < n_c; c += NN2_THREADS)
      if (cand_s[c] <= fbound) {
        const double dn = exact_dist(cand_node[c]);
        if (dn < 10000.0 && lex_less(dn, cand_node[c], d, i)) { d = dn; i = cand_node[c]; }
      }
  } else {
    // a flood of near ties: exact rescan of this CTA's range from global memory (fp32 copy).  HALF: the exact nearest
    // node is within sqrt(run_min) + eps of the query, which bounds the fp32 distances worth an exact look.
    float fb32 = fbound;
    if constexpr (HALF) {
      const float m = sqrtf(__uint_as_float(run_min)) * (1.0f + 1e-5f) + eps16;
      fb32 = nn_candidate_bound(m * m * (1.0f + 1e-5f));
    }
    for (int64_t node = lo + tid; node < hi; node += NN2_THREADS) {
      float sv = 0.f;
      for (int j = 0; j < NWB_NJ; ++j) {
        const float e = (float)qd[j] - soa[soa_index(node, j)];
        sv = fmaf(e, e, sv);
      }
      if (sv <= fb32) {
        const double dn = exact_dist(node);
        if (dn < 10000.0 && lex_less(dn, (int)node, d, i)) { d = dn; i = (int)node; }
      }
    }
  }
  nn_lex_warp_min(d, i);
  if (lane == 0) { red_d[warp] = d; red_i[warp] = i; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < NN2_THREADS / 32; ++w)
      if (red_i[w] >= 0 && lex_less(red_d[w], red_i[w], d, i)) { d = red_d[w]; i = red_i[w]; }
    part_d[blockIdx.x] = d;
    part_i[blockIdx.x] = i;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  d = 10000.0;
  i = -1;
  for (int b = tid; b < (int)gridDim.x; b += NN2_THREADS) {
    const double db = part_d[b];
    const int ib = part_i[b];
    if (ib >= 0 && lex_less(db, ib, d, i)) { d = db; i = ib; }
  }
  nn_lex_warp_min(d, i);
  if (lane == 0) { red_d[warp] = d; red_i[warp] = i; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < NN2_THREADS / 32; ++w)
      if (red_i[w] >= 0 && lex_less(red_d[w], red_i[w], d, i)) { d = red_d[w]; i = red_i[w]; }
    idx_out[0] = i;
    if (dist_out) dist_out[0] = d;
    *ticket = 0u;
  }
}

}  // namespace nwb

using namespace nwb;

namespace {
#define TREE_CUDA(t, call)                                                           \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess) {                                                         \
      (t)->ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);            \
      return NWB_E_CUDA;                                                             \
    }                                                                                \
  } while (0)

int tree_reserve(nwb_tree* t, int64_t need) {
  if (need <= t->cap) return NWB_OK;
  int64_t cap = std::max<int64_t>(1024, t->cap);
  while (cap < need) cap *= 2;
  cap = (cap + 63) & ~(int64_t)63;
  cap = (cap + TC_N - 1) / TC_N * TC_N;                 // whole tiles of the tensor-core scan copy
  float* soa = nullptr;
  uint16_t* h16 = nullptr;
  double* aos = nullptr;
  int32_t* pred = nullptr;
  unsigned char* tc = nullptr;
  const size_t tc_bytes = (size_t)(cap / TC_N + 1) * TC_TILE_BYTES;   // + one spare tile: the scan reads tiles in pairs
  TREE_CUDA(t, cudaMalloc(&soa, sizeof(float) * NWB_NJ * cap));
  TREE_CUDA(t, cudaMalloc(&h16, sizeof(uint16_t) * NWB_NJ * cap));
  TREE_CUDA(t, cudaMalloc(&aos, sizeof(double) * NWB_NJ * cap));
  TREE_CUDA(t, cudaMalloc(&pred, sizeof(int32_t) * cap));
  TREE_CUDA(t, cudaMalloc(&tc, tc_bytes));
  cudaStream_t s = t->ctx->stream;
  TREE_CUDA(t, cudaMemsetAsync(tc, 0, tc_bytes, s));    // rows past the end of the tree hold finite values
  TREE_CUDA(t, cudaMemsetAsync(soa, 0, sizeof(float) * NWB_NJ * cap, s));
  TREE_CUDA(t, cudaMemsetAsync(h16, 0, sizeof(uint16_t) * NWB_NJ * cap, s));
  if (!t->max_nn) {
    TREE_CUDA(t, cudaMalloc(&t->max_nn, sizeof(unsigned)));
    TREE_CUDA(t, cudaMemsetAsync(t->max_nn, 0, sizeof(unsigned), s));
  }
  if (t->size > 0) {
    TREE_CUDA(t, cudaMemcpyAsync(tc, t->tc, (size_t)((t->size + TC_N - 1) / TC_N) * TC_TILE_BYTES, cudaMemcpyDeviceToDevice, s));
    TREE_CUDA(t, cudaMemcpyAsync(soa, t->soa, sizeof(float) * SOA_BLK_FLOATS * (size_t)((t->size + SOA_BLK - 1) / SOA_BLK),
                                 cudaMemcpyDeviceToDevice, s));
    TREE_CUDA(t, cudaMemcpyAsync(h16, t->h16, sizeof(uint16_t) * SOA_BLK_FLOATS * (size_t)((t->size + SOA_BLK - 1) / SOA_BLK),
                                 cudaMemcpyDeviceToDevice, s));
    TREE_CUDA(t, cudaMemcpyAsync(aos, t->aos, sizeof(double) * NWB_NJ * t->size, cudaMemcpyDeviceToDevice, s));
    TREE_CUDA(t, cudaMemcpyAsync(pred, t->pred, sizeof(int32_t) * t->size, cudaMemcpyDeviceToDevice, s));
  }
  TREE_CUDA(t, cudaStreamSynchronize(s));
  cudaFree(t->soa);
  cudaFree(t->h16);
  cudaFree(t->aos);
  cudaFree(t->pred);
  cudaFree(t->tc);
  t->soa = soa;
  t->h16 = h16;
  t->aos = aos;
  t->pred = pred;
  t->tc = tc;
  t->cap = cap;
  return NWB_OK;
}

int stage_reserve(nwb_tree* t, size_t bytes) {
  if (bytes <= t->stage_bytes) return NWB_OK;
  if (t->stage) cudaFree(t->stage);
  t->stage = nullptr;
  t->stage_bytes = 0;
  size_t cap = std::max(bytes, (size_t)1 << 16);
  TREE_CUDA(t, cudaMalloc(&t->stage, cap));
  t->stage_bytes = cap;
  return NWB_OK;
}

int nn_grid_blocks(const nwb_tree* t) {
  int64_t want = (t->size + NN_TILE - 1) / NN_TILE;
  int64_t cap = (int64_t)t->ctx->sm_count;                             // persistent: one CTA per SM
  return (int)std::max<int64_t>(1, std::min(want, cap));
}

int nn_launch(nwb_tree* t, const double* q_dev, int64_t nq, int32_t* idx_dev, double* dist_dev) {
  const int nb = nn_grid_blocks(t);
  size_t need = std::max((size_t)nq * nb, (size_t)(kNnSmallMax / NN1_THREADS + 1));   // the small-tree kernel: one partial per block
  if (need > t->part_cap) {
    if (t->part_d) cudaFree(t->part_d);
    if (t->part_i) cudaFree(t->part_i);
    t->part_d = nullptr;
    t->part_i = nullptr;
    t->part_cap = 0;
    TREE_CUDA(t, cudaMalloc(&t->part_d, sizeof(double) * need));
    TREE_CUDA(t, cudaMalloc(&t->part_i, sizeof(int32_t) * need));
    t->part_cap = need;
  }
  {
    const size_t groups = (size_t)nq;                 // >= number of query groups for either QT
    if (groups > t->ticket_cap) {
      if (t->tickets) cudaFree(t->tickets);
      t->tickets = nullptr;
      t->ticket_cap = 0;
      TREE_CUDA(t, cudaMalloc(&t->tickets, sizeof(unsigned) * groups));
      TREE_CUDA(t, cudaMemsetAsync(t->tickets, 0, sizeof(unsigned) * groups, t->ctx->stream));
      t->ticket_cap = groups;
    }
  }
  cudaStream_t s = t->ctx->stream;
  constexpr size_t kNnSmem = sizeof(float) * NN_STAGES * NWB_NJ * NN_TILE + 8 * NN_STAGES;
  static std::atomic<bool> configured[64];             // the attribute is per device (contexts may live on several GPUs)
  const int dev = t->ctx->device >= 0 && t->ctx->device < 64 ? t->ctx->device : 0;
  if (!configured[dev].load(std::memory_order_relaxed)) {
    TREE_CUDA(t, cudaFuncSetAttribute(nn_scan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNnSmem));
    TREE_CUDA(t, cudaFuncSetAttribute(nn_scan_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNnSmem));
    TREE_CUDA(t, cudaFuncSetAttribute(nn_scan_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNnSmem));
    configured[dev].store(true, std::memory_order_relaxed);
  }
  // batches of 32 queries and more: the distance matrix is a contraction and runs on the tensor cores (nn_tensor.cu);
  // NWB_NN_NO_TC=1 keeps the CUDA-core scan for A/B measurements
  static const bool no_tc = getenv("NWB_NN_NO_TC") != nullptr;
  if (nq >= 32 && !no_tc) {
    TREE_CUDA(t, ev_record(t->ctx, t->ctx->ev0, s));
    int launches = 0;
    TREE_CUDA(t, nn_tc_launch(t, q_dev, nq, idx_dev, dist_dev, nullptr, &launches));
    TREE_CUDA(t, ev_record(t->ctx, t->ctx->ev1, s));
    t->ctx->last_ms = -1;
    t->ctx->last_launches = launches;
    t->ctx->last_nn_passes = (nq + TC_M - 1) / TC_M;
    return NWB_OK;
  }
  // one query (the RRT loop's own call pattern): exact fp64 blocks for small trees, the fp32 stream for large ones;
  // NWB_NN_OLD=1 keeps the round-1 tile-interleaved kernel for A/B measurements
  static const bool old_single = getenv("NWB_NN_OLD") != nullptr;
  if (nq == 1 && !old_single) {
    TREE_CUDA(t, ev_record(t->ctx, t->ctx->ev0, s));
    if (t->size <= kNnSmallMax) {
      const unsigned blocks = (unsigned)std::max<int64_t>(1, (t->size + NN1_THREADS - 1) / NN1_THREADS);
      if (t->size == 0) {
        nn_small_kernel<<<1, NN1_THREADS, 0, s>>>(t->aos, 0, q_dev, t->part_d, t->part_i, t->tickets, idx_dev, dist_dev);
      } else {
        nn_small_kernel<<<blocks, NN1_THREADS, 0, s>>>(t->aos, t->size, q_dev, t->part_d, t->part_i, t->tickets, idx_dev, dist_dev);
      }
    } else {
      constexpr size_t kSmem2 = sizeof(float) * NN2_STAGES * NWB_NJ * NN2_TILE + 8 * NN2_STAGES_H;
      static std::atomic<bool> cfg2[64];
      if (!cfg2[dev].load(std::memory_order_relaxed)) {
        TREE_CUDA(t, cudaFuncSetAttribute(nn_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem2));
        TREE_CUDA(t, cudaFuncSetAttribute(nn_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem2));
        cfg2[dev].store(true, std::memory_order_relaxed);
      }
      const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(t->ctx->sm_count, (t->size + NN2_TILE - 1) / NN2_TILE));
      // the fp16 filter copy halves the stream (42 B/node); NWB_NN_NO_F16=1 keeps the fp32 stream for A/B measurements
      static const bool no_f16 = getenv("NWB_NN_NO_F16") != nullptr;
      if (no_f16)
        nn_stream_kernel<false><<<blocks, NN2_THREADS, kSmem2, s>>>(t->soa, t->h16, t->max_nn, t->aos, t->cap, t->size, q_dev, t->part_d,
                                                                    t->part_i, t->tickets, idx_dev, dist_dev);
      else
        nn_stream_kernel<true><<<blocks, NN2_THREADS, kSmem2, s>>>(t->soa, t->h16, t->max_nn, t->aos, t->cap, t->size, q_dev, t->part_d,
                                                                   t->part_i, t->tickets, idx_dev, dist_dev);
    }
    TREE_CUDA(t, cudaGetLastError());
    TREE_CUDA(t, ev_record(t->ctx, t->ctx->ev1, s));
    t->ctx->last_ms = -1;
    t->ctx->last_launches = 1;
    t->ctx->last_nn_passes = 1;
    return NWB_OK;
  }
  TREE_CUDA(t, ev_record(t->ctx, t->ctx->ev0, s));
  // batches below 32 queries: 8 or 16 queries share each pass over the nodes on the FP32 pipe
  static const int qt_batch = getenv("NWB_NN_QT") ? atoi(getenv("NWB_NN_QT")) : 16;
  const int QT = nq == 1 ? 1 : (nq <= 8 || qt_batch == 8 ? 8 : 16);
  const int64_t qtiles = (nq + QT - 1) / QT;
  for (int64_t y0 = 0; y0 < qtiles; y0 += 65535) {       // gridDim.y limit
    const int64_t ny = std::min<int64_t>(65535, qtiles - y0);
    const int64_t qoff = y0 * QT;
    if (QT == 1)
      nn_scan_kernel<1><<<dim3(nb, (unsigned)ny), NN_THREADS, kNnSmem, s>>>(t->soa, t->aos, t->cap, t->size, q_dev + qoff * NWB_NJ,
                                                                            nq - qoff, t->part_d + qoff * nb, t->part_i + qoff * nb,
                                                                            t->tickets + y0, idx_dev + qoff, dist_dev ? dist_dev + qoff : nullptr);
    else if (QT == 8)
      nn_scan_kernel<8><<<dim3(nb, (unsigned)ny), NN_THREADS, kNnSmem, s>>>(t->soa, t->aos, t->cap, t->size, q_dev + qoff * NWB_NJ,
                                                                            nq - qoff, t->part_d + qoff * nb, t->part_i + qoff * nb,
                                                                            t->tickets + y0, idx_dev + qoff, dist_dev ? dist_dev + qoff : nullptr);
    else
      nn_scan_kernel<16><<<dim3(nb, (unsigned)ny), NN_THREADS, kNnSmem, s>>>(t->soa, t->aos, t->cap, t->size, q_dev + qoff * NWB_NJ,
                                                                             nq - qoff, t->part_d + qoff * nb, t->part_i + qoff * nb,
                                                                             t->tickets + y0, idx_dev + qoff, dist_dev ? dist_dev + qoff : nullptr);
  }
  TREE_CUDA(t, cudaGetLastError());
  TREE_CUDA(t, ev_record(t->ctx, t->ctx->ev1, s));
  t->ctx->last_ms = -1;
  t->ctx->last_launches = (qtiles + 65534) / 65535;
  t->ctx->last_nn_passes = qtiles;
  return NWB_OK;
}
}  // namespace

extern "C" {

int nwb_tree_create(nwb_ctx* ctx, int64_t capacity, nwb_tree** out) {
  if (!ctx || !out || capacity < 0) return NWB_E_INVALID;
  *out = nullptr;
  nwb_tree* t = new (std::nothrow) nwb_tree();
  if (!t) return NWB_E_NOMEM;
  t->ctx = ctx;
  cudaSetDevice(ctx->device);
  int rc = tree_reserve(t, std::max<int64_t>(capacity, 1));
  if (rc != NWB_OK) {
    delete t;
    return rc;
  }
  *out = t;
  return NWB_OK;
}

void nwb_tree_destroy(nwb_tree* t) {
  if (!t) return;
  cudaSetDevice(t->ctx->device);
  cudaStreamSynchronize(t->ctx->stream);
  cudaFree(t->soa);
  cudaFree(t->h16);
  cudaFree(t->aos);
  cudaFree(t->pred);
  cudaFree(t->tc);
  cudaFree(t->max_nn);
  cudaFree(t->tc_scratch);
  cudaFree(t->part_d);
  cudaFree(t->part_i);
  cudaFree(t->tickets);
  cudaFree(t->stage);
  delete t;
}

int nwb_tree_clear(nwb_tree* t) {
  if (!t) return NWB_E_INVALID;
  t->size = 0;
  if (t->max_nn) {
    cudaSetDevice(t->ctx->device);
    TREE_CUDA(t, cudaMemsetAsync(t->max_nn, 0, sizeof(unsigned), t->ctx->stream));
  }
  return NWB_OK;
}

int64_t nwb_tree_size(const nwb_tree* t) { return t ? t->size : NWB_E_INVALID; }

int nwb_tree_add_dev(nwb_tree* t, const double* q_dev, const int32_t* pred_dev, int64_t n) {
  if (!t || n < 0 || (n > 0 && !q_dev)) return NWB_E_INVALID;
  if (n == 0) return NWB_OK;
  cudaSetDevice(t->ctx->device);
  int rc = tree_reserve(t, t->size + n);
  if (rc != NWB_OK) return rc;
  const int64_t e = n * NWB_NJ;
  tree_append_kernel<<<(unsigned)((e + 255) / 256), 256, 0, t->ctx->stream>>>(q_dev, pred_dev, n, t->soa, t->h16, t->aos, t->pred,
                                                                               t->cap, t->size);
  tree_append_tc_kernel<<<(unsigned)((n + 127) / 128), 128, 0, t->ctx->stream>>>(q_dev, n, t->tc, t->size, t->max_nn);
  TREE_CUDA(t, cudaGetLastError());
  t->size += n;
  return NWB_OK;
}

int nwb_tree_add(nwb_tree* t, const double* q, const int32_t* predecessor, int64_t n) {
  if (!t || n < 0 || (n > 0 && !q)) return NWB_E_INVALID;
  if (n == 0) return NWB_OK;
  cudaSetDevice(t->ctx->device);
  const size_t qb = sizeof(double) * NWB_NJ * n, pb = sizeof(int32_t) * n;
  int rc = stage_reserve(t, qb + pb);
  if (rc != NWB_OK) return rc;
  cudaStream_t s = t->ctx->stream;
  TREE_CUDA(t, cudaMemcpyAsync(t->stage, q, qb, cudaMemcpyHostToDevice, s));
  int32_t* pd = nullptr;
  if (predecessor) {
    pd = (int32_t*)((char*)t->stage + qb);
    TREE_CUDA(t, cudaMemcpyAsync(pd, predecessor, pb, cudaMemcpyHostToDevice, s));
  }
  rc = nwb_tree_add_dev(t, (const double*)t->stage, pd, n);
  if (rc != NWB_OK) return rc;
  TREE_CUDA(t, cudaStreamSynchronize(s));   // the staging buffer is reused by the next call
  return NWB_OK;
}

int nwb_tree_get(nwb_tree* t, int64_t first, int64_t n, double* q_out, int32_t* pred_out) {
  if (!t || first < 0 || n < 0 || first + n > t->size) return NWB_E_INVALID;
  if (n == 0) return NWB_OK;
  cudaSetDevice(t->ctx->device);
  cudaStream_t s = t->ctx->stream;
  if (q_out) TREE_CUDA(t, cudaMemcpyAsync(q_out, t->aos + first * NWB_NJ, sizeof(double) * NWB_NJ * n, cudaMemcpyDeviceToHost, s));
  if (pred_out) TREE_CUDA(t, cudaMemcpyAsync(pred_out, t->pred + first, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
  TREE_CUDA(t, cudaStreamSynchronize(s));
  return NWB_OK;
}

int64_t nwb_nn_scan_owner(const nwb_tree* t, int64_t nq, int64_t node) {
  if (!t || nq < 1 || node < 0 || node >= t->size) return NWB_E_INVALID;
  if (nq == 1) {
    if (t->size <= kNnSmallMax) return node;                              // exact kernel: one node per thread
    const int64_t blocks = std::max<int64_t>(1, std::min<int64_t>(t->ctx->sm_count, (t->size + NN2_TILE - 1) / NN2_TILE));
    const int64_t nblk = (t->size + SOA_BLK - 1) / SOA_BLK;
    int64_t b = 0;                                                        // contiguous ranges (nn_stream_kernel)
    while (b + 1 < blocks && (nblk * (b + 1) / blocks) * SOA_BLK <= node) ++b;
    const int64_t lo = (nblk * b / blocks) * SOA_BLK;
    return b * NN2_THREADS + ((node - lo) % NN2_TILE) / 2;
  }
  if (nq >= 32) return (node % TC_N) / 32 * TC_M + 0;                     // tensor-core path: a thread per (query, 32-column chunk)
  const int64_t npt = NnCfg<16>::NPT, tile = NN_THREADS * npt;
  const int64_t nb = nn_grid_blocks(t);
  return ((node / tile) % nb) * NN_THREADS + (node % tile) / npt;
}

int nwb_nn_debug_tensor_tile(nwb_tree* t, const double* q, int64_t nq, float* tile_out) {
  if (!t || !q || nq < 1 || nq > TC_M || !tile_out || t->size < 1) return NWB_E_INVALID;
  cudaSetDevice(t->ctx->device);
  const size_t qb = sizeof(double) * NWB_NJ * nq, ob = sizeof(float) * TC_M * TC_N, ib = sizeof(int32_t) * nq;
  int rc = stage_reserve(t, ((qb + 255) & ~(size_t)255) + ob + ib);
  if (rc != NWB_OK) return rc;
  cudaStream_t s = t->ctx->stream;
  double* qd = (double*)t->stage;
  float* od = (float*)((char*)t->stage + ((qb + 255) & ~(size_t)255));
  int32_t* id = (int32_t*)((char*)od + ob);
  TREE_CUDA(t, cudaMemcpyAsync(qd, q, qb, cudaMemcpyHostToDevice, s));
  TREE_CUDA(t, cudaMemsetAsync(od, 0, ob, s));
  TREE_CUDA(t, nn_tc_launch(t, qd, nq, id, nullptr, od, nullptr));
  TREE_CUDA(t, cudaMemcpyAsync(tile_out, od, ob, cudaMemcpyDeviceToHost, s));
  TREE_CUDA(t, cudaStreamSynchronize(s));
  return NWB_OK;
}

int nwb_nearest_neighbour_batch_dev(nwb_tree* t, const double* q_dev, int64_t nq, int32_t* idx_dev, double* dist_dev) {
  if (!t || nq < 0 || (nq > 0 && (!q_dev || !idx_dev))) return NWB_E_INVALID;
  if (nq == 0) return NWB_OK;
  cudaSetDevice(t->ctx->device);
  return nn_launch(t, q_dev, nq, idx_dev, dist_dev);
}

int nwb_nearest_neighbour_batch(nwb_tree* t, const double* q, int64_t nq, int32_t* idx, double* dist) {
  if (!t || nq < 0 || (nq > 0 && (!q || !idx))) return NWB_E_INVALID;
  if (nq == 0) return NWB_OK;
  cudaSetDevice(t->ctx->device);
  const size_t qb = sizeof(double) * NWB_NJ * nq, ib = sizeof(int32_t) * nq, db = sizeof(double) * nq;
  const size_t qb_a = (qb + 255) & ~(size_t)255, db_a = (db + 255) & ~(size_t)255;
  int rc = stage_reserve(t, qb_a + db_a + ib);
  if (rc != NWB_OK) return rc;
  cudaStream_t s = t->ctx->stream;
  double* qd = (double*)t->stage;
  double* dd = (double*)((char*)t->stage + qb_a);
  int32_t* id = (int32_t*)((char*)t->stage + qb_a + db_a);
  TREE_CUDA(t, cudaMemcpyAsync(qd, q, qb, cudaMemcpyHostToDevice, s));
  rc = nn_launch(t, qd, nq, id, dd);
  if (rc != NWB_OK) return rc;
  TREE_CUDA(t, cudaMemcpyAsync(idx, id, ib, cudaMemcpyDeviceToHost, s));
  if (dist) TREE_CUDA(t, cudaMemcpyAsync(dist, dd, db, cudaMemcpyDeviceToHost, s));
  TREE_CUDA(t, cudaStreamSynchronize(s));
  return NWB_OK;
}

}  // extern "C"
