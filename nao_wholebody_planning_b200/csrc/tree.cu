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
// reductions (warp shuffle -> block -> grid) then reproduce "first strict minimum wins".
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <new>

#include "nwb_internal.h"

namespace nwb {

constexpr int NN_THREADS = 256;
constexpr int NN_QT = 4;          // queries per pass over the nodes (registers: 4 x {best, idx, threshold})

__global__ void tree_append_kernel(const double* __restrict__ src, const int32_t* __restrict__ pred_src, int64_t n,
                                   float* __restrict__ soa, double* __restrict__ aos, int32_t* __restrict__ pred,
                                   int64_t cap, int64_t size) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * NWB_NJ) return;
  int64_t i = e / NWB_NJ;
  int j = (int)(e - i * NWB_NJ);
  double v = src[e];
  aos[(size + i) * NWB_NJ + j] = v;
  soa[(int64_t)j * cap + size + i] = (float)v;
  if (j == 0) pred[size + i] = pred_src ? pred_src[i] : -1;
}

__device__ __forceinline__ bool lex_less(double da, int ia, double db, int ib) {
  // smaller distance first; on equal distance the lower index (a found node beats "none" = -1 only
  // through a strictly smaller distance, like the reference's best_norm = 10000.0 start)
  return (da < db) || (da == db && (unsigned)ia < (unsigned)ib);
}

// one pass over the nodes for NN_QT queries; writes one (dist, idx) partial per (query, block)
__global__ void __launch_bounds__(NN_THREADS)
nn_scan_kernel(const float* __restrict__ soa, const double* __restrict__ aos, int64_t cap, int64_t N,
               const double* __restrict__ q, int64_t nq, double* __restrict__ part_d, int32_t* __restrict__ part_i) {
  __shared__ double qd[NN_QT][NWB_NJ];
  __shared__ float qf[NN_QT][NWB_NJ];
  __shared__ double red_d[NN_QT][NN_THREADS / 32];
  __shared__ int red_i[NN_QT][NN_THREADS / 32];
  const int tid = threadIdx.x;
  const int64_t q0 = (int64_t)blockIdx.y * NN_QT;
  for (int e = tid; e < NN_QT * NWB_NJ; e += NN_THREADS) {
    int k = e / NWB_NJ, j = e % NWB_NJ;
    double v = (q0 + k < nq) ? q[(q0 + k) * NWB_NJ + j] : 0.0;
    qd[k][j] = v;
    qf[k][j] = (float)v;
  }
  __syncthreads();

  double best[NN_QT];
  int besti[NN_QT];
  float thr[NN_QT];
#pragma unroll
  for (int k = 0; k < NN_QT; ++k) {
    best[k] = 10000.0;       // RP:347
    besti[k] = -1;
    thr[k] = 3.0e38f;
  }
  for (int64_t node = (int64_t)blockIdx.x * NN_THREADS + tid; node < N; node += (int64_t)gridDim.x * NN_THREADS) {
    float x[NWB_NJ];
#pragma unroll
    for (int j = 0; j < NWB_NJ; ++j) x[j] = __ldg(soa + (int64_t)j * cap + node);   // 21 loads in flight
#pragma unroll
    for (int k = 0; k < NN_QT; ++k) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < NWB_NJ; ++j) {
        float d = qf[k][j] - x[j];
        s = fmaf(d, d, s);
      }
      if (s <= thr[k]) {     // rare: within the fp32 error bound of this thread's running minimum
        double sum = 0.0;
        const double* row = aos + node * NWB_NJ;
        for (int j = 0; j < NWB_NJ; ++j) {
          double d = __dsub_rn(qd[k][j], row[j]);
          sum = __dadd_rn(sum, __dmul_rn(d, d));      // RP:369: separately rounded multiply and add
        }
        double dn = sqrt(sum);                        // RP:373
        if (dn < best[k]) {                           // RP:377 strict
          best[k] = dn;
          besti[k] = (int)node;
        }
        // any node whose exact distance can still be < best has fp32 s below this bound
        double b = best[k];
        thr[k] = __double2float_ru((b * b + 1.2e-5 * b + 1e-10) * (1.0 + 1e-6));
      }
    }
  }
  // block reduction per query
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int k = 0; k < NN_QT; ++k) {
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
  if (tid < NN_QT) {
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

__global__ void nn_final_kernel(const double* __restrict__ part_d, const int32_t* __restrict__ part_i, int nb, int64_t nq,
                                int32_t* __restrict__ idx, double* __restrict__ dist) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nq) return;
  double d = part_d[k * nb];
  int i = part_i[k * nb];
  for (int b = 1; b < nb; ++b)
    if (lex_less(part_d[k * nb + b], part_i[k * nb + b], d, i)) { d = part_d[k * nb + b]; i = part_i[k * nb + b]; }
  idx[k] = i;
  if (dist) dist[k] = d;
}

}  // namespace nwb

using namespace nwb;

struct nwb_tree {
  nwb_ctx* ctx = nullptr;
  int64_t size = 0, cap = 0;
  float* soa = nullptr;
  double* aos = nullptr;
  int32_t* pred = nullptr;
  // scratch
  double* part_d = nullptr;
  int32_t* part_i = nullptr;
  size_t part_cap = 0;
  void* stage = nullptr;      // device staging for host-pointer calls
  size_t stage_bytes = 0;
};

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
  float* soa = nullptr;
  double* aos = nullptr;
  int32_t* pred = nullptr;
  TREE_CUDA(t, cudaMalloc(&soa, sizeof(float) * NWB_NJ * cap));
  TREE_CUDA(t, cudaMalloc(&aos, sizeof(double) * NWB_NJ * cap));
  TREE_CUDA(t, cudaMalloc(&pred, sizeof(int32_t) * cap));
  cudaStream_t s = t->ctx->stream;
  if (t->size > 0) {
    TREE_CUDA(t, cudaMemcpy2DAsync(soa, sizeof(float) * cap, t->soa, sizeof(float) * t->cap, sizeof(float) * t->size, NWB_NJ,
                                   cudaMemcpyDeviceToDevice, s));
    TREE_CUDA(t, cudaMemcpyAsync(aos, t->aos, sizeof(double) * NWB_NJ * t->size, cudaMemcpyDeviceToDevice, s));
    TREE_CUDA(t, cudaMemcpyAsync(pred, t->pred, sizeof(int32_t) * t->size, cudaMemcpyDeviceToDevice, s));
  }
  TREE_CUDA(t, cudaStreamSynchronize(s));
  cudaFree(t->soa);
  cudaFree(t->aos);
  cudaFree(t->pred);
  t->soa = soa;
  t->aos = aos;
  t->pred = pred;
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
  int64_t want = (t->size + NN_THREADS - 1) / NN_THREADS;
  int64_t cap = (int64_t)t->ctx->sm_count * 8;     // 8 resident blocks of 256 threads per SM
  return (int)std::max<int64_t>(1, std::min(want, cap));
}

int nn_launch(nwb_tree* t, const double* q_dev, int64_t nq, int32_t* idx_dev, double* dist_dev) {
  const int nb = nn_grid_blocks(t);
  size_t need = (size_t)nq * nb;
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
  cudaStream_t s = t->ctx->stream;
  TREE_CUDA(t, cudaEventRecord(t->ctx->ev0, s));
  const int64_t qtiles = (nq + NN_QT - 1) / NN_QT;
  for (int64_t y0 = 0; y0 < qtiles; y0 += 65535) {       // gridDim.y limit
    const int64_t ny = std::min<int64_t>(65535, qtiles - y0);
    const int64_t qoff = y0 * NN_QT;
    nn_scan_kernel<<<dim3(nb, (unsigned)ny), NN_THREADS, 0, s>>>(t->soa, t->aos, t->cap, t->size, q_dev + qoff * NWB_NJ,
                                                                  nq - qoff, t->part_d + qoff * nb, t->part_i + qoff * nb);
  }
  nn_final_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, s>>>(t->part_d, t->part_i, nb, nq, idx_dev, dist_dev);
  TREE_CUDA(t, cudaGetLastError());
  TREE_CUDA(t, cudaEventRecord(t->ctx->ev1, s));
  t->ctx->last_ms = -1;
  t->ctx->last_launches = (qtiles + 65534) / 65535 + 1;
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
  cudaFree(t->aos);
  cudaFree(t->pred);
  cudaFree(t->part_d);
  cudaFree(t->part_i);
  cudaFree(t->stage);
  delete t;
}

int nwb_tree_clear(nwb_tree* t) {
  if (!t) return NWB_E_INVALID;
  t->size = 0;
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
  tree_append_kernel<<<(unsigned)((e + 255) / 256), 256, 0, t->ctx->stream>>>(q_dev, pred_dev, n, t->soa, t->aos, t->pred, t->cap,
                                                                               t->size);
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
