// planner_abi.cu - extern "C" entry points of the local-planner operators (include/nwb.h).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "nwb_internal.h"
#include "planner_core.cuh"

using namespace nwb;

namespace {
#define P_CUDA(ctx, call) NWB_TRY_CUDA(ctx, call)

int bad(nwb_ctx* ctx, const char* msg) {
  if (ctx) ctx->err = msg;
  return NWB_E_INVALID;
}

using Arena = nwb::DeviceArena;

bool static_pairs(const nwb_ctx* ctx) { return ctx->params.srdf_trim != 0 && !ctx->force_table_pairs; }
}  // namespace

namespace nwb {
void make_plan_const(const nwb_ctx* ctx, const double* weights, double step, PlanConst& pc) {
  for (int j = 0; j < NWB_NJ; ++j) {
    pc.jmin[j] = nwbm::kJointMin[j];
    pc.jmax[j] = nwbm::kJointMax[j];
    pc.weights[j] = weights ? weights[j] : 1.0;
  }
  pc.step = step;
  pc.ds_tolerance = ctx->params.ds_tolerance;
  pc.ik_eps = ctx->params.ik_eps;
  pc.pinv_eps = ctx->params.pinv_eps;
  pc.ik_maxiter = ctx->params.ik_maxiter;
  pc.quirks = (unsigned)ctx->params.quirks;
}
}  // namespace nwb

extern "C" {

int nwb_steer_project_batch_dev(nwb_ctx* ctx, const double* q_near, const double* q_target, int64_t n, const double* weights,
                                double step, double* q_out, int32_t* status, int32_t* ds_status) {
  if (!ctx) return NWB_E_INVALID;
  if (n < 0 || (n > 0 && (!q_near || !q_target || !q_out || !status))) return bad(ctx, "nwb_steer_project_batch_dev: bad argument");
  P_CUDA(ctx, cudaSetDevice(ctx->device));
  PlanConst pc;
  make_plan_const(ctx, weights, step, pc);
  const size_t work_need = steer_scratch_bytes(n);
  if (n >= 4096 && ctx->d_work_bytes < work_need) {
    if (ctx->d_work) cudaFree(ctx->d_work);
    ctx->d_work = nullptr;
    ctx->d_work_bytes = 0;
    P_CUDA(ctx, cudaMalloc(&ctx->d_work, work_need));
    ctx->d_work_bytes = work_need;
  }
  P_CUDA(ctx, ev_record(ctx, ctx->ev0, ctx->stream));
  P_CUDA(ctx, launch_steer_project(pc, q_near, q_target, n, q_out, status, ds_status, nullptr, n >= 4096 ? ctx->d_work : nullptr,
                                   ctx->stream));
  P_CUDA(ctx, ev_record(ctx, ctx->ev1, ctx->stream));
  ctx->last_ms = -1;
  ctx->last_launches = n >= 4096 ? 3 : (n > 0);
  return NWB_OK;
}

int nwb_steer_project_batch(nwb_ctx* ctx, const double* q_near, const double* q_target, int64_t n, const double* weights,
                            double step, double* q_out, int32_t* status, int32_t* ds_status) {
  if (!ctx) return NWB_E_INVALID;
  if (n < 0 || (n > 0 && (!q_near || !q_target || !q_out || !status))) return bad(ctx, "nwb_steer_project_batch: bad argument");
  if (n == 0) return NWB_OK;
  P_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t qb = sizeof(double) * NWB_NJ * n, ib = sizeof(int32_t) * n;
  Arena A(ctx);
  size_t o_near = A.add(qb), o_tgt = A.add(qb), o_out = A.add(qb), o_st = A.add(ib), o_ds = A.add(ib);
  int rc = A.commit();
  if (rc != NWB_OK) return rc;
  cudaStream_t s = ctx->stream;
  P_CUDA(ctx, cudaMemcpyAsync(A.at<double>(o_near), q_near, qb, cudaMemcpyHostToDevice, s));
  P_CUDA(ctx, cudaMemcpyAsync(A.at<double>(o_tgt), q_target, qb, cudaMemcpyHostToDevice, s));
  // rows of TRAPPED entries are left untouched on the host: copy the caller's buffer in first
  P_CUDA(ctx, cudaMemcpyAsync(A.at<double>(o_out), q_out, qb, cudaMemcpyHostToDevice, s));
  rc = nwb_steer_project_batch_dev(ctx, A.at<double>(o_near), A.at<double>(o_tgt), n, weights, step, A.at<double>(o_out),
                                   A.at<int32_t>(o_st), A.at<int32_t>(o_ds));
  if (rc != NWB_OK) return rc;
  P_CUDA(ctx, cudaMemcpyAsync(q_out, A.at<double>(o_out), qb, cudaMemcpyDeviceToHost, s));
  P_CUDA(ctx, cudaMemcpyAsync(status, A.at<int32_t>(o_st), ib, cudaMemcpyDeviceToHost, s));
  if (ds_status) P_CUDA(ctx, cudaMemcpyAsync(ds_status, A.at<int32_t>(o_ds), ib, cudaMemcpyDeviceToHost, s));
  P_CUDA(ctx, cudaStreamSynchronize(s));
  return NWB_OK;
}

int nwb_connect_batch(nwb_ctx* ctx, const double* q_start, const double* q_connect, int64_t n_edges, const double* weights,
                      double step, int32_t max_steps, double* nodes_out, int32_t* n_added, int32_t* status) {
  if (!ctx) return NWB_E_INVALID;
  if (n_edges < 0 || max_steps <= 0 || (n_edges > 0 && (!q_start || !q_connect || !nodes_out || !n_added || !status)))
    return bad(ctx, "nwb_connect_batch: bad argument");
  if (n_edges == 0) return NWB_OK;
  P_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t qb = sizeof(double) * NWB_NJ * n_edges, nb = qb * max_steps, ib = sizeof(int32_t) * n_edges;
  Arena A(ctx);
  size_t o_s = A.add(qb), o_c = A.add(qb), o_n = A.add(nb), o_a = A.add(ib), o_st = A.add(ib), o_off = A.add(sizeof(int64_t) * n_edges),
         o_p = A.add(nb / 2 + 256);                        // packed rows: used only while they fill under half of o_n
  int rc = A.commit();
  if (rc != NWB_OK) return rc;
  cudaStream_t s = ctx->stream;
  PlanConst pc;
  make_plan_const(ctx, weights, step, pc);
  P_CUDA(ctx, cudaMemcpyAsync(A.at<double>(o_s), q_start, qb, cudaMemcpyHostToDevice, s));
  P_CUDA(ctx, cudaMemcpyAsync(A.at<double>(o_c), q_connect, qb, cudaMemcpyHostToDevice, s));
  P_CUDA(ctx, ev_record(ctx, ctx->ev0, s));
  P_CUDA(ctx, launch_connect(pc, ctx->tab_group, static_pairs(ctx), A.at<double>(o_s), A.at<double>(o_c), n_edges, max_steps,
                             A.at<double>(o_n), A.at<int32_t>(o_a), A.at<int32_t>(o_st), s));
  P_CUDA(ctx, ev_record(ctx, ctx->ev1, s));
  ctx->last_ms = -1;
  P_CUDA(ctx, cudaMemcpyAsync(n_added, A.at<int32_t>(o_a), ib, cudaMemcpyDeviceToHost, s));
  P_CUDA(ctx, cudaMemcpyAsync(status, A.at<int32_t>(o_st), ib, cudaMemcpyDeviceToHost, s));
  P_CUDA(ctx, cudaStreamSynchronize(s));
  // Only the appended rows travel back: the walks fill a few per cent of [n_edges][max_steps][21] (1024 walks over a
  // 10^5-node tree: 4956 of 131 072 rows), so the rows are packed on the device and scattered into the caller's layout
  // on the host.  Rows past n_added[e] of the caller's buffer are left as they were.
  std::vector<int64_t> off((size_t)n_edges);
  int64_t total = 0;
  for (int64_t e = 0; e < n_edges; ++e) {
    off[(size_t)e] = total;
    total += std::max(0, std::min(n_added[e], max_steps));
  }
  ctx->last_launches = 1;
  if (total == 0) return NWB_OK;
  if ((size_t)total * 2 >= (size_t)n_edges * max_steps) {      // dense result: the straight copy is as cheap
    P_CUDA(ctx, cudaMemcpyAsync(nodes_out, A.at<double>(o_n), nb, cudaMemcpyDeviceToHost, s));
    P_CUDA(ctx, cudaStreamSynchronize(s));
    return NWB_OK;
  }
  const size_t pb = sizeof(double) * NWB_NJ * (size_t)total;
  double* d_packed = A.at<double>(o_p);
  P_CUDA(ctx, cudaMemcpyAsync(A.at<int64_t>(o_off), off.data(), sizeof(int64_t) * n_edges, cudaMemcpyHostToDevice, s));
  P_CUDA(ctx, launch_pack_rows(A.at<double>(o_n), A.at<int32_t>(o_a), A.at<int64_t>(o_off), n_edges, max_steps, d_packed, s));
  std::vector<double> packed((size_t)total * NWB_NJ);
  P_CUDA(ctx, cudaMemcpyAsync(packed.data(), d_packed, pb, cudaMemcpyDeviceToHost, s));
  P_CUDA(ctx, cudaStreamSynchronize(s));
  for (int64_t e = 0; e < n_edges; ++e) {
    const int64_t k = std::max(0, std::min(n_added[e], max_steps));
    if (k) std::memcpy(nodes_out + (size_t)e * max_steps * NWB_NJ, packed.data() + (size_t)off[(size_t)e] * NWB_NJ, sizeof(double) * NWB_NJ * (size_t)k);
  }
  ctx->last_launches = 2;
  return NWB_OK;
}

int nwb_is_path_valid_batch_dev(nwb_ctx* ctx, const double* qa, const double* qb, int64_t n_edges, int32_t k, uint8_t* ok,
                                int32_t* first_bad) {
  if (!ctx) return NWB_E_INVALID;
  if (n_edges < 0 || k < 0 || (n_edges > 0 && (!qa || !qb || !first_bad))) return bad(ctx, "nwb_is_path_valid_batch_dev: bad argument");
  P_CUDA(ctx, cudaSetDevice(ctx->device));
  P_CUDA(ctx, ev_record(ctx, ctx->ev0, ctx->stream));
  P_CUDA(ctx, launch_edge_check(ctx->tab_all, static_pairs(ctx), qa, qb, n_edges, k, ok, first_bad, ctx->stream));
  P_CUDA(ctx, ev_record(ctx, ctx->ev1, ctx->stream));
  ctx->last_ms = -1;
  ctx->last_launches = n_edges > 0 ? 1 : 0;
  return NWB_OK;
}

int nwb_is_path_valid_batch(nwb_ctx* ctx, const double* qa, const double* qb, int64_t n_edges, int32_t k, uint8_t* ok,
                            int32_t* first_bad) {
  if (!ctx) return NWB_E_INVALID;
  if (n_edges < 0 || k < 0 || (n_edges > 0 && (!qa || !qb || !ok))) return bad(ctx, "nwb_is_path_valid_batch: bad argument");
  if (n_edges == 0) return NWB_OK;
  P_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t qbytes = sizeof(double) * NWB_NJ * n_edges;
  Arena A(ctx);
  size_t o_a = A.add(qbytes), o_b = A.add(qbytes), o_ok = A.add(n_edges), o_fb = A.add(sizeof(int32_t) * n_edges);
  int rc = A.commit();
  if (rc != NWB_OK) return rc;
  cudaStream_t s = ctx->stream;
  P_CUDA(ctx, cudaMemcpyAsync(A.at<double>(o_a), qa, qbytes, cudaMemcpyHostToDevice, s));
  P_CUDA(ctx, cudaMemcpyAsync(A.at<double>(o_b), qb, qbytes, cudaMemcpyHostToDevice, s));
  rc = nwb_is_path_valid_batch_dev(ctx, A.at<double>(o_a), A.at<double>(o_b), n_edges, k, A.at<uint8_t>(o_ok), A.at<int32_t>(o_fb));
  if (rc != NWB_OK) return rc;
  P_CUDA(ctx, cudaMemcpyAsync(ok, A.at<uint8_t>(o_ok), n_edges, cudaMemcpyDeviceToHost, s));
  if (first_bad) P_CUDA(ctx, cudaMemcpyAsync(first_bad, A.at<int32_t>(o_fb), sizeof(int32_t) * n_edges, cudaMemcpyDeviceToHost, s));
  P_CUDA(ctx, cudaStreamSynchronize(s));
  return NWB_OK;
}

int nwb_sample_stable_configs_dev(nwb_ctx* ctx, uint64_t seed, uint64_t first, int64_t count, int32_t max_ik_iter,
                                  double* rows, uint64_t* index, int64_t cap, unsigned long long* n_out, uint8_t* stage) {
  if (!ctx) return NWB_E_INVALID;
  if (count < 0 || cap < 0 || (count > 0 && (!rows || !index || !n_out))) return bad(ctx, "nwb_sample_stable_configs_dev: bad argument");
  P_CUDA(ctx, cudaSetDevice(ctx->device));
  if (count > (int64_t)0xffffffff) return bad(ctx, "nwb_sample_stable_configs_dev: at most 2^32 - 1 samples per call");
  PlanConst pc;
  make_plan_const(ctx, nullptr, ctx->params.planner_step_factor, pc);
  const size_t work_need = sampler_scratch_bytes(count);
  if (ctx->d_work_bytes < work_need) {
    if (ctx->d_work) cudaFree(ctx->d_work);
    ctx->d_work = nullptr;
    ctx->d_work_bytes = 0;
    P_CUDA(ctx, cudaMalloc(&ctx->d_work, work_need));
    ctx->d_work_bytes = work_need;
  }
  P_CUDA(ctx, ev_record(ctx, ctx->ev0, ctx->stream));
  P_CUDA(ctx, launch_sample(pc, ctx->tab_group, static_pairs(ctx), seed, first, count, max_ik_iter, rows, index, cap, n_out, stage,
                            ctx->d_work, ctx->stream));
  P_CUDA(ctx, ev_record(ctx, ctx->ev1, ctx->stream));
  ctx->last_ms = -1;
  ctx->last_launches = count > 0 ? 3 * (int)((count + ((int64_t)1 << 22) - 1) >> 22) : 0;
  return NWB_OK;
}

int nwb_sample_stable_configs(nwb_ctx* ctx, uint64_t seed, uint64_t first, int64_t count, int32_t max_ik_iter, double* rows_out,
                              int64_t cap, int64_t* n_out, uint64_t* index_out, uint8_t* stage_out) {
  if (!ctx) return NWB_E_INVALID;
  if (count < 0 || cap < 0 || !n_out || (cap > 0 && !rows_out)) return bad(ctx, "nwb_sample_stable_configs: bad argument");
  *n_out = 0;
  if (count == 0) return NWB_OK;
  P_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const int64_t chunk = std::min<int64_t>(count, (int64_t)1 << 20);
  Arena A(ctx);
  size_t o_rows = A.add(sizeof(double) * NWB_NJ * chunk), o_idx = A.add(sizeof(uint64_t) * chunk), o_n = A.add(8),
         o_stage = A.add(chunk);
  int rc = A.commit();
  if (rc != NWB_OK) return rc;
  std::vector<double> rows;
  std::vector<uint64_t> idx;
  int64_t launches = 0;
  for (int64_t off = 0; off < count; off += chunk) {
    const int64_t cn = std::min(chunk, count - off);
    P_CUDA(ctx, cudaMemsetAsync(A.at<char>(o_n), 0, 8, s));
    rc = nwb_sample_stable_configs_dev(ctx, seed, first + (uint64_t)off, cn, max_ik_iter, A.at<double>(o_rows), A.at<uint64_t>(o_idx),
                                       cn, A.at<unsigned long long>(o_n), stage_out ? A.at<uint8_t>(o_stage) : nullptr);
    if (rc != NWB_OK) return rc;
    ++launches;
    unsigned long long got = 0;
    P_CUDA(ctx, cudaMemcpyAsync(&got, A.at<char>(o_n), 8, cudaMemcpyDeviceToHost, s));
    if (stage_out) P_CUDA(ctx, cudaMemcpyAsync(stage_out + off, A.at<uint8_t>(o_stage), cn, cudaMemcpyDeviceToHost, s));
    P_CUDA(ctx, cudaStreamSynchronize(s));
    const size_t base = idx.size();
    rows.resize((base + got) * NWB_NJ);
    idx.resize(base + got);
    if (got) {
      P_CUDA(ctx, cudaMemcpyAsync(rows.data() + base * NWB_NJ, A.at<double>(o_rows), sizeof(double) * NWB_NJ * got, cudaMemcpyDeviceToHost, s));
      P_CUDA(ctx, cudaMemcpyAsync(idx.data() + base, A.at<uint64_t>(o_idx), sizeof(uint64_t) * got, cudaMemcpyDeviceToHost, s));
      P_CUDA(ctx, cudaStreamSynchronize(s));
    }
  }
  // arrival order is not deterministic: restore ascending sample index (the order the reference's
  // serial loop writes its file in, SCG:463-468)
  std::vector<size_t> order(idx.size());
  std::iota(order.begin(), order.end(), (size_t)0);
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return idx[a] < idx[b]; });
  *n_out = (int64_t)idx.size();
  const int64_t m = std::min<int64_t>(cap, (int64_t)idx.size());
  for (int64_t r = 0; r < m; ++r) {
    std::memcpy(rows_out + r * NWB_NJ, rows.data() + order[r] * NWB_NJ, sizeof(double) * NWB_NJ);
    if (index_out) index_out[r] = idx[order[r]];
  }
  ctx->last_launches = launches;
  if ((int64_t)idx.size() > cap) {
    ctx->err = "nwb_sample_stable_configs: more rows accepted than the output capacity";
    return NWB_E_NOMEM;
  }
  return NWB_OK;
}

// ---- on-disk formats -------------------------------------------------------------------------------------
int nwb_write_configs(const char* path, const double* rows, int64_t n_rows) {
  if (!path || n_rows < 0 || (n_rows > 0 && !rows)) return NWB_E_INVALID;
  std::remove(path);                                    // SCG:191, RP:1428: the old file is removed first
  FILE* f = std::fopen(path, "w");
  if (!f) return NWB_E_IO;
  for (int64_t r = 0; r < n_rows; ++r) {
    for (int j = 0; j < NWB_NJ; ++j) std::fprintf(f, "%g ", rows[r * NWB_NJ + j]);   // ostream default: 6 significant digits
    std::fputc('\n', f);
  }
  return std::fclose(f) == 0 ? NWB_OK : NWB_E_IO;
}

int nwb_load_ds_database(const char* path, double* rows, int64_t cap, int64_t* n_rows) {
  if (!path || !n_rows || cap < 0 || (cap > 0 && !rows)) return NWB_E_INVALID;
  *n_rows = 0;
  FILE* f = std::fopen(path, "r");
  if (!f) return NWB_E_IO;
  std::string line;
  int c;
  int64_t n = 0;
  auto flush = [&]() {
    const char* p = line.c_str();
    char* end = nullptr;
    double v[NWB_NJ];
    int got = 0;
    for (; got < NWB_NJ; ++got) {
      v[got] = std::strtod(p, &end);
      if (end == p) break;
      p = end;
    }
    if (got == NWB_NJ) {                                // blank / short lines are not configurations
      if (n < cap) std::memcpy(rows + n * NWB_NJ, v, sizeof(v));
      ++n;
    }
    line.clear();
  };
  while ((c = std::fgetc(f)) != EOF) {
    if (c == '\n') flush();
    else line.push_back((char)c);
  }
  if (!line.empty()) flush();
  std::fclose(f);
  *n_rows = n;
  return NWB_OK;
}

int nwb_generate_ds_database(nwb_ctx* ctx, const nwb_generate_ds_configs_request* req, nwb_generate_ds_configs_response* resp,
                             const char* path, uint64_t seed) {
  if (!ctx) return NWB_E_INVALID;
  if (!req || !resp || !path || req->max_samples < 0) return bad(ctx, "nwb_generate_ds_database: bad argument");
  resp->num_configs_generated = 0;
  std::vector<double> rows((size_t)std::max(req->max_samples, 1) * NWB_NJ);
  int64_t n = 0;
  int rc = nwb_sample_stable_configs(ctx, seed, 0, req->max_samples, req->max_ik_iterations, rows.data(), req->max_samples, &n,
                                     nullptr, nullptr);
  if (rc != NWB_OK) return rc;
  rc = nwb_write_configs(path, rows.data(), n);
  if (rc != NWB_OK) {
    ctx->err = std::string("nwb_generate_ds_database: cannot write ") + path;
    return rc;
  }
  resp->num_configs_generated = (int32_t)n;
  ctx->ds_db.assign(rows.begin(), rows.begin() + (size_t)n * NWB_NJ);   // planner_->load_DS_database(SSC_DS_CONFIGS), NPC:144
  ++ctx->ds_db_version;
  for (double& v : ctx->ds_db) {                                        // ... which reads the 6-digit text back
    char buf[64];
    std::snprintf(buf, sizeof buf, "%g", v);
    v = std::strtod(buf, nullptr);
  }
  return NWB_OK;
}

}  // extern "C"
