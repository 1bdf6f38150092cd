// nwb_abi.cu - the extern "C" boundary (include/nwb.h): context, scene, batch operators.
//
// Host code is C++ that only marshals buffers and enqueues kernels; there is no CPU compute path:
// every operator ends in a kernel launch on the context's device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

#include "nwb_internal.h"
#include "nwb_tables.h"

using namespace nwb;

namespace {
thread_local std::string g_create_error;

#define NWB_CUDA(ctx, call) NWB_TRY_CUDA(ctx, call)

int fail(Ctx* ctx, int code, const char* msg) {
  if (ctx) ctx->err = msg;
  return code;
}

int ensure(Ctx* ctx, void** buf, size_t* have, size_t need) {
  if (*have >= need) return NWB_OK;
  if (*buf) cudaFree(*buf);
  *buf = nullptr;
  *have = 0;
  size_t cap = std::max(need, (size_t)1 << 20);
  NWB_CUDA(ctx, cudaMalloc(buf, cap));
  *have = cap;
  return NWB_OK;
}

}  // namespace

namespace nwb {
// Scenes of kCulledMinBoxes boxes and more: hierarchy + boxes in one device block [nodes | boxes], rebuilt on every change
// of the scene.  The stream is drained first: kernels in flight may still read the previous block.
static void upload_scene(Ctx* ctx, SceneView& view) {
  view = SceneView{nullptr, nullptr, 0, 0};
  if (ctx->boxes.size() < (size_t)kCulledMinBoxes) return;
  build_scene_bvh(ctx->boxes, ctx->bvh);
  const size_t nb = sizeof(SceneNode) * ctx->bvh.size(), bb = sizeof(SceneBox) * ctx->boxes.size();
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->d_scene_bytes < nb + bb) {
    if (ctx->d_scene) cudaFree(ctx->d_scene);
    ctx->d_scene = nullptr;
    ctx->d_scene_bytes = 0;
    const size_t cap = std::max(nb + bb, (size_t)1 << 16) * 2;
    if (cudaMalloc(&ctx->d_scene, cap) != cudaSuccess) {
      ctx->err = "scene upload: out of device memory";
      return;
    }
    ctx->d_scene_bytes = cap;
  }
  char* base = static_cast<char*>(ctx->d_scene);
  cudaMemcpy(base, ctx->bvh.data(), nb, cudaMemcpyHostToDevice);
  cudaMemcpy(base + nb, ctx->boxes.data(), bb, cudaMemcpyHostToDevice);
  view.nodes = reinterpret_cast<const SceneNode*>(base);
  view.boxes = reinterpret_cast<const SceneBox*>(base + nb);
  view.n_nodes = (int)ctx->bvh.size();
  view.n_boxes = (int)ctx->boxes.size();
}

void build_tables(Ctx* ctx) {
  build_validity_tables(ctx->params, ctx->boxes, true, ctx->tab_group);
  build_validity_tables(ctx->params, ctx->boxes, false, ctx->tab_all);
  SceneView view;
  upload_scene(ctx, view);
  const bool big = ctx->boxes.size() > (size_t)kMaxBoxes;    // small scenes keep their inline boxes for every other kernel
  ctx->tab_group.scene = ctx->tab_all.scene = big ? view : SceneView{nullptr, nullptr, 0, 0};
  ctx->tab_group.culled = ctx->tab_all.culled = view;
  ctx->scene_dirty = false;
  if (getenv("NWB_ENV_NO_CULL")) {                           // full evaluation of every (link, box) pair (roofline runs)
    ctx->tab_group.env_cull = ctx->tab_all.env_cull = 0;
    if (!big) ctx->tab_group.culled = ctx->tab_all.culled = SceneView{nullptr, nullptr, 0, 0};
  }
}
}  // namespace nwb

extern "C" {

int nwb_abi_version(void) { return NWB_ABI_VERSION; }

int nwb_params_default(nwb_params* p) {
  if (!p) return NWB_E_INVALID;
  std::memset(p, 0, sizeof(*p));
  p->device = 0;
  p->quirks = NWB_QUIRKS_DEFAULT;
  p->srdf_trim = 1;
  p->scale_origin = nwbm::kScaleOriginDefault;
  p->scale_sp = nwbm::kSupportScaleDefault;
  p->ds_tolerance = 0.01;
  p->ik_eps = 1e-3;
  p->ik_maxiter = 100;
  p->scg_max_samples = 5000;
  p->scg_max_ik_iterations = 5;
  p->planner_max_iterations = 8000;
  p->planner_step_factor = 0.1;
  p->pinv_eps = 1e-5;
  p->shortcut_waypoints = 100;
  p->shortcut_max_rounds = 500;
  p->r_arm_weight_approach = 1.0; p->l_arm_weight_approach = 1.0; p->legs_weight_approach = 1.0;
  p->r_arm_weight_manipulate = 1.0; p->l_arm_weight_manipulate = 0.0; p->legs_weight_manipulate = 1.0;
  p->manipulate_waypoints = 4;
  p->object_interp_points = 20;
  p->manipulation_velocity = 0.02;
  p->time_parameterization = 1;
  p->iptp_max_iterations = 100;
  p->iptp_max_time_change = 0.01;
  // planner joint order (RP:2003-2023): R leg (AnkleRoll, AnklePitch, KneePitch, HipPitch, HipRoll), L arm, R arm
  // (ShoulderPitch, ShoulderRoll, ElbowYaw, ElbowRoll, WristYaw), LHipYawPitch, L leg (HipRoll, HipPitch, KneePitch,
  // AnklePitch, AnkleRoll).  NAO V4 URDF velocity limits [rad/s], from memory (UNVERIFIED).
  static const double vmax[NWB_NJ] = {4.16174, 6.40239, 6.40239, 6.40239, 4.16174, 8.26797, 7.19407, 8.26797, 7.19407, 24.6229,
                                      8.26797, 7.19407, 8.26797, 7.19407, 24.6229, 4.16174, 4.16174, 6.40239, 6.40239, 6.40239,
                                      4.16174};
  for (int j = 0; j < NWB_NJ; ++j) {
    p->joint_max_velocity[j] = vmax[j];
    p->joint_max_acceleration[j] = 1.0;
  }
  return NWB_OK;
}

int nwb_ctx_create(const nwb_params* params, nwb_ctx** out) {
  if (!out) return NWB_E_INVALID;
  *out = nullptr;
  nwb_params p;
  if (params) p = *params; else nwb_params_default(&p);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    g_create_error = std::string("no usable CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback";
    return NWB_E_CUDA;
  }
  if (p.device < 0 || p.device >= ndev) {
    g_create_error = "device ordinal out of range";
    return NWB_E_INVALID;
  }
  nwb_ctx* ctx = new (std::nothrow) nwb_ctx();
  if (!ctx) return NWB_E_NOMEM;
  ctx->params = p;
  ctx->device = p.device;
  auto bail = [&](const char* what, cudaError_t ce) {
    g_create_error = std::string(what) + ": " + cudaGetErrorString(ce);
    nwb_ctx_destroy(ctx);
    return NWB_E_CUDA;
  };
  if ((e = cudaSetDevice(p.device)) != cudaSuccess) return bail("cudaSetDevice", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  ctx->stream = ctx->own_stream;
  if ((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess) return bail("cudaEventCreate", e);
  if ((e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return bail("cudaEventCreate", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  for (int i = 0; i < 2; ++i) {
    if ((e = cudaEventCreateWithFlags(&ctx->in_done[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->k_done[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
  }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, p.device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
  ctx->sm_count = prop.multiProcessorCount;
  if (prop.major < 10) {
    g_create_error = "this library is built for sm_100a (B200) only; found sm_" + std::to_string(prop.major * 10 + prop.minor);
    nwb_ctx_destroy(ctx);
    return NWB_E_CUDA;
  }
  ctx->force_table_pairs = getenv("NWB_FORCE_TABLE_PAIRS") != nullptr;
  build_tables(ctx);
  *out = ctx;
  return NWB_OK;
}

void nwb_ctx_destroy(nwb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
  if (ctx->d_in) cudaFree(ctx->d_in);
  if (ctx->d_scene) cudaFree(ctx->d_scene);
  if (ctx->d_out) cudaFree(ctx->d_out);
  if (ctx->d_plan) cudaFree(ctx->d_plan);
  if (ctx->d_forest) cudaFree(ctx->d_forest);
  if (ctx->d_rrt) cudaFree(ctx->d_rrt);
  if (ctx->h_flag) cudaFreeHost(ctx->h_flag);
  if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
  if (ctx->d_work) cudaFree(ctx->d_work);
  if (ctx->d_dbcache) cudaFree(ctx->d_dbcache);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  for (int i = 0; i < 2; ++i) {
    if (ctx->in_done[i]) cudaEventDestroy(ctx->in_done[i]);
    if (ctx->k_done[i]) cudaEventDestroy(ctx->k_done[i]);
  }
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

const char* nwb_last_error(const nwb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int nwb_ctx_synchronize(nwb_ctx* ctx) {
  if (!ctx) return NWB_E_INVALID;
  NWB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NWB_OK;
}

void* nwb_ctx_stream(nwb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int nwb_ctx_set_stream(nwb_ctx* ctx, void* cuda_stream) {
  if (!ctx) return NWB_E_INVALID;
  NWB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return NWB_OK;
}

int nwb_scene_clear(nwb_ctx* ctx) {
  if (!ctx) return NWB_E_INVALID;
  ctx->boxes.clear();
  build_tables(ctx);
  return NWB_OK;
}

int nwb_scene_add_box(nwb_ctx* ctx, const double* center, const double* size, const double* rot) {
  if (!ctx || !center || !size) return fail(ctx, NWB_E_INVALID, "nwb_scene_add_box: bad argument");
  if (ctx->boxes.size() >= (size_t)kSceneBoxLimit) return fail(ctx, NWB_E_INVALID, "nwb_scene_add_box: scene is full (2^20 boxes)");
  if (!(size[0] > 0 && size[1] > 0 && size[2] > 0)) return fail(ctx, NWB_E_INVALID, "nwb_scene_add_box: size must be > 0");
  SceneBox b{};
  for (int k = 0; k < 3; ++k) {
    b.c[k] = (float)center[k];
    b.h[k] = (float)(0.5 * size[k]);
  }
  for (int k = 0; k < 9; ++k) b.r[k] = rot ? (float)rot[k] : (k % 4 == 0 ? 1.f : 0.f);
  ctx->boxes.push_back(b);
  build_tables(ctx);
  return ctx->boxes.size() > (size_t)kMaxBoxes && !ctx->tab_group.scene.nodes ? NWB_E_NOMEM : NWB_OK;
}

int nwb_scene_add_boxes(nwb_ctx* ctx, const double* boxes15, int64_t n) {
  if (!ctx || n < 0 || (n > 0 && !boxes15)) return fail(ctx, NWB_E_INVALID, "nwb_scene_add_boxes: bad argument");
  if (ctx->boxes.size() + (size_t)n > (size_t)kSceneBoxLimit) return fail(ctx, NWB_E_INVALID, "nwb_scene_add_boxes: scene is full (2^20 boxes)");
  for (int64_t i = 0; i < n; ++i) {
    const double* v = boxes15 + 15 * i;
    if (!(v[3] > 0 && v[4] > 0 && v[5] > 0)) return fail(ctx, NWB_E_INVALID, "nwb_scene_add_boxes: size must be > 0");
  }
  for (int64_t i = 0; i < n; ++i) {
    const double* v = boxes15 + 15 * i;
    SceneBox b{};
    for (int k = 0; k < 3; ++k) {
      b.c[k] = (float)v[k];
      b.h[k] = (float)(0.5 * v[3 + k]);
    }
    for (int k = 0; k < 9; ++k) b.r[k] = (float)v[6 + k];
    ctx->boxes.push_back(b);
  }
  build_tables(ctx);                         // one hierarchy build + upload for the whole set
  return ctx->boxes.size() > (size_t)kMaxBoxes && !ctx->tab_group.scene.nodes ? NWB_E_NOMEM : NWB_OK;
}

int nwb_scene_num_boxes(const nwb_ctx* ctx) { return ctx ? (int)ctx->boxes.size() : NWB_E_INVALID; }

int nwb_scale_support_polygon(nwb_ctx* ctx, double scale_sp) {
  if (!ctx || !(scale_sp > 0)) return fail(ctx, NWB_E_INVALID, "scale_sp must be > 0");
  ctx->params.scale_sp = scale_sp;
  build_tables(ctx);
  return NWB_OK;
}

int nwb_get_joint_limits(const nwb_ctx* ctx, double* mn, double* mx) {
  if (!mn || !mx) return NWB_E_INVALID;
  for (int i = 0; i < NWB_NJ; ++i) {
    mn[i] = nwbm::kJointMin[i];
    mx[i] = nwbm::kJointMax[i];
  }
  return NWB_OK;
}

const char* nwb_frame_name(int i) { return (i >= 0 && i < nwbm::kNumFrames) ? nwbm::kFrameNames[i] : nullptr; }

void* nwb_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
  return p;
}
void nwb_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

// ---- validity -----------------------------------------------------------------------------------
static int feasible_dev(nwb_ctx* ctx, const ValidityTables& tab, const void* q, int dtype, int64_t n, uint8_t* valid,
                        uint8_t* reason, float* margins, bool collision_only) {
  NWB_CUDA(ctx, ev_record(ctx, ctx->ev0, ctx->stream));
  const GatherPeers* gp = (!collision_only && ctx->gather.n > 0) ? &ctx->gather : nullptr;
  NWB_CUDA(ctx, launch_validity(tab, q, dtype, n, valid, reason, margins, collision_only, ctx->params.srdf_trim != 0 && !ctx->force_table_pairs, ctx->stream, gp));
  NWB_CUDA(ctx, ev_record(ctx, ctx->ev1, ctx->stream));
  ctx->last_launches = n > 0 ? 1 : 0;
  ctx->last_ms = -1;   // resolved lazily by nwb_last_kernel_ms
  return NWB_OK;
}

int nwb_is_feasible_batch_dev(nwb_ctx* ctx, const void* q_dev, int dtype, int64_t n, uint8_t* valid_dev,
                              uint8_t* reason_dev, float* margins_dev) {
  if (!ctx) return NWB_E_INVALID;
  if (n < 0 || (n > 0 && (!q_dev || !valid_dev)) || (dtype != NWB_F64 && dtype != NWB_F32))
    return fail(ctx, NWB_E_INVALID, "nwb_is_feasible_batch_dev: bad argument");
  NWB_CUDA(ctx, cudaSetDevice(ctx->device));
  return feasible_dev(ctx, ctx->tab_group, q_dev, dtype, n, valid_dev, reason_dev, margins_dev, false);
}

// Host-buffer operator: chunks of the batch are copied in, checked and copied out on the context's
// stream.  With pinned caller buffers (nwb_host_alloc) the copies run at full PCIe speed.
static int feasible_host(nwb_ctx* ctx, const ValidityTables& tab, const void* q, int dtype, int64_t n, uint8_t* valid,
                         uint8_t* reason, float* margins, bool collision_only) {
  if (n == 0) return NWB_OK;
  NWB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int mcols = collision_only ? 1 : 2;
  const int64_t chunk = std::min<int64_t>(n, (int64_t)1 << 18);
  const size_t elem = dtype == NWB_F32 ? sizeof(float) : sizeof(double);
  const size_t in_bytes = (size_t)chunk * NWB_NJ * elem;
  const size_t chunk_a = ((size_t)chunk + 255) & ~(size_t)255;          // keeps every section 256 B aligned
  const size_t out_bytes = chunk_a * (2 + sizeof(float) * mcols);
  int rc;
  if ((rc = ensure(ctx, &ctx->d_in, &ctx->d_in_bytes, 2 * in_bytes)) != NWB_OK) return rc;
  if ((rc = ensure(ctx, &ctx->d_out, &ctx->d_out_bytes, 2 * out_bytes)) != NWB_OK) return rc;
  double total_ms = 0;
  int64_t launches = 0;
  // two buffers, one stream: copy-in of chunk k+1 is queued behind chunk k's kernel; a second
  // (copy) stream lets the DMA engines overlap with the kernel.
  cudaStream_t copy_stream = ctx->copy_stream;
  cudaEvent_t* in_done = ctx->in_done;
  cudaEvent_t* k_done = ctx->k_done;
  int64_t nchunks = (n + chunk - 1) / chunk;
  for (int64_t c = 0; c < nchunks; ++c) {
    const int s = (int)(c & 1);
    const int64_t off = c * chunk, cn = std::min(chunk, n - off);
    char* din = (char*)ctx->d_in + (size_t)s * in_bytes;
    char* dout = (char*)ctx->d_out + (size_t)s * out_bytes;
    uint8_t* d_valid = (uint8_t*)dout;
    uint8_t* d_reason = d_valid + chunk_a;
    float* d_marg = (float*)(dout + 2 * chunk_a);
    // the buffer pair s was last used by chunk c-2: its kernel / copy-out are ordered on ctx->stream
    if (c >= 2) NWB_CUDA(ctx, cudaStreamWaitEvent(copy_stream, k_done[s], 0));
    NWB_CUDA(ctx, cudaMemcpyAsync(din, static_cast<const char*>(q) + (size_t)off * NWB_NJ * elem, (size_t)cn * NWB_NJ * elem,
                                  cudaMemcpyHostToDevice, copy_stream));
    NWB_CUDA(ctx, cudaEventRecord(in_done[s], copy_stream));
    NWB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, in_done[s], 0));
    NWB_CUDA(ctx, launch_validity(tab, din, dtype, cn, d_valid, reason ? d_reason : nullptr, margins ? d_marg : nullptr,
                                  collision_only, ctx->params.srdf_trim != 0 && !ctx->force_table_pairs, ctx->stream));
    ++launches;
    NWB_CUDA(ctx, cudaMemcpyAsync(valid + off, d_valid, (size_t)cn, cudaMemcpyDeviceToHost, ctx->stream));
    if (reason) NWB_CUDA(ctx, cudaMemcpyAsync(reason + off, d_reason, (size_t)cn, cudaMemcpyDeviceToHost, ctx->stream));
    if (margins)
      NWB_CUDA(ctx, cudaMemcpyAsync(margins + off * mcols, d_marg, (size_t)cn * mcols * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    NWB_CUDA(ctx, cudaEventRecord(k_done[s], ctx->stream));
  }
  NWB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->last_launches = launches;
  ctx->last_ms = total_ms;
  return NWB_OK;
}

int nwb_is_feasible_batch(nwb_ctx* ctx, const double* q, int64_t n, uint8_t* valid, uint8_t* reason, float* margins) {
  if (!ctx) return NWB_E_INVALID;
  if (n < 0 || (n > 0 && (!q || !valid))) return fail(ctx, NWB_E_INVALID, "nwb_is_feasible_batch: bad argument");
  return feasible_host(ctx, ctx->tab_group, q, NWB_F64, n, valid, reason, margins, false);
}

int nwb_is_feasible_batch_f32(nwb_ctx* ctx, const float* q, int64_t n, uint8_t* valid, uint8_t* reason, float* margins) {
  if (!ctx) return NWB_E_INVALID;
  if (n < 0 || (n > 0 && (!q || !valid))) return fail(ctx, NWB_E_INVALID, "nwb_is_feasible_batch_f32: bad argument");
  return feasible_host(ctx, ctx->tab_group, q, NWB_F32, n, valid, reason, margins, false);
}

int nwb_is_state_colliding_batch(nwb_ctx* ctx, const double* q, int64_t n, uint8_t* colliding, float* min_separation) {
  if (!ctx) return NWB_E_INVALID;
  if (n < 0 || (n > 0 && (!q || !colliding))) return fail(ctx, NWB_E_INVALID, "nwb_is_state_colliding_batch: bad argument");
  return feasible_host(ctx, ctx->tab_all, q, NWB_F64, n, colliding, nullptr, min_separation, true);
}

// ---- FK -----------------------------------------------------------------------------------------
int nwb_fk_batch(nwb_ctx* ctx, const double* q, int64_t n, double* link_poses, double* com) {
  if (!ctx) return NWB_E_INVALID;
  if (n < 0 || (n > 0 && !q)) return fail(ctx, NWB_E_INVALID, "nwb_fk_batch: bad argument");
  if (n == 0) return NWB_OK;
  NWB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t chunk = std::min<int64_t>(n, 1 << 16);
  const size_t pose_per = (size_t)nwbm::kNumFrames * 12 * sizeof(double);
  int rc;
  if ((rc = ensure(ctx, &ctx->d_in, &ctx->d_in_bytes, (size_t)chunk * NWB_NJ * sizeof(double))) != NWB_OK) return rc;
  if ((rc = ensure(ctx, &ctx->d_out, &ctx->d_out_bytes, (size_t)chunk * (pose_per + 3 * sizeof(double)))) != NWB_OK) return rc;
  int64_t launches = 0;
  for (int64_t off = 0; off < n; off += chunk) {
    const int64_t cn = std::min(chunk, n - off);
    double* d_pose = (double*)ctx->d_out;
    double* d_com = (double*)((char*)ctx->d_out + (size_t)chunk * pose_per);
    NWB_CUDA(ctx, cudaMemcpyAsync(ctx->d_in, q + off * NWB_NJ, (size_t)cn * NWB_NJ * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NWB_CUDA(ctx, launch_fk(ctx->d_in, NWB_F64, cn, link_poses ? d_pose : nullptr, com ? d_com : nullptr, ctx->stream));
    ++launches;
    if (link_poses)
      NWB_CUDA(ctx, cudaMemcpyAsync(link_poses + (size_t)off * nwbm::kNumFrames * 12, d_pose, (size_t)cn * pose_per, cudaMemcpyDeviceToHost, ctx->stream));
    if (com) NWB_CUDA(ctx, cudaMemcpyAsync(com + off * 3, d_com, (size_t)cn * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    NWB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  ctx->last_launches = launches;
  return NWB_OK;
}

// ---- multi-GPU: peer-mapped result gather ------------------------------------------------------------
int nwb_dev_alloc(nwb_ctx* ctx, size_t bytes, void** out) {
  if (!ctx || !out) return NWB_E_INVALID;
  NWB_CUDA(ctx, cudaSetDevice(ctx->device));
  NWB_CUDA(ctx, cudaMalloc(out, bytes ? bytes : 1));
  return NWB_OK;
}
int nwb_dev_free(nwb_ctx* ctx, void* p) {
  if (!ctx) return NWB_E_INVALID;
  NWB_CUDA(ctx, cudaSetDevice(ctx->device));
  NWB_CUDA(ctx, cudaFree(p));
  return NWB_OK;
}
int nwb_ipc_export(nwb_ctx* ctx, void* dev_ptr, unsigned char* handle64) {
  if (!ctx || !dev_ptr || !handle64) return NWB_E_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  NWB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  NWB_CUDA(ctx, cudaIpcGetMemHandle(&h, dev_ptr));
  std::memcpy(handle64, &h, 64);
  return NWB_OK;
}
int nwb_ipc_import(nwb_ctx* ctx, const unsigned char* handle64, void** out) {
  if (!ctx || !handle64 || !out) return NWB_E_INVALID;
  NWB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  NWB_CUDA(ctx, cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return NWB_OK;
}
int nwb_ipc_close(nwb_ctx* ctx, void* p) {
  if (!ctx) return NWB_E_INVALID;
  NWB_CUDA(ctx, cudaSetDevice(ctx->device));
  NWB_CUDA(ctx, cudaIpcCloseMemHandle(p));
  return NWB_OK;
}
int nwb_set_gather_peers(nwb_ctx* ctx, void* const* gathered, int32_t n_ranks, int64_t my_offset) {
  if (!ctx || n_ranks < 0 || n_ranks > 8 || (n_ranks > 0 && !gathered) || my_offset < 0)
    return fail(ctx, NWB_E_INVALID, "nwb_set_gather_peers: bad argument (at most 8 ranks)");
  ctx->gather = GatherPeers{};
  for (int r = 0; r < n_ranks; ++r) ctx->gather.ptr[r] = static_cast<uint8_t*>(gathered[r]);
  ctx->gather.n = n_ranks;
  ctx->gather.offset = my_offset;
  return NWB_OK;
}

// ---- measurement helpers --------------------------------------------------------------------------
int nwb_ffma_peak(nwb_ctx* ctx, int iters, double* tflops_out, double* ms_out) {
  if (!ctx || iters <= 0) return fail(ctx, NWB_E_INVALID, "nwb_ffma_peak: bad argument");
  NWB_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc;
  if ((rc = ensure(ctx, &ctx->d_out, &ctx->d_out_bytes, 256)) != NWB_OK) return rc;
  double flops = 0;
  NWB_CUDA(ctx, launch_ffma_peak(ctx->sm_count, 16, (float*)ctx->d_out, ctx->stream, nullptr));   // warm-up
  NWB_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  NWB_CUDA(ctx, launch_ffma_peak(ctx->sm_count, iters, (float*)ctx->d_out, ctx->stream, &flops));
  NWB_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  NWB_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
  float ms = 0;
  NWB_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  if (ms_out) *ms_out = ms;
  if (tflops_out) *tflops_out = flops / (ms * 1e-3) / 1e12;
  ctx->last_launches = 2;
  return NWB_OK;
}

int nwb_set_kernel_timing(nwb_ctx* ctx, int enabled) {
  if (!ctx) return NWB_E_INVALID;
  ctx->timing = enabled != 0;
  ctx->last_ms = 0;
  return NWB_OK;
}

int nwb_last_kernel_ms(nwb_ctx* ctx, double* ms_out, int64_t* launches_out) {
  if (!ctx) return NWB_E_INVALID;
  if (!ctx->timing) {                       // no events were recorded: only the launch count is known
    if (ms_out) *ms_out = -1.0;
    if (launches_out) *launches_out = ctx->last_launches;
    return NWB_OK;
  }
  if (ctx->last_ms < 0) {
    NWB_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    float ms = 0;
    NWB_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
  }
  if (ms_out) *ms_out = ctx->last_ms;
  if (launches_out) *launches_out = ctx->last_launches;
  return NWB_OK;
}

}  // extern "C"
