"""Thin Python handle over the C ABI (include/nwb.h) for tests and benchmarks.

All compute happens in libnwb.so (C++ host code + CUDA kernels); this module only marshals NumPy
arrays / raw device pointers.  Method names follow the reference's operator surface:
  isFeasible            RRT_Planner::isFeasible            rrt_planner.h:73
  scale_support_polygon RRT_Planner::scale_support_polygon rrt_planner.h:65
"""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import NJ, NFRAMES, NwbError, default_params, load_library


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """nwb_ctx: robot model + planning scene + parameters on one CUDA device."""

    def __init__(self, params=None, **overrides):
        self.lib = load_library()
        self.params = params or default_params(**overrides)
        h = C.c_void_p()
        rc = self.lib.nwb_ctx_create(C.byref(self.params), C.byref(h))
        if rc != 0:
            raise NwbError(f"nwb_ctx_create failed ({rc}): {self.lib.nwb_last_error(None).decode()}")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.nwb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            raise NwbError(f"{what} failed ({rc}): {self.lib.nwb_last_error(self.h).decode()}")

    # ---- parameters -----------------------------------------------------------------------------
    def scale_support_polygon(self, scale_sp):
        self._check(self.lib.nwb_scale_support_polygon(self.h, float(scale_sp)), "nwb_scale_support_polygon")

    def joint_limits(self):
        lo, hi = np.empty(NJ), np.empty(NJ)
        self._check(self.lib.nwb_get_joint_limits(self.h, lo.ctypes.data_as(C.POINTER(C.c_double)),
                                                  hi.ctypes.data_as(C.POINTER(C.c_double))), "nwb_get_joint_limits")
        return lo, hi

    def synchronize(self):
        self._check(self.lib.nwb_ctx_synchronize(self.h), "nwb_ctx_synchronize")

    @property
    def stream(self):
        return self.lib.nwb_ctx_stream(self.h)

    def set_stream(self, cuda_stream):
        """Launch on a caller-owned cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); None restores."""
        self._check(self.lib.nwb_ctx_set_stream(self.h, cuda_stream), "nwb_ctx_set_stream")

    # ---- planning scene -------------------------------------------------------------------------
    def scene_clear(self):
        self._check(self.lib.nwb_scene_clear(self.h), "nwb_scene_clear")

    def scene_add_box(self, center, size, rot=None):
        pd = C.POINTER(C.c_double)
        c = np.ascontiguousarray(center, dtype=np.float64)
        s = np.ascontiguousarray(size, dtype=np.float64)
        r = None if rot is None else np.ascontiguousarray(rot, dtype=np.float64).reshape(9)
        self._check(self.lib.nwb_scene_add_box(self.h, c.ctypes.data_as(pd), s.ctypes.data_as(pd),
                                               None if r is None else r.ctypes.data_as(pd)), "nwb_scene_add_box")

    def scene_add_boxes(self, boxes15):
        """rows of [cx,cy,cz, sx,sy,sz, R row-major] (tests/oracle_lib.scene_boxes layout)."""
        for b in np.asarray(boxes15, dtype=np.float64).reshape(-1, 15):
            self.scene_add_box(b[0:3], b[3:6], b[6:15])

    @property
    def num_boxes(self):
        return self.lib.nwb_scene_num_boxes(self.h)

    # ---- validity -------------------------------------------------------------------------------
    def isFeasible(self, q, want_reason=True, want_margins=False):
        """Batched RRT_Planner::isFeasible.  q [n,21] float64 (host).  Returns (valid, reason, margins)."""
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, NJ)
        n = len(q)
        valid = np.empty(n, np.uint8)
        reason = np.empty(n, np.uint8) if want_reason else None
        margins = np.empty((n, 2), np.float32) if want_margins else None
        self._check(self.lib.nwb_is_feasible_batch(self.h, _p(q), n, _p(valid), _p(reason), _p(margins)),
                    "nwb_is_feasible_batch")
        return valid, reason, margins

    def is_feasible_host_buffers(self, q_ptr, n, valid_ptr, reason_ptr=None, margins_ptr=None):
        """Raw-pointer form (pinned buffers from host_alloc) used by the end-to-end benchmark."""
        self._check(self.lib.nwb_is_feasible_batch(self.h, q_ptr, n, valid_ptr, reason_ptr, margins_ptr),
                    "nwb_is_feasible_batch")

    def is_feasible_dev(self, q_dev, dtype, n, valid_dev, reason_dev=None, margins_dev=None):
        """Device-pointer form: only enqueues on the context's stream."""
        self._check(self.lib.nwb_is_feasible_batch_dev(self.h, q_dev, dtype, n, valid_dev, reason_dev, margins_dev),
                    "nwb_is_feasible_batch_dev")

    def is_state_colliding(self, q, want_separation=False):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, NJ)
        n = len(q)
        col = np.empty(n, np.uint8)
        sep = np.empty(n, np.float32) if want_separation else None
        self._check(self.lib.nwb_is_state_colliding_batch(self.h, _p(q), n, _p(col), _p(sep)),
                    "nwb_is_state_colliding_batch")
        return col, sep

    def fk(self, q, want_poses=True, want_com=True):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, NJ)
        n = len(q)
        poses = np.empty((n, NFRAMES, 3, 4)) if want_poses else None
        com = np.empty((n, 3)) if want_com else None
        self._check(self.lib.nwb_fk_batch(self.h, _p(q), n, _p(poses), _p(com)), "nwb_fk_batch")
        return poses, com

    # ---- local planner operators ----------------------------------------------------------------
    @staticmethod
    def _weights(w):
        return None if w is None else np.ascontiguousarray(w, dtype=np.float64).reshape(NJ)

    def steer_project(self, q_near, q_target, weights=None, step=0.1):
        """generate_q_new + enforce_DS_Constraints (rrt_planner.cpp:424-593) -> (q_out, status, ds_status)."""
        q_near = np.ascontiguousarray(q_near, dtype=np.float64).reshape(-1, NJ)
        q_target = np.ascontiguousarray(q_target, dtype=np.float64).reshape(-1, NJ)
        n = len(q_near)
        w = self._weights(weights)
        out = np.full((n, NJ), np.nan)
        st = np.empty(n, np.int32)
        ds = np.empty(n, np.int32)
        self._check(self.lib.nwb_steer_project_batch(self.h, _p(q_near), _p(q_target), n, _p(w), float(step), _p(out), _p(st),
                                                     _p(ds)), "nwb_steer_project_batch")
        return out, st, ds

    def connect(self, q_start, q_connect, weights=None, step=0.1, max_steps=128):
        """connectTree edge walks (rrt_planner.cpp:926-1031) -> (nodes [n,max_steps,21], n_added, status)."""
        q_start = np.ascontiguousarray(q_start, dtype=np.float64).reshape(-1, NJ)
        q_connect = np.ascontiguousarray(q_connect, dtype=np.float64).reshape(-1, NJ)
        n = len(q_start)
        w = self._weights(weights)
        nodes = np.full((n, max_steps, NJ), np.nan)
        added = np.empty(n, np.int32)
        st = np.empty(n, np.int32)
        self._check(self.lib.nwb_connect_batch(self.h, _p(q_start), _p(q_connect), n, _p(w), float(step), int(max_steps),
                                               _p(nodes), _p(added), _p(st)), "nwb_connect_batch")
        return nodes, added, st

    def isPathValid(self, qa, qb, k=100):
        """interpolateWaypoints + PlanningScene::isPathValid (rrt_planner.cpp:1876) -> (ok, first_bad)."""
        qa = np.ascontiguousarray(qa, dtype=np.float64).reshape(-1, NJ)
        qb = np.ascontiguousarray(qb, dtype=np.float64).reshape(-1, NJ)
        n = len(qa)
        ok = np.empty(n, np.uint8)
        fb = np.empty(n, np.int32)
        self._check(self.lib.nwb_is_path_valid_batch(self.h, _p(qa), _p(qb), n, int(k), _p(ok), _p(fb)),
                    "nwb_is_path_valid_batch")
        return ok, fb

    def sample_stable_configs(self, seed, first, count, max_ik_iter=5, cap=None, want_stage=True):
        """sample_stable_configs database branch (stable_config_generator.cpp:278-480) -> (rows, index, stage)."""
        cap = count if cap is None else cap
        rows = np.empty((cap, NJ))
        idx = np.empty(cap, np.uint64)
        stage = np.empty(count, np.uint8) if want_stage else None
        n = C.c_int64(0)
        self._check(self.lib.nwb_sample_stable_configs(self.h, int(seed), int(first), int(count), int(max_ik_iter), _p(rows),
                                                       int(cap), C.byref(n), _p(idx), _p(stage)), "nwb_sample_stable_configs")
        return rows[:n.value].copy(), idx[:n.value].copy(), stage

    # ---- RRT-Connect driver ---------------------------------------------------------------------
    def solveQuery(self, starts, goals, db, seeds, weights=None, max_iter=8000, step=0.1, tree_capacity=4096,
                   raw_path_capacity=1024):
        """setStartGoalConfigs + solveQuery + raw path extraction for a batch of queries.
        Returns (list of result dicts, list of raw paths [len,21])."""
        starts = np.ascontiguousarray(starts, dtype=np.float64).reshape(-1, NJ)
        goals = np.ascontiguousarray(goals, dtype=np.float64).reshape(-1, NJ)
        db = np.ascontiguousarray(db, dtype=np.float64).reshape(-1, NJ)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        nq = len(starts)
        w = self._weights(weights)
        res = (_abi.NwbPlanResult * nq)()
        raw = np.full((nq, raw_path_capacity, NJ), np.nan)
        self._check(self.lib.nwb_solve_query_batch(self.h, _p(starts), _p(goals), nq, _p(db), len(db), _p(w), _p(seeds),
                                                   int(max_iter), float(step), int(tree_capacity), C.byref(res), _p(raw),
                                                   int(raw_path_capacity)), "nwb_solve_query_batch")
        out = [{n: getattr(r, n) for n, _ in _abi.NwbPlanResult._fields_} for r in res]
        paths = [raw[k, :min(out[k]["raw_path_len"], raw_path_capacity)].copy() for k in range(nq)]
        return out, paths

    def shortcut_path(self, raw):
        raw = np.ascontiguousarray(raw, dtype=np.float64).reshape(-1, NJ)
        out = np.empty((len(raw) + 4, NJ))
        edges = C.c_int32(0)
        n = self.lib.nwb_shortcut_path(self.h, _p(raw), len(raw), _p(out), len(out), C.byref(edges))
        if n == -100:
            return None, edges.value
        if n < 0:
            self._check(n, "nwb_shortcut_path")
        return out[:n].copy(), edges.value

    def interpolate_path(self, path, k):
        path = np.ascontiguousarray(path, dtype=np.float64).reshape(-1, NJ)
        cap = (len(path) - 1) * (k + 1) + 1
        out = np.empty((cap, NJ))
        n = self.lib.nwb_interpolate_path(_p(path), len(path), int(k), _p(out), cap)
        assert n == cap
        return out

    # ---- multi-GPU fused gather ------------------------------------------------------------------
    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        self._check(self.lib.nwb_dev_alloc(self.h, int(nbytes), C.byref(p)), "nwb_dev_alloc")
        return p.value

    def dev_free(self, ptr):
        self._check(self.lib.nwb_dev_free(self.h, ptr), "nwb_dev_free")

    def ipc_export(self, ptr):
        buf = (C.c_ubyte * 64)()
        self._check(self.lib.nwb_ipc_export(self.h, ptr, buf), "nwb_ipc_export")
        return bytes(buf)

    def ipc_import(self, handle):
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        self._check(self.lib.nwb_ipc_import(self.h, buf, C.byref(p)), "nwb_ipc_import")
        return p.value

    def ipc_close(self, ptr):
        self._check(self.lib.nwb_ipc_close(self.h, ptr), "nwb_ipc_close")

    def set_gather_peers(self, ptrs, my_offset):
        """ptrs: gathered-array base pointer of every rank, in rank order ([] switches the fused gather off)."""
        arr = (C.c_void_p * max(len(ptrs), 1))(*ptrs)
        self._check(self.lib.nwb_set_gather_peers(self.h, arr, len(ptrs), int(my_offset)), "nwb_set_gather_peers")

    # ---- database service + files -----------------------------------------------------------------
    def generate_ds_database(self, max_samples, max_ik_iterations, path, seed):
        """generate_ds_database service (nao_planner_control.cpp:137-147) -> num_configs_generated."""
        req = _abi.GenDsRequest(int(max_samples), int(max_ik_iterations))
        resp = _abi.GenDsResponse()
        self._check(self.lib.nwb_generate_ds_database(self.h, C.byref(req), C.byref(resp), str(path).encode(), int(seed)),
                    "nwb_generate_ds_database")
        return resp.num_configs_generated

    # ---- measurement ----------------------------------------------------------------------------
    def ffma_peak(self, iters=4096):
        t, ms = C.c_double(0), C.c_double(0)
        self._check(self.lib.nwb_ffma_peak(self.h, iters, C.byref(t), C.byref(ms)), "nwb_ffma_peak")
        return t.value, ms.value

    def last_kernel_ms(self):
        ms, n = C.c_double(0), C.c_int64(0)
        self._check(self.lib.nwb_last_kernel_ms(self.h, C.byref(ms), C.byref(n)), "nwb_last_kernel_ms")
        return ms.value, n.value


class Tree:
    """nwb_tree: device-resident RRT tree (struct Tree, rrt_planner.h:38-44) with the nearest-neighbour scan."""

    def __init__(self, ctx, capacity_hint=1024):
        self.ctx = ctx
        self.lib = ctx.lib
        h = C.c_void_p()
        ctx._check(self.lib.nwb_tree_create(ctx.h, int(capacity_hint), C.byref(h)), "nwb_tree_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.nwb_tree_destroy(self.h)
            self.h = None

    def __len__(self):
        return int(self.lib.nwb_tree_size(self.h))

    def clear(self):
        self.ctx._check(self.lib.nwb_tree_clear(self.h), "nwb_tree_clear")

    def add(self, q, predecessor=None):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, NJ)
        p = None if predecessor is None else np.ascontiguousarray(predecessor, dtype=np.int32)
        self.ctx._check(self.lib.nwb_tree_add(self.h, _p(q), _p(p), len(q)), "nwb_tree_add")

    def add_dev(self, q_dev, n, predecessor_dev=None):
        self.ctx._check(self.lib.nwb_tree_add_dev(self.h, q_dev, predecessor_dev, n), "nwb_tree_add_dev")

    def get(self, first=0, n=None):
        n = len(self) - first if n is None else n
        q = np.empty((n, NJ))
        p = np.empty(n, np.int32)
        self.ctx._check(self.lib.nwb_tree_get(self.h, first, n, _p(q), _p(p)), "nwb_tree_get")
        return q, p

    def findNearestNeighbour(self, q):
        """RRT_Planner::findNearestNeighbour for a batch of queries: (idx int32 [nq], dist float64 [nq])."""
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, NJ)
        idx = np.empty(len(q), np.int32)
        dist = np.empty(len(q))
        self.ctx._check(self.lib.nwb_nearest_neighbour_batch(self.h, _p(q), len(q), _p(idx), _p(dist)),
                        "nwb_nearest_neighbour_batch")
        return idx, dist

    def nearest_dev(self, q_dev, nq, idx_dev, dist_dev=None):
        self.ctx._check(self.lib.nwb_nearest_neighbour_batch_dev(self.h, q_dev, nq, idx_dev, dist_dev),
                        "nwb_nearest_neighbour_batch_dev")


def load_ds_database(path, cap=1 << 20):
    """RRT_Planner::load_DS_database (rrt_planner.cpp:121-191): text file -> [n,21] float64."""
    lib = load_library()
    n = C.c_int64(0)
    rc = lib.nwb_load_ds_database(str(path).encode(), None, 0, C.byref(n))
    if rc != 0:
        raise NwbError(f"nwb_load_ds_database({path}) failed ({rc})")
    rows = np.empty((n.value, NJ))
    rc = lib.nwb_load_ds_database(str(path).encode(), _p(rows), n.value, C.byref(n))
    if rc != 0:
        raise NwbError(f"nwb_load_ds_database({path}) failed ({rc})")
    return rows


def write_configs(path, rows):
    rows = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, NJ)
    rc = load_library().nwb_write_configs(str(path).encode(), _p(rows), len(rows))
    if rc != 0:
        raise NwbError(f"nwb_write_configs({path}) failed ({rc})")


class PinnedArray:
    """NumPy view over nwb_host_alloc memory (page-locked)."""

    def __init__(self, shape, dtype):
        self.lib = load_library()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = self.lib.nwb_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise NwbError("nwb_host_alloc failed")
        buf = (C.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.nwb_host_free(self.ptr)
            self.ptr = None
