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


_DTYPE_CACHE = {}


def _struct_dtype(ct):
    """numpy structured dtype with the layout of a ctypes Structure (pointers as uint64, nested structs recursively)."""
    if ct in _DTYPE_CACHE:
        return _DTYPE_CACHE[ct]
    names, formats, offsets = [], [], []
    for name, ftype in ct._fields_:
        names.append(name)
        offsets.append(getattr(ct, name).offset)
        if isinstance(ftype, type) and issubclass(ftype, C.Structure):
            formats.append(_struct_dtype(ftype))
        elif isinstance(ftype, type) and issubclass(ftype, C._Pointer):
            formats.append(np.dtype(np.uint64))
        elif isinstance(ftype, type) and issubclass(ftype, C.Array):
            formats.append(np.dtype((np.dtype(ftype._type_), (ftype._length_,))))
        else:
            formats.append(np.dtype(ftype))
    dt = np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": C.sizeof(ct)})
    _DTYPE_CACHE[ct] = dt
    return dt



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
        b = np.ascontiguousarray(boxes15, dtype=np.float64).reshape(-1, 15)
        self._check(self.lib.nwb_scene_add_boxes(self.h, _p(b), len(b)), "nwb_scene_add_boxes")

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

    def is_feasible_host_buffers_f32(self, q_ptr, n, valid_ptr, reason_ptr=None, margins_ptr=None):
        """nwb_is_feasible_batch_f32 on raw host pointers (float32 configurations)."""
        self._check(self.lib.nwb_is_feasible_batch_f32(self.h, q_ptr, int(n), valid_ptr, reason_ptr, margins_ptr),
                    "nwb_is_feasible_batch_f32")

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

    def connect(self, q_start, q_connect, weights=None, step=0.1, max_steps=128, out=None):
        """connectTree edge walks (rrt_planner.cpp:926-1031) -> (nodes [n,max_steps,21], n_added, status).
        out = (nodes, n_added, status) reuses the caller's arrays (a fresh 22 MB array costs ~3 ms of page faults per
        call, twice the walk itself); rows past n_added are not written."""
        q_start = np.ascontiguousarray(q_start, dtype=np.float64).reshape(-1, NJ)
        q_connect = np.ascontiguousarray(q_connect, dtype=np.float64).reshape(-1, NJ)
        n = len(q_start)
        w = self._weights(weights)
        if out is not None:
            nodes, added, st = out
            assert nodes.shape == (n, max_steps, NJ) and nodes.dtype == np.float64 and nodes.flags.c_contiguous
            assert added.shape == (n,) and added.dtype == np.int32 and st.shape == (n,) and st.dtype == np.int32
        else:
            nodes = np.zeros((n, max_steps, NJ))      # rows past n_added stay zero (only the appended rows are copied back)
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
                   raw_path_capacity=1024, hand_trajectories=None):
        """setStartGoalConfigs + solveQuery + raw path extraction for a batch of queries.
        hand_trajectories [nq][n_points][6]: activate_drawer_constraint / activate_door_constraint per query.
        Returns (list of result dicts, list of raw paths [len,21])."""
        starts = np.ascontiguousarray(starts, dtype=np.float64).reshape(-1, NJ)
        goals = np.ascontiguousarray(goals, dtype=np.float64).reshape(-1, NJ)
        db = np.ascontiguousarray(db, dtype=np.float64).reshape(-1, NJ)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        nq = len(starts)
        w = self._weights(weights)
        res = (_abi.NwbPlanResult * nq)()
        raw = np.full((nq, raw_path_capacity, NJ), np.nan)
        if hand_trajectories is None:
            self._check(self.lib.nwb_solve_query_batch(self.h, _p(starts), _p(goals), nq, _p(db), len(db), _p(w), _p(seeds),
                                                       int(max_iter), float(step), int(tree_capacity), C.byref(res), _p(raw),
                                                       int(raw_path_capacity)), "nwb_solve_query_batch")
        else:
            ht = np.ascontiguousarray(hand_trajectories, dtype=np.float64)
            ht = ht.reshape(nq, -1, 6)
            self._check(self.lib.nwb_solve_query_constrained_batch(
                self.h, _p(starts), _p(goals), nq, _p(db), len(db), _p(w), _p(seeds), int(max_iter), float(step),
                int(tree_capacity), _p(ht), ht.shape[1], C.byref(res), _p(raw), int(raw_path_capacity)),
                "nwb_solve_query_constrained_batch")
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

    # ---- right arm / articulated objects / goal configurations ----------------------------------------
    def arm_ik(self, targets):
        """IKSolver::ik (ikfast_TranslationDirection5D.cpp:341) for [n][6] torso-frame targets ->
        (solutions [n][8][5] NaN padded, counts [n])."""
        t = np.ascontiguousarray(targets, dtype=np.float64).reshape(-1, 6)
        sols = np.empty((len(t), 8, 5))
        cnt = np.empty(len(t), np.int32)
        self._check(self.lib.nwb_arm_ik_batch(self.h, _p(t), len(t), _p(sols), _p(cnt)), "nwb_arm_ik_batch")
        return sols, cnt

    def computeRightArmIK(self, q, hand_pose6, seed=0, first_index=0):
        """ConstraintKinematics::computeRightArmIK (nao_constraint_kinematics.cpp:224) for whole-body rows."""
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, NJ)
        pose = np.ascontiguousarray(hand_pose6, dtype=np.float64)
        n = len(q)
        ok, arm, erg = np.empty(n, np.uint8), np.empty((n, 5)), np.empty(n)
        self._check(self.lib.nwb_goal_arm_ik_batch(self.h, _p(q), n, _p(pose), int(seed), int(first_index), _p(ok), _p(arm),
                                                   _p(erg)), "nwb_goal_arm_ik_batch")
        return ok, arm, erg

    def hand_trajectory(self, kind, start6, goal6, num_interp_points=20, door4=None):
        a = np.ascontiguousarray(start6, dtype=np.float64)
        b = np.ascontiguousarray(goal6, dtype=np.float64)
        d = None if door4 is None else np.ascontiguousarray(door4, dtype=np.float64)
        out = np.empty((num_interp_points + 2, 6))
        rc = self.lib.nwb_hand_trajectory(int(kind), _p(a), _p(b), int(num_interp_points), _p(d), _p(out))
        if rc != 0:
            raise NwbError(f"nwb_hand_trajectory failed ({rc})")
        return out

    def enforce_MA_Constraints(self, pts, start_tree, near_hand_index, q_near, q_new_ds):
        """RRT_Planner::enforce_MA_Constraints (rrt_planner.cpp:597) -> (q_out, status); TRAPPED rows are NaN."""
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 6)
        q_near = np.ascontiguousarray(q_near, dtype=np.float64).reshape(-1, NJ)
        q_ds = np.ascontiguousarray(q_new_ds, dtype=np.float64).reshape(-1, NJ)
        hi = np.ascontiguousarray(near_hand_index, dtype=np.int32)
        out = np.full_like(q_ds, np.nan)
        st = np.empty(len(q_ds), np.int32)
        self._check(self.lib.nwb_enforce_ma_batch(self.h, _p(pts), len(pts), int(bool(start_tree)), _p(hi), _p(q_near), _p(q_ds),
                                                  len(q_ds), _p(out), _p(st)), "nwb_enforce_ma_batch")
        return out, st

    def goal_sample_batch(self, seed, first, count, hand_pose6, max_ik_iter=5):
        pose = np.ascontiguousarray(hand_pose6, dtype=np.float64)
        stage, erg, rows = np.empty(count, np.uint8), np.empty(count), np.empty((count, NJ))
        self._check(self.lib.nwb_goal_sample_batch(self.h, int(seed), int(first), int(count), _p(pose), int(max_ik_iter), _p(stage),
                                                   _p(erg), _p(rows)), "nwb_goal_sample_batch")
        return stage, erg, rows

    def sample_goal_configs(self, seed, hand_pose6, max_samples=5000, max_ik_iter=5):
        """sample_stable_configs(goal branch) + sort_goal_configs_by_ergonomy -> (rows, ergonomics, index, drawn)."""
        pose = np.ascontiguousarray(hand_pose6, dtype=np.float64)
        rows, erg, idx = np.zeros((6, NJ)), np.zeros(6), np.zeros(6, np.uint64)
        found, drawn = C.c_int32(0), C.c_int64(0)
        self._check(self.lib.nwb_sample_goal_configs(self.h, int(seed), _p(pose), int(max_samples), int(max_ik_iter), _p(rows),
                                                     _p(erg), _p(idx), C.byref(found), C.byref(drawn)), "nwb_sample_goal_configs")
        n = found.value
        return rows[:n].copy(), erg[:n].copy(), idx[:n].copy(), drawn.value

    # ---- planning services (nao_planner_control.cpp:150-870) ---------------------------------------------
    def set_ds_database(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, NJ)
        self._check(self.lib.nwb_ctx_set_ds_database(self.h, _p(rows), len(rows)), "nwb_ctx_set_ds_database")

    def load_ds_database(self, path):
        self._check(self.lib.nwb_ctx_load_ds_database(self.h, str(path).encode()), "nwb_ctx_load_ds_database")
        return self.lib.nwb_ctx_ds_database_size(self.h)

    @staticmethod
    def _traj(cap):
        pos, tfs = np.empty((cap, NJ)), np.empty(cap)       # rows beyond n_points are never read
        t = _abi.Trajectory(cap, 0, pos.ctypes.data_as(C.POINTER(C.c_double)), tfs.ctypes.data_as(C.POINTER(C.c_double)))
        return t, pos, tfs

    def time_parameterize(self, positions):
        """generateTrajectory's time parameterisation (rrt_planner.cpp:1596-1618) -> time_from_start [n]."""
        positions = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, NJ)
        out = np.empty(len(positions))
        self._check(self.lib.nwb_time_parameterize(self.h, _p(positions), len(positions), _p(out)), "nwb_time_parameterize")
        return out

    def compute_goal_config(self, position, direction, seed):
        req = _abi.GoalConfigRequest(_abi.Point3(*position), _abi.Point3(*direction))
        resp = _abi.GoalConfigResponse()
        self._check(self.lib.nwb_compute_goal_config(self.h, C.byref(req), C.byref(resp), int(seed)), "nwb_compute_goal_config")
        return np.array(resp.wb_goal_config), resp.num_goal_configs_found

    def compute_motion_plan(self, position, direction, wb_start_config, seed, execute_trajectory=False, cap=4096):
        start = np.ascontiguousarray(wb_start_config, dtype=np.float64)
        req = _abi.MotionPlanRequest(_abi.Point3(*position), _abi.Point3(*direction), start.ctypes.data_as(C.POINTER(C.c_double)),
                                     len(start), int(execute_trajectory))
        resp = _abi.MotionPlanResponse()
        resp.motion_plan, pos, tfs = self._traj(cap)
        self._check(self.lib.nwb_compute_motion_plan(self.h, C.byref(req), C.byref(resp), int(seed)), "nwb_compute_motion_plan")
        n = min(resp.motion_plan.n_points, cap)
        return dict(success=bool(resp.success), wb_goal_config=np.array(resp.wb_goal_config), time_goal_config=resp.time_goal_config,
                    num_nodes_generated=resp.num_nodes_generated, search_time=resp.search_time, trajectory=pos[:n].copy(),
                    time_from_start=tfs[:n].copy())

    def compute_motion_plan_batch(self, positions, directions, wb_start_configs, seeds, execute_trajectory=False, cap=2048,
                                  reuse_buffers=False):
        """n compute_motion_plan requests in one call -> list of result dicts (as compute_motion_plan).
        The request / response arrays of the C ABI are filled and read through numpy views of the ctypes arrays (one
        vector operation per field instead of a Python loop over 1681 structs).  reuse_buffers: the trajectory output
        arrays are kept on the context and handed to the next call of the same shape, as a C++ caller would keep its
        buffers - the library then writes into resident pages instead of first-touching ~7 pages per request; the
        returned trajectories are views that the NEXT such call overwrites."""
        positions = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
        directions = np.ascontiguousarray(directions, dtype=np.float64).reshape(-1, 3)
        starts = np.ascontiguousarray(wb_start_configs, dtype=np.float64).reshape(-1, NJ)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        n = len(positions)
        reqs = (_abi.MotionPlanRequest * n)()
        resps = (_abi.MotionPlanResponse * n)()
        cached = getattr(self, "_mp_buffers", None)
        if reuse_buffers and cached is not None and cached[0].shape == (n, cap, NJ):
            pos, tfs = cached
        else:
            pos = np.empty((n, cap, NJ))                 # rows beyond n_points are never read
            tfs = np.empty((n, cap))
            if reuse_buffers:
                self._mp_buffers = (pos, tfs)
        if n:
            rq = np.frombuffer(reqs, dtype=_struct_dtype(_abi.MotionPlanRequest))
            rs = np.frombuffer(resps, dtype=_struct_dtype(_abi.MotionPlanResponse))
            k = np.arange(n, dtype=np.uint64)
            for c, name in enumerate("xyz"):
                rq["right_hand_position"][name] = positions[:, c]
                rq["right_hand_direction"][name] = directions[:, c]
            rq["wb_start_config"] = starts.ctypes.data + k * np.uint64(NJ * 8)
            rq["wb_start_config_len"] = NJ
            rq["execute_trajectory"] = int(execute_trajectory)
            mp = rs["motion_plan"]
            mp["capacity"] = cap
            mp["n_points"] = 0
            mp["positions"] = pos.ctypes.data + k * np.uint64(cap * NJ * 8)
            mp["time_from_start"] = tfs.ctypes.data + k * np.uint64(cap * 8)
        self._check(self.lib.nwb_compute_motion_plan_batch(self.h, reqs, resps, n, _p(seeds)), "nwb_compute_motion_plan_batch")
        out = []
        if n:
            m_all = np.minimum(rs["motion_plan"]["n_points"], cap)
            ok, goal, tg = rs["success"].astype(bool), rs["wb_goal_config"].copy(), rs["time_goal_config"].copy()
            nodes, st = rs["num_nodes_generated"].copy(), rs["search_time"].copy()
            for i in range(n):
                m = int(m_all[i])
                out.append(dict(success=bool(ok[i]), wb_goal_config=goal[i], time_goal_config=float(tg[i]),
                                num_nodes_generated=int(nodes[i]), search_time=float(st[i]), trajectory=pos[i, :m],
                                time_from_start=tfs[i, :m]))       # views into the call's two output arrays
        return out

    def _manipulation_result(self, resp, pa, ta, pm, tm, cap):
        na, nm = min(resp.motion_plan_approach.n_points, cap), min(resp.motion_plan_manipulation.n_points, cap)
        return dict(success_approach=bool(resp.success_approach), success_manipulation=bool(resp.success_manipulation),
                    wb_goal_config_approach=np.array(resp.wb_goal_config_approach),
                    wb_goal_config_manipulate=np.array(resp.wb_goal_config_manipulate),
                    time_goal_config_approach=resp.time_goal_config_approach,
                    time_goal_config_manipulate=resp.time_goal_config_manipulate,
                    num_nodes_generated_approach=resp.num_nodes_generated_approach,
                    num_nodes_generated_manipulate=resp.num_nodes_generated_manipulate,
                    search_time_approach=resp.search_time_approach, search_time_manipulate=resp.search_time_manipulate,
                    traj_approach=pa[:na].copy(), tfs_approach=ta[:na].copy(), traj_manipulation=pm[:nm].copy(),
                    tfs_manipulation=tm[:nm].copy())

    def compute_linear_manipulation_plan(self, position, direction, wb_start_config, relative_angle, object_translation_length,
                                         seed, execute_trajectory=False, cap=4096):
        start = np.ascontiguousarray(wb_start_config, dtype=np.float64)
        req = _abi.LinearManipulationRequest(_abi.Point3(*position), _abi.Point3(*direction),
                                             start.ctypes.data_as(C.POINTER(C.c_double)), len(start), float(relative_angle),
                                             float(object_translation_length), int(execute_trajectory))
        resp = _abi.ManipulationPlanResponse()
        resp.motion_plan_approach, pa, ta = self._traj(cap)
        resp.motion_plan_manipulation, pm, tm = self._traj(cap)
        self._check(self.lib.nwb_compute_linear_manipulation_plan(self.h, C.byref(req), C.byref(resp), int(seed)),
                    "nwb_compute_linear_manipulation_plan")
        return self._manipulation_result(resp, pa, ta, pm, tm, cap)

    def compute_circular_manipulation_plan(self, position, direction, wb_start_config, relative_angle, rotation_radius,
                                           rotation_angle, clearance_handle, seed, execute_trajectory=False, cap=4096):
        start = np.ascontiguousarray(wb_start_config, dtype=np.float64)
        req = _abi.CircularManipulationRequest(_abi.Point3(*position), _abi.Point3(*direction),
                                               start.ctypes.data_as(C.POINTER(C.c_double)), len(start), float(relative_angle),
                                               float(rotation_radius), float(rotation_angle), float(clearance_handle),
                                               int(execute_trajectory))
        resp = _abi.ManipulationPlanResponse()
        resp.motion_plan_approach, pa, ta = self._traj(cap)
        resp.motion_plan_manipulation, pm, tm = self._traj(cap)
        self._check(self.lib.nwb_compute_circular_manipulation_plan(self.h, C.byref(req), C.byref(resp), int(seed)),
                    "nwb_compute_circular_manipulation_plan")
        return self._manipulation_result(resp, pa, ta, pm, tm, cap)

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

    def set_kernel_timing(self, enabled):
        """Timing events around every operator's kernels (default on); off for pipelined back-to-back calls."""
        self._check(self.lib.nwb_set_kernel_timing(self.h, int(bool(enabled))), "nwb_set_kernel_timing")

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

    def scan_owner(self, nq, node):
        """Diagnostic: id of the scan thread that evaluates `node` for a call with nq queries."""
        return int(self.lib.nwb_nn_scan_owner(self.h, nq, node))

    def debug_tensor_tile(self, q):
        """Diagnostic: raw tcgen05 accumulator [128 queries][128 nodes] of the first node tile."""
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, NJ)
        out = np.zeros((128, 128), np.float32)
        self.ctx._check(self.lib.nwb_nn_debug_tensor_tile(self.h, _p(q), len(q), _p(out)), "nwb_nn_debug_tensor_tile")
        return out

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
