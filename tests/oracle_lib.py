"""ctypes wrapper of the CPU oracle (oracle/_build/libnwb_oracle.so) - test infrastructure only."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from nao_wholebody_planning_b200._abi import NJ, NFRAMES, NwbParams, default_params

ROOT = Path(__file__).resolve().parents[1]
ORACLE_SO = ROOT / "oracle" / "_build" / "libnwb_oracle.so"
GOLDEN = ROOT / "tests" / "golden"


def build_oracle():
    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle")], check=True)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self):
        if not ORACLE_SO.exists():
            build_oracle()
        self.lib = C.CDLL(str(ORACLE_SO))
        L = self.lib
        L.nwbo_seg_seg_distance.restype = C.c_double
        L.nwbo_seg_box_distance.restype = C.c_double
        L.nwbo_stream_uniform.restype = C.c_double
        L.nwbo_stream_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32]
        L.nwbo_sample_stable_configs.restype = C.c_int64
        L.nwbo_ergonomics_index.restype = C.c_double
        self.threads = L.nwbo_hardware_threads()

    @staticmethod
    def _boxes(boxes):
        if boxes is None or len(boxes) == 0:
            return None, 0
        b = np.ascontiguousarray(boxes, dtype=np.float64).reshape(-1, 15)
        return b, len(b)

    def fk(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, NJ)
        n = len(q)
        poses = np.empty((n, NFRAMES, 3, 4))
        com = np.empty((n, 3))
        self.lib.nwbo_fk_batch(_ptr(q), C.c_int64(n), _ptr(poses), _ptr(com))
        return poses, com

    def is_feasible(self, q, params=None, boxes=None, threads=1):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, NJ)
        n = len(q)
        p = params or default_params()
        b, nb = self._boxes(boxes)
        valid = np.empty(n, np.uint8)
        reason = np.empty(n, np.uint8)
        margins = np.empty((n, 2))
        self.lib.nwbo_is_feasible_batch(C.byref(p), _ptr(b), nb, _ptr(q), C.c_int64(n), _ptr(valid), _ptr(reason),
                                        _ptr(margins), threads)
        return valid, reason, margins

    def is_state_colliding(self, q, params=None, boxes=None, group_only=False):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, NJ)
        n = len(q)
        p = params or default_params()
        b, nb = self._boxes(boxes)
        col = np.empty(n, np.uint8)
        sep = np.empty(n)
        self.lib.nwbo_is_state_colliding_batch(C.byref(p), _ptr(b), nb, _ptr(q), C.c_int64(n), _ptr(col), _ptr(sep),
                                               int(group_only))
        return col, sep

    def support_hull(self, q, params=None):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(NJ)
        p = params or default_params()
        out = np.empty((64, 2))
        n = self.lib.nwbo_support_hull(C.byref(p), _ptr(q), _ptr(out), 64)
        return out[:n].copy()

    def chain_fk(self, which, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        out = np.empty((3, 4))
        self.lib.nwbo_chain_fk(which, _ptr(q), _ptr(out))
        return out

    def right_hand_pose(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(NJ)
        out = np.empty((3, 4))
        self.lib.nwbo_right_hand_pose(_ptr(q), _ptr(out))
        return out

    def chain_jacobian(self, which, q, nj):
        q = np.ascontiguousarray(q, dtype=np.float64)
        out = np.empty((6, nj))
        self.lib.nwbo_chain_jacobian(which, _ptr(q), _ptr(out))
        return out

    def check_joint_limits(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(NJ)
        return bool(self.lib.nwbo_check_joint_limits(_ptr(q)))

    def check_left_leg_pose(self, legs10, params=None):
        legs10 = np.ascontiguousarray(legs10, dtype=np.float64)
        p = params or default_params()
        return bool(self.lib.nwbo_check_left_leg_pose(C.byref(p), _ptr(legs10)))

    def left_leg_ik_local(self, legs10, params=None):
        legs10 = np.ascontiguousarray(legs10, dtype=np.float64)
        p = params or default_params()
        out = np.zeros(5)
        it = C.c_int(0)
        ok = self.lib.nwbo_left_leg_ik_local(C.byref(p), _ptr(legs10), _ptr(out), C.byref(it))
        return bool(ok), out, it.value

    def left_leg_ik_sampler(self, rleg5, max_ik_iter=5, params=None):
        rleg5 = np.ascontiguousarray(rleg5, dtype=np.float64)
        p = params or default_params()
        out = np.zeros(5)
        it = C.c_int(0)
        ok = self.lib.nwbo_left_leg_ik_sampler(C.byref(p), _ptr(rleg5), max_ik_iter, _ptr(out), C.byref(it))
        return bool(ok), out, it.value

    def nearest_neighbour(self, nodes, q, as_shipped=False, threads=1):
        nodes = np.ascontiguousarray(nodes, dtype=np.float64).reshape(-1, NJ)
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, NJ)
        idx = np.empty(len(q), np.int32)
        dist = np.empty(len(q))
        self.lib.nwbo_nearest_neighbour_batch(_ptr(nodes), C.c_int64(len(nodes)), _ptr(q), C.c_int64(len(q)), _ptr(idx),
                                              _ptr(dist), int(as_shipped), threads)
        return idx, dist

    def generate_q_new(self, q_near, q_rand, w, step):
        q_near = np.ascontiguousarray(q_near, dtype=np.float64).reshape(-1, NJ)
        q_rand = np.ascontiguousarray(q_rand, dtype=np.float64).reshape(-1, NJ)
        w = np.ascontiguousarray(w, dtype=np.float64)
        out = np.empty_like(q_near)
        st = np.empty(len(q_near), np.int32)
        self.lib.nwbo_generate_q_new_batch(_ptr(q_near), _ptr(q_rand), C.c_int64(len(q_near)), _ptr(w), C.c_double(step),
                                           _ptr(out), _ptr(st))
        return out, st

    def steer_project(self, q_near, q_target, w, step, params=None, threads=1):
        q_near = np.ascontiguousarray(q_near, dtype=np.float64).reshape(-1, NJ)
        q_target = np.ascontiguousarray(q_target, dtype=np.float64).reshape(-1, NJ)
        w = np.ascontiguousarray(w, dtype=np.float64)
        p = params or default_params()
        out = np.full_like(q_near, np.nan)
        st = np.empty(len(q_near), np.int32)
        ds = np.empty(len(q_near), np.int32)
        self.lib.nwbo_steer_project_batch(C.byref(p), _ptr(q_near), _ptr(q_target), C.c_int64(len(q_near)), _ptr(w),
                                          C.c_double(step), _ptr(out), _ptr(st), _ptr(ds), threads)
        return out, st, ds

    def connect_edge(self, q_start, q_connect, w, step, params=None, boxes=None, cap=4096):
        q_start = np.ascontiguousarray(q_start, dtype=np.float64)
        q_connect = np.ascontiguousarray(q_connect, dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        p = params or default_params()
        b, nb = self._boxes(boxes)
        nodes = np.empty((cap, NJ))
        n = C.c_int(0)
        st = self.lib.nwbo_connect_edge(C.byref(p), _ptr(b), nb, _ptr(q_start), _ptr(q_connect), _ptr(w), C.c_double(step),
                                        _ptr(nodes), cap, C.byref(n))
        return st, nodes[:min(n.value, cap)].copy()

    def interpolate(self, a, b, k):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        out = np.empty((k + 2, NJ))
        self.lib.nwbo_interpolate_waypoints(_ptr(a), _ptr(b), k, _ptr(out))
        return out

    def is_path_valid(self, qa, qb, k=100, params=None, boxes=None, threads=1):
        qa = np.ascontiguousarray(qa, dtype=np.float64).reshape(-1, NJ)
        qb = np.ascontiguousarray(qb, dtype=np.float64).reshape(-1, NJ)
        p = params or default_params()
        b, nb = self._boxes(boxes)
        ok = np.empty(len(qa), np.uint8)
        fb = np.empty(len(qa), np.int32)
        self.lib.nwbo_is_path_valid_batch(C.byref(p), _ptr(b), nb, _ptr(qa), _ptr(qb), C.c_int64(len(qa)), k, _ptr(ok),
                                          _ptr(fb), threads)
        return ok, fb

    def shortcut_path(self, raw, params=None, boxes=None):
        raw = np.ascontiguousarray(raw, dtype=np.float64).reshape(-1, NJ)
        p = params or default_params()
        b, nb = self._boxes(boxes)
        out = np.empty((len(raw) + 8, NJ))
        edges = C.c_int(0)
        n = self.lib.nwbo_shortcut_path(C.byref(p), _ptr(b), nb, _ptr(raw), len(raw), _ptr(out), len(out), C.byref(edges))
        return (None if n < 0 else out[:n].copy()), edges.value

    def interpolate_short_path(self, path, k):
        path = np.ascontiguousarray(path, dtype=np.float64).reshape(-1, NJ)
        cap = (len(path) - 1) * (k + 1) + 1
        out = np.empty((cap, NJ))
        n = self.lib.nwbo_interpolate_short_path(_ptr(path), len(path), k, _ptr(out), cap)
        assert n == cap
        return out

    def sample_stable_configs(self, seed, first, count, max_ik_iter=5, params=None, boxes=None, threads=1):
        p = params or default_params()
        b, nb = self._boxes(boxes)
        rows = np.empty((count, NJ))
        idx = np.empty(count, np.uint64)
        stage = np.empty(count, np.uint8)
        n = self.lib.nwbo_sample_stable_configs(C.byref(p), _ptr(b), nb, C.c_uint64(seed), C.c_uint64(first),
                                                C.c_uint64(count), max_ik_iter, _ptr(rows), C.c_int64(count), _ptr(idx),
                                                _ptr(stage), threads)
        return rows[:n].copy(), idx[:n].copy(), stage

    def random_db_index(self, seed, iteration, n_db, params=None):
        p = params or default_params()
        return self.lib.nwbo_random_db_index(C.byref(p), C.c_uint64(seed), C.c_uint64(iteration), n_db)

    def solve_query(self, start, goal, db, w, seed, max_iter=8000, step=0.1, params=None, boxes=None, cap=8192):
        start = np.ascontiguousarray(start, dtype=np.float64)
        goal = np.ascontiguousarray(goal, dtype=np.float64)
        db = np.ascontiguousarray(db, dtype=np.float64).reshape(-1, NJ)
        w = np.ascontiguousarray(w, dtype=np.float64)
        p = params or default_params()
        b, nb = self._boxes(boxes)
        raw = np.empty((cap, NJ))
        stats = np.zeros(5, np.int32)
        n = self.lib.nwbo_solve_query(C.byref(p), _ptr(b), nb, _ptr(start), _ptr(goal), _ptr(db), len(db), _ptr(w),
                                      C.c_uint64(seed), max_iter, C.c_double(step), _ptr(raw), cap, _ptr(stats))
        return raw[:min(n, cap)].copy(), dict(success=bool(stats[0]), iterations=int(stats[1]), num_nodes=int(stats[2]),
                                               nodes_start=int(stats[3]), nodes_goal=int(stats[4]))

    # ---- right arm, articulated-object constraint, goal sampling, services ----
    def arm5d_fk(self, q5):
        q5 = np.ascontiguousarray(q5, dtype=np.float64)
        pos, d = np.empty(3), np.empty(3)
        self.lib.nwbo_arm5d_fk(_ptr(q5), _ptr(pos), _ptr(d))
        return pos, d

    def arm5d_ik(self, pos, d):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        d = np.ascontiguousarray(d, dtype=np.float64)
        out = np.empty((8, 5))
        n = self.lib.nwbo_arm5d_ik(_ptr(pos), _ptr(d), _ptr(out))
        return out[:n].copy()

    def hand_trajectory(self, kind, start6, goal6, num_interp=20, door4=None):
        start6 = np.ascontiguousarray(start6, dtype=np.float64)
        goal6 = np.ascontiguousarray(goal6, dtype=np.float64)
        door4 = None if door4 is None else np.ascontiguousarray(door4, dtype=np.float64)
        out = np.empty((num_interp + 2, 6))
        self.lib.nwbo_hand_trajectory(kind, _ptr(start6), _ptr(goal6), num_interp, _ptr(door4), _ptr(out))
        return out

    def ergonomics_index(self, arm5):
        arm5 = np.ascontiguousarray(arm5, dtype=np.float64)
        return self.lib.nwbo_ergonomics_index(_ptr(arm5))

    def right_hand_fk(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(NJ)
        pos, z = np.empty(3), np.empty(3)
        self.lib.nwbo_right_hand_fk(_ptr(q), _ptr(pos), _ptr(z))
        return pos, z

    def goal_right_arm_ik(self, q, pose6, seed=0, first_index=0, params=None):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, NJ)
        pose6 = np.ascontiguousarray(pose6, dtype=np.float64)
        p = params or default_params()
        n = len(q)
        ok, arm, erg = np.empty(n, np.uint8), np.empty((n, 5)), np.empty(n)
        self.lib.nwbo_goal_right_arm_ik_batch(C.byref(p), _ptr(q), C.c_int64(n), _ptr(pose6), C.c_uint64(seed),
                                              C.c_uint64(first_index), _ptr(ok), _ptr(arm), _ptr(erg))
        return ok, arm, erg

    def enforce_ma(self, pts, start_tree, near_hand_index, q_near, q_ds):
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 6)
        q_near = np.ascontiguousarray(q_near, dtype=np.float64).reshape(-1, NJ)
        q_ds = np.ascontiguousarray(q_ds, dtype=np.float64).reshape(-1, NJ)
        hi = np.ascontiguousarray(near_hand_index, dtype=np.int32)
        out = np.empty_like(q_ds)
        st = np.empty(len(q_ds), np.int32)
        self.lib.nwbo_enforce_ma_batch(_ptr(pts), len(pts), int(start_tree), _ptr(hi), _ptr(q_near), _ptr(q_ds),
                                       C.c_int64(len(q_ds)), _ptr(out), _ptr(st))
        return out, st

    def goal_sample_batch(self, seed, first, count, pose6, max_ik_iter=5, params=None, boxes=None, threads=1):
        p = params or default_params()
        b, nb = self._boxes(boxes)
        pose6 = np.ascontiguousarray(pose6, dtype=np.float64)
        stage, erg, rows = np.empty(count, np.uint8), np.empty(count), np.empty((count, NJ))
        self.lib.nwbo_goal_sample_batch(C.byref(p), _ptr(b), nb, C.c_uint64(seed), C.c_uint64(first), C.c_uint64(count),
                                        _ptr(pose6), max_ik_iter, _ptr(stage), _ptr(erg), _ptr(rows), threads)
        return stage, erg, rows

    def sample_goal_configs(self, seed, pose6, max_samples=5000, max_ik_iter=5, params=None, boxes=None):
        p = params or default_params()
        b, nb = self._boxes(boxes)
        pose6 = np.ascontiguousarray(pose6, dtype=np.float64)
        rows, erg, idx = np.zeros((6, NJ)), np.zeros(6), np.zeros(6, np.uint64)
        found, drawn = C.c_int32(0), C.c_uint64(0)
        self.lib.nwbo_sample_goal_configs(C.byref(p), _ptr(b), nb, C.c_uint64(seed), _ptr(pose6), max_samples, max_ik_iter,
                                          _ptr(rows), _ptr(erg), _ptr(idx), C.byref(found), C.byref(drawn))
        n = found.value
        return rows[:n].copy(), erg[:n].copy(), idx[:n].copy(), drawn.value

    def solve_query_ma(self, start, goal, db, w, seed, pts, max_iter=8000, step=0.1, params=None, boxes=None, cap=8192):
        start = np.ascontiguousarray(start, dtype=np.float64)
        goal = np.ascontiguousarray(goal, dtype=np.float64)
        db = np.ascontiguousarray(db, dtype=np.float64).reshape(-1, NJ)
        w = np.ascontiguousarray(w, dtype=np.float64)
        pts = None if pts is None else np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 6)
        p = params or default_params()
        b, nb = self._boxes(boxes)
        raw = np.empty((cap, NJ))
        stats = np.zeros(5, np.int32)
        n = self.lib.nwbo_solve_query_ma(C.byref(p), _ptr(b), nb, _ptr(start), _ptr(goal), _ptr(db), len(db), _ptr(w),
                                         C.c_uint64(seed), max_iter, C.c_double(step), _ptr(pts), 0 if pts is None else len(pts),
                                         _ptr(raw), cap, _ptr(stats))
        return raw[:min(n, cap)].copy(), dict(success=bool(stats[0]), iterations=int(stats[1]), num_nodes=int(stats[2]),
                                               nodes_start=int(stats[3]), nodes_goal=int(stats[4]))

    @staticmethod
    def _service(service):
        return None if service is None else np.ascontiguousarray(service, dtype=np.float64)

    def compute_motion_plan(self, db, pose6, start, seed, execute=False, service=None, params=None, boxes=None, cap=16384):
        p = params or default_params()
        b, nb = self._boxes(boxes)
        db = np.ascontiguousarray(db, dtype=np.float64).reshape(-1, NJ)
        pose6 = np.ascontiguousarray(pose6, dtype=np.float64)
        start = np.ascontiguousarray(start, dtype=np.float64)
        sv = self._service(service)
        goal, traj, tfs, stats = np.zeros(NJ), np.empty((cap, NJ)), np.empty(cap), np.zeros(4, np.int32)
        self.lib.nwbo_compute_motion_plan(C.byref(p), _ptr(b), nb, _ptr(sv), _ptr(db), len(db), _ptr(pose6), _ptr(start),
                                          int(execute), C.c_uint64(seed), _ptr(goal), _ptr(traj), _ptr(tfs), cap, _ptr(stats))
        n = min(int(stats[3]), cap)
        return dict(success=bool(stats[0]), num_nodes=int(stats[1]), goals_found=int(stats[2]), goal_config=goal,
                    trajectory=traj[:n].copy(), time_from_start=tfs[:n].copy())

    def compute_manipulation_plan(self, db, pose6, start, seed, circular, obj, execute=False, service=None, params=None,
                                  boxes=None, cap=16384):
        p = params or default_params()
        b, nb = self._boxes(boxes)
        db = np.ascontiguousarray(db, dtype=np.float64).reshape(-1, NJ)
        pose6 = np.ascontiguousarray(pose6, dtype=np.float64)
        start = np.ascontiguousarray(start, dtype=np.float64)
        obj = np.ascontiguousarray(obj, dtype=np.float64)
        sv = self._service(service)
        ga, gm = np.zeros(NJ), np.zeros(NJ)
        ta, tm, fa, fm = np.empty((cap, NJ)), np.empty((cap, NJ)), np.empty(cap), np.empty(cap)
        stats = np.zeros(8, np.int32)
        self.lib.nwbo_compute_manipulation_plan(C.byref(p), _ptr(b), nb, _ptr(sv), _ptr(db), len(db), _ptr(pose6), _ptr(start),
                                                int(circular), _ptr(obj), int(execute), C.c_uint64(seed), _ptr(ga), _ptr(gm),
                                                _ptr(ta), _ptr(fa), _ptr(tm), _ptr(fm), cap, _ptr(stats))
        na, nm = min(int(stats[6]), cap), min(int(stats[7]), cap)
        return dict(success_approach=bool(stats[0]), success_manipulation=bool(stats[1]), nodes_approach=int(stats[2]),
                    nodes_manipulate=int(stats[3]), goals_approach=int(stats[4]), goals_manipulate=int(stats[5]),
                    goal_approach=ga, goal_manipulate=gm, traj_approach=ta[:na].copy(), tfs_approach=fa[:na].copy(),
                    traj_manipulation=tm[:nm].copy(), tfs_manipulation=fm[:nm].copy())

    def time_parameterize(self, positions, params=None):
        positions = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, NJ)
        p = params or default_params()
        out = np.empty(len(positions))
        self.lib.nwbo_time_parameterize(C.byref(p), _ptr(positions), len(positions), _ptr(out))
        return out

    def philox(self, ctr, key):
        ctr = np.asarray(ctr, dtype=np.uint32)
        key = np.asarray(key, dtype=np.uint32)
        out = np.empty(4, np.uint32)
        self.lib.nwbo_philox4x32_10(_ptr(ctr), _ptr(key), _ptr(out))
        return out

    def seg_seg(self, p1, q1, p2, q2):
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (p1, q1, p2, q2)]
        return self.lib.nwbo_seg_seg_distance(*[_ptr(x) for x in a])

    def seg_box(self, a, b, box15):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        box15 = np.ascontiguousarray(box15, dtype=np.float64)
        return self.lib.nwbo_seg_box_distance(_ptr(a), _ptr(b), _ptr(box15))


def load_dat(path, drop_ergonomics=True):
    rows = []
    for line in Path(path).read_text().splitlines():
        v = [float(x) for x in line.split()]
        if v:
            rows.append(v)
    a = np.array(rows, dtype=np.float64)
    if drop_ergonomics and a.shape[1] == NJ + 1:   # goal files: ergonomics index in column 0 (SCG:409)
        a = a[:, 1:]
    return a


def golden_db():
    return load_dat(GOLDEN / "config_database" / "ssc_double_support.dat")


def golden_valid_rows():
    """Every configuration the reference shipped as valid."""
    cd = GOLDEN / "config_database"
    names = ["ssc_double_support.dat"] + sorted(p.name for p in cd.glob("ssc_double_support_goal_*.dat"))
    rows = [load_dat(cd / n) for n in names]
    rows.append(load_dat(GOLDEN / "solutions" / "solution_path_first.dat"))
    rows.append(load_dat(GOLDEN / "solutions" / "solution_path_second.dat"))
    return np.concatenate(rows, axis=0)


def joint_limits():
    import json
    m = json.loads((ROOT / "nao_wholebody_planning_b200" / "models" / "nao_h25_rfoot.json").read_text())
    lim = np.array(m["joint_limits"], dtype=np.float64)
    return lim[:, 0].copy(), lim[:, 1].copy()


def uniform_configs(n, seed=20121001, lhyp_zero=False):
    """V1 / V1b synthetic inputs of SURVEY.md section 8d."""
    lo, hi = joint_limits()
    rng = np.random.default_rng(seed)
    q = lo + (hi - lo) * rng.random((n, NJ))
    if lhyp_zero:
        q[:, 15] = 0.0
    return q


def noisy_db_configs(n, seed=20121004, sigma=0.05):
    """V1d: golden DB rows tiled with N(0, sigma^2) joint noise - dense near the validity boundary."""
    db = golden_db()
    rng = np.random.default_rng(seed)
    q = db[rng.integers(0, len(db), n)] + rng.normal(0, sigma, (n, NJ))
    q[:, 15] = 0.0
    return q


# Appendix D scenes: list of boxes [cx,cy,cz, sx,sy,sz (full extents), R row-major]
def scene_boxes(name):
    eye = [1, 0, 0, 0, 1, 0, 0, 0, 1]
    if name == "S0":
        return np.zeros((0, 15))
    if name == "S1":   # shelf (initScene.cpp:120-158; SIM:322-356), shelf_x 0.33, shelf_y -0.06, shelf_z 0.02
        sx, sy, sz = 0.33, -0.06, 0.02
        b = [[sx, sy, sz, 0.27, 0.31, 0.04],
             [sx + 0.0175, sy + 0.132, sz + 0.28, 0.235, 0.014, 0.52],
             [sx + 0.0175, sy - 0.132, sz + 0.28, 0.235, 0.014, 0.52],
             [sx, sy, sz + 0.5485, 0.27, 0.31, 0.017]]
        for dz in (0.147, 0.2775, 0.408):
            b.append([sx + 0.025, sy, sz + dz, 0.22, 0.25, 0.014])
        return np.array([r + eye for r in b], dtype=np.float64)
    if name == "S2":   # table (EXE:376-377)
        return np.array([[0.35, 0.06, 0.32, 0.3, 0.6, 0.03] + eye], dtype=np.float64)
    if name == "S3":   # drawer + beam (PAR:465-472)
        return np.array([[0.15, -0.08, 0.125, 0.06, 0.08, 0.25] + eye, [0.39, -0.15, 0.25, 0.16, 0.17, 0.5] + eye])
    if name == "S4":   # pole (initScene.cpp:175-176)
        return np.array([[0.15, -0.17, 0.02, 0.02, 1.0, 0.02] + eye], dtype=np.float64)
    if name == "S5":   # bottle table (PAR:484-485)
        return np.array([[0.34, 0.18, 0.21, 0.3, 0.4, 0.1] + eye], dtype=np.float64)
    if name == "ALL":  # every object the reference's scene builders ever insert, at once (21 boxes > the 16 inline slots):
        shelf30 = scene_boxes("S1").copy()                 # the shelf at both shelf_x values (0.33 in the examples, 0.30 in initScene.cpp:120)
        shelf30[:, 0] -= 0.03
        tables = np.array([[0.32, 0.05, 0.333, 0.3, 0.3, 0.03] + eye, [0.35, 0.06, 0.35, 0.3, 0.6, 0.03] + eye])   # initScene.cpp:572-573, nao_rrt_planner.cpp:158-159
        return np.concatenate([scene_boxes("S1"), shelf30, scene_boxes("S2"), tables, scene_boxes("S3"), scene_boxes("S4"), scene_boxes("S5")])
    raise KeyError(name)


def generated_scene(n, seed=20121006, keep_out=0.22):
    """n oriented boxes scattered through the robot's workspace (r_sole frame): centres uniform in
    [-0.6, 0.8] x [-0.7, 0.7] x [0, 1.0] m outside a vertical keep-out cylinder of radius `keep_out` around the stance
    (so that part of the configurations stays collision-free), full extents 2-12 cm, every second box rotated."""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        c = np.array([rng.uniform(-0.6, 0.8), rng.uniform(-0.7, 0.7), rng.uniform(0.0, 1.0)])
        size = rng.uniform(0.02, 0.12, 3)
        if np.hypot(c[0] - 0.02, c[1] - 0.05) < keep_out + 0.5 * np.linalg.norm(size):
            continue
        if len(out) % 2:
            ax = rng.normal(size=3)
            ax /= np.linalg.norm(ax)
            ang = rng.uniform(-np.pi, np.pi)
            K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
            R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
        else:
            R = np.eye(3)
        out.append(np.concatenate([c, size, R.reshape(9)]))
    return np.array(out)
