"""Pins the oracle's right-arm / articulated-object / goal-sampling restatement (oracle/arm_restated.h,
oracle/planner_restated.h) on what the reference itself provides:

  * tests/golden/ikfast_vectors.npz - outputs of the reference's generated solver, produced by running it
    (tests/golden/make_ikfast_golden.py);
  * oracle/_ref/libikfast_ref.so    - the same solver, live, when it has been built (`make -C oracle ref`);
  * the shipped goal-configuration files: arm columns = OUR selected solution.
"""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as OL

GOLD = OL.GOLDEN / "config_database"
ARM_LO = np.array([-2.0857, -1.3265, -2.0857, 0.0349, -1.8238])
ARM_HI = np.array([2.0857, 0.3142, 2.0857, 1.5446, 1.8238])


def ang_diff(a, b):
    return np.abs(np.angle(np.exp(1j * (a - b))))


def test_closed_form_solver_reproduces_the_reference_solver_golden_vectors(oracle):
    g = np.load(OL.GOLDEN / "ikfast_vectors.npz")
    worst, dropped = 0.0, 0
    for t, q, k, sols in zip(g["targets"], g["joints"], g["count"], g["solutions"]):
        pos, d = oracle.arm5d_fk(q)
        assert np.abs(pos - t[:3]).max() < 1e-12 and np.abs(d - t[3:]).max() < 1e-12   # same arm model
        mine = oracle.arm5d_ik(t[:3], t[3:])
        assert len(mine) == 8
        assert ang_diff(mine, q).max(1).min() < 1e-9                                   # the generating posture is found
        dropped += k < 8                # the reference's polynomial root search loses roots now and then
        for s in sols[:k]:              # ... but every solution it returns is one of ours
            worst = max(worst, ang_diff(mine, s).max(1).min())
    assert worst < 1e-6, worst
    assert dropped <= 4                 # counted: 1 of 400 in the committed vectors


def test_closed_form_solver_against_the_live_reference_solver(oracle):
    so = OL.ROOT / "oracle" / "_ref" / "libikfast_ref.so"
    if not so.exists():
        pytest.skip("oracle/_ref not built (needs /root/reference: make -C oracle ref)")
    lib = C.CDLL(str(so))
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    rng = np.random.default_rng(3)
    fewer = 0
    for i in range(3000):
        q = rng.uniform(ARM_LO, ARM_HI)
        pos, d = oracle.arm5d_fk(q)
        out = np.zeros((16, 5))
        k = lib.ikf_ref_solve(P(pos), P(d), P(out), 16)
        mine = oracle.arm5d_ik(pos, d)
        assert len(mine) == 8
        fewer += k < 8
        for s in out[:k]:
            assert ang_diff(mine, s).max(1).min() < 1e-6
    assert fewer < 30                   # ~0.3 % of targets: roots the generated solver misses


def test_unreachable_target_has_no_solution(oracle):
    assert len(oracle.arm5d_ik([0.6, -0.1, 0.1], [0, -1, 0])) == 0


@pytest.mark.parametrize("fname,pose,tol", [
    ("ssc_double_support_goal_second.dat", [0.17, -0.065, 0.30, 0, -1, 0], 1e-4),      # EXE:617-625: 0.24 - 0.07
    ("ssc_double_support_goal_first.dat", [0.23, -0.065, 0.20, 0, -1, 0], 3e-4),       # EXE:542-547
])
def test_shipped_goal_files_are_reproduced_by_goal_arm_ik(oracle, fname, pose, tol):
    """The files hold what computeRightArmIK (NCK:224-506) returned: from the file's own legs (hip pose)
    and the requested hand pose, OUR solver + ergonomics selection must return the file's arm columns
    (to the 6 significant digits the file keeps)."""
    rows = OL.load_dat(GOLD / fname)
    ok, arm, erg = oracle.goal_right_arm_ik(rows, pose)
    assert ok.all()
    assert np.abs(arm - rows[:, 10:15]).max() < tol
    assert (erg > 0).all()


def test_shipped_goal_file_is_sorted_by_our_ergonomics_index(oracle):
    """sort_goal_configs_by_ergonomy (SCG:510-661) wrote goal_second.dat in descending ergonomics order:
    pins the quirky 'determinant' of compute_ergonomics_index (NCK:612-659)."""
    rows = OL.load_dat(GOLD / "ssc_double_support_goal_second.dat")
    erg = np.array([oracle.ergonomics_index(r[10:15]) for r in rows])
    assert len(rows) == 6 and (np.diff(erg) < 0).all() and (erg > 0).all()
    assert np.allclose(rows[:, 5:10], [1.5708, 0.17, 0.0, -0.036, 0.0])               # SCG:296-300 left arm fixed


def test_older_goal_files_put_the_hand_on_the_request(oracle):
    """Files of older code vintages (other selection rule, request only known to ~mm): the sole->hand
    KDL chain (AOC:42-56) still puts the grasp frame on the drawer handle, z-axis along -y."""
    for fname, pose in [("ssc_double_support_goal_first_drawer.dat", [0.20, -0.065, 0.20, 0, -1, 0]),
                        ("ssc_double_support_goal_second_drawer.dat", [0.10, -0.065, 0.20, 0, -1, 0])]:
        rows = OL.load_dat(GOLD / fname)
        for r in rows:
            pos, z = oracle.right_hand_fk(r)                        # KDL chain: within ~1.5 mm of the request
            assert np.abs(pos - pose[:3]).max() < 6e-3 and np.abs(z - pose[3:]).max() < 5e-3


def test_drawer_trajectory_points(oracle):
    a = np.array([0.24, -0.065, 0.30, 0, -1, 0.0])
    b = np.array([0.14, -0.065, 0.30, 0, -1, 0.0])
    pts = oracle.hand_trajectory(0, a, b, 20)
    assert pts.shape == (22, 6) and np.allclose(pts[0], a) and np.allclose(pts[-1], b)
    assert np.allclose(np.diff(pts[:, 0]), -0.1 / 21)


def test_door_trajectory_points_end_at_the_service_goal_pose(oracle):
    """computeDoorTrajectoryPoints (AOC:126-189) must end where NPC:762-790 puts the manipulation goal."""
    req = dict(pos=[0.27, 0.017, 0.2775], d=[0, 0, -1.0], relative_angle=10.0, radius=0.3, angle=30.0, clearance=0.03)
    rel = np.deg2rad(req["relative_angle"])
    x_door = req["pos"][0] + req["clearance"] * np.cos(rel)
    y_door = req["pos"][1] - req["clearance"] * np.sin(rel)
    x_h, y_h = x_door - req["radius"] * np.sin(rel), y_door - req["radius"] * np.cos(rel)
    t = np.deg2rad(req["angle"] - req["relative_angle"])
    x_door, y_door = x_h - req["radius"] * np.sin(t), y_h + req["radius"] * np.cos(t)
    t = 1.5708 - t
    goal = np.array([x_door - req["clearance"] * np.sin(t), y_door - req["clearance"] * np.cos(t), req["pos"][2], 0, 0, -1.0])
    start = np.array(req["pos"] + req["d"])
    pts = oracle.hand_trajectory(1, start, goal, 20, [req["radius"], req["angle"], req["relative_angle"], req["clearance"]])
    assert np.abs(pts[-1] - goal).max() < 1e-12
    assert np.abs(pts[0, :3] - start[:3]).max() < 1e-5          # 1.5708 vs pi/2 (AOC:176)
    hinge = np.array([x_h, y_h])
    # the door edge (hand + clearance along the door normal) stays on the circle about the hinge
    assert np.ptp(np.hypot(*(pts[:, :2] - hinge).T)) < 2e-3


def test_solution_path_second_follows_the_drawer_constraint(oracle):
    """E8: the shipped manipulation path (22 nodes x 5 - 4 rows) of the scenario-3 request (SIM:496-510:
    hand (0.24,-0.065,0.30), pull 0.10): node k keeps its hand within right_hand_at_pose's 5 mm / 0.007
    (AOC:247-262) of trajectory point k - an IK is only triggered beyond that, hence the uneven spacing."""
    path = OL.load_dat(OL.GOLDEN / "solutions" / "solution_path_second.dat")
    nodes = path[::5]
    assert len(nodes) == 22
    pts = oracle.hand_trajectory(0, [0.24, -0.065, 0.30, 0, -1, 0], [0.14, -0.065, 0.30, 0, -1, 0], 20)
    for k, n in enumerate(nodes):
        pos, z = oracle.right_hand_fk(n)
        assert np.abs(pos - pts[k, :3]).max() < 5e-3 + 1e-4 and np.abs(z - [0, -1, 0]).max() < 7e-3 + 1e-4
        r_leg, r_arm = -n[:5], n[10:15]
    # and the restated predicate itself accepts every node at its index
    hi = np.arange(22, dtype=np.int32)
    out, st = oracle.enforce_ma(pts, True, hi - 1, nodes, nodes)     # desired = (k-1)+1 = k (k = 0: goal-tree rule)
    assert (st[1:] == 1).all()                                        # REACHED: "Already at correct pose"


def test_enforce_ma_constraints_semantics(oracle):
    rows = OL.load_dat(GOLD / "ssc_double_support_goal_second.dat")
    q = rows[0]
    pos, z = oracle.right_hand_fk(q)
    # KDL chain and solver model differ by ~1 mm: build the trajectory from the solver-side pose
    a = np.r_[0.17, -0.065, 0.30, 0, -1, 0]
    b = np.r_[0.10, -0.065, 0.30, 0, -1, 0]
    pts = oracle.hand_trajectory(0, a, b, 5)                     # 7 cm towards the body in 6 steps of 11.7 mm
    n = 40
    near = np.tile(q, (n, 1))
    ds = near.copy()
    hi = np.zeros(n, np.int32)
    out, st = oracle.enforce_ma(pts, True, hi, near, ds)        # start tree: desired index 1, beyond 5 mm: IK
    assert (st == 0).all()                                       # ADVANCED
    for r in out[:4]:
        pos, z = oracle.right_hand_fk(r)
        assert np.abs(pos - pts[1, :3]).max() < 3e-3
    assert np.abs(out[:, :10] - ds[:, :10]).max() == 0 and np.abs(out[:, 15:] - ds[:, 15:]).max() == 0
    out2, st2 = oracle.enforce_ma(pts, False, hi, near, ds)      # goal tree at index 0: stays at 0 -> already there
    assert (st2 == 1).all() and (out2 == ds).all()               # REACHED
    hi[:] = 6
    out3, st3 = oracle.enforce_ma(pts, True, hi, near, ds)       # start tree at the last index: stays 6, 7 cm away
    assert (st3 == 0).all() or (st3 == 2).all()
    if (st3 == 0).all():
        pos, z = oracle.right_hand_fk(out3[0])
        assert np.abs(pos - pts[6, :3]).max() < 3e-3


def test_goal_sampler_finds_six_sorted_feasible_goals(oracle):
    pose = [0.24, -0.065, 0.30, 0, -1, 0]
    rows, erg, idx, drawn = oracle.sample_goal_configs(20121003, pose, 5000, 5)
    assert len(rows) == 6 and (np.diff(erg) <= 0).all() and drawn == idx.max() + 1
    valid, _, _ = oracle.is_feasible(rows)
    assert valid.sum() >= 5                                       # text rounding may push a boundary row over
    for r in rows:
        pos, z = oracle.right_hand_fk(r)
        assert np.abs(pos - pose[:3]).max() < 4e-3 and np.abs(z - pose[3:]).max() < 5e-3
        assert np.allclose(r[5:10], [1.5708, 0.17, 0.0, -0.036, 0.0]) and r[15] == 0 and r[10] > 0
    stage, e, q = oracle.goal_sample_batch(20121003, 0, drawn, pose, threads=oracle.threads)
    assert (np.nonzero(stage == 2)[0] == np.sort(idx)).all()


def test_scenario_3_linear_manipulation_plan(oracle):
    """V0 (SIM:496-510): approach + drawer pull of 10 cm; both stages succeed and the manipulation
    trajectory has the shape of the shipped solution_path_second.dat (22 nodes x 5 - 4 = 106 rows)."""
    db = OL.golden_db()
    start = np.zeros(21)
    start[8], start[13] = -0.5, 0.5
    pose = [0.24, -0.065, 0.30, 0, -1, 0]
    r = oracle.compute_manipulation_plan(db, pose, start, 20121003, False, [0.0, 0.10])
    assert r["success_approach"] and r["success_manipulation"]
    assert r["goals_approach"] == 6 and r["goals_manipulate"] == 6
    tm = r["traj_manipulation"]
    assert len(tm) == 106
    assert np.allclose(tm[:, 5:10], [1.5708, 0.17, 0.0, -0.036, 0.0], atol=1e-6)      # l_arm weight 0: frozen
    for k, n in enumerate(tm[::5]):
        pos, z = oracle.right_hand_fk(n)
        assert abs(pos[0] - (0.24 - 0.10 * k / 21)) < 6.5e-3 and abs(pos[1] + 0.065) < 6.5e-3 and abs(pos[2] - 0.30) < 6.5e-3   # 5 mm predicate + ~1 mm chain/solver model offset
    tfs = r["tfs_manipulation"]
    assert tfs[0] == 2.0 and (np.diff(tfs) > 0).all() and len(tfs) == len(tm)                    # time-parameterised
    raw = oracle.compute_manipulation_plan(db, pose, start, 20121003, False, [0.0, 0.10],
                                           params=OL.default_params(time_parameterization=0))
    assert np.allclose(raw["tfs_manipulation"][:2], [2.0, 2.0 + (0.10 / 0.02) / 22], atol=1e-6)   # RP:1532-1553 stamps
    assert (tm[0] == r["goal_approach"]).all() and np.abs(tm[-1] - r["goal_manipulate"]).max() < 1e-12
    ta = r["traj_approach"]
    assert (ta[0] == start).all() and np.abs(ta[-1] - r["goal_approach"]).max() < 1e-12 and (len(ta) - 1) % 101 == 0


def test_time_parameterization_known_answers(oracle):
    """IterativeParabolicTimeParameterization as generateTrajectory calls it (RP:1596-1618): one joint moving
    1 rad between two way-points needs T with 2/T^2 <= a_max + 0.01 (mirrored end points), reached in 0.01 s
    steps; durations respect the velocity limits; stamps start at 2.0 s and increase."""
    path = np.zeros((2, 21))
    path[1, 10] = 1.0
    t = oracle.time_parameterize(path)
    T = t[1] - t[0]
    assert t[0] == 2.0 and np.sqrt(2 / 1.01) <= T <= np.sqrt(2 / 1.01) + 0.02
    sol = OL.load_dat(OL.GOLDEN / "solutions" / "solution_path_second.dat")
    t = oracle.time_parameterize(sol)
    dt = np.diff(t)
    assert (dt > 0).all()
    v = np.abs(np.diff(sol, axis=0)) / dt[:, None]
    assert (v <= np.array(nwb_vmax()) * (1 + 1e-12)).all()
    # interior accelerations of the parabolic blends are within the limit + rounding threshold once converged
    a = 2 * (np.diff(sol, axis=0)[1:] / dt[1:, None] - np.diff(sol, axis=0)[:-1] / dt[:-1, None]) / (dt[1:] + dt[:-1])[:, None]
    assert np.abs(a).max() <= 1.0 + 0.01 + 1e-9
    assert oracle.time_parameterize(sol[:1])[0] == 2.0


def test_time_parameterization_three_way_points_by_hand(oracle):
    """Two hand-computed three-way-point answers.
    (1) Velocity regime - accelerations unbounded, so computeTimeStamps reduces to applyVelocityConstraints: every
        interval lasts max_j |dq_j| / v_max_j exactly, stamps = 2.0 + running sum (RP:1614-1618 adds the 2 s head start).
    (2) Acceleration regime - one joint going 0 -> D -> 0: at all three points the blend acceleration is 2 D / T^2 in
        magnitude (the end points mirror their only neighbour), so both intervals must reach T >= sqrt(2 D / (a_max +
        0.01)) and the 0.01 s growth steps stop within two steps of it."""
    vmax = np.array(nwb_vmax())
    p = OL.default_params()
    p.joint_max_acceleration[:] = [1e9] * 21
    path = np.zeros((3, 21))
    path[1, 10], path[1, 11] = 0.5, -0.2
    path[2, 10], path[2, 11] = 0.6, 0.2
    t = oracle.time_parameterize(path, params=p)
    dt0 = max(0.5 / vmax[10], 0.2 / vmax[11])
    dt1 = max(0.1 / vmax[10], 0.4 / vmax[11])
    assert np.allclose(t, [2.0, 2.0 + dt0, 2.0 + dt0 + dt1], rtol=0, atol=1e-15)
    D = 0.4
    path = np.zeros((3, 21))
    path[1, 12] = D
    t = oracle.time_parameterize(path)                                  # default limits: a_max = 1.0
    T = np.sqrt(2 * D / 1.01)
    assert t[0] == 2.0
    for dt in np.diff(t):
        assert T <= dt <= T + 0.02, (dt, T)


def nwb_vmax():
    from nao_wholebody_planning_b200._abi import NAO_MAX_VELOCITY
    return NAO_MAX_VELOCITY
