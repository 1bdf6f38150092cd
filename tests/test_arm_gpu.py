"""GPU parity (through the C ABI) of the right-arm operators, the articulated-object constraint, goal
sampling and the planning services against the CPU oracle and the reference solver's golden vectors."""
import numpy as np
import pytest

import nao_wholebody_planning_b200 as nwb
import oracle_lib as OL

pytestmark = pytest.mark.gpu
POSE_DRAWER = [0.24, -0.065, 0.30, 0, -1, 0]        # SIM:496-501
POSE_DOOR = [0.27, 0.02, 0.35, 0, 0, -1]            # SIM:551-556
START = np.zeros(21)
START[8], START[13] = -0.5, 0.5                     # SIM:232-252


@pytest.fixture(scope="module")
def ctx():
    c = nwb.Context()
    c.set_ds_database(OL.golden_db())
    yield c
    c.close()


def ang_diff(a, b):
    return np.abs(np.angle(np.exp(1j * (a - b))))


def test_arm_ik_matches_reference_solver_golden_vectors(ctx, oracle):
    g = np.load(OL.GOLDEN / "ikfast_vectors.npz")
    sols, cnt = ctx.arm_ik(g["targets"])
    assert (cnt == 8).all()
    worst = 0.0
    for mine, q, k, ref in zip(sols, g["joints"], g["count"], g["solutions"]):
        assert ang_diff(mine, q).max(1).min() < 1e-9
        for s in ref[:k]:
            worst = max(worst, ang_diff(mine, s).max(1).min())
    assert worst < 1e-6                                # the generated solver's own root-polishing accuracy
    # against the oracle: same solutions in the same order
    for t, mine in zip(g["targets"][:100], sols[:100]):
        o = oracle.arm5d_ik(t[:3], t[3:])
        assert len(o) == 8 and ang_diff(mine, o).max() < 1e-9


def test_arm_ik_edge_cases(ctx, oracle):
    t = np.array([[0.6, -0.1, 0.1, 0, -1, 0],          # out of reach
                  [0.0, -0.098, 0.1, 0, -1, 0],        # at the shoulder
                  [0.2, -0.1, 0.1, 1, 0, 0]])
    sols, cnt = ctx.arm_ik(t)
    for i in range(len(t)):
        assert cnt[i] == len(oracle.arm5d_ik(t[i, :3], t[i, 3:]))
    assert cnt[0] == 0 and np.isnan(sols[0]).all()


def test_goal_arm_ik_matches_oracle_and_shipped_goal_file(ctx, oracle):
    rows = OL.load_dat(OL.GOLDEN / "config_database" / "ssc_double_support_goal_second.dat")
    ok, arm, erg = ctx.computeRightArmIK(rows, [0.17, -0.065, 0.30, 0, -1, 0])
    assert ok.all() and np.abs(arm - rows[:, 10:15]).max() < 1e-4          # the file keeps 6 digits
    db = OL.golden_db()
    for pose in (POSE_DRAWER, POSE_DOOR, [0.2, -0.065, 0.2, 0, -1, 0]):
        ok, arm, erg = ctx.computeRightArmIK(db, pose)
        ook, oarm, oerg = oracle.goal_right_arm_ik(db, pose)
        assert (ok == ook).all() and ok.sum() > 10
        m = ok == 1
        assert np.abs(arm[m] - oarm[m]).max() < 1e-9
        assert np.abs(erg[m] - oerg[m]).max() < 1e-12 * max(1.0, np.abs(oerg).max()) + 1e-15


def test_position_only_arm_ik_branch_matches_oracle(ctx, oracle):
    """NCK:291-372: a hand direction of exactly (0,0,0) selects KDL NR on the arm chain towards Frame(position) -
    identity rotation, so it almost never succeeds; statuses and solutions must match the oracle, the seeds come
    from the injected stream."""
    db = OL.golden_db()
    rows = db[:300].copy()
    # give some rows a target they can reach: hand pose of the row itself, orientation forced to identity
    for i in range(0, 300, 3):
        rows[i, 10:15] = [0.3, -0.2, -1.2, 0.04, -0.37]
    pose = [0.2, -0.1, 0.25, 0, 0, 0]
    ok, arm, erg = ctx.computeRightArmIK(rows, pose, seed=7, first_index=100)
    ook, oarm, oerg = oracle.goal_right_arm_ik(rows, pose, seed=7, first_index=100)
    assert (ok == ook).all() and (erg == 0).all() and (oerg == 0).all()
    if ok.any():
        assert np.abs(arm[ok == 1] - oarm[ok == 1]).max() < 1e-8
    # through the sampler: no goal configuration, no error
    rows6, e6, idx6, drawn = ctx.sample_goal_configs(3, pose, max_samples=600)
    orows6, _, oidx6, odrawn = oracle.sample_goal_configs(3, pose, max_samples=600)
    assert len(rows6) == len(orows6) and drawn == odrawn
    r = ctx.compute_motion_plan(pose[:3], pose[3:], START, 1)
    assert r["success"] in (False, True)


def test_hand_trajectories_match_oracle(ctx, oracle):
    a, b = np.array(POSE_DRAWER, float), np.array([0.14, -0.065, 0.30, 0, -1, 0.0])
    assert (ctx.hand_trajectory(0, a, b, 20) == oracle.hand_trajectory(0, a, b, 20)).all()
    door = [0.25, 40.0, 0.0, 0.05]
    g = np.array([0.1, 0.1, 0.35, 0, 0, -1.0])
    assert np.abs(ctx.hand_trajectory(1, POSE_DOOR, g, 20, door) - oracle.hand_trajectory(1, POSE_DOOR, g, 20, door)).max() < 1e-15


@pytest.mark.parametrize("start_tree", [True, False])
def test_enforce_ma_constraints_matches_oracle(ctx, oracle, start_tree):
    rng = np.random.default_rng(7)
    db = OL.golden_db()
    # rows whose arm reaches the drawer handle: solve the goal IK first, then perturb
    ok, arm, _ = oracle.goal_right_arm_ik(db, [0.19, -0.065, 0.30, 0, -1, 0])
    base = db[ok == 1].copy()
    base[:, 10:15] = arm[ok == 1]
    near = base[rng.integers(0, len(base), 600)]
    ds = near + rng.normal(0, 0.01, near.shape)
    pts = oracle.hand_trajectory(0, [0.24, -0.065, 0.30, 0, -1, 0], [0.14, -0.065, 0.30, 0, -1, 0], 20)
    hi = rng.integers(0, 22, len(near)).astype(np.int32)
    q, st = ctx.enforce_MA_Constraints(pts, start_tree, hi, near, ds)
    oq, ost = oracle.enforce_ma(pts, start_tree, hi, near, ds)
    assert (st == ost).mean() > 0.995                   # a 5 mm / 0.007 predicate flips only on its boundary
    same = st == ost
    live = same & (st != nwb.TRAPPED)
    assert np.abs(q[live] - oq[live]).max() < 1e-9
    assert np.isnan(q[st == nwb.TRAPPED]).all()
    assert {0, 1, 2} <= set(np.unique(st))


def test_goal_sampler_matches_oracle(ctx, oracle):
    seed, count = 20121003, 6000
    stage, erg, rows = ctx.goal_sample_batch(seed, 0, count, POSE_DRAWER)
    ostage, oerg, orows = oracle.goal_sample_batch(seed, 0, count, POSE_DRAWER, threads=oracle.threads)
    assert (stage != ostage).sum() <= 2
    acc = (stage == 2) & (ostage == 2)
    assert acc.sum() >= 6
    assert np.abs(rows[acc] - orows[acc]).max() < 1e-8
    assert np.abs(erg[acc] - oerg[acc]).max() < 1e-12
    # partition independence
    s2, e2, r2 = ctx.goal_sample_batch(seed, 1000, 500, POSE_DRAWER)
    assert (s2 == stage[1000:1500]).all() and (r2[s2 == 2] == rows[1000:1500][s2 == 2]).all()


@pytest.mark.parametrize("pose,seed", [(POSE_DRAWER, 20121003), (POSE_DOOR, 5), ([0.14, -0.065, 0.30, 0, -1, 0], 11)])
def test_sample_goal_configs_matches_oracle(ctx, oracle, pose, seed):
    rows, erg, idx, drawn = ctx.sample_goal_configs(seed, pose)
    orows, oerg, oidx, odrawn = oracle.sample_goal_configs(seed, pose)
    assert len(rows) == len(orows) and (idx == oidx).all() and drawn == odrawn
    assert np.abs(rows - orows).max() < 2e-6            # both sides pass through 6 significant digits
    assert (np.diff(erg) <= 0).all()
    valid, _, _ = ctx.isFeasible(rows)
    assert valid.sum() >= len(rows) - 1


def test_constrained_solve_query_matches_oracle(ctx, oracle):
    db = OL.golden_db()
    ga, _, _, _ = oracle.sample_goal_configs(20121003, POSE_DRAWER)
    pm = [0.14, -0.065, 0.30, 0, -1, 0]
    gm, _, _, _ = oracle.sample_goal_configs(20121004, pm)
    pts = oracle.hand_trajectory(0, POSE_DRAWER, pm, 20)
    w = np.ones(21)
    w[5:10] = 0.0
    w[15] = 0.0
    nq = 6
    seeds = np.arange(100, 100 + nq, dtype=np.uint64)
    starts = np.tile(ga[0], (nq, 1))
    goals = np.tile(gm[0], (nq, 1))
    starts[1], goals[2] = ga[1], gm[1]
    res, paths = ctx.solveQuery(starts, goals, db, seeds, w, hand_trajectories=np.tile(pts, (nq, 1, 1)))
    n_ok = 0
    for k in range(nq):
        raw, st = oracle.solve_query_ma(starts[k], goals[k], db, w, int(seeds[k]), pts)
        assert res[k]["success"] == st["success"]
        if not st["success"]:
            continue
        n_ok += 1
        assert res[k]["iterations"] == st["iterations"] and res[k]["num_nodes_generated"] == st["num_nodes"]
        assert len(paths[k]) == len(raw) and np.abs(paths[k] - raw).max() < 1e-8
        assert len(raw) == 22                                       # one node per trajectory point
    assert n_ok >= 4


def test_linear_manipulation_plan_service_matches_oracle(ctx, oracle):
    """Scenario 3 (V0, SIM:496-510)."""
    db = OL.golden_db()
    for seed in (20121003, 20121004):
        r = ctx.compute_linear_manipulation_plan(POSE_DRAWER[:3], POSE_DRAWER[3:], START, 0.0, 0.10, seed)
        o = oracle.compute_manipulation_plan(db, POSE_DRAWER, START, seed, False, [0.0, 0.10])
        assert r["success_approach"] == o["success_approach"] and r["success_manipulation"] == o["success_manipulation"]
        assert r["success_approach"] and r["success_manipulation"]
        assert np.abs(r["wb_goal_config_approach"] - o["goal_approach"]).max() < 2e-6
        assert np.abs(r["wb_goal_config_manipulate"] - o["goal_manipulate"]).max() < 2e-6
        assert r["num_nodes_generated_approach"] == o["nodes_approach"]
        assert r["num_nodes_generated_manipulate"] == o["nodes_manipulate"]
        assert r["traj_approach"].shape == o["traj_approach"].shape
        assert np.abs(r["traj_approach"] - o["traj_approach"]).max() < 1e-5
        assert r["traj_manipulation"].shape == o["traj_manipulation"].shape == (106, 21)
        assert np.abs(r["traj_manipulation"] - o["traj_manipulation"]).max() < 1e-5
        # time stamps: the parameterisation runs on way-points that agree to ~1e-9 between device and oracle
        assert np.abs(r["tfs_approach"] - o["tfs_approach"]).max() < 1e-6
        assert np.abs(r["tfs_manipulation"] - o["tfs_manipulation"]).max() < 1e-6
        assert r["time_goal_config_approach"] > 0 and r["search_time_manipulate"] > 0


def test_linear_manipulation_plan_for_execution(ctx, oracle):
    """setTrajectoryMode(true) (NPC:362): the articulated plan is returned raw - one way-point per tree node,
    stamps 2.0 + i * time_steps_manipulation (RP:1417, 1547-1575), no time parameterisation - and the approach
    plan keeps only the shortcut way-points (RP:1941-1945)."""
    db = OL.golden_db()
    r = ctx.compute_linear_manipulation_plan(POSE_DRAWER[:3], POSE_DRAWER[3:], START, 0.0, 0.10, 20121003, execute_trajectory=True)
    o = oracle.compute_manipulation_plan(db, POSE_DRAWER, START, 20121003, False, [0.0, 0.10], execute=True)
    assert r["success_approach"] and r["success_manipulation"] and o["success_manipulation"]
    assert r["traj_manipulation"].shape == o["traj_manipulation"].shape == (22, 21)
    assert np.abs(r["traj_manipulation"] - o["traj_manipulation"]).max() < 1e-8
    step = np.float32(np.float32(0.10) / np.float32(0.02)) / np.float32(22)
    assert np.allclose(np.diff(r["tfs_manipulation"]), step, atol=1e-6) and r["tfs_manipulation"][0] == 2.0
    assert (r["tfs_manipulation"] == o["tfs_manipulation"]).all()
    assert r["traj_approach"].shape == o["traj_approach"].shape and len(r["traj_approach"]) < 20
    assert np.abs(r["traj_approach"] - o["traj_approach"]).max() < 1e-8


def test_circular_manipulation_plan_service_matches_oracle(ctx, oracle):
    """Door opening (SIM:551-561)."""
    db = OL.golden_db()
    r = ctx.compute_circular_manipulation_plan(POSE_DOOR[:3], POSE_DOOR[3:], START, 0.0, 0.25, 40.0, 0.05, 1)
    o = oracle.compute_manipulation_plan(db, POSE_DOOR, START, 1, True, [0.0, 0.25, 40.0, 0.05])
    assert r["success_approach"] and r["success_manipulation"] and o["success_manipulation"]
    assert np.abs(r["wb_goal_config_manipulate"] - o["goal_manipulate"]).max() < 2e-6
    assert r["num_nodes_generated_manipulate"] == o["nodes_manipulate"]
    assert r["traj_manipulation"].shape == o["traj_manipulation"].shape
    assert np.abs(r["traj_manipulation"] - o["traj_manipulation"]).max() < 1e-5
    assert np.abs(r["tfs_manipulation"] - o["tfs_manipulation"]).max() < 1e-6


def test_motion_plan_and_goal_config_services(ctx, oracle):
    db = OL.golden_db()
    pose = [0.24, -0.065, 0.3, 0, -1, 0]                            # SIM:441-446
    goal, n = ctx.compute_goal_config(pose[:3], pose[3:], 42)
    orows, _, _, _ = oracle.sample_goal_configs(42, pose)
    assert n == len(orows) and np.abs(goal - orows[0]).max() < 2e-6
    for execute in (False, True):
        r = ctx.compute_motion_plan(pose[:3], pose[3:], START, 42, execute_trajectory=execute)
        o = oracle.compute_motion_plan(db, pose, START, 42, execute=execute)
        assert r["success"] and o["success"]
        assert r["num_nodes_generated"] == o["num_nodes"]
        assert r["trajectory"].shape == o["trajectory"].shape
        assert np.abs(r["trajectory"] - o["trajectory"]).max() < 1e-5
        assert np.abs(r["time_from_start"] - o["time_from_start"]).max() < 1e-6
    # unreachable hand pose: no goal configuration, success false, no error (NPC:331-334)
    r = ctx.compute_motion_plan([0.6, 0.0, 0.3], [0, -1, 0], START, 1)
    assert not r["success"] and len(r["trajectory"]) == 0
    # short start vector (NPC:242 reads out of bounds; here: an argument error)
    with pytest.raises(nwb.NwbError, match="bad argument"):
        ctx.compute_motion_plan(pose[:3], pose[3:], START[:10], 1)


def test_phased_goal_sampler_equals_the_one_kernel_form(ctx, monkeypatch):
    """The services run the goal sampler as three launches over packed work lists; NWB_GOAL_ONE_KERNEL selects the
    one-thread-per-sample kernel (the operator form the oracle test above checks).  Same goal sets, bit for bit -
    direction given, position only (KDL branch), an unreachable pose, and a batch with host-thread finishing."""
    poses = [POSE_DRAWER, POSE_DOOR, [0.14, -0.065, 0.30, 0, -1, 0], [0.20, -0.10, 0.28, 0, 0, 0], [0.9, 0.0, 0.3, 0, -1, 0]]

    def run():
        sets = [ctx.sample_goal_configs(40 + k, p, max_samples=3000 if k == 3 else 5000) for k, p in enumerate(poses)]
        pos = np.array([[-0.08 + i * 0.4 / 15, -0.30 + j * 0.4 / 15, 0.2] for i in (5, 8, 11, 14) for j in (2, 6, 10, 14)])
        start = np.zeros(21)
        batch = ctx.compute_motion_plan_batch(pos, np.tile([0.0, -1.0, 0.0], (16, 1)), np.tile(start, (16, 1)),
                                              300 + np.arange(16, dtype=np.uint64))
        return sets, batch

    monkeypatch.delenv("NWB_GOAL_ONE_KERNEL", raising=False)
    monkeypatch.setenv("NWB_HOST_THREADS", "4")
    sets_a, batch_a = run()
    monkeypatch.setenv("NWB_GOAL_ONE_KERNEL", "1")
    monkeypatch.setenv("NWB_HOST_THREADS", "1")
    sets_b, batch_b = run()
    assert sum(len(a[0]) for a in sets_a) >= 12 and len(sets_a[4][0]) == 0
    for a, b in zip(sets_a, sets_b):
        assert len(a[0]) == len(b[0]) and a[3] == b[3]
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (a[2] == b[2]).all()
    assert sum(r["success"] for r in batch_a) >= 4
    for a, b in zip(batch_a, batch_b):
        assert a["success"] == b["success"] and a["num_nodes_generated"] == b["num_nodes_generated"]
        assert (a["wb_goal_config"] == b["wb_goal_config"]).all()
        assert a["trajectory"].shape == b["trajectory"].shape and (a["trajectory"] == b["trajectory"]).all()
        assert (a["time_from_start"] == b["time_from_start"]).all()


def test_motion_plan_batch_equals_single_calls(ctx):
    """V4: the grasp-pose grid of SIM:778-788 (a 3 x 3 corner of it) from the crouched pose SIM:199-219."""
    start = np.array([0, -0.349066, 0.698132, -0.436332, 0, 1.39626, 0.349066, -1.39626, -0.5, 0, 1.39626, -0.349066, 1.39626,
                      0.5, 0, 0, 0, -0.436332, 0.698132, -0.349066, 0])
    pos = np.array([[-0.08 + i * 0.4 / 15, -0.30 + j * 0.4 / 15, 0.2] for i in (7, 9, 11) for j in (6, 8, 10)] + [[0.6, 0.0, 0.3]])
    dirs = np.tile([0.0, -1.0, 0.0], (len(pos), 1))
    seeds = 20121100 + np.arange(len(pos), dtype=np.uint64)
    batch = ctx.compute_motion_plan_batch(pos, dirs, np.tile(start, (len(pos), 1)), seeds)
    n_ok = 0
    for k in range(len(pos)):
        one = ctx.compute_motion_plan(pos[k], dirs[k], start, int(seeds[k]))
        assert batch[k]["success"] == one["success"]
        assert (batch[k]["wb_goal_config"] == one["wb_goal_config"]).all() or not one["success"]
        assert batch[k]["num_nodes_generated"] == one["num_nodes_generated"]
        assert batch[k]["trajectory"].shape == one["trajectory"].shape
        if one["success"]:
            n_ok += 1
            assert (batch[k]["trajectory"] == one["trajectory"]).all()
    assert n_ok >= 3 and not batch[-1]["success"]


def test_time_parameterization_matches_oracle(ctx, oracle):
    sol = OL.load_dat(OL.GOLDEN / "solutions" / "solution_path_second.dat")
    for path in (sol, sol[::7], sol[:2], sol[:1]):
        assert (ctx.time_parameterize(path) == oracle.time_parameterize(path)).all()
    with nwb.Context(time_parameterization=0) as c:
        c.set_ds_database(OL.golden_db())
        r = c.compute_linear_manipulation_plan(POSE_DRAWER[:3], POSE_DRAWER[3:], START, 0.0, 0.10, 20121003)
        assert np.allclose(r["tfs_manipulation"][:2], [2.0, 2.0 + (0.10 / 0.02) / 22], atol=1e-6)
        assert np.allclose(np.diff(r["tfs_approach"]), 1.0)


def test_motion_plan_service_with_scene_boxes(oracle):
    """The drawer + beam scene (PAR:465-472) in front of the robot: goal sampling rejects colliding goal postures,
    the search and the shortcutter see the boxes; same plan as the oracle."""
    db = OL.golden_db()
    boxes = OL.scene_boxes("S3")
    pose = [0.24, -0.065, 0.30, 0, -1, 0]
    with nwb.Context() as c:
        c.scene_add_boxes(boxes)
        c.set_ds_database(db)
        r = c.compute_motion_plan(pose[:3], pose[3:], START, 20121003)
        free = nwb.Context()
        free.set_ds_database(db)
        r0 = free.compute_motion_plan(pose[:3], pose[3:], START, 20121003)
        free.close()
    o = oracle.compute_motion_plan(db, pose, START, 20121003, boxes=boxes)
    assert r["success"] == o["success"] and r["num_nodes_generated"] == o["num_nodes"]
    assert np.abs(r["wb_goal_config"] - o["goal_config"]).max() < 2e-6
    assert r["trajectory"].shape == o["trajectory"].shape
    if o["success"]:
        assert np.abs(r["trajectory"] - o["trajectory"]).max() < 1e-5
        col, _ = oracle.is_state_colliding(r["trajectory"][::10], boxes=boxes)
        assert not col.any()
    # a hand pose inside the shelf of scene S1 (between two boards): every goal posture collides -> no plan, no error
    with nwb.Context() as c:
        c.scene_add_boxes(OL.scene_boxes("S1"))
        c.set_ds_database(db)
        r1 = c.compute_motion_plan([0.30, -0.06, 0.30], [0, -1, 0], START, 7)
    o1 = oracle.compute_motion_plan(db, [0.30, -0.06, 0.30, 0, -1, 0], START, 7, boxes=OL.scene_boxes("S1"))
    assert r1["success"] == o1["success"] and len(r1["trajectory"]) == len(o1["trajectory"])
    assert r0["success"]


def test_services_need_a_database(oracle):
    with nwb.Context() as c:
        with pytest.raises(nwb.NwbError, match="database"):
            c.compute_motion_plan(POSE_DRAWER[:3], POSE_DRAWER[3:], START, 20121003)
