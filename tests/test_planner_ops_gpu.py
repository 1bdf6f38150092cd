"""GPU parity: steer + double-support projection, connect walks, edge checks and the stable-config
sampler (through the C ABI) against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

import nao_wholebody_planning_b200 as nwb
import oracle_lib as OL

pytestmark = pytest.mark.gpu
W1 = np.ones(21)
W_APPROACH = np.ones(21)
W_APPROACH[15] = 0.0          # nao_planner_control.cpp:301: LHipYawPitch weight 0


@pytest.fixture(scope="module")
def ctx():
    c = nwb.Context()
    yield c
    c.close()


def steer_inputs(n, seed):
    rng = np.random.default_rng(seed)
    db = OL.golden_db()
    near = db[rng.integers(0, len(db), n)]
    target = db[rng.integers(0, len(db), n)]
    return near, target


@pytest.mark.parametrize("weights,step", [(W1, 0.1), (W_APPROACH, 0.1), (W1, 0.5)])
def test_steer_project_matches_oracle(ctx, oracle, weights, step):
    near, target = steer_inputs(4000, 1)
    target[:50] = near[:50]                       # q_rand == q_near -> REACHED through the NaN direction
    target[50:100] = near[50:100] + 0.01          # closer than one step -> REACHED
    q, st, ds = ctx.steer_project(near, target, weights, step)
    oq, ost, ods = oracle.steer_project(near, target, weights, step, threads=oracle.threads)
    assert (st == ost).all() and (ds == ods).all()
    live = st != nwb.TRAPPED
    assert np.abs(q[live] - oq[live]).max() < 1e-9            # fp64 on both sides; libm differences only
    assert np.isnan(q[~live]).all()                           # TRAPPED rows are left untouched
    assert {0, 1} <= set(np.unique(st))


def test_steer_project_random_configs_exercise_trapped_and_ik(ctx, oracle):
    near = OL.uniform_configs(6000, seed=5, lhyp_zero=True)
    target = OL.uniform_configs(6000, seed=6, lhyp_zero=True)
    q, st, ds = ctx.steer_project(near, target, W1, 0.1)
    oq, ost, ods = oracle.steer_project(near, target, W1, 0.1, threads=oracle.threads)
    # the NR iteration can flip its convergence decision within ~1e-12 of eps: allow a handful
    diff = (st != ost) | (ds != ods)
    assert diff.sum() <= 2, diff.sum()
    live = (st != nwb.TRAPPED) & ~diff
    assert np.abs(q[live] - oq[live]).max() < 1e-8
    assert (ds == nwb.TRAPPED).sum() > 100 and (ds == nwb.ADVANCED).sum() > 100


def test_quirk_c2_signed_vs_absolute_foot_error(oracle):
    near, target = steer_inputs(1500, 9)
    near = near.copy()
    near[:, 18] += 0.15                            # bend the left knee: foot off its target
    for quirks in (nwb._abi.QUIRK_SIGNED_FOOT_ERROR | nwb._abi.QUIRK_DB_INDEX_FLOOR, nwb._abi.QUIRK_DB_INDEX_FLOOR):
        with nwb.Context(quirks=quirks) as c:
            q, st, ds = c.steer_project(near, target, W1, 0.1)
        oq, ost, ods = oracle.steer_project(near, target, W1, 0.1, params=OL.default_params(quirks=quirks))
        assert (st == ost).all() and (ds == ods).all()


def test_connect_walks_match_oracle(ctx, oracle):
    near, target = steer_inputs(96, 3)
    nodes, added, st = ctx.connect(near, target, W_APPROACH, 0.1, max_steps=128)
    n_reached = 0
    for e in range(len(near)):
        ost, onodes = oracle.connect_edge(near[e], target[e], W_APPROACH, 0.1)
        assert st[e] == ost, e
        assert added[e] == len(onodes), e
        if len(onodes):
            assert np.abs(nodes[e, :added[e]] - onodes).max() < 1e-8
        n_reached += ost == nwb.REACHED
    assert n_reached > 5 and (st == nwb.TRAPPED).sum() > 5


def test_connect_step_budget(ctx):
    near, target = steer_inputs(16, 4)
    nodes, added, st = ctx.connect(near, target, W1, 0.01, max_steps=3)
    assert (added <= 3).all() and ((st == -1) == (added == 3)).all()


@pytest.mark.parametrize("k", [100, 4, 0])
def test_is_path_valid_matches_oracle(ctx, oracle, k):
    rng = np.random.default_rng(11)
    db = OL.golden_db()
    qa = db[rng.integers(0, len(db), 600)]
    qb = db[rng.integers(0, len(db), 600)]
    qb[:20] = qa[:20]                               # zero-length edges
    ok, fb = ctx.isPathValid(qa, qb, k)
    ook, ofb = oracle.is_path_valid(qa, qb, k, threads=oracle.threads)
    bad = np.nonzero((ok != ook) | (fb != ofb))[0]
    # a way-point within 1e-5 m of contact may flip: verify each disagreement is such a case
    for e in bad:
        way = oracle.interpolate(qa[e], qb[e], k)
        _, sep = oracle.is_state_colliding(way, group_only=False)
        cand = [x for x in (fb[e], ofb[e]) if x >= 0]
        assert min(abs(sep[c]) for c in cand) < 1e-5
    assert len(bad) <= 3
    assert ok.sum() > 0 and (k == 0 or ok.sum() < len(ok))   # k = 0: only the (valid) end points are checked


def test_is_path_valid_with_scene(oracle):
    boxes = OL.scene_boxes("S2")
    rng = np.random.default_rng(12)
    db = OL.golden_db()
    qa, qb = db[rng.integers(0, len(db), 200)], db[rng.integers(0, len(db), 200)]
    with nwb.Context() as c:
        c.scene_add_boxes(boxes)
        ok, fb = c.isPathValid(qa, qb, 100)
    ook, ofb = oracle.is_path_valid(qa, qb, 100, boxes=boxes, threads=oracle.threads)
    assert ((ok == ook) & (fb == ofb)).mean() > 0.98
    assert ook.sum() < oracle.is_path_valid(qa, qb, 100, threads=oracle.threads)[0].sum()


def test_interpolation_endpoints_are_exact(ctx, oracle):
    """RP:1763: the first way-point is waypoint_start itself; a colliding start fails at index 0."""
    q_bad = OL.uniform_configs(4000, seed=2)
    col, _ = oracle.is_state_colliding(q_bad, group_only=False)
    q_bad = q_bad[col == 1][:50]
    db = OL.golden_db()[:50]
    ok, fb = ctx.isPathValid(q_bad, db, 100)
    assert (ok == 0).all() and (fb == 0).all()
    ok, fb = ctx.isPathValid(db, q_bad, 100)
    assert (ok == 0).all() and (fb <= 101).all() and (fb > 0).all()


def test_sampler_matches_oracle_and_is_partition_independent(ctx, oracle):
    seed, count = 20121005, 20000
    rows, idx, stage = ctx.sample_stable_configs(seed, 0, count)
    orows, oidx, ostage = oracle.sample_stable_configs(seed, 0, count, threads=oracle.threads)
    flips = (stage != ostage).sum()
    assert flips <= 2                                      # boundary-band configurations only
    common = np.intersect1d(idx, oidx)
    assert len(common) >= len(oidx) - 2
    a = rows[np.isin(idx, common)]
    b = orows[np.isin(oidx, common)]
    assert np.abs(a - b).max() < 1e-8
    assert (np.diff(idx.astype(np.int64)) > 0).all()       # ascending sample index
    assert 0.005 < len(idx) / count < 0.5
    # partition independence: two halves and an offset window give the same rows
    r1, i1, _ = ctx.sample_stable_configs(seed, 0, count // 2)
    r2, i2, _ = ctx.sample_stable_configs(seed, count // 2, count - count // 2)
    assert (np.concatenate([i1, i2]) == idx).all() and (np.concatenate([r1, r2]) == rows).all()
    # every accepted row is in double support and valid (through the GPU validity operator)
    valid, _, _ = ctx.isFeasible(rows)
    assert valid.all()
    poses, _ = ctx.fk(rows[:2000])
    # KDL's NR applies the pinv update BEFORE its convergence test: near a singular Jacobian
    # (sigma_min just above the 1e-5 threshold) the last update can throw the foot off again.
    # Faithfully reproduced (the oracle returns the same rows); ~0.5 % of accepted samples.
    foot_err = np.abs(poses[:, 28, :, 3] - [0, 0.1, 0]).max(1)
    assert (foot_err < 2e-3).mean() > 0.98 and np.median(foot_err) < 1e-3
    assert (rows[:, 15] == 0).all()


def test_sampler_quirk_c1_keeps_the_infeasible_samples(oracle):
    seed, count = 7, 8000
    q = nwb._abi.QUIRK_DB_KEEPS_INFEASIBLE | nwb._abi.QUIRK_SIGNED_FOOT_ERROR | nwb._abi.QUIRK_DB_INDEX_FLOOR
    with nwb.Context(quirks=q) as c:
        rows, idx, stage = c.sample_stable_configs(seed, 0, count)
        valid, _, _ = c.isFeasible(rows)
    assert (stage[idx.astype(np.int64)] == 1).all() and valid.sum() == 0
    orows, oidx, ostage = oracle.sample_stable_configs(seed, 0, count, params=OL.default_params(quirks=q))
    assert abs(len(oidx) - len(idx)) <= 2


def test_sampler_capacity_overflow_is_reported(ctx):
    with pytest.raises(nwb.NwbError, match="capacity"):
        ctx.sample_stable_configs(1, 0, 20000, cap=10)


def test_sampler_ranks_append_into_one_database(oracle):
    """Sharded generation with the gather fused into the sampler, emulated on one GPU: two 'ranks' (contexts) sample
    their halves of the index range straight into ONE set of output arrays with a shared slot counter (across
    processes the arrays are CUDA-IPC mappings of rank 0's memory - bench_multi.py runs that path)."""
    import torch
    seed, count = 20121005, 30000
    cap = count // 4
    rows = torch.zeros((cap, 21), dtype=torch.float64, device="cuda")
    idx = torch.zeros(cap, dtype=torch.int64, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctxs = [nwb.Context(), nwb.Context()]
    for r, c in enumerate(ctxs):
        lo, hi = r * count // 2, (r + 1) * count // 2
        c._check(c.lib.nwb_sample_stable_configs_dev(c.h, seed, lo, hi - lo, 5, rows.data_ptr(), idx.data_ptr(), cap,
                                                     cnt.data_ptr(), None), "nwb_sample_stable_configs_dev")
    for c in ctxs:
        c.synchronize()
    n = int(cnt.item())
    order = torch.argsort(idx[:n])
    ref_rows, ref_idx, _ = ctxs[0].sample_stable_configs(seed, 0, count)
    assert n == len(ref_idx) and (idx[:n][order].cpu().numpy() == ref_idx.astype(np.int64)).all()
    assert (rows[:n][order].cpu().numpy() == ref_rows).all()
    for c in ctxs:
        c.close()


def test_generate_ds_database_service(ctx, oracle, tmp_path):
    """generate_ds_database (nao_planner_control.cpp:137-147) with the reference's default budget
    (5000 samples, 5 IK attempts): file written in the reference format, loadable, rows as the oracle's."""
    path = tmp_path / "ssc_double_support.dat"
    n = ctx.generate_ds_database(5000, 5, path, seed=20121005)
    rows = nwb.load_ds_database(path)
    assert n == len(rows) and 20 < n < 2500
    orows, oidx, _ = oracle.sample_stable_configs(20121005, 0, 5000)
    assert abs(len(orows) - n) <= 1
    if len(orows) == n:
        assert np.abs(rows - orows).max() < 1e-5          # the file keeps 6 significant digits
    valid, _, _ = ctx.isFeasible(rows)
    assert valid.mean() > 0.99                            # rounding to 6 digits may push a boundary row over
    # the generated database drives a planning query
    res, paths = ctx.solveQuery(rows[:4], rows[4:8], rows, np.arange(4, dtype=np.uint64), W_APPROACH, max_iter=200)
    assert sum(r["success"] for r in res) >= 3
