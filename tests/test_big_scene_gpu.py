"""GPU parity for scenes beyond the 16 inline boxes (north_star item 3): the boxes live in device
memory under a flattened bounding-volume hierarchy; leaves run the same exact capsule-box test, so every operator must
agree with the oracle, which simply tests every box."""
import numpy as np
import pytest

import nao_wholebody_planning_b200 as nwb
import oracle_lib as OL

pytestmark = pytest.mark.gpu
BAND = 1e-5
W = np.ones(21)
W[15] = 0.0


def configs(n, seed):
    return np.concatenate([OL.uniform_configs(n // 2, seed=seed), OL.noisy_db_configs(n - n // 2, seed=seed + 1)])


@pytest.mark.parametrize("n_boxes", [17, 64, 256, 1024])
def test_generated_scene_validity_and_margins(oracle, n_boxes):
    boxes = OL.generated_scene(n_boxes)
    q = configs(12000, 60 + n_boxes)
    o_valid, o_reason, o_marg = oracle.is_feasible(q, boxes=boxes, threads=oracle.threads)
    o_free = oracle.is_feasible(q, threads=oracle.threads)[1]
    band = (np.abs(o_marg) < BAND).any(1)
    with nwb.Context() as c:
        c.scene_add_boxes(boxes)
        assert c.num_boxes == n_boxes
        valid, reason, _ = c.isFeasible(q)
        assert (((valid == o_valid) & (reason == o_reason)) | band).all()
        valid_m, reason_m, marg = c.isFeasible(q, want_margins=True)
        assert (((valid_m == o_valid) & (reason_m == o_reason)) | band).all()
        assert np.abs(marg - o_marg).max() < 3e-6                # the branch-and-bound walk keeps the minimum exact
        col, sep = c.is_state_colliding(q, want_separation=True)
        o_col, o_sep = oracle.is_state_colliding(q, boxes=boxes, group_only=False)
        assert ((col == o_col) | (np.abs(o_sep) < BAND)).all() and np.abs(sep - o_sep).max() < 3e-6
        # one-by-one insertion (the way the reference's clients add objects, SIM:310-357) builds the same scene
        c.scene_clear()
        for b in boxes[:40]:
            c.scene_add_box(b[0:3], b[3:6], b[6:15])
        v40, r40, _ = c.isFeasible(q[:3000])
        ov40, or40, om40 = oracle.is_feasible(q[:3000], boxes=boxes[:40], threads=oracle.threads)
        assert (((v40 == ov40) & (r40 == or40)) | (np.abs(om40) < BAND).any(1)).all()
        c.scene_clear()
        assert c.num_boxes == 0
        v0, r0, _ = c.isFeasible(q)
        assert ((r0 == o_free) | band | (np.abs(oracle.is_feasible(q, threads=oracle.threads)[2]) < BAND).any(1)).all()
    added = ((o_reason & 1) != 0).sum() - ((o_free & 1) != 0).sum()
    print(f"\n{n_boxes} boxes: valid {o_valid.mean():.4f}, collisions added by the scene {int(added)}, in-band {int(band.sum())}")
    assert added > 0 and o_valid.sum() > 0


@pytest.mark.parametrize("n_boxes", [40, 300])
def test_work_queue_kernel_every_instantiation(oracle, n_boxes):
    """The verdict-only kernel of large scenes (hierarchy walk per thread, exact tests compacted through a CTA queue) in
    each of its forms: float32 input, the table pair list (srdf_trim = 0), the whole-robot collision predicate, a ragged
    last block, and against the per-thread walk that the margins path still uses."""
    boxes = OL.generated_scene(n_boxes, seed=21)
    q = configs(128 * 77 + 53, 900 + n_boxes)                   # ragged tail
    o_valid, o_reason, o_marg = oracle.is_feasible(q, boxes=boxes, threads=oracle.threads)
    band = (np.abs(o_marg) < BAND).any(1)
    with nwb.Context() as c:
        c.scene_add_boxes(boxes)
        valid, reason, _ = c.isFeasible(q)
        assert (((valid == o_valid) & (reason == o_reason)) | band).all()
        valid_m, reason_m, _ = c.isFeasible(q, want_margins=True)   # per-thread walk
        assert (valid == valid_m).all() and (reason == reason_m).all()
        qf = np.ascontiguousarray(q.astype(np.float32))
        vf, rf = np.empty(len(q), np.uint8), np.empty(len(q), np.uint8)
        c.is_feasible_host_buffers_f32(qf.ctypes.data, len(q), vf.ctypes.data, rf.ctypes.data)
        v64, r64, _ = c.isFeasible(qf.astype(np.float64))
        assert (vf == v64).all() and (rf == r64).all()
        col = c.is_state_colliding(q)[0]
        o_col, o_sep = oracle.is_state_colliding(q, boxes=boxes, group_only=False)
        assert ((col == o_col) | (np.abs(o_sep) < BAND)).all()
        assert (col == c.is_state_colliding(q, want_separation=True)[0]).all()
        # a single configuration and an empty batch
        v1, r1, _ = c.isFeasible(q[:1])
        assert v1[0] == valid[0] and r1[0] == reason[0]
        assert len(c.isFeasible(q[:0])[0]) == 0
    pn = OL.default_params(srdf_trim=0)
    on_valid, on_reason, on_marg = oracle.is_feasible(q, params=pn, boxes=boxes, threads=oracle.threads)
    bandn = (np.abs(on_marg) < BAND).any(1)
    with nwb.Context(srdf_trim=0) as c:
        c.scene_add_boxes(boxes)
        valid, reason, _ = c.isFeasible(q)
        assert (((valid == on_valid) & (reason == on_reason)) | bandn).all()
        col = c.is_state_colliding(q)[0]
        o_col, o_sep = oracle.is_state_colliding(q, params=pn, boxes=boxes, group_only=False)
        assert ((col == o_col) | (np.abs(o_sep) < BAND)).all()
    assert (on_reason != o_reason).any()                           # the two SRDF readings do differ on this batch


def test_every_reference_scene_object_at_once(oracle):
    """initScene.cpp:120-158,172-176,569-573 + the clients' objects (SIM:322-356, EXE:376-377, PAR:465-485): 21 boxes."""
    boxes = OL.scene_boxes("ALL")
    assert len(boxes) == 21
    q = configs(16000, 7)
    o_valid, o_reason, o_marg = oracle.is_feasible(q, boxes=boxes, threads=oracle.threads)
    band = (np.abs(o_marg) < BAND).any(1)
    with nwb.Context() as c:
        c.scene_add_boxes(boxes)
        valid, reason, marg = c.isFeasible(q, want_margins=True)
    assert (((valid == o_valid) & (reason == o_reason)) | band).all()
    assert np.abs(marg - o_marg).max() < 3e-6


def test_planner_operators_in_a_big_scene(oracle):
    """Edge checks, connect walks, the sampler and the RRT driver read the same hierarchy."""
    boxes = OL.generated_scene(64, seed=5)
    db = OL.golden_db()
    rng = np.random.default_rng(8)
    qa, qb = db[rng.integers(0, len(db), 400)], db[rng.integers(0, len(db), 400)]
    with nwb.Context() as c:
        c.scene_add_boxes(boxes)
        ok, fb = c.isPathValid(qa, qb, 100)
        ook, ofb = oracle.is_path_valid(qa, qb, 100, boxes=boxes, threads=oracle.threads)
        assert ((ok == ook) & (fb == ofb)).mean() > 0.98
        assert ook.sum() < oracle.is_path_valid(qa, qb, 100, threads=oracle.threads)[0].sum()
        nodes, added, st = c.connect(qa[:64], qb[:64], W, 0.1, 128)
        same = 0
        for k in range(64):
            o_st, o_nodes = oracle.connect_edge(qa[k], qb[k], W, 0.1, boxes=boxes)
            same += int(o_st == st[k] and len(o_nodes) == added[k])
        assert same >= 62
        rows, idx, stage = c.sample_stable_configs(20121005, 0, 20000)
        orows, oidx, ostage = oracle.sample_stable_configs(20121005, 0, 20000, boxes=boxes, threads=oracle.threads)
        assert (stage != ostage).sum() <= 2
        starts, goals = db[rng.integers(0, len(db), 12)], db[rng.integers(0, len(db), 12)]
        seeds = np.arange(12, dtype=np.uint64) + 300
        res, paths = c.solveQuery(starts, goals, db, seeds, W, max_iter=200)
        n_same = 0
        for k in range(12):
            oraw, ost = oracle.solve_query(starts[k], goals[k], db, W, int(seeds[k]), max_iter=200, boxes=boxes)
            n_same += int(res[k]["success"] == ost["success"] and res[k]["iterations"] == ost["iterations"] and len(paths[k]) == len(oraw))
        assert n_same >= 11
