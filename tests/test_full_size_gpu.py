"""GPU parity at BASELINE.json's FULL sizes against the CPU oracle on all host threads.

  V1 / V1b / V1c / V1d at 2^20 configurations, scenes S0 (empty: what the shipped planner checks) and S1 (shelf)
  nearest neighbour: 10^6 nodes x 1024 queries
  edge checks: 4096 edges x 102 way-points

Bar (north_star): validity bit-exact outside the 1e-5 m band around a contact / support-polygon boundary (the in-band
count is printed), NN indices exact, first colliding way-point exact outside the band.  Inputs: SURVEY.md section 8d.
"""
import numpy as np
import pytest

import nao_wholebody_planning_b200 as nwb
import oracle_lib as OL

pytestmark = pytest.mark.gpu
BAND = 1e-5
B = 1 << 20


@pytest.fixture(scope="module")
def ctx():
    c = nwb.Context()
    yield c
    c.close()


@pytest.fixture(scope="module")
def inputs(ctx):
    """V1 (uniform), V1b (LHipYawPitch = 0, SCG:291), V1c (V1b after the double-support projection RP:498-593;
    rows the projection rejects keep their V1b values), V1d (database rows + N(0, 0.05^2) noise)."""
    v1 = OL.uniform_configs(B, seed=20121001)
    v1b = OL.uniform_configs(B, seed=20121001, lhyp_zero=True)
    out, st, ds = ctx.steer_project(np.zeros_like(v1b), v1b, step=1e9)      # one step that REACHES q_rand, then the projection
    ok = (st != nwb.TRAPPED) & np.isfinite(out).all(1)
    v1c = np.where(ok[:, None], out, v1b)
    v1d = OL.noisy_db_configs(B, seed=20121004)
    return {"V1": v1, "V1b": v1b, "V1c": v1c, "V1d": v1d, "V1c_projected": int(ok.sum())}


@pytest.mark.parametrize("tag", ["V1", "V1d"])
def test_validity_full_batch_in_a_256_box_scene(oracle, inputs, tag):
    """2^20 configurations against 256 generated boxes (bounding-volume hierarchy): the verdict-only work-queue kernel
    against the oracle, which tests every box, and against the per-thread walk of the margins path."""
    q = inputs[tag]
    boxes = OL.generated_scene(256)
    o_valid, o_reason, o_marg = oracle.is_feasible(q, boxes=boxes, threads=oracle.threads)
    band = (np.abs(o_marg) < BAND).any(1)
    with nwb.Context() as c:
        c.scene_add_boxes(boxes)
        valid, reason, _ = c.isFeasible(q)
        valid_m, reason_m, marg = c.isFeasible(q, want_margins=True)
    bad = ((valid != o_valid) | (reason != o_reason)) & ~band
    print(f"\n{tag}/256 boxes: n={len(q)} valid={o_valid.mean():.5f} collision={(o_reason & 1).mean():.4f} "
          f"in-band={int(band.sum())} mismatches-outside-band={int(bad.sum())}")
    err = np.abs(marg - o_marg)
    print(f"margins: max |product - oracle| = {err.max():.3e}, 99.99th percentile {np.quantile(err, 0.9999):.3e}")
    assert bad.sum() == 0
    assert (valid == valid_m).all() and (reason == reason_m).all()
    # fp32 product against the fp64 oracle: a separation is sqrt(d^2) - r, so where two axes all but cross (d^2 -> 0) the
    # root amplifies the fp32 rounding of d^2; over 10^6 configurations a handful of such pairs occur
    assert np.quantile(err, 0.9999) < 3e-6 and err.max() < 1e-3


@pytest.mark.parametrize("scene", ["S0", "S1"])
@pytest.mark.parametrize("tag", ["V1", "V1b", "V1c", "V1d"])
def test_validity_full_batch_against_oracle(oracle, inputs, tag, scene):
    q = inputs[tag]
    boxes = OL.scene_boxes(scene) if scene != "S0" else None
    o_valid, o_reason, o_marg = oracle.is_feasible(q, boxes=boxes, threads=oracle.threads)
    band = (np.abs(o_marg) < BAND).any(1)
    with nwb.Context() as c:
        if boxes is not None:
            c.scene_add_boxes(boxes)
        valid, reason, _ = c.isFeasible(q)
    bad = ((valid != o_valid) | (reason != o_reason)) & ~band
    flips_in_band = int(((valid != o_valid) & band).sum())
    print(f"\n{tag}/{scene}: n={len(q)} valid={o_valid.mean():.5f} collision={(o_reason & 1).mean():.4f} "
          f"unstable={((o_reason & 2) != 0).mean():.4f} in-band={int(band.sum())} flips-in-band={flips_in_band} "
          f"mismatches-outside-band={int(bad.sum())}")
    assert bad.sum() == 0
    if tag == "V1c":
        assert o_valid.mean() > 0.01             # the projection makes the valid fraction non-trivial


def test_v1c_projection_against_oracle(ctx, oracle, inputs):
    """The generator of V1c itself, on 2^17 rows: same projected rows and status as the oracle's
    generate_q_new + enforce_DS_Constraints."""
    n = 1 << 17
    v1b = inputs["V1b"][:n]
    out, st, ds = ctx.steer_project(np.zeros_like(v1b), v1b, step=1e9)
    o_out, o_st, o_ds = oracle.steer_project(np.zeros_like(v1b), v1b, np.ones(21), 1e9, threads=oracle.threads)
    assert (st == o_st).mean() > 0.9999 and (ds == o_ds).mean() > 0.9999
    same = (st == o_st) & (st != nwb.TRAPPED)
    assert np.abs(out[same] - o_out[same]).max() < 1e-8
    print(f"\nV1c: projected {inputs['V1c_projected']} of {B} rows")


def test_nn_million_nodes_1024_queries(ctx, oracle, inputs):
    """V2: nodes = first 10^6 rows of V1c, queries = 1024 database rows drawn as RP:319-320."""
    nodes = np.ascontiguousarray(inputs["V1c"][:1_000_000])
    db = OL.golden_db()
    rng = np.random.default_rng(20121002)
    q = db[np.floor(rng.random(1024) * (len(db) - 1)).astype(int)]
    t = nwb.Tree(ctx, len(nodes))
    t.add(nodes)
    idx, dist = t.findNearestNeighbour(q)
    i1 = np.array([t.findNearestNeighbour(q[k:k + 1])[0][0] for k in range(0, 1024, 64)])      # the single-query form
    t.close()
    o_idx, o_dist = oracle.nearest_neighbour(nodes, q, threads=oracle.threads)
    assert (idx == o_idx).all() and np.abs(dist - o_dist).max() < 1e-12
    assert (i1 == o_idx[::64]).all()
    for n in (10_000, 100_000):                   # the smaller trees of the sweep, all 1024 queries
        t = nwb.Tree(ctx, n)
        t.add(nodes[:n])
        idx, dist = t.findNearestNeighbour(q)
        t.close()
        o_idx, o_dist = oracle.nearest_neighbour(nodes[:n], q, threads=oracle.threads)
        assert (idx == o_idx).all() and np.abs(dist - o_dist).max() < 1e-12


@pytest.mark.parametrize("scene", ["S0", "S1"])
def test_edge_checks_4096_x_102(oracle, inputs, scene):
    """V2: shortcut-style edge checks, 4096 random node pairs x 102 way-points, collision only (RP:1876)."""
    nodes = inputs["V1c"]
    rng = np.random.default_rng(20121002)
    pairs = rng.integers(0, 200_000, (4096, 2))
    qa, qb = nodes[pairs[:, 0]], nodes[pairs[:, 1]]
    db = OL.golden_db()
    dbp = rng.integers(0, len(db), (2048, 2))
    qa = np.concatenate([qa, db[dbp[:, 0]]])          # + 2048 database-row edges, where many edges are collision-free
    qb = np.concatenate([qb, db[dbp[:, 1]]])
    boxes = OL.scene_boxes(scene) if scene != "S0" else None
    with nwb.Context() as c:
        if boxes is not None:
            c.scene_add_boxes(boxes)
        ok, fb = c.isPathValid(qa, qb, 100)
    ook, ofb = oracle.is_path_valid(qa, qb, 100, boxes=boxes, threads=oracle.threads)
    bad = np.nonzero((ok != ook) | (fb != ofb))[0]
    for e in bad:                                      # a way-point within 1e-5 m of contact may flip, nothing else
        way = oracle.interpolate(qa[e], qb[e], 100)
        _, sep = oracle.is_state_colliding(way, boxes=boxes, group_only=False)
        cand = [x for x in (fb[e], ofb[e]) if x >= 0]
        assert min(abs(sep[c]) for c in cand) < BAND
    print(f"\nedges/{scene}: {len(qa)} edges, collision-free {int(ook.sum())}, in-band disagreements {len(bad)}")
    assert len(bad) <= 12
