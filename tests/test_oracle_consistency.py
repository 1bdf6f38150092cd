"""Second, independent check of the C++ oracle: a NumPy evaluation of the same authored model
(tools/np_model.py) and brute-force geometry.  CPU only."""
import numpy as np

import np_model as M
import oracle_lib as OL


def test_fk_com_match_numpy_model(oracle):
    model = M.load_model()
    q = OL.uniform_configs(300, seed=7)
    poses, com = oracle.fk(q)
    fr = M.fk(model, q)
    for i, f in enumerate(model["frames"]):
        R, p = fr[f["name"]]
        assert np.abs(poses[:, i, :, :3] - R).max() < 1e-12
        assert np.abs(poses[:, i, :, 3] - p).max() < 1e-12
    assert np.abs(com - M.com(model, fr)).max() < 1e-12


def test_margins_match_numpy_model(oracle):
    model = M.load_model()
    q = np.concatenate([OL.uniform_configs(400, seed=11), OL.noisy_db_configs(400, seed=12)])
    valid, reason, margins = oracle.is_feasible(q)
    fr = M.fk(model, q)
    sep = M.pair_separations(model, fr, M.enabled_pairs("enabled_group")).min(1)
    assert np.abs(margins[:, 0] - sep).max() < 1e-9
    stab = M.stability_margin(model, fr)
    inside = stab > 0
    # inside: distance to the nearest edge line (identical definitions); outside: the oracle reports
    # the true distance to the hull, numpy the most violated edge line -> only the sign is compared
    assert np.abs(margins[inside, 1] - stab[inside]).max() < 1e-9
    assert ((margins[:, 1] > 0) == inside).all()
    assert (((reason & 1) != 0) == (sep < 0)).all()
    assert (((reason & 2) != 0) == ~inside).all()
    assert (valid == ((sep >= 0) & inside)).all()
    assert 0.02 < valid.mean() < 0.6      # the mix has both outcomes


def test_pairs_all_adds_only_the_fixed_head_torso_pair(oracle):
    q = OL.uniform_configs(200, seed=3)
    _, sep_group = oracle.is_state_colliding(q, group_only=True)
    _, sep_all = oracle.is_state_colliding(q, group_only=False)
    assert (sep_all <= sep_group + 1e-15).all()
    # HeadPitch_link - torso is rigid (head joints stay at 0): constant separation
    model = M.load_model()
    fr = M.fk(model, q[:5])
    s = M.pair_separations(model, fr, [["torso", "HeadPitch_link"]])
    assert np.ptp(s) < 1e-12 and s.min() > 0.005


def test_seg_seg_against_sampling(oracle):
    rng = np.random.default_rng(5)
    t = np.linspace(0, 1, 201)
    for _ in range(200):
        p1, q1, p2, q2 = rng.normal(0, 0.1, (4, 3))
        if rng.random() < 0.2:
            q1 = p1.copy()          # degenerate: sphere
        if rng.random() < 0.2:
            q2 = p2 + 1e-3 * (q1 - p1)   # nearly parallel
        a = p1 + t[:, None] * (q1 - p1)
        b = p2 + t[:, None] * (q2 - p2)
        brute = np.sqrt(((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)).min()
        d = oracle.seg_seg(p1, q1, p2, q2)
        assert d <= brute + 1e-12 and brute - d < 2e-3
        assert abs(d - M.seg_seg_dist(p1[None], q1[None], p2[None], q2[None])[0]) < 1e-12


def test_seg_box_against_sampling(oracle):
    rng = np.random.default_rng(6)
    t = np.linspace(0, 1, 4001)
    for _ in range(300):
        a, b = rng.normal(0, 0.2, (2, 3))
        c = rng.normal(0, 0.1, 3)
        size = rng.uniform(0.02, 0.3, 3)
        ang = rng.uniform(-1, 1)
        R = M.rot_axis([0, 0, 1.0], np.array([ang]))[0] if rng.random() < 0.5 else np.eye(3)
        box = np.concatenate([c, size, R.reshape(-1)])
        pts = a + t[:, None] * (b - a)
        loc = (pts - c) @ R
        ex = np.maximum(np.abs(loc) - size / 2, 0)
        brute = np.sqrt((ex ** 2).sum(1)).min()
        d = oracle.seg_box(a, b, box)
        assert d <= brute + 1e-12
        assert brute - d < 1e-4
