"""Pin the CPU oracle against the reference's own data files.

The reference has no tests; its shipped .dat files and the logged support polygon are the only
known answers for this path.  These run on CPU (`-m "not gpu"`).
"""
import numpy as np
import pytest

import oracle_lib as OL
from oracle_lib import GOLDEN, golden_db, load_dat

CD = GOLDEN / "config_database"


def legs_config(q):
    return np.concatenate([-q[:5], q[16:21]])


def test_philox_known_answers(oracle):
    # Random123 known-answer vectors for philox4x32-10
    assert list(oracle.philox([0, 0, 0, 0], [0, 0])) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert list(oracle.philox([0xffffffff] * 4, [0xffffffff] * 2)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert list(oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])) == [
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_E1_legs_fk_puts_left_sole_at_target(oracle):
    db = golden_db()
    assert db.shape == (463, 21)
    worst_p, worst_r = 0.0, 0.0
    for q in db:
        f = oracle.chain_fk(2, legs_config(q))
        worst_p = max(worst_p, np.abs(f[:, 3] - [0, 0.1, 0]).max())
        worst_r = max(worst_r, np.abs(f[:, :3] - np.eye(3)).max())
        assert oracle.check_left_leg_pose(legs_config(q))
    assert worst_p < 1.1e-3          # observed 1.04e-3 (= IK eps)
    assert worst_r < 1.1e-3


def test_E1_full_body_fk_agrees_with_kdl_chain(oracle):
    db = golden_db()
    poses, _ = oracle.fk(db)
    lsole = poses[:, 28]
    for i, q in enumerate(db):
        f = oracle.chain_fk(2, legs_config(q))
        assert np.abs(lsole[i] - f).max() < 1e-12
    # hip frame of the KDL right-leg chain = torso frame shifted down by HipOffsetZ
    for i in (0, 100, 462):
        hip = oracle.chain_fk(0, -db[i, :5])
        torso = poses[i, 7]
        assert np.abs(torso[:, :3] - hip[:, :3]).max() < 1e-12
        assert np.abs(torso[:, 3] - (hip[:, 3] + hip[:, :3] @ [0, 0, 0.085])).max() < 1e-12


def test_E2_sampler_ik_reproduces_left_leg_columns(oracle):
    db = golden_db()
    err, iters = [], []
    for q in db:
        ok, sol, it = oracle.left_leg_ik_sampler(-q[:5], 5)
        assert ok
        err.append(np.abs(sol - q[16:21]).max())
        iters.append(it)
    err = np.array(err)
    assert np.median(err) < 5e-6       # observed 2.5e-6 (file precision: 6 significant digits)
    assert err.max() < 2.5e-5          # observed 1.6e-5
    assert (err < 1e-5).mean() > 0.97
    assert max(iters) <= 25 and np.median(iters) <= 4


def test_E3_lhipyawpitch_is_zero_everywhere():
    n = 0
    for f in sorted(CD.glob("*.dat")):
        d = load_dat(f)
        assert (d[:, 15] == 0).all(), f.name
        n += len(d)
    assert n == 1425


@pytest.mark.parametrize("name", ["ssc_double_support_backup.dat", "ssc_double_support_current.dat",
                                  "ssc_double_support_with_wrong_KDL_IK.dat"])
def test_E4_negative_fixtures_fail_double_support(oracle, name):
    d = load_dat(CD / name)
    worst = max(np.abs(oracle.chain_fk(2, legs_config(q))[:, 3] - [0, 0.1, 0]).max() for q in d)
    assert worst > 0.16                       # observed 0.169 - 0.184


def test_E5_goal_first_hand_pose(oracle):
    q = load_dat(CD / "ssc_double_support_goal_first.dat")[0]
    sol = load_dat(GOLDEN / "solutions" / "solution_path_first.dat")
    assert np.allclose(q, sol[1], atol=1e-6)                       # = row 2 of the approach solution
    f = oracle.right_hand_pose(q)
    assert np.abs(f[:, 3] - [0.23, -0.065, 0.2]).max() < 1.5e-3    # request of EXE:542-547
    assert np.abs(f[:, 2] - [0, -1, 0]).max() < 5e-3               # hand z-axis = requested direction


def test_E6_E7_goal_files_share_one_hand_pose(oracle):
    expect = {
        "ssc_double_support_goal_second.dat": ([0.1708, -0.065, 0.3005], [0, -1, 0]),
        "ssc_double_support_goal_first_drawer.dat": ([0.2008, -0.0651, 0.2003], [0, -1, 0]),
        "ssc_double_support_goal_second_drawer.dat": ([0.103, -0.0647, 0.199], [0, -1, 0]),
        "ssc_double_support_goal_first_door.dat": ([0.2702, 0.0167, 0.2775], [0, 0, -1]),
        "ssc_double_support_goal_second_door.dat": ([0.131, -0.026, 0.277], [0, 0, -1]),
    }
    for name, (pos, direction) in expect.items():
        d = load_dat(CD / name)
        P = np.array([oracle.right_hand_pose(q)[:, 3] for q in d])
        Z = np.array([oracle.right_hand_pose(q)[:, 2] for q in d])
        assert np.abs(P.mean(0) - pos).max() < 2e-3, name
        assert P.std(0).max() < 2e-3, name
        assert np.abs(Z.mean(0) - direction).max() < 2e-2, name


def test_E8_scenario3_manipulation_path(oracle):
    d = load_dat(GOLDEN / "solutions" / "solution_path_second.dat")
    assert d.shape == (106, 21)                 # 22 nodes * 5 - 4 rows (RP:1397,1413)
    P = np.array([oracle.right_hand_pose(q)[:, 3] for q in d])
    Z = np.array([oracle.right_hand_pose(q)[:, 2] for q in d])
    assert abs(P[0, 0] - 0.2391) < 1e-3 and abs(P[-1, 0] - 0.1409) < 1e-3
    # the drawer line: y, z constant within the reference's own hand tolerance (AOC:299-303: 5 mm)
    assert np.abs(P[:, 1] + 0.0653).max() < 1.5e-3 and np.abs(P[:, 2] - 0.3005).max() < 1.5e-3
    assert np.median(np.abs(P[:, 1] + 0.0653)) < 2e-4
    assert np.abs(Z - [0, -1, 0]).max() < 1e-2        # 0.007 rad tolerance of AOC:303
    assert np.abs(d[:, 5:10] - [1.5708, 0.17, 0, -0.036, 0]).max() < 1e-6      # left arm frozen
    for q in d:
        assert np.abs(oracle.chain_fk(2, legs_config(q))[:, 3] - [0, 0.1, 0]).max() < 4e-4


def test_E9_crouched_init_pose():
    sol = load_dat(GOLDEN / "solutions" / "solution_path_first.dat")
    assert sol.shape == (2, 21)
    init_pose = [0, -0.349066, 0.698132, -0.436332, 0, 1.39626, 0.349066, -1.39626, -0.5, 0,
                 1.39626, -0.349066, 1.39626, 0.5, 0, 0, 0, -0.436332, 0.698132, -0.349066, 0]   # EXE:323-343
    assert np.abs(sol[0] - init_pose).max() < 1e-6


def test_E10_support_hull_matches_logged_polygon(oracle):
    log = np.loadtxt(GOLDEN / "polygon_log.txt")
    assert np.allclose(log[0], log[-1])
    log = log[:-1]
    area = 0.5 * np.sum(log[:, 0] * np.roll(log[:, 1], -1) - np.roll(log[:, 0], -1) * log[:, 1])
    assert abs(area + 0.03032) < 1e-4            # clockwise
    # a perfect double-support stance at scale 1.0
    q = np.zeros(21)
    hull = oracle.support_hull(q, OL.default_params(scale_sp=1.0))
    assert len(hull) == 20
    harea = 0.5 * np.sum(hull[:, 0] * np.roll(hull[:, 1], -1) - np.roll(hull[:, 0], -1) * hull[:, 1])
    assert harea < 0 and abs(harea - area) < 2e-4
    d = np.array([np.hypot(*(hull - v).T).min() for v in log])
    assert d.max() < 3e-4                        # mirror residual of the logged pose (2.8e-4)


def test_E12_everything_the_reference_shipped_as_valid_is_valid(oracle):
    """All 679 shipped rows passed the reference's isFeasible (RP:1275, SCG:146): collision-free AND statically
    stable at scale 0.8.  The foot centre of the model is calibrated on exactly this (models/…json foot_center_doc);
    the gate is 100 %, and the heel and toe edges of the hull are pinned from both sides: the window the rows leave
    for the scale centre is 2.2 mm wide."""
    q = OL.golden_valid_rows()
    assert len(q) == 679
    valid, reason, margins = oracle.is_feasible(q)
    assert (reason & 1).sum() == 0                                # no proxy collision on any shipped row
    assert margins[:, 0].min() > 1e-3
    assert (reason & 2).sum() == 0 and valid.all()                # every shipped row is statically stable at 0.8
    assert margins[:, 1].min() > 1e-4                             # hull margin of the worst row (0.22 mm)
    col, sep = oracle.is_state_colliding(q, group_only=False)     # whole-robot pair set of isPathValid
    assert col.sum() == 0
    # the main database alone
    v463, r463, _ = oracle.is_feasible(golden_db())
    assert v463.all() and (r463 == 0).all()
    # two-sided: moving the scale centre 2 mm either way along x pushes shipped rows over the heel or the toe edge,
    # and so does the round-1 reading (mean of the 12 polygon vertices: 22 rows out)
    _, r_hull, _ = oracle.is_feasible(q, OL.default_params(scale_origin=2))
    assert ((r_hull & 2) != 0).sum() == 22
    _, r_sole, _ = oracle.is_feasible(q, OL.default_params(scale_origin=0))
    assert ((r_sole & 2) != 0).sum() == 6


def test_pair_counts(oracle):
    assert oracle.lib.nwbo_num_pairs(1, 1) == 108 and oracle.lib.nwbo_num_pairs(1, 0) == 109
    assert oracle.lib.nwbo_num_pairs(0, 1) == 111 and oracle.lib.nwbo_num_pairs(0, 0) == 112
