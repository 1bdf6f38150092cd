"""The exact rejection rule of the left-leg Newton-Raphson solve (csrc/planner_core.cuh, left_leg_target_feasible;
DESIGN.md 4.3) against the CPU oracle, which runs every solve to the end: a target the rule rejects must never be
one the oracle converges to.  The rule is restated here in numpy, formula for formula, so that the claim "exact" is
checked on every CPU run; the device code itself is covered by the GPU parity tests (verdicts vs oracle)."""
import numpy as np

import oracle_lib as OL

EPS = 1e-3                                   # nwb_params.ik_eps default (local_planner.cpp:41)
FOOT, LEGS = 0.04519, 0.1 + 0.1029


def target_of(hip):                          # getDesiredLeftLegPose (local_planner.cpp:59-110): hip^-1 * (0, 0.1, 0)
    R, p = hip[:, :3], hip[:, 3]
    return R.T, R.T @ (np.array([0.0, 0.1, 0.0]) - p)


def feasible(Rt, pt, eps=EPS):
    t = np.sqrt(3.0) * eps
    d = pt - np.array([0.0, 0.05, 0.0])
    if d @ d > (LEGS + FOOT + t + 1e-6) ** 2:
        return False, 1
    a = d + FOOT * Rt[:, 2]
    da = (1.0 + FOOT) * t
    if a @ a > (LEGS + da + 1e-6) ** 2:
        return False, 2
    uy, uz = -Rt[2, 0], Rt[1, 0]
    un = np.hypot(uy, uz)
    if un > t:
        r = t / un
        dn = r / np.sqrt(0.5 * (1.0 + np.sqrt(1.0 - r * r)))
        arm = min(LEGS, np.sqrt(a @ a) + da)
        if abs(uy * a[1] + uz * a[2]) > 1.1 * (un * (da + dn * arm)) + 1e-12:
            return False, 3
    return True, 0


def test_rejected_targets_are_never_converged_to(oracle):
    lo, hi = OL.joint_limits()
    rng = np.random.default_rng(20121005)
    n, converged, rejected, by_rule = 5000, 0, 0, [0, 0, 0, 0]
    for _ in range(n):
        q = lo + (hi - lo) * rng.random(21)
        rleg = -q[:5]                                        # SCG:307-313
        Rt, pt = target_of(oracle.chain_fk(0, rleg))
        ok_rule, why = feasible(Rt, pt)
        ok, _, iters = oracle.left_leg_ik_sampler(rleg)
        nr_converged = ok or iters < 100                     # ok also asks for the joint limits; iters < 100 = NR converged
        converged += nr_converged
        if not ok_rule:
            rejected += 1
            by_rule[why] += 1
            assert not nr_converged, (q, why, iters)
    failing = n - converged
    assert converged > 0.08 * n                              # ~12.6 % of uniform samples converge
    assert rejected > 0.8 * failing                          # the rule removes ~87 % of the failing solves
    assert by_rule[1] > 0.3 * failing and by_rule[3] > 0.3 * failing


def test_reachable_poses_pass_the_rule(oracle):
    """FK poses of the left leg itself (exactly reachable, any joint values) always pass, also when pushed to the
    edge of the convergence tolerance."""
    rng = np.random.default_rng(5)
    for _ in range(2000):
        ll = rng.uniform(-2.5, 2.5, 5)
        f = oracle.chain_fk(1, ll)
        assert feasible(f[:, :3], f[:, 3])[0]
        shift = rng.uniform(-1, 1, 3) * 0.999e-3             # every twist component below eps: NR would accept it
        w = rng.uniform(-1, 1, 3) * 0.999e-3
        if rng.random() < 0.5:                               # the corners of the tolerance box
            shift, w = np.sign(shift) * 0.999e-3, np.sign(w) * 0.999e-3
        th = np.linalg.norm(w)
        K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]) / th
        Rw = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K          # rotation vector w in the base frame
        assert feasible(Rw @ f[:, :3], f[:, 3] + shift)[0]
