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


def plane_terms(Rt, pt):
    """F = u . (A - c) and the gradient pieces of rule (3): u (w.r.t. -dp) and g (w.r.t. -w)."""
    a = pt - np.array([0.0, 0.05, 0.0]) + FOOT * Rt[:, 2]
    x, z = Rt[:, 0], Rt[:, 2]
    u = np.cross([1.0, 0.0, 0.0], x)
    g = FOOT * np.cross(z, u) + np.cross(x, np.cross(a, [1.0, 0.0, 0.0]))
    return float(u @ a), u, g, a


def feasible(Rt, pt, eps=EPS):
    t = np.sqrt(3.0) * eps
    d = pt - np.array([0.0, 0.05, 0.0])
    if d @ d > (LEGS + FOOT + t + 1e-6) ** 2:
        return False, 1
    a = d + FOOT * Rt[:, 2]
    da = (1.0 + FOOT) * t
    if a @ a > (LEGS + da + 1e-6) ** 2:
        return False, 2
    F0, u, g, _ = plane_terms(Rt, pt)
    B1 = eps * (np.abs(u).sum() + np.abs(g).sum())
    h = 0.51 * t * t
    R2 = t * t * (1.0 + FOOT) + h * (np.sqrt(a @ a) + da + FOOT * h) + (np.linalg.norm(u) + t) * FOOT * h
    if abs(F0) > 1.0001 * (B1 + R2) + 1e-12:
        return False, 3
    return True, 0


def rotation(w):
    th = np.linalg.norm(w)
    if th == 0:
        return np.eye(3)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]) / th
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def test_first_order_expansion_and_remainder_of_the_plane_rule():
    """F(target perturbed by (dp, w)) = F - u.dp - w.g up to the stated second-order bound: checked by evaluating F on
    perturbed frames, inside the tolerance box and at its corners."""
    rng = np.random.default_rng(11)
    t = np.sqrt(3.0) * EPS
    h = 0.51 * t * t
    worst = 0.0
    for _ in range(4000):
        R = rotation(rng.uniform(-3, 3, 3))
        p = rng.uniform(-0.25, 0.25, 3)
        F0, u, g, a = plane_terms(R, p)
        if np.linalg.norm(a) > 0.25:
            continue
        dp, w = rng.uniform(-1, 1, 3) * EPS, rng.uniform(-1, 1, 3) * EPS
        if rng.random() < 0.5:
            dp, w = np.sign(dp) * EPS, np.sign(w) * EPS
        F1 = plane_terms(rotation(-w) @ R, p - dp)[0]                     # the reachable frame: target = Exp(w) frame + dp
        lin = F0 - u @ dp - w @ g
        R2 = t * t * (1.0 + FOOT) + h * (np.linalg.norm(a) + (1.0 + FOOT) * t + FOOT * h) + (np.linalg.norm(u) + t) * FOOT * h
        assert abs(F1 - lin) <= R2
        assert abs(F1 - F0) <= EPS * (np.abs(u).sum() + np.abs(g).sum()) + R2
        worst = max(worst, abs(F1 - lin) / R2)
    assert worst > 0.05                                                    # the remainder bound is not absurdly loose


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
    assert rejected > 0.93 * failing                         # the rule removes ~96 % of the failing solves
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
        Rw = rotation(w)                                     # rotation vector w in the base frame
        assert feasible(Rw @ f[:, :3], f[:, 3] + shift)[0]
