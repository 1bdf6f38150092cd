#!/usr/bin/env python3
"""Gate the authored model against the reference's shipped configurations.

Prints, for the golden rows the reference judged valid (E12):
  * left-sole pose error of the full-body FK (E1),
  * static-stability margins at scale 0.8 (gate: >= 95 % stable),
  * per-pair minimum separation of the capsule proxies (gate: all > 0),
  * the double-support hull at scale 1.0 against polygon_log.txt (E10),
  * validity statistics of uniformly random configurations (how selective the proxies are).
"""
import sys

import numpy as np

sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent))
import np_model as M  # noqa: E402


def golden_valid_rows():
    rows = []
    cd = M.GOLDEN / "config_database"
    names = ["ssc_double_support.dat"] + sorted(
        p.name for p in cd.glob("ssc_double_support_goal_*.dat"))
    for n in names:
        d = M.load_dat(cd / n)
        if d.shape[1] == 22:      # goal files carry the ergonomics index in column 0
            d = d[:, 1:]
        rows.append(d)
    rows.append(M.load_dat(M.GOLDEN / "solutions" / "solution_path_first.dat"))
    rows.append(M.load_dat(M.GOLDEN / "solutions" / "solution_path_second.dat"))
    return np.concatenate(rows, axis=0)


def main():
    model = M.load_model()
    q = golden_valid_rows()
    print("golden valid rows:", q.shape)
    fr = M.fk(model, q)
    Rl, pl = fr["l_sole"]
    err = np.abs(pl - np.array(model["left_sole_target"])).max(1)
    print("E1 l_sole pos err: max %.3e  rows>2e-3: %d ; rot dev max %.3e" % (
        err.max(), (err > 2e-3).sum(), np.abs(Rl - np.eye(3)).max()))

    mar = M.stability_margin(model, fr)
    print("stability @0.8: stable %d / %d (%.1f %%), min margin %.4f, p5 %.4f" % (
        (mar > 0).sum(), len(mar), 100 * (mar > 0).mean(), mar.min(), np.percentile(mar, 5)))
    c = M.com(model, fr)
    print("COM x range [%.3f, %.3f], y range [%.3f, %.3f], z mean %.3f" % (
        c[:, 0].min(), c[:, 0].max(), c[:, 1].min(), c[:, 1].max(), c[:, 2].mean()))

    # window the shipped rows leave for the scale centre of the foot polygon (foot_center_right in the model file)
    import copy
    trial = copy.deepcopy(model)
    trial["support_scale_origin"] = "center"
    print("scale-centre window (rows unstable / min margin):")
    for cx in np.arange(0.0075, 0.0121, 0.0005):
        cells = []
        for cy in (-0.017, -0.010, -0.006, -0.002, 0.002):
            trial["foot_center_right"] = [float(cx), cy]
            mm = M.stability_margin(trial, fr)
            cells.append("%2d/%+.5f" % ((mm <= 0).sum(), mm.min()))
        print("   cx %.4f | cy -0.017 .. 0.002: %s" % (cx, "  ".join(cells)))

    for kind in ("enabled_group", "enabled_all"):
        pairs = M.enabled_pairs(kind)
        sep = M.pair_separations(model, fr, pairs)
        worst = sep.min(0)
        order = np.argsort(worst)
        print("%s: %d pairs; rows colliding %d; worst pairs:" % (kind, len(pairs), (sep.min(1) <= 0).sum()))
        for k in order[:12]:
            print("   %-22s %-22s min sep %+.4f  (rows<=0: %d)" % (
                pairs[k][0], pairs[k][1], worst[k], (sep[:, k] <= 0).sum()))

    # E10: hull at scale 1.0 vs the logged polygon
    log = np.loadtxt(M.GOLDEN / "polygon_log.txt")[:-1]
    right, left = M.scaled_foot_polygons(model, 1.0)
    pts = np.concatenate([right, left + np.array([0.0, 0.1])])
    hull = M.monotone_chain(pts)
    d = np.array([np.hypot(*(hull - v).T).min() for v in log])
    print("E10 hull vertices %d; max distance of a logged vertex to the model hull vertices %.2e" % (
        len(hull), d.max()))

    # selectivity on uniform random configurations (V1, V1b)
    rng = np.random.default_rng(20121001)
    lim = np.array(model["joint_limits"])
    B = 20000
    u = lim[:, 0] + (lim[:, 1] - lim[:, 0]) * rng.random((B, 21))
    for tag, qq in (("V1 uniform", u), ("V1b LHipYawPitch=0", np.where(np.arange(21) == 15, 0.0, u))):
        fr = M.fk(model, qq)
        sep = M.pair_separations(model, fr, M.enabled_pairs("enabled_group"))
        free = sep.min(1) > 0
        stab = M.stability_margin(model, fr) > 0
        print("%s: collision-free %.3f  stable %.3f  valid %.4f" % (
            tag, free.mean(), stab.mean(), (free & stab).mean()))


if __name__ == "__main__":
    main()
