"""NumPy evaluation of the authored NAO model (float64, vectorised over a batch).

Used by tools/calibrate_model.py to choose proxy radii / check the mass table against the
golden configurations, and by the CPU tests as a second, independent check of the C++ oracle.
It is tooling, not product: nothing in nao_wholebody_planning_b200/ imports it.
"""
import json
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
MODEL_JSON = ROOT / "nao_wholebody_planning_b200" / "models" / "nao_h25_rfoot.json"
GOLDEN = ROOT / "tests" / "golden"


def load_model(path=MODEL_JSON):
    return json.loads(Path(path).read_text())


def load_dat(path):
    rows = []
    for line in Path(path).read_text().splitlines():
        v = [float(x) for x in line.split()]
        if v:
            rows.append(v)
    return np.array(rows, dtype=np.float64)


def rot_axis(axis, ang):
    """Rodrigues rotation, batch of angles -> [B,3,3]."""
    if axis == "x":
        a = np.array([1.0, 0, 0])
    elif axis == "y":
        a = np.array([0, 1.0, 0])
    elif axis == "z":
        a = np.array([0, 0, 1.0])
    else:
        a = np.asarray(axis, dtype=np.float64)
    c, s = np.cos(ang), np.sin(ang)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    eye = np.eye(3)
    return (c[:, None, None] * eye + s[:, None, None] * K
            + (1 - c)[:, None, None] * np.outer(a, a))


def fk(model, q):
    """q [B,21] -> dict name -> (R [B,3,3], p [B,3]) for every frame."""
    q = np.atleast_2d(np.asarray(q, dtype=np.float64))
    B = q.shape[0]
    out = []
    for f in model["frames"]:
        if f["parent"] < 0:
            R = np.tile(np.eye(3), (B, 1, 1))
            p = np.zeros((B, 3))
        else:
            Rp, pp = out[f["parent"]]
            R, p = Rp, pp + Rp @ np.asarray(f["pre"], dtype=np.float64)
            if f["axis"] is not None and f["joint"] >= 0:
                R = R @ rot_axis(f["axis"], f["sign"] * q[:, f["joint"]])
            p = p + R @ np.asarray(f["post"], dtype=np.float64)
        out.append((R, p))
    return {f["name"]: out[i] for i, f in enumerate(model["frames"])}


def com(model, frames):
    tot = 0.0
    acc = 0.0
    for m in model["masses"]:
        R, p = frames[m["link"]]
        acc = acc + m["mass"] * (p + R @ np.asarray(m["com"]))
        tot += m["mass"]
    return acc / tot


def scaled_foot_polygons(model, scale=None):
    """(right [n,2] in r_sole, left [n,2] in l_sole), scaled."""
    s = model["support_scale"] if scale is None else scale
    right = np.asarray(model["foot_polygon_right"], dtype=np.float64)
    left = right * np.array([1.0, -1.0])
    if model["support_scale_origin"] == "center":
        cr = np.asarray(model["foot_center_right"], dtype=np.float64)
        cl = cr * np.array([1.0, -1.0])
    elif model["support_scale_origin"] == "hull_centroid":
        cr, cl = right.mean(0), left.mean(0)
    else:
        cr = cl = np.zeros(2)
    return cr + s * (right - cr), cl + s * (left - cl)


def monotone_chain(pts):
    """Andrew monotone chain, returns hull CCW without the closing point."""
    pts = sorted(set(map(tuple, pts)))
    if len(pts) < 3:
        return np.array(pts)

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])
    lower, upper = [], []
    for p in pts:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], p) <= 0:
            lower.pop()
        lower.append(p)
    for p in reversed(pts):
        while len(upper) >= 2 and cross(upper[-2], upper[-1], p) <= 0:
            upper.pop()
        upper.append(p)
    return np.array(lower[:-1] + upper[:-1])


def support_points(model, frames, scale=None):
    """2-D support point set in the r_sole frame: [B, 2n, 2]."""
    right, left = scaled_foot_polygons(model, scale)
    Rl, pl = frames["l_sole"]
    l3 = np.concatenate([left, np.zeros((len(left), 1))], axis=1)      # [n,3]
    lw = np.einsum("bij,nj->bni", Rl, l3) + pl[:, None, :]
    B = Rl.shape[0]
    return np.concatenate([np.tile(right, (B, 1, 1)), lw[:, :, :2]], axis=1)


def stability_margin(model, frames, scale=None):
    """Signed distance of the projected COM to the support hull boundary (>0 inside)."""
    c = com(model, frames)
    pts = support_points(model, frames, scale)
    out = np.empty(len(c))
    for b in range(len(c)):
        h = monotone_chain(pts[b])
        if len(h) < 3:
            out[b] = -np.inf
            continue
        nxt = np.roll(h, -1, axis=0)
        e = nxt - h
        t = (e[:, 0] * (c[b, 1] - h[:, 1]) - e[:, 1] * (c[b, 0] - h[:, 0])) / np.hypot(e[:, 0], e[:, 1])
        out[b] = t.min()
    return out


def seg_seg_dist(p1, q1, p2, q2):
    """Closest distance between segments, batch [B,3] each (Ericson, clamped)."""
    d1, d2, r = q1 - p1, q2 - p2, p1 - p2
    a = (d1 * d1).sum(-1)
    e = (d2 * d2).sum(-1)
    f = (d2 * r).sum(-1)
    c = (d1 * r).sum(-1)
    b = (d1 * d2).sum(-1)
    eps = 1e-18
    denom = a * e - b * b
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.where(denom > eps, np.clip((b * f - c * e) / np.where(denom > eps, denom, 1), 0, 1), 0.0)
        t = np.where(e > eps, (b * s + f) / np.where(e > eps, e, 1), 0.0)
        # second segment degenerate (sphere): closest point of the first segment to that point
        s = np.where((e <= eps) & (a > eps), np.clip(-c / np.where(a > eps, a, 1), 0, 1), s)
        s = np.where(t < 0, np.where(a > eps, np.clip(-c / np.where(a > eps, a, 1), 0, 1), 0.0), s)
        s = np.where(t > 1, np.where(a > eps, np.clip((b - c) / np.where(a > eps, a, 1), 0, 1), 0.0), s)
        t = np.clip(t, 0, 1)
    d = r + d1 * s[:, None] - d2 * t[:, None]
    return np.sqrt((d * d).sum(-1))


def proxy_world(model, frames):
    """name -> (a [B,3], b [B,3], r)."""
    out = {}
    for pr in model["proxies"]:
        R, p = frames[pr["link"]]
        out[pr["link"]] = (p + R @ np.asarray(pr["a"]), p + R @ np.asarray(pr["b"]), pr["r"])
    return out


def pair_separations(model, frames, pairs):
    """[B, npairs] signed separation (distance - r1 - r2)."""
    pw = proxy_world(model, frames)
    cols = []
    for a, b in pairs:
        a1, b1, r1 = pw[a]
        a2, b2, r2 = pw[b]
        cols.append(seg_seg_dist(a1, b1, a2, b2) - r1 - r2)
    return np.stack(cols, axis=1)


def enabled_pairs(kind="enabled_group", trim=True):
    d = json.loads((GOLDEN / "srdf_pairs.json").read_text())
    return d["trim" if trim else "notrim"][kind]
