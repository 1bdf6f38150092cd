"""One launch of the fused RRT-Connect driver on a few hard queries (scene S3, 1500 iterations): the profile of a search
that runs out its budget.  Under ncu: -k regex:rrt_fused_kernel."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nao_wholebody_planning_b200 as nwb  # noqa: E402
import oracle_lib as OL  # noqa: E402


def main():
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    db = OL.golden_db()
    rng = np.random.default_rng(3)
    starts, goals = db[rng.integers(0, len(db), 256)], db[rng.integers(0, len(db), 256)]
    seeds = np.arange(256, dtype=np.uint64) + 1000
    w = np.ones(21)
    w[15] = 0.0
    with nwb.Context() as c:
        c.scene_add_boxes(OL.scene_boxes("S3"))
        res, _ = c.solveQuery(starts, goals, db, seeds, w, max_iter=500)
        hard = [k for k in range(256) if not res[k]["success"]][:nq]
        print("unsolved at 500 iterations:", len([k for k in range(256) if not res[k]["success"]]), "profiling", hard)
        t0 = time.perf_counter()
        res2, _ = c.solveQuery(starts[hard], goals[hard], db, seeds[hard], w, max_iter=1500)
        dt = time.perf_counter() - t0
        its = [r["iterations"] for r in res2]
        print(f"{len(hard)} queries, iterations {its}, wall {dt * 1e3:.1f} ms, kernel {c.last_kernel_ms()[0]:.1f} ms "
              f"-> {c.last_kernel_ms()[0] * 1e3 / max(its):.1f} us per iteration of the longest search")


if __name__ == "__main__":
    main()
