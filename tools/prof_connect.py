"""Where the wall time of nwb_connect_batch goes: the same call with a fresh (untouched) output array and with one whose
pages are already resident."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nao_wholebody_planning_b200 as nwb  # noqa: E402
import bench_sections as BS  # noqa: E402
import oracle_lib as OL  # noqa: E402
from nao_wholebody_planning_b200.context import _p  # noqa: E402


def main():
    c = nwb.Context()
    nodes_all, _ = BS.v1c_nodes(nwb, c, OL, 100_000)
    queries = BS.db_queries(OL, 1024)
    tree = nwb.Tree(c, 100_000)
    tree.add(nodes_all)
    idx, _ = tree.findNearestNeighbour(queries)
    q_start = np.ascontiguousarray(nodes_all[idx])
    queries = np.ascontiguousarray(queries)
    w = np.ones(21)
    w[15] = 0.0
    n, max_steps = 1024, 128
    added, st = np.empty(n, np.int32), np.empty(n, np.int32)
    c.connect(q_start, queries, w, 0.1, max_steps)
    for label in ("fresh zeros", "same array again", "same array again", "fresh zeros"):
        if label.startswith("fresh"):
            nodes = np.zeros((n, max_steps, 21))
        t0 = time.perf_counter()
        rc = c.lib.nwb_connect_batch(c.h, _p(q_start), _p(queries), n, _p(w), 0.1, max_steps, _p(nodes), _p(added), _p(st))
        dt = time.perf_counter() - t0
        print(f"{label:18s} rc={rc} wall {dt * 1e3:.3f} ms  kernel {c.last_kernel_ms()[0]:.3f} ms  rows {int(added.sum())}")


if __name__ == "__main__":
    main()
