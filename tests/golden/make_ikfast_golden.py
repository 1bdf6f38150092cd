#!/usr/bin/env python3
"""Generate tests/golden/ikfast_vectors.npz by RUNNING the reference's own right-arm solver.

The solver (rrt_connect_planner/src/ikfast_TranslationDirection5D.cpp) is the one file of the
reference's hot path that compiles standalone; `make -C oracle ref` builds it where it lies into
oracle/_ref/libikfast_ref.so (nothing is copied).  This script feeds it seeded targets and stores its
answers so that the CPU tests can pin oracle/arm_restated.h without /root/reference being present.

  targets   [N][6]     hand position + unit direction in the solver's torso frame (from its own fk)
  joints    [N][5]     the joint vector the target was generated from
  count     [N]        number of solutions the reference solver returned
  solutions [N][8][5]  its solutions in its own order (NaN padded)
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
LO = np.array([-2.0857, -1.3265, -2.0857, 0.0349, -1.8238])   # right-arm limits, planner joints 10..14
HI = np.array([2.0857, 0.3142, 2.0857, 1.5446, 1.8238])


def main(n=400, seed=20121006):
    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle"), "ref"], check=True)
    lib = C.CDLL(str(ROOT / "oracle" / "_ref" / "libikfast_ref.so"))
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    rng = np.random.default_rng(seed)
    targets, joints, count, sols = np.zeros((n, 6)), np.zeros((n, 5)), np.zeros(n, np.int32), np.full((n, 8, 5), np.nan)
    for i in range(n):
        q = rng.uniform(LO, HI) if i % 4 else rng.uniform(-3.1, 3.1, 5)   # every 4th target outside the limits
        pos, d, out = np.zeros(3), np.zeros(3), np.zeros((16, 5))
        lib.ikf_ref_fk(P(q), P(pos), P(d))
        k = lib.ikf_ref_solve(P(pos), P(d), P(out), 16)
        assert k <= 8
        targets[i, :3], targets[i, 3:], joints[i], count[i] = pos, d, q, k
        sols[i, :k] = out[:k]
    np.savez_compressed(Path(__file__).resolve().parent / "ikfast_vectors.npz", targets=targets, joints=joints, count=count,
                        solutions=sols)
    print("wrote", n, "targets; solution-count histogram", np.bincount(count))


if __name__ == "__main__":
    main()
