#!/usr/bin/env python3
"""Profiling driver for the nearest-neighbour scans: builds a tree of N nodes and runs the scan for nq queries a few
times (ncu: `ncu --set full -k regex:nn_ --launch-skip 2 --launch-count 1 python tools/prof_nn.py 1000000 1024`)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import torch  # noqa: E402

import nao_wholebody_planning_b200 as nwb  # noqa: E402
import oracle_lib as OL  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
ctx = nwb.Context()
nodes = OL.uniform_configs(n, seed=20121001, lhyp_zero=True)
db = OL.golden_db()
q = db[np.floor(np.random.default_rng(20121002).random(max(nq, 1)) * (len(db) - 1)).astype(int)]
tree = nwb.Tree(ctx, n)
tree.add(nodes)
qd = torch.from_numpy(q).cuda()
idx = torch.empty(nq, dtype=torch.int32, device="cuda")
dist = torch.empty(nq, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    tree.nearest_dev(qd.data_ptr(), nq, idx.data_ptr(), dist.data_ptr())
    ts.append(ctx.last_kernel_ms()[0])
print(f"nodes {n} queries {nq}: ms per call {ts}")
