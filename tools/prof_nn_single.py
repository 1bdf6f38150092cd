#!/usr/bin/env python3
"""Where the fixed cost of a single-query nearest-neighbour call sits: event-pair floor (empty tree), per-call event
times, and K back-to-back calls inside ONE event pair."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import torch  # noqa: E402

import nao_wholebody_planning_b200 as nwb  # noqa: E402
import oracle_lib as OL  # noqa: E402

ctx = nwb.Context()
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
q = torch.from_numpy(OL.golden_db()[:1].copy()).cuda()
idx = torch.empty(1, dtype=torch.int32, device="cuda")
dist = torch.empty(1, dtype=torch.float64, device="cuda")
nodes = OL.uniform_configs(1_000_000, seed=20121001, lhyp_zero=True)
sizes = [int(a) for a in sys.argv[1:]] or [0, 10_000, 100_000, 200_000, 300_000, 1_000_000]
for n in sizes:
    tree = nwb.Tree(ctx, max(n, 1))
    if n:
        tree.add(nodes[:n])
    torch.cuda.synchronize()
    ts = []
    for _ in range(12):
        tree.nearest_dev(q.data_ptr(), 1, idx.data_ptr(), dist.data_ptr())
        ts.append(ctx.last_kernel_ms()[0] * 1e3)
    K = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(K):
        tree.nearest_dev(q.data_ptr(), 1, idx.data_ptr(), dist.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    print(f"nodes {n:8d}: per-call event pair median {np.median(ts[2:]):7.2f} us (min {min(ts):.2f}); "
          f"{K} back-to-back calls: {e0.elapsed_time(e1) * 1e3 / K:7.2f} us per call")
    tree.close()
