#!/usr/bin/env python3
"""Single-query nearest neighbour over 10^6 nodes with FOUR trees rotating (336 MB of fp32 scan data between two visits
of a tree: every call reads HBM), 40 calls back to back between one event pair.  NWB_NN_NO_F16=1: the fp32 stream."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import torch  # noqa: E402

import nao_wholebody_planning_b200 as nwb  # noqa: E402
import oracle_lib as OL  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx = nwb.Context()
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
q = torch.from_numpy(OL.golden_db()[:1].copy()).cuda()
idx = torch.empty(1, dtype=torch.int32, device="cuda")
dist = torch.empty(1, dtype=torch.float64, device="cuda")
nodes = OL.uniform_configs(n, seed=20121001, lhyp_zero=True)
trees = []
for _ in range(4):
    t = nwb.Tree(ctx, n)
    t.add(nodes)
    trees.append(t)
ctx.set_kernel_timing(False)
for k in range(8):
    trees[k % 4].nearest_dev(q.data_ptr(), 1, idx.data_ptr(), dist.data_ptr())
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream.synchronize()
    e0.record(stream)
    for k in range(40):
        trees[k % 4].nearest_dev(q.data_ptr(), 1, idx.data_ptr(), dist.data_ptr())
    e1.record(stream)
    stream.synchronize()
    print(f"nodes {n}: {e0.elapsed_time(e1) / 40 * 1e3:.2f} us per call (rotating trees), index {int(idx[0])}")
