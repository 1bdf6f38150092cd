#!/usr/bin/env python3
"""Quick device-time check of the two FP64 planner operators (subset of bench_ops.py, same inputs and timer):
database sampler over 2^20 samples and steer+project over 2^18 uniform pairs.  Usage: python tools/bench_sampler_steer.py"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench_ops as B  # noqa: E402
import nao_wholebody_planning_b200 as nwb  # noqa: E402
import oracle_lib as OL  # noqa: E402  (input generator only)

ctx = nwb.Context()
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
ctx.set_stream(stream.cuda_stream)
count = 1 << 20
with torch.cuda.stream(stream):
    rows_d = torch.empty((count, 21), dtype=torch.float64, device=dev)
    idx_d = torch.empty(count, dtype=torch.int64, device=dev)
    n_d = torch.zeros(1, dtype=torch.int64, device=dev)


def gen():
    with torch.cuda.stream(stream):
        n_d.zero_()
    ctx.lib.nwb_sample_stable_configs_dev(ctx.h, 20121005, 0, count, 5, rows_d.data_ptr(), idx_d.data_ptr(), count, n_d.data_ptr(), None)


ms = B.dev_time(gen, stream, reps=3, warm=1)
print("sampler %.2f ms %.1f M samples/s accepted %d" % (ms, count / ms / 1e3, int(n_d.item())))
n = 1 << 18
u = OL.uniform_configs(n, seed=5, lhyp_zero=True)
t = OL.uniform_configs(n, seed=6, lhyp_zero=True)
with torch.cuda.stream(stream):
    a, b = torch.from_numpy(u).to(dev), torch.from_numpy(t).to(dev)
    o = torch.empty_like(a)
    st = torch.empty(n, dtype=torch.int32, device=dev)
w = np.ones(21)
ms = B.dev_time(lambda: ctx.lib.nwb_steer_project_batch_dev(ctx.h, a.data_ptr(), b.data_ptr(), n, w.ctypes.data, 0.1, o.data_ptr(),
                                                            st.data_ptr(), None), stream, reps=3, warm=1)
print("steer+project uniform %.2f ms %.1f M pairs/s" % (ms, n / ms / 1e3))
