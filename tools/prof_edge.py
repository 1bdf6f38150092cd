"""Edge checks, 4096 edges x 102 way-points (every way-point evaluated): back-to-back timing of nwb_is_path_valid_batch_dev."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("NWB_EDGE_FULL", "1")
import nao_wholebody_planning_b200 as nwb  # noqa: E402
import bench_sections as BS  # noqa: E402
import oracle_lib as OL  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    c0 = nwb.Context()
    nodes_all, _ = BS.v1c_nodes(nwb, c0, OL, 200_000)
    c0.close()
    rng = np.random.default_rng(20121002)
    pairs = rng.integers(0, 200_000, (4096, 2))
    qa, qb = np.ascontiguousarray(nodes_all[pairs[:, 0]]), np.ascontiguousarray(nodes_all[pairs[:, 1]])
    st = torch.cuda.Stream()
    c = nwb.Context()
    c.set_stream(st.cuda_stream)
    with torch.cuda.stream(st):
        a_d, b_d = torch.from_numpy(qa).to(dev), torch.from_numpy(qb).to(dev)
        ok_d = torch.empty(4096, dtype=torch.uint8, device=dev)
        fb_d = torch.empty(4096, dtype=torch.int32, device=dev)
    st.synchronize()

    def run():
        c.lib.nwb_is_path_valid_batch_dev(c.h, a_d.data_ptr(), b_d.data_ptr(), 4096, 100, ok_d.data_ptr(), fb_d.data_ptr())
    for _ in range(5):
        run()
    c.set_kernel_timing(False)
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(50):
            run()
        e1.record(st)
        st.synchronize()
        print(f"edge check back to back: {e0.elapsed_time(e1) / 50 * 1e3:.2f} us per call, ok fraction {float(ok_d.float().mean()):.4f}")


if __name__ == "__main__":
    main()
