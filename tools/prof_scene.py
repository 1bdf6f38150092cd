"""Large-scene validity: time the kernel over 2^20 uniform configurations for a few generated scenes.
NWB_SCENE_NO_QUEUE=1 selects the per-thread walk; the default is the CTA work-queue kernel.  Prints verdict digests so
two runs can be compared."""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import nao_wholebody_planning_b200 as nwb  # noqa: E402
import oracle_lib as OL  # noqa: E402


def main():
    n = 1 << 20
    dev = torch.device("cuda:0")
    q = OL.uniform_configs(n, seed=20121001)
    q_dev = torch.from_numpy(q).to(dev)
    v = torch.empty(n, dtype=torch.uint8, device=dev)
    r = torch.empty(n, dtype=torch.uint8, device=dev)
    sizes = sys.argv[1:] or ["17", "64", "256", "4096"]
    for nb in sizes:
        c = nwb.Context()
        boxes = OL.scene_boxes(nb) if nb.startswith("S") else OL.generated_scene(int(nb))
        c.scene_add_boxes(boxes)
        nb = len(boxes)
        ts = []
        for it in range(6):
            c.is_feasible_dev(q_dev.data_ptr(), nwb.F64, n, v.data_ptr(), r.data_ptr())
            ts.append(c.last_kernel_ms()[0])
        torch.cuda.synchronize()
        dig = hashlib.sha256(v.cpu().numpy().tobytes() + r.cpu().numpy().tobytes()).hexdigest()[:16]
        print(f"boxes {nb:6d}  ms {np.median(ts[2:]):8.3f}  configs/s {n / np.median(ts[2:]) * 1e3:.3e}  "
              f"collision-free {float(((r & 1) == 0).float().mean()):.4f}  digest {dig}", flush=True)
        c.close()


if __name__ == "__main__":
    main()
