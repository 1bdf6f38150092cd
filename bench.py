#!/usr/bin/env python3
"""bench.py - headline benchmark: validated configurations / second (FK + COM + collision).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1], SURVEY.md section 8d "V1/S0"): B = 1 048 576 uniformly random NAO
whole-body configurations per GPU (seed 20121001 + rank), scene S0 (the empty scene the shipped
planner actually checks), right-foot support, support polygon scale 0.8.  One step = one pass of
the validity hot path over the batch (FK + COM-in-support-polygon + 108 self-collision pairs, both
tests always evaluated, one kernel launch).

  value   device-timed throughput with the batch resident in HBM in the reference API's layout
          ([n][21] float64, 176 MB > the 126 MB L2, so no inter-iteration reuse)
  e2e     the same metric through the C-ABI call nwb_is_feasible_batch with pinned HOST buffers:
          H2D of the configurations and D2H of the verdicts inside the timed region
  roofline  FP32-pipe: algorithmic flops (profiles/roofline.json, frozen by tools/count_flops.cpp)
            / mean kernel duration (CUDA events on the launching stream) vs the FFMA peak measured
            live by nwb_ffma_peak in this same process
  cpu_baseline  the CPU oracle (a port: the reference cannot be built here) on a bounded sample,
            one thread, on this box's host cores

Extra keys of the same line (bench_sections.py; BASELINE configs 3-5, inputs of SURVEY.md section 8d): `validity_scenes`
(S1 shelf and a generated 256-box scene), `nn_scan` {1e4,1e5,1e6} x {1,1024}, `connect`, `edge_check` 4096 x 102, `dbgen`
(2^20 on one GPU + 2^26 strong-scaled over the ranks, fused peer append, digest equal at every N), `grasp_queries` 256 / 1681
sharded round-robin with a packed NCCL gather (digest equal at every N).  `--no-sections` skips them.

Multi-GPU: the batch shards by index, no data-path collective; the 1 B/config verdicts reach every
rank because the validity kernel stores them into all ranks' gathered arrays through peer-mapped
memory (NVLink / NVSwitch) - a fused compute + gather.  NCCL is used for the barrier, the max-over-ranks
timing and, after the timed region, to verify the fused gather against a plain all_gather.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

B_PER_GPU = 1 << 20
SEED = 20121001
METRIC = "validated_configs_per_sec"
UNIT = "configs/s"
WORKLOAD = "V1/S0: 1,048,576 uniform random NAO configs per GPU vs the scenario scene (empty), right-foot support, scale_sp 0.8"


def load_json(path, default=None):
    try:
        return json.loads(Path(path).read_text())
    except Exception:
        return default


def make_batch(n, seed):
    import oracle_lib as OL   # input generator only (joint limits + PCG64 stream of SURVEY.md 8d)
    return OL.uniform_configs(n, seed=seed)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


S3_POS, S3_DIR, S3_PULL = [0.24, -0.065, 0.30], [0.0, -1.0, 0.0], 0.10     # planning_example_simulation.cpp:496-510


def s3_start():
    q = np.zeros(21)
    q[8], q[13] = -0.5, 0.5                                                 # planning_example_simulation.cpp:232-252
    return q


def scenario3_gpu(reps=101):
    """Second half of BASELINE.json's metric: scenario-3 plan time p50 (compute_linear_manipulation_plan through
    the C ABI, 101 seeded repetitions, host wall clock around the call - it returns host-side trajectories)."""
    import oracle_lib as OL
    import nao_wholebody_planning_b200 as nwb
    c = nwb.Context()
    c.set_ds_database(OL.golden_db())
    c.compute_linear_manipulation_plan(S3_POS, S3_DIR, s3_start(), 0.0, S3_PULL, 1)    # warm-up (allocations)
    lat, ok = [], 0
    for i in range(reps):
        t0 = time.perf_counter()
        r = c.compute_linear_manipulation_plan(S3_POS, S3_DIR, s3_start(), 0.0, S3_PULL, 20121003 + i)
        lat.append(time.perf_counter() - t0)
        ok += bool(r["success_approach"] and r["success_manipulation"])
    c.close()
    return {"p50_ms": 1e3 * float(np.median(lat)), "p90_ms": 1e3 * float(np.percentile(lat, 90)), "repetitions": reps,
            "both_plans_found": ok, "api": "nwb_compute_linear_manipulation_plan", "seeds": "20121003+i"}


def scenario3_cpu(reps=3):
    import oracle_lib as OL
    orc = OL.Oracle()
    db = OL.golden_db()
    lat = []
    for i in range(reps):
        t0 = time.perf_counter()
        orc.compute_manipulation_plan(db, S3_POS + S3_DIR, s3_start(), 20121003 + i, False, [0.0, S3_PULL])
        lat.append(time.perf_counter() - t0)
    return {"p50_ms": 1e3 * float(np.median(lat)), "repetitions": reps, "cores": 1, "kind": "port",
            "sample": f"the first {reps} of the 101 seeds; the planner services are single-threaded (ros::spin)"}


def cpu_baseline(q, seconds=12.0, threads=1):
    """Time the CPU oracle on a bounded sample of the same batch (about `seconds` of CPU work)."""
    import oracle_lib as OL
    orc = OL.Oracle()
    probe = 20000
    t0 = time.perf_counter()
    orc.is_feasible(q[:probe], threads=threads)
    rate = probe / (time.perf_counter() - t0)
    n = int(min(len(q), max(probe, rate * seconds)))
    done, dt, valid, band = 0, 0.0, None, None
    t0 = time.perf_counter()
    while dt < seconds * 0.8:                      # whole passes over the first n configurations
        valid, _, marg = orc.is_feasible(q[:n], threads=threads)
        done += n
        dt = time.perf_counter() - t0
    band = (np.abs(marg) < 1e-5).any(1)            # within 1e-5 m of a decision boundary: fp32 product vs fp64 oracle may differ
    return {"value": done / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{done // n} pass(es) over the first {n} configurations of the same batch, {dt:.1f} s, oracle/ "
                      f"(fp64 C++ restatement; the reference itself needs ROS/MoveIt/FCL/KDL and cannot be built here)"}, valid, n, band


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port), all host threads, bounded sample per step."""
    if rank != 0:
        return
    import oracle_lib as OL
    orc = OL.Oracle()
    threads = max(1, orc.threads)
    q = make_batch(B_PER_GPU, SEED)
    # size the per-step sample so the whole run stays within a few minutes
    t0 = time.perf_counter()
    orc.is_feasible(q[:50000], threads=threads)
    rate = 50000 / (time.perf_counter() - t0)
    budget_s = 150.0 / max(1, args.steps + args.warmup)
    n = int(min(B_PER_GPU, max(50000, rate * min(budget_s, 20.0))))
    for _ in range(args.warmup):
        orc.is_feasible(q[:n], threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.is_feasible(q[:n], threads=threads)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = f"first {n} configurations of the batch per step, {threads} host threads"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B_PER_GPU},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if args.gpus == 1 and not args.no_scenario3:
        line["scenario3"] = scenario3_cpu(3)
    print(json.dumps(line))


def run_sections(args, torch, dist, nwb, ctx, dev, rank, world, barrier, ffma_tflops):
    """BASELINE configs 3-5 as extra keys (bench_sections.py).  A section that fails reports {"error": ...} instead of
    taking the headline down with it.  Single-GPU operators (scenes, nn_scan, connect, edge_check) run at N = 1 only;
    dbgen and grasp_queries run at every N (strong scaling, digests equal across N)."""
    import bench_sections as BS
    import oracle_lib as OL
    from nao_wholebody_planning_b200 import sharding
    out = {}
    orc = OL.Oracle() if rank == 0 else None
    cpu = not args.no_cpu_baseline
    rf = load_json(ROOT / "profiles" / "roofline.json", {})
    peaks = load_json(ROOT / "MEASURED_PEAKS.json", {})
    hbm = peaks.get("hbm_gbs") or 6538.9

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def guarded(name, fn):
        try:
            v = fn()
            if rank == 0 and v is not None:
                out[name] = v
        except Exception as e:                                   # noqa: BLE001 - reported, never hidden
            import traceback
            traceback.print_exc()
            if rank == 0:
                out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
            if world > 1:
                raise                                            # a rank out of step would hang the others: fail loudly

    if world == 1:
        stream = torch.cuda.Stream(device=dev)
        ctx.set_stream(stream.cuda_stream)
        flush = BS.L2Flush(torch, dev, stream)
        t0 = time.perf_counter()
        nodes_all, n_proj = BS.v1c_nodes(nwb, ctx, OL, 1_000_000)
        prep_s = time.perf_counter() - t0
        guarded("validity_scenes", lambda: BS.section_validity_scenes(torch, nwb, OL, orc, stream, dev, ffma_tflops, rf, cpu))
        guarded("nn_scan", lambda: BS.section_nn_scan(torch, nwb, ctx, OL, orc, stream, dev, nodes_all, hbm, peaks.get("bf16_tflops") or 1650.7, flush))
        guarded("edge_check", lambda: BS.section_edge_check(torch, nwb, ctx, OL, orc, stream, dev, nodes_all, ffma_tflops,
                                                            rf.get("collision_only_S0_all_pairs_flops_per_config", 0)))
        ctx.set_stream(None)
        guarded("connect", lambda: BS.section_connect(nwb, ctx, OL, orc, nodes_all))
        out["v2_inputs"] = {"nodes": "first N rows of V1c (uniform, LHipYawPitch = 0, after the double-support projection)",
                            "rows_changed_by_the_projection": n_proj, "prepare_s": prep_s,
                            "queries": "1024 database rows, index stream seed 20121002 (RP:319-320)"}
        del nodes_all
    ctx.set_stream(None)
    ctx.set_ds_database(OL.golden_db())
    guarded("dbgen", lambda: BS.section_dbgen(torch, dist, nwb, sharding, ctx, orc, dev, rank, world, barrier, max_over_ranks, cpu=cpu))
    guarded("grasp_queries", lambda: BS.section_grasp_queries(torch, dist, nwb, sharding, ctx, OL, orc, dev, rank, world, barrier,
                                                              max_over_ranks, cpu=cpu))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scenario3", action="store_true", help="skip the scenario-3 plan-time measurement (N = 1 only)")
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="configurations per GPU (default = the headline size)")
    ap.add_argument("--no-sections", action="store_true", help="skip the extra keys (configs 3-5: nn_scan, edge_check, connect, dbgen, grasp_queries)")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the process to the GPU's NUMA node before allocating pinned memory")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import nao_wholebody_planning_b200 as nwb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("NWB_EDGE_FULL", "1")     # edge checks evaluate every way-point: the roofline counts all of them
    # pinned staging buffers NUMA-local to this GPU's PCIe root: bind BEFORE anything allocates pinned memory
    numa = None
    if not args.no_numa_bind:
        from nao_wholebody_planning_b200 import sharding as _sh
        pr = torch.cuda.get_device_properties(local_rank)
        try:
            bus_id = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        except AttributeError:
            bus_id = ""
        numa = _sh.bind_to_gpu_numa_node(bus_id) if bus_id else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n = args.batch
    q_host = make_batch(n, SEED + rank)
    ctx = nwb.Context(device=local_rank)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)          # the library launches on this torch stream

    with torch.cuda.stream(stream):
        q_dev = torch.from_numpy(q_host).to(dev)                       # [n,21] f64, resident in HBM
        valid = torch.empty(n, dtype=torch.uint8, device=dev)
        reason = torch.empty(n, dtype=torch.uint8, device=dev)
    stream.synchronize()

    # N > 1: the verdicts of every shard land in every rank's gathered array.  No collective in the step:
    # the validity kernel stores them through peer-mapped memory (NVLink / NVSwitch) as it produces them.
    gathered_ptr, peer_ptrs = None, []
    if world > 1:
        gathered_ptr = ctx.dev_alloc(n * world)
        handles = [None] * world
        dist.all_gather_object(handles, ctx.ipc_export(gathered_ptr))
        peer_ptrs = [gathered_ptr if r == rank else ctx.ipc_import(handles[r]) for r in range(world)]
        ctx.set_gather_peers(peer_ptrs, rank * n)

    def step():
        """One pass of the hot path over this rank's shard (one kernel launch; at N > 1 it also writes the
        verdicts into all ranks' gathered arrays)."""
        ctx.is_feasible_dev(q_dev.data_ptr(), nwb.F64, n, valid.data_ptr(), reason.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # measured FP32 peak (roofline denominator), before the timed region
    ffma_tflops = max(ctx.ffma_peak(8192)[0] for _ in range(3))

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step()
        barrier()
        clocks = ClockSampler(local_rank)
        clocks.start()
        time.sleep(0.25)
        # The K launches go back to back with ONE event on either side: an event record between two launches costs
        # ~2.5 us of stream time (measured: 131.1 us per launch bare, 136.3 with the library's own timing pair, 142.9
        # with a third, per-step event), so the library's per-call timing events are switched off for the timed region.
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ctx.set_kernel_timing(False)
        barrier()
        ev[0].record(stream)
        for k in range(args.steps):
            step()
        ev[1].record(stream)
        barrier()                                       # all ranks done => every gathered array is complete
        # keep the GPU busy a little longer so the clock sampler sees the loaded state
        t_end = time.time() + 0.6
        while time.time() < t_end:
            step()
        torch.cuda.synchronize(dev)
        clk = clocks.stop()
        ctx.set_kernel_timing(True)
    total_ms = ev[0].elapsed_time(ev[1])
    step_ms = [total_ms / args.steps]
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = n * world * args.steps / (total_ms * 1e-3)

    # kernel-only duration (one launch per step): events bracket exactly the kernel when N = 1
    kern_ms = float(np.mean(step_ms)) if world == 1 else None
    if world > 1:
        # time the kernel alone on this rank for the roofline figure
        with torch.cuda.stream(stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ctx.set_kernel_timing(False)
            e0.record(stream)
            for _ in range(args.steps):
                ctx.is_feasible_dev(q_dev.data_ptr(), nwb.F64, n, valid.data_ptr(), reason.data_ptr())
            e1.record(stream)
            torch.cuda.synchronize(dev)
            ctx.set_kernel_timing(True)
            kern_ms = e0.elapsed_time(e1) / args.steps

    gather_ok = None
    if world > 1:                                       # check the fused gather against an NCCL all_gather (untimed)
        ref = torch.empty(n * world, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(ref, valid)

        class _Raw:                                     # view of the library-allocated gathered array
            __cuda_array_interface__ = {"shape": (n * world,), "typestr": "|u1", "data": (int(gathered_ptr), False), "version": 3}
        mine = torch.as_tensor(_Raw(), device=dev)
        torch.cuda.synchronize(dev)
        gather_ok = bool((mine == ref).all().item())
        assert gather_ok, "fused peer-memory gather disagrees with the NCCL all_gather"
        ctx.set_gather_peers([], 0)
        dist.barrier()

    # ---- end to end through the C ABI with pinned host buffers ------------------------------------
    ctx.set_stream(None)
    pin_q = nwb.PinnedArray((n, nwb.NJ), np.float64)
    pin_v = nwb.PinnedArray((n,), np.uint8)
    pin_q.array[:] = q_host
    for _ in range(2):
        ctx.is_feasible_host_buffers(pin_q.ptr, n, pin_v.ptr)
    barrier()
    # the wire under the end-to-end number: this rank's pinned H2D copy rate while ALL ranks copy at the same time
    # (the SAME pinned buffer the end-to-end call reads: a second allocation may sit on other host pages)
    h2d_src = torch.from_numpy(pin_q.array).view(-1)
    h2d_dst = torch.empty(n * nwb.NJ, dtype=torch.float64, device=dev)
    h2d_dst.copy_(h2d_src, non_blocking=True)
    barrier()
    ha, hb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ha.record()
    for _ in range(4):
        h2d_dst.copy_(h2d_src, non_blocking=True)
    hb.record()
    torch.cuda.synchronize(dev)
    h2d_gbs = 4 * n * nwb.NJ * 8 / (ha.elapsed_time(hb) * 1e-3) / 1e9
    h2d_all = [h2d_gbs]
    if world > 1:
        t = torch.tensor([h2d_gbs], device=dev, dtype=torch.float64)
        parts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        h2d_all = [float(x.item()) for x in parts]
    del h2d_src, h2d_dst
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.is_feasible_host_buffers(pin_q.ptr, n, pin_v.ptr)
    torch.cuda.synchronize(dev)
    e2e_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e_value = n * world * e2e_steps / e2e_dt
    e2e_valid = pin_v.array.copy()
    assert (e2e_valid == valid.cpu().numpy()).all(), "host-buffer path and device path disagree"
    # supplementary: the same call for a caller that keeps its configurations in float32 (half the PCIe bytes)
    pin_qf = nwb.PinnedArray((n, nwb.NJ), np.float32)
    pin_vf = nwb.PinnedArray((n,), np.uint8)
    pin_qf.array[:] = q_host.astype(np.float32)
    for _ in range(2):
        ctx.is_feasible_host_buffers_f32(pin_qf.ptr, n, pin_vf.ptr)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.is_feasible_host_buffers_f32(pin_qf.ptr, n, pin_vf.ptr)
    torch.cuda.synchronize(dev)
    e2e32_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e32_dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e32_dt = float(t.item())
    e2e32_value = n * world * e2e_steps / e2e32_dt
    e2e32_same = int((pin_vf.array == e2e_valid).sum())
    pin_qf.free()
    pin_vf.free()

    sections = {} if args.no_sections else run_sections(args, torch, dist, nwb, ctx, dev, rank, world, barrier, ffma_tflops)

    if rank == 0:
        rf = load_json(ROOT / "profiles" / "roofline.json", {})
        flops_cfg = rf.get("validity_S0_flops_per_config")
        peaks = load_json(ROOT / "MEASURED_PEAKS.json", {})
        ncu = load_json(ROOT / "profiles" / "validity_ncu_summary.json", {})
        achieved = flops_cfg * n / (kern_ms * 1e-3) / 1e12 if flops_cfg else None
        roofline = {"bound": "fp32", "kernel": "nwb::validity_persistent_kernel<double, MARGINS=false, COLLISION_ONLY=false, PAIRS_STATIC_GROUP>",
                    "achieved": achieved, "peak": ffma_tflops, "unit": "TFLOP/s",
                    "frac": (achieved / ffma_tflops) if achieved else None,
                    "peak_source": "nwb_ffma_peak measured live in this run (MEASURED_PEAKS.json has no FP32 figure; nominal 74.4)",
                    "flops_per_config": flops_cfg, "kernel_ms": kern_ms,
                    "traffic": ncu.get("dram_bytes_per_launch"),
                    "traffic_source": "profiles/validity_ncu_summary.json (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full "
                                      "capture of this kernel, per launch; not re-measured in this run)",
                    "algorithmic_bytes_per_launch": n * 170,
                    "hbm_view": {"achieved_gbs": n * 170 / (kern_ms * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                                 "note": "170 B/config (168 in, 2 out): far from the HBM bound, as designed"}}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "batch_per_gpu": n, "input_layout": "[n][21] float64 resident in HBM",
                           "l2_policy": "inputs (176 MB per GPU) larger than L2 (126 MB)",
                           "parallelism": f"shard{world}" if world > 1 else "single",
                           "result_gather": ("verdicts stored into every rank's gathered array by the validity kernel itself "
                                             "(peer-mapped memory over NVLink/NVSwitch, CUDA IPC); checked against an NCCL "
                                             "all_gather after the timed region: %s" % gather_ok) if world > 1 else None,
                           "valid_fraction": float(valid.float().mean().item())},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * nwb.NJ * 8, "d2h_bytes_per_step": n,
                        "steps": e2e_steps, "h2d_gbs_per_rank_all_ranks_copying": h2d_all,
                        "wire_bound_configs_per_s": sum(h2d_all) * 1e9 / (nwb.NJ * 8),
                        "numa_bind": ({"node": numa[0], "cpus": numa[1]} if numa else None), "api": "nwb_is_feasible_batch (pinned host buffers, float64 as the reference keeps them)"},
                "e2e_f32_input": {"value": e2e32_value, "unit": UNIT, "h2d_bytes_per_step": n * nwb.NJ * 4, "d2h_bytes_per_step": n,
                                  "api": "nwb_is_feasible_batch_f32 (supplementary: callers holding float32 configurations)",
                                  "verdicts_equal_to_f64_input": e2e32_same, "of": n},
                "gpu_launches": args.steps, "clocks": clk, "roofline": roofline}
        if not args.no_cpu_baseline:
            cb, cpu_valid, cn, band = cpu_baseline(q_host)
            line["cpu_baseline"] = cb
            diff = cpu_valid != e2e_valid[:cn]
            line["config"]["oracle_mismatches_in_cpu_sample"] = {"outside_band": int((diff & ~band).sum()), "inside_band": int((diff & band).sum()),
                                                                 "band_m": 1e-5, "configs_in_band": int(band.sum()), "sample": int(cn)}
        if world == 1 and not args.no_scenario3:
            line["scenario3"] = scenario3_gpu(101)
            if not args.no_cpu_baseline:
                line["scenario3"]["cpu_baseline"] = scenario3_cpu(3)
        line.update(sections)
        print(json.dumps(line))
    pin_q.free()
    pin_v.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
