#!/usr/bin/env python3
"""bench_ops.py - secondary measurements of the hot-path operators.

bench.py carries the headline metric (V1 validity); this script times the other operators on one
GPU and prints one JSON line per measurement (also appended to --out):
  V2  nearest-neighbour scan over trees of 1e4 / 1e5 / 1e6 nodes (HBM roofline), connect walks and
      4096 x 102-state edge checks
  V3  stable-configuration database generation (2^20 samples)
  V0  single planning query latency (101 seeded repetitions -> p50), V4: 256 queries in lock step
Each line carries the CPU oracle timed on a bounded sample of the same work on this box.
"""
import argparse
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

import nao_wholebody_planning_b200 as nwb  # noqa: E402
import oracle_lib as OL  # noqa: E402


def emit(out, **kw):
    line = json.dumps(kw)
    print(line, flush=True)
    if out:
        with open(out, "a") as f:
            f.write(line + "\n")


def dev_time(fn, stream, reps=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


CROUCH = np.array([0, -0.349066, 0.698132, -0.436332, 0, 1.39626, 0.349066, -1.39626, -0.5, 0, 1.39626, -0.349066, 1.39626,
                   0.5, 0, 0, 0, -0.436332, 0.698132, -0.349066, 0])     # planning_example_simulation.cpp:199-219


def bench_services(args):
    """V0: scenario 3 (compute_linear_manipulation_plan, SIM:496-510), 101 seeded repetitions -> p50.
    V4: 256 grasp-pose compute_motion_plan requests (SIM:778-788 window) in one batched call."""
    orc = OL.Oracle()
    db = OL.golden_db()
    ctx = nwb.Context()
    ctx.set_ds_database(db)
    start = np.zeros(21)
    start[8], start[13] = -0.5, 0.5                                       # SIM:232-252
    P, D = [0.24, -0.065, 0.30], [0.0, -1.0, 0.0]
    ctx.compute_linear_manipulation_plan(P, D, start, 0.0, 0.10, 1)       # warm-up: allocations
    lat, parts, ok = [], [], 0
    reps = 21 if args.quick else 101
    for i in range(reps):
        t0 = time.perf_counter()
        r = ctx.compute_linear_manipulation_plan(P, D, start, 0.0, 0.10, 20121003 + i)
        lat.append(time.perf_counter() - t0)
        ok += r["success_approach"] and r["success_manipulation"]
        parts.append((r["time_goal_config_approach"], r["search_time_approach"], r["search_time_manipulate"]))
    ncpu = 3
    cpu, cpu_ok = [], 0
    for i in range(ncpu):
        t0 = time.perf_counter()
        o = orc.compute_manipulation_plan(db, P + D, start, 20121003 + i, False, [0.0, 0.10])
        cpu.append(time.perf_counter() - t0)
        cpu_ok += o["success_approach"] and o["success_manipulation"]
    parts = 1e3 * np.median(np.array(parts), 0)
    emit(args.out, op="scenario3_linear_manipulation_plan", repetitions=reps, success=int(ok), p50_ms=1e3 * float(np.median(lat)),
         p90_ms=1e3 * float(np.percentile(lat, 90)), max_ms=1e3 * float(np.max(lat)),
         median_parts_ms=dict(goal_sampling_both=float(parts[0]), search_approach=float(parts[1]), search_manipulate=float(parts[2])),
         cpu_oracle_s=[float(x) for x in cpu], cpu_oracle_p50_ms=1e3 * float(np.median(cpu)), cpu_repetitions=ncpu,
         cpu_success=int(cpu_ok), cpu_threads=1,
         note="request SIM:496-510 (hand (0.24,-0.065,0.30), dir (0,-1,0), pull 0.10 m), start SIM:232-252, shipped database, "
              "seeds 20121003+i; both goal samplings share one launch; the p90 tail = approach searches that use most of the "
              "8000 iterations (7 launches each)")
    # ---- V4 ----
    n = 256
    pos = np.array([[-0.08 + i * 0.4 / 15, -0.30 + j * 0.4 / 15, 0.2] for i in range(16) for j in range(16)])
    dirs = np.tile([0.0, -1.0, 0.0], (n, 1))
    seeds = (20121100 + np.arange(n)).astype(np.uint64)
    starts = np.tile(CROUCH, (n, 1))
    ctx.compute_motion_plan_batch(pos[:8], dirs[:8], starts[:8], seeds[:8])
    t0 = time.perf_counter()
    res = ctx.compute_motion_plan_batch(pos, dirs, starts, seeds)
    dt = time.perf_counter() - t0
    solved = [k for k in range(n) if res[k]["success"]]
    t0 = time.perf_counter()
    one_by_one = [ctx.compute_motion_plan(pos[k], dirs[k], CROUCH, int(seeds[k])) for k in range(n)]
    dt_seq = time.perf_counter() - t0
    same = sum(a["success"] == b["success"] and a["trajectory"].shape == b["trajectory"].shape for a, b in zip(res, one_by_one))
    # CPU: a bounded sample of the same requests (every 16th), extrapolated
    sample = list(range(0, n, 17))            # a diagonal of the grid
    t0 = time.perf_counter()
    cpu_solved = 0
    for k in sample:
        o = orc.compute_motion_plan(db, list(pos[k]) + list(dirs[k]), CROUCH, int(seeds[k]))
        cpu_solved += o["success"]
    cpu_dt = (time.perf_counter() - t0) / len(sample) * n
    emit(args.out, op="grasp_pose_queries_256", queries=n, solved=len(solved), wall_s_batched=dt, wall_s_one_by_one=dt_seq,
         batch_equals_single=int(same), time_goal_config_s=res[0]["time_goal_config"], search_time_s=max(r["search_time"] for r in res),
         cpu_oracle_1thread_s_extrapolated=cpu_dt, cpu_sample=len(sample), cpu_sample_solved=int(cpu_solved),
         note="16x16 grid of the evaluation window SIM:778-788 (z = 0.2, dir (0,-1,0)), crouched start SIM:199-219, "
              "compute_motion_plan semantics; seeds 20121100+k")
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="", help="comma list: ops,services (default: both)")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    if not only or "services" in only:
        bench_services(args)
    if only and "ops" not in only:
        return
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    rf = json.loads((ROOT / "profiles" / "roofline.json").read_text())
    orc = OL.Oracle()
    dev = torch.device("cuda", 0)
    ctx = nwb.Context()
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    ffma = max(ctx.ffma_peak(8192)[0] for _ in range(3))
    db = OL.golden_db()
    w = np.ones(21)
    w[15] = 0.0

    # ---- V1 / S1: validity against the 7-box shelf scene -------------------------
    nv = 1 << (18 if args.quick else 20)
    qv = OL.uniform_configs(nv, seed=20121001)
    c1 = nwb.Context()
    c1.set_stream(stream.cuda_stream)
    c1.scene_add_boxes(OL.scene_boxes("S1"))
    with torch.cuda.stream(stream):
        qv_dev = torch.from_numpy(qv).to(dev)
        v_dev = torch.empty(nv, dtype=torch.uint8, device=dev)
    ms = dev_time(lambda: c1.is_feasible_dev(qv_dev.data_ptr(), nwb._abi.F64, nv, v_dev.data_ptr()), stream, reps=10)
    ns = 40000
    t0 = time.perf_counter()
    cv, _, _ = orc.is_feasible(qv[:ns], boxes=OL.scene_boxes("S1"))
    cpu1 = ns / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    orc.is_feasible(qv[:ns * 8], boxes=OL.scene_boxes("S1"), threads=orc.threads)
    cpun = ns * 8 / (time.perf_counter() - t0)
    fl = rf["validity_S1_shelf_flops_per_config"]
    emit(args.out, op="validity_S1_shelf", configs=nv, boxes=c1.num_boxes, ms=ms, configs_per_s=nv / (ms * 1e-3),
         valid_fraction=float(v_dev.float().mean().item()), mismatches_vs_oracle=int((v_dev[:ns].cpu().numpy() != cv).sum()),
         oracle_sample=ns, cpu_configs_per_s_1thread=cpu1, cpu_configs_per_s_all_threads=cpun, cpu_threads=orc.threads,
         roofline={"bound": "fp32", "achieved": fl * nv / (ms * 1e-3) / 1e12, "peak": ffma, "unit": "TFLOP/s",
                   "frac": fl * nv / (ms * 1e-3) / 1e12 / ffma, "flops_per_config": fl})
    c1.close()

    # ---- V2: nearest neighbour ------------------------------------------------------------------
    nodes_all = OL.uniform_configs(1_000_000, seed=20121001, lhyp_zero=True)
    rng = np.random.default_rng(20121002)
    queries = db[np.floor(rng.random(1024) * (len(db) - 1)).astype(int)]
    with torch.cuda.stream(stream):
        q_dev = torch.from_numpy(queries).to(dev)
        idx_dev = torch.empty(1024, dtype=torch.int32, device=dev)
        dist_dev = torch.empty(1024, dtype=torch.float64, device=dev)
    for n in (10_000, 100_000, 1_000_000):
        tree = nwb.Tree(ctx, n)
        tree.add(nodes_all[:n])
        for nq in (1, 4, 1024):
            ts = []
            for _ in range(12):                        # device time of the scan + final kernels (library events)
                tree.nearest_dev(q_dev.data_ptr(), nq, idx_dev.data_ptr(), dist_dev.data_ptr())
                ts.append(ctx.last_kernel_ms()[0])
            ms = float(np.median(ts[2:]))
            passes = 1 if nq == 1 else -(-nq // 8)     # queries share a pass over the nodes in groups of 8
            gbs = 84.0 * n * passes / (ms * 1e-3) / 1e9
            sample_q = min(nq, 8 if n == 1_000_000 else 64)
            t0 = time.perf_counter()
            orc.nearest_neighbour(nodes_all[:n], queries[:sample_q])
            cpu = sample_q / (time.perf_counter() - t0)
            cpu_shipped = None
            if nq == 1 and n <= 100_000:
                t0 = time.perf_counter()
                orc.nearest_neighbour(nodes_all[:n], queries[:2], as_shipped=True)
                cpu_shipped = 2 / (time.perf_counter() - t0)
            emit(args.out, op="nearest_neighbour", nodes=n, queries=nq, ms=ms, queries_per_s=nq / (ms * 1e-3),
                 roofline={"bound": "hbm", "bytes": 84.0 * n * passes, "achieved_gbs": gbs, "peak_gbs": hbm, "frac": gbs / hbm,
                           "note": "84 B/node per pass (fp32 SoA scan copy, 1 query or a group of 8 per pass); later passes of a batch are served from L2; trees below ~1e5 nodes are bound by the ~13 us fixed cost of two launches"},
                 cpu_queries_per_s_1thread=cpu, cpu_as_shipped_tree_by_value_queries_per_s=cpu_shipped)
        tree.close()

    # ---- V2: edge checks (4096 random node pairs x 102 states, collision only) ----------------------------
    pairs = rng.integers(0, 200_000, (4096, 2))
    qa, qb = nodes_all[pairs[:, 0]], nodes_all[pairs[:, 1]]
    dbp = rng.integers(0, len(db), (4096, 2))
    for tag, a, b in (("random_nodes", qa, qb), ("db_rows", db[dbp[:, 0]], db[dbp[:, 1]])):
        with torch.cuda.stream(stream):
            a_d, b_d = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
            ok_d = torch.empty(4096, dtype=torch.uint8, device=dev)
            fb_d = torch.empty(4096, dtype=torch.int32, device=dev)
        lib, h = ctx.lib, ctx.h
        ms = dev_time(lambda: lib.nwb_is_path_valid_batch_dev(h, a_d.data_ptr(), b_d.data_ptr(), 4096, 100, ok_d.data_ptr(),
                                                              fb_d.data_ptr()), stream, reps=20)
        flops = rf["collision_only_S0_all_pairs_flops_per_config"] * 102 * 4096
        t0 = time.perf_counter()
        orc.is_path_valid(a[:256], b[:256], 100)
        cpu = 256 / (time.perf_counter() - t0)
        emit(args.out, op="edge_check", inputs=tag, edges=4096, states_per_edge=102, ms=ms, edges_per_s=4096 / (ms * 1e-3),
             valid_fraction=float(ok_d.float().mean().item()),
             roofline={"bound": "fp32", "achieved_tflops": flops / (ms * 1e-3) / 1e12, "peak_tflops": ffma,
                       "frac": flops / (ms * 1e-3) / 1e12 / ffma, "note": "full evaluation of all 102 way-points, no early exit"},
             cpu_edges_per_s_1thread=cpu, cpu_note="the serial CPU path stops at the first colliding way-point")

    # ---- steer + project ------------------------------------------------------------------------------------
    n = 1 << 18
    wp = w.ctypes.data_as(C.c_void_p)
    big = OL.uniform_configs(2 << 20, seed=20121009, lhyp_zero=True)
    for tag, near, target in (("db_rows", db[rng.integers(0, len(db), n)], db[rng.integers(0, len(db), n)]),
                              ("uniform_configs", nodes_all[:n], nodes_all[n:2 * n]),
                              ("uniform_configs_2^20", big[:1 << 20], big[1 << 20:])):
        n = len(near)
        with torch.cuda.stream(stream):
            near_d, tgt_d = torch.from_numpy(np.ascontiguousarray(near)).to(dev), torch.from_numpy(np.ascontiguousarray(target)).to(dev)
            out_d = torch.empty_like(near_d)
            st_d = torch.empty(n, dtype=torch.int32, device=dev)
            ds_d = torch.empty(n, dtype=torch.int32, device=dev)
        ms = dev_time(lambda: ctx.lib.nwb_steer_project_batch_dev(ctx.h, near_d.data_ptr(), tgt_d.data_ptr(), n, wp, 0.1,
                                                                  out_d.data_ptr(), st_d.data_ptr(), ds_d.data_ptr()), stream, reps=5)
        t0 = time.perf_counter()
        orc.steer_project(near[:20000], target[:20000], w, 0.1)
        cpu = 20000 / (time.perf_counter() - t0)
        emit(args.out, op="steer_project", inputs=tag, n=n, ms=ms, per_s=n / (ms * 1e-3),
             ik_solved_fraction=float((ds_d == 0).float().mean().item()), trapped_fraction=float((ds_d == 2).float().mean().item()),
             cpu_per_s_1thread=cpu, note="generate_q_new + joint limits + legs FK + fp64 NR/pinv IK where the foot is off target")

    # ---- V3: database generation -----------------------------------------------------------------------------
    count = 1 << (16 if args.quick else 20)
    with torch.cuda.stream(stream):
        rows_d = torch.empty((count, 21), dtype=torch.float64, device=dev)
        idx_d = torch.empty(count, dtype=torch.int64, device=dev)
        n_d = torch.zeros(1, dtype=torch.int64, device=dev)

    def gen():
        with torch.cuda.stream(stream):                 # the counter must be cleared on the stream the sampler runs on
            n_d.zero_()
        ctx.lib.nwb_sample_stable_configs_dev(ctx.h, 20121005, 0, count, 5, rows_d.data_ptr(), idx_d.data_ptr(), count,
                                              n_d.data_ptr(), None)
    ms = dev_time(gen, stream, reps=3, warm=1)
    accepted = int(n_d.item())
    t0 = time.perf_counter()
    orc.sample_stable_configs(20121005, 0, 20000)
    cpu1 = 20000 / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    _, oidx, ostage = orc.sample_stable_configs(20121005, 0, 200000, threads=orc.threads)
    cpun = 200000 / (time.perf_counter() - t0)
    # the oracle runs every Newton-Raphson solve to the end; the device rejects the unreachable targets up front
    _, gidx, gstage = ctx.sample_stable_configs(20121005, 0, 200000)
    emit(args.out, op="sample_stable_configs", samples=count, accepted=accepted, ms=ms, samples_per_s=count / (ms * 1e-3),
         cpu_samples_per_s_1thread=cpu1, cpu_samples_per_s_all_threads=cpun, cpu_threads=orc.threads,
         oracle_sample=200000, ik_verdict_mismatches_vs_oracle=int(((gstage == 0) != (ostage == 0)).sum()),
         stage_mismatches_vs_oracle=int((gstage != ostage).sum()), accepted_oracle_sample=[int(len(gidx)), int(len(oidx))],
         note="Philox sample -> hip FK -> reach / leg-plane rejection -> NR/pinv IK (fp64) on the packed survivors -> validity "
              "(fp32) -> atomic compaction")

    # ---- V0 / V4: planning ----------------------------------------------------------------------------------------
    ctx.set_stream(None)
    sol = OL.load_dat(OL.GOLDEN / "solutions" / "solution_path_first.dat")
    start0 = np.zeros(21)
    start0[8], start0[13] = -0.5, 0.5                         # SIM:232-252 (scenario 3 start pose)
    goal = sol[1]                                              # shipped goal configuration of the approach plan
    lat, lat_cpu, ok_all = [], [], 0
    for i in range(101):
        seed = np.array([20121003 + i], dtype=np.uint64)
        t0 = time.perf_counter()
        res, paths = ctx.solveQuery(start0[None], goal[None], db, seed, w, max_iter=8000, step=0.1)
        short, _ = ctx.shortcut_path(paths[0]) if res[0]["success"] else (None, 0)
        lat.append(time.perf_counter() - t0)
        ok_all += res[0]["success"]
        t0 = time.perf_counter()
        oraw, ost = orc.solve_query(start0, goal, db, w, int(seed[0]))
        if ost["success"]:
            orc.shortcut_path(oraw)
        lat_cpu.append(time.perf_counter() - t0)
    emit(args.out, op="plan_single_query", scene="S0", repetitions=101, success=ok_all, p50_ms=1e3 * float(np.median(lat)),
         p90_ms=1e3 * float(np.percentile(lat, 90)), cpu_p50_ms=1e3 * float(np.median(lat_cpu)),
         note="approach plan of scenario 3 (start SIM:232-252 -> shipped goal config) incl. shortcutter; one query is "
              "latency bound on a GPU: RRT-Connect is serial by construction")
    nq = 256
    starts = np.tile(start0, (nq, 1))
    goals = db[rng.integers(0, len(db), nq)]
    seeds = (20121100 + np.arange(nq)).astype(np.uint64)
    for scene in ("S0", "S3"):
        c2 = nwb.Context()
        c2.scene_add_boxes(OL.scene_boxes(scene))
        c2.solveQuery(starts, goals, db, seeds, w, max_iter=500)                   # warm-up (allocates the forest)
        t0 = time.perf_counter()
        res, paths = c2.solveQuery(starts, goals, db, seeds, w, max_iter=500, step=0.1)
        dt = time.perf_counter() - t0
        ns = 32
        t0 = time.perf_counter()
        for k in range(ns):
            orc.solve_query(starts[k], goals[k], db, w, int(seeds[k]), max_iter=500, boxes=OL.scene_boxes(scene))
        cpu_dt = (time.perf_counter() - t0) / ns * nq
        emit(args.out, op="plan_batch", scene=scene, queries=nq, max_iter=500, wall_s=dt, solved=sum(r["success"] for r in res),
             iterations_max=max(r["iterations"] for r in res), cpu_1thread_s_extrapolated=cpu_dt,
             note="256 queries advanced in lock step on one GPU vs the oracle run query by query")
        c2.close()
    ctx.close()


if __name__ == "__main__":
    main()
