#!/usr/bin/env python3
"""bench_multi.py - the sharded workloads of SURVEY.md section 8e other than the headline validity batch:

  V3  stable-configuration database generation: sample-index ranges per rank (the Philox counter is the
      GLOBAL sample index, so the accepted set does not depend on the number of GPUs), rows gathered
      and merged in ascending sample index
  V2  nearest neighbour over ONE big tree (2^24 nodes): node ranges per rank, every rank scans its shard,
      all_gather of (distance, global index) per query + lexicographic minimum (the one exchange step)
  V4  256 grasp-pose compute_motion_plan requests: round-robin request -> rank, every rank runs its
      share as ONE batched call, results gathered

Launch like bench.py: `python bench_multi.py` (1 GPU) or
`python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench_multi.py`.
One JSON line per workload from rank 0; timing = barrier + synchronize on both sides, max over ranks.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

CROUCH = np.array([0, -0.349066, 0.698132, -0.436332, 0, 1.39626, 0.349066, -1.39626, -0.5, 0, 1.39626, -0.349066, 1.39626,
                   0.5, 0, 0, 0, -0.436332, 0.698132, -0.349066, 0])     # planning_example_simulation.cpp:199-219


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples-log2", type=int, default=22)
    ap.add_argument("--check-log2", type=int, default=16, help="size of the partition-independence check")
    ap.add_argument("--nodes-log2", type=int, default=24, help="nodes of the one big tree of the NN workload")
    ap.add_argument("--grasp-grid", type=int, default=16, help="V4: G x G hand targets over the evaluation window (the reference's design is 41)")
    ap.add_argument("--skip-nn", action="store_true", help="leave out the one-big-tree workload")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import nao_wholebody_planning_b200 as nwb
    from nao_wholebody_planning_b200 import sharding
    import oracle_lib as OL                       # golden database + joint limits only

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench_multi.py needs CUDA devices: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ctx = nwb.Context(device=local)
    ctx.set_ds_database(OL.golden_db())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(dt):
        if world == 1:
            return dt
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def emit(**kw):
        if rank == 0:
            line = json.dumps(kw)
            print(line, flush=True)
            if args.out:
                with open(args.out, "a") as f:
                    f.write(line + "\n")

    # ---- V3 ---------------------------------------------------------------------------------------------------
    seed = 20121005

    def generate(total):
        lo, hi = sharding.shard_bounds(total, world)[rank]
        rows, idx, _ = ctx.sample_stable_configs(seed, lo, hi - lo, max_ik_iter=5, cap=max(1024, (hi - lo) // 6), want_stage=False)
        r = torch.from_numpy(rows).to(dev)
        i = torch.from_numpy(idx.astype(np.int64)).to(dev)
        if world > 1:
            r, i = sharding.gather_sampler_rows(r, i)
        return r, i

    small = 1 << args.check_log2
    r_small, i_small = generate(small)
    ref_rows, ref_idx, _ = ctx.sample_stable_configs(seed, 0, small, max_ik_iter=5, cap=max(1024, small // 6), want_stage=False)   # unsharded, this rank alone
    same = bool(len(ref_idx) == len(i_small) and (i_small.cpu().numpy() == ref_idx.astype(np.int64)).all()
                and (r_small.cpu().numpy() == ref_rows).all())
    total = 1 << args.samples_log2
    generate(total >> 2)                                                    # warm-up
    barrier()
    t0 = time.perf_counter()
    rows, idx = generate(total)
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    emit(op="db_generation_sharded", n_gpus=world, samples=total, accepted=int(len(idx)), wall_s=dt, samples_per_s=total / dt,
         sharded_equals_unsharded=same, check_samples=small, scaling="strong",
         note="host API per rank (rows come back to the host, are re-uploaded and gathered with NCCL); accepted set identical "
              "to the single-GPU run by construction (global Philox counter)")

    # ---- V3, device-resident with the gather fused into the sampler: every rank appends its accepted rows straight
    # into rank 0's database (peer-mapped memory, system-scope atomic slot counter) - no host round trip, no collective
    total = 1 << args.samples_log2
    cap = total // 6
    blk_bytes = 256 + 8 * cap + 8 * 21 * cap                               # counter | index [cap] | rows [cap][21]
    root_ptr = ctx.dev_alloc(blk_bytes) if rank == 0 else None
    if world > 1:
        h = [ctx.ipc_export(root_ptr) if rank == 0 else None]
        dist.broadcast_object_list(h, src=0)
        base = root_ptr if rank == 0 else ctx.ipc_import(h[0])
    else:
        base = root_ptr
    p_cnt, p_idx, p_rows = base, base + 256, base + 256 + 8 * cap

    class _Raw:
        def __init__(self, ptr, shape, typestr):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (int(ptr), False), "version": 3}

    def generate_fused(n_total):
        if rank == 0:
            torch.as_tensor(_Raw(p_cnt, (1,), "<u8"), device=dev).zero_()
            torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        lo, hi = sharding.shard_bounds(n_total, world)[rank]
        ctx._check(ctx.lib.nwb_sample_stable_configs_dev(ctx.h, seed, lo, hi - lo, 5, p_rows, p_idx, cap, p_cnt, None),
                   "nwb_sample_stable_configs_dev")
        ctx.synchronize()
        barrier()
        return max_over_ranks(time.perf_counter() - t0)

    generate_fused(total >> 2)
    dt = generate_fused(total)
    fused_ok, n_acc = None, None
    if rank == 0:
        n_acc = int(torch.as_tensor(_Raw(p_cnt, (1,), "<u8"), device=dev).cpu().numpy().view(np.uint64)[0])
        g_idx = torch.as_tensor(_Raw(p_idx, (n_acc,), "<u8"), device=dev).cpu().numpy().view(np.uint64).astype(np.int64)
        g_rows = torch.as_tensor(_Raw(p_rows, (n_acc, 21), "<f8"), device=dev).cpu().numpy()
        order = np.argsort(g_idx, kind="stable")
        fused_ok = bool(len(idx) == n_acc and (g_idx[order] == idx.cpu().numpy()).all() and (g_rows[order] == rows.cpu().numpy()).all())
    emit(op="db_generation_fused_gather", n_gpus=world, samples=total, accepted=n_acc, wall_s=dt, samples_per_s=total / dt,
         equals_host_path=fused_ok, scaling="strong",
         note="device-resident: all ranks append into rank 0's arrays through peer-mapped memory inside the sampler kernel "
              "(system-scope atomic slot counter); rows compared with the host-API run above after sorting by sample index")
    if world > 1 and rank != 0:
        ctx.ipc_close(base)
    barrier()
    if rank == 0:
        ctx.dev_free(root_ptr)

    if not args.skip_nn:
        # ---- V2: one big tree, node-range shards ---------------------------------------------------------------------
        lo_j, hi_j = OL.joint_limits()
        lo_t, hi_t = torch.tensor(lo_j, device=dev), torch.tensor(hi_j, device=dev)
        BLK = 1 << 20
        n_blocks = 1 << max(0, args.nodes_log2 - 20)
        b0, b1 = sharding.shard_bounds(n_blocks, world)[rank]                 # contiguous blocks of 2^20 nodes per rank
        tree = nwb.Tree(ctx, max(1, (b1 - b0)) * BLK)
        for b in range(b0, b1):                                                 # block b is the same whatever the GPU count
            g = torch.Generator(device=dev)
            g.manual_seed(20121001 + b)
            blk = lo_t + (hi_t - lo_t) * torch.rand((BLK, 21), generator=g, device=dev, dtype=torch.float64)
            torch.cuda.synchronize()
            tree.add_dev(blk.data_ptr(), BLK)
            ctx.synchronize()
        nq = 64
        db = OL.golden_db()
        queries = db[np.floor(np.random.default_rng(20121002).random(nq) * (len(db) - 1)).astype(int)]
        q_dev = torch.from_numpy(queries).to(dev)
        idx_dev = torch.empty(nq, dtype=torch.int32, device=dev)
        dist_dev = torch.empty(nq, dtype=torch.float64, device=dev)

        def nn():
            if b1 > b0:
                tree.nearest_dev(q_dev.data_ptr(), nq, idx_dev.data_ptr(), dist_dev.data_ptr())
                ctx.synchronize()
                gidx = torch.where(idx_dev >= 0, idx_dev.to(torch.int64) + b0 * BLK, idx_dev.to(torch.int64))
                d = dist_dev
            else:                                                               # more ranks than blocks: an empty shard
                gidx = torch.full((nq,), -1, dtype=torch.int64, device=dev)
                d = torch.full((nq,), 10000.0, dtype=torch.float64, device=dev)
            if world > 1:
                return sharding.nn_reduce(d, gidx)
            return d, gidx

        nn()
        barrier()
        reps = 10
        t0 = time.perf_counter()
        for _ in range(reps):
            dmin, imin = nn()
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0) / reps
        total_nodes = n_blocks * BLK
        emit(op="nn_one_big_tree_sharded", n_gpus=world, nodes=total_nodes, queries=nq, wall_ms_per_batch=1e3 * dt,
             scan_gbs=84.0 * total_nodes * (nq / 8) / dt / 1e9, digest=int(imin.sum().item()), dist_digest=float(dmin.sum().item()),
             scaling="strong",
             note="64 queries (8 share a pass): each rank scans its node range, then all_gather + lexicographic min; host wall "
                  "clock per batch incl. the exchange; the digests (winning global indices / distances) must not depend on the "
                  "number of GPUs")
        tree.close()

    # ---- V4 ---------------------------------------------------------------------------------------------------
    G = args.grasp_grid
    n = G * G
    pos = np.array([[-0.08 + i * 0.4 / (G - 1), -0.30 + j * 0.4 / (G - 1), 0.2] for i in range(G) for j in range(G)])
    dirs = np.tile([0.0, -1.0, 0.0], (n, 1))
    seeds = (20121100 + np.arange(n)).astype(np.uint64)
    mine = sharding.round_robin(n, rank, world)

    def plan():
        res = ctx.compute_motion_plan_batch(pos[mine], dirs[mine], np.tile(CROUCH, (len(mine), 1)), seeds[mine], cap=2048)
        out = [(k, r["success"], r["num_nodes_generated"], r["trajectory"]) for k, r in zip(mine, res)]
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, out)
            out = sorted((x for p in parts for x in p), key=lambda x: x[0])
        return out

    plan()                                                                  # warm-up (allocations)
    barrier()
    t0 = time.perf_counter()
    out = plan()
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    digest = int(sum(int(s) * (k + 1) + int(nn) for k, s, nn, _ in out))    # compare across GPU counts
    emit(op=f"grasp_pose_queries_{n}_sharded", n_gpus=world, queries=n, solved=int(sum(s for _, s, _, _ in out)), wall_s=dt,
         digest=digest, scaling="strong",
         note="round-robin request -> rank, one batched compute_motion_plan call per rank, results gathered as objects; the "
              "digest (solved flags and node counts) must not depend on the number of GPUs")
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
