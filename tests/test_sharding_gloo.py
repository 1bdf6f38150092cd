"""world_size-2 `gloo` tests (CPU) of the multi-GPU host logic: shard bounds, ragged gathers, the
partition independence of the sampler (via the oracle) and the sharded nearest-neighbour reduction."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nao_wholebody_planning_b200 import sharding


def test_shard_bounds():
    assert sharding.shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert sharding.shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert sharding.shard_bounds(0, 2) == [(0, 0), (0, 0)]
    for n in (1, 7, 1 << 20, 1000003):
        for w in (1, 2, 4, 8):
            b = sharding.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
    assert sharding.round_robin(7, 1, 3) == [1, 4]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__)))
    import oracle_lib as OL
    orc = OL.Oracle()

    # 1. validity batch sharded by index: gather == unsharded result
    n = 3001
    q = OL.noisy_db_configs(n, seed=5)
    lo, hi = sharding.my_shard(n)
    v_local, _, _ = orc.is_feasible(q[lo:hi])
    full = sharding.gather_validity(torch.from_numpy(v_local), n)
    v_ref, _, _ = orc.is_feasible(q)
    assert (full.numpy() == v_ref).all()

    # 2. sampler: sample-index ranges, merged in ascending index == single-range run
    count, seed = 6000, 20121005
    lo, hi = sharding.my_shard(count)
    rows, idx, _ = orc.sample_stable_configs(seed, lo, hi - lo)
    g_rows, g_idx = sharding.gather_sampler_rows(torch.from_numpy(rows), torch.from_numpy(idx.astype(np.int64)))
    r_ref, i_ref, _ = orc.sample_stable_configs(seed, 0, count)
    assert (g_idx.numpy() == i_ref.astype(np.int64)).all() and (g_rows.numpy() == r_ref).all()

    # 3. NN over node ranges: lexicographic reduction == scan of the whole tree (with duplicates -> ties)
    nodes = OL.uniform_configs(4000, seed=9)
    nodes[3000:3100] = nodes[100:200]                      # exact duplicates across the shard boundary
    queries = np.concatenate([nodes[100:110], OL.golden_db()[:30]])
    lo, hi = sharding.my_shard(len(nodes))
    i_loc, d_loc = orc.nearest_neighbour(nodes[lo:hi], queries)
    i_glob = np.where(i_loc >= 0, i_loc + lo, -1).astype(np.int64)
    d, i = sharding.nn_reduce(torch.from_numpy(d_loc), torch.from_numpy(i_glob))
    i_ref, d_ref = orc.nearest_neighbour(nodes, queries)
    assert (i.numpy() == i_ref).all() and (d.numpy() == d_ref).all()

    # 4. planning requests: round-robin request -> rank, objects gathered and re-ordered == the serial loop
    poses = [[0.24, -0.065, 0.30, 0, -1, 0], [0.14, -0.065, 0.30, 0, -1, 0], [0.6, 0.0, 0.3, 0, -1, 0]]
    mine = sharding.round_robin(len(poses))
    part = [(k, orc.sample_goal_configs(20121100 + k, poses[k], max_samples=400)) for k in mine]
    parts = [None] * world
    dist.all_gather_object(parts, part)
    merged = sorted((x for p in parts for x in p), key=lambda x: x[0])
    assert [k for k, _ in merged] == [0, 1, 2]
    for k, (rows, erg, idx, drawn) in merged:
        r_ref, e_ref, i_ref, d_ref = orc.sample_goal_configs(20121100 + k, poses[k], max_samples=400)
        assert (rows == r_ref).all() and (idx == i_ref).all() and drawn == d_ref
    assert len(merged[2][1][0]) == 0                      # the unreachable pose finds no goal configuration

    # 4b. the same requests' trajectories as ONE packed buffer per rank (no pickled objects): request order restored
    n_req = 7
    trajs = {k: (np.full((k % 3 + (k != 4), 21), float(k)) if k != 4 else None) for k in range(n_req)}   # request 4: no plan
    mine = sharding.round_robin(n_req)
    rows_l, len_l = sharding.pack_trajectories([trajs[k] for k in mine])
    meta_l = np.array([[int(trajs[k] is not None), 100 + k] for k in mine], dtype=np.int64).reshape(-1, 2)
    rows, offsets, meta = sharding.gather_round_robin_trajectories(torch.from_numpy(rows_l), torch.from_numpy(len_l),
                                                                   torch.from_numpy(meta_l), n_req)
    assert offsets.tolist() == np.concatenate([[0], np.cumsum([0 if trajs[k] is None else len(trajs[k]) for k in range(n_req)])]).tolist()
    for k in range(n_req):
        seg = rows[offsets[k]:offsets[k + 1]].numpy()
        assert (len(seg) == 0 and (trajs[k] is None or len(trajs[k]) == 0)) or (seg == trajs[k]).all()
        assert meta[k].tolist() == [int(trajs[k] is not None), 100 + k]

    # 5. ragged gather with an empty rank
    local = torch.arange(5 if rank == 0 else 0, dtype=torch.float64).reshape(-1, 1)
    out, counts = sharding.all_gather_ragged(local)
    assert counts == [5, 0] and out.shape == (5, 1)
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")


def test_world_size_2_gloo(tmp_path):
    from oracle_lib import build_oracle
    build_oracle()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
