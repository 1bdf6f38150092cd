"""Multi-GPU plumbing: how the independent units of the hot path are split over ranks and how the
results come back.  One process per GPU, `torch.distributed` (NCCL on the GPU
box, gloo in the CPU tests); the data path itself has no collective - only result gathers:

  validity / edge batches   contiguous index ranges, all_gather of the 1 B/config verdicts
  stable-config sampler     sample-index ranges (the Philox counter is the global sample index, so the
                            accepted set is independent of the partition); all_gather of the per-rank
                            counts, then of the padded rows; merged in ascending sample index
  NN over one big tree      node ranges; all_gather of (distance, global index) per query, then the
                            lexicographic minimum (ties -> lowest global index, as RP:377)
  planning queries          round-robin query -> rank, gather of the per-query results
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world):
    """Contiguous ranges of ceil(n / world) units: [(lo, hi)] * world (trailing ranks may be empty)."""
    per = -(-n // world) if world > 0 else n
    return [(min(r * per, n), min((r + 1) * per, n)) for r in range(world)]


def my_shard(n, rank=None, world=None):
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    return shard_bounds(n, world)[rank]


def round_robin(n, rank=None, world=None):
    """Indices of the units (planning queries) owned by this rank."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    return list(range(rank, n, world))


def all_gather_ragged(local, group=None):
    """Concatenate per-rank tensors of different length (same trailing shape) in rank order."""
    world = dist.get_world_size(group)
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(counts) if counts else 0
    padded = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0), counts


def gather_validity(valid_local, n_total, group=None):
    """Verdicts of a batch sharded with shard_bounds(n_total, world) -> full [n_total] tensor on every rank."""
    out, counts = all_gather_ragged(valid_local, group)
    assert out.shape[0] == n_total, (out.shape, n_total, counts)
    return out


def gather_sampler_rows(rows_local, index_local, group=None):
    """Accepted sampler rows of every rank merged in ascending sample index (the order the reference's
    serial loop writes its file in, stable_config_generator.cpp:463-468)."""
    rows, _ = all_gather_ragged(rows_local, group)
    idx, _ = all_gather_ragged(index_local, group)
    order = torch.argsort(idx, stable=True)
    return rows[order], idx[order]


def nn_reduce(dist_local, index_local_global, group=None):
    """Per-rank nearest neighbours (distance [nq], GLOBAL node index [nq], -1 = none) -> global winner.
    Lexicographic (distance, index): the lowest index wins ties, as the serial scan's strict < does."""
    world = dist.get_world_size(group)
    d_parts = [torch.empty_like(dist_local) for _ in range(world)]
    i_parts = [torch.empty_like(index_local_global) for _ in range(world)]
    dist.all_gather(d_parts, dist_local, group=group)
    dist.all_gather(i_parts, index_local_global, group=group)
    d = torch.stack(d_parts)                                  # [world, nq]
    i = torch.stack(i_parts).to(torch.int64)
    big = torch.iinfo(torch.int64).max
    key_i = torch.where(i < 0, torch.full_like(i, big), i)    # "none" loses every tie
    dmin = d.min(dim=0).values
    cand = torch.where(d == dmin, key_i, torch.full_like(key_i, big))
    best_i = cand.min(dim=0).values
    best_i = torch.where(best_i == big, torch.full_like(best_i, -1), best_i)
    return dmin, best_i.to(index_local_global.dtype)


def pack_trajectories(trajectories):
    """Per-request results -> one packed buffer: (rows [sum n_k, 21] float64, lengths [len] int64).  A request
    without a trajectory contributes length 0."""
    import numpy as np
    lengths = np.array([0 if t is None else len(t) for t in trajectories], dtype=np.int64)
    rows = (np.concatenate([np.asarray(t, dtype=np.float64).reshape(-1, 21) for t in trajectories if t is not None and len(t)])
            if lengths.sum() else np.zeros((0, 21)))
    return rows, lengths


def gather_round_robin_trajectories(rows_local, lengths_local, meta_local, n_total, group=None):
    """Planning requests sharded with round_robin(n_total): every rank contributes ONE packed row buffer, its
    per-request lengths and a small per-request int64 record (success flag, node count ...); two ragged
    all_gathers of numbers replace the pickled-object gather.  Returns, in request order, (rows [sum, 21],
    offsets [n_total + 1], meta [n_total, m]) on every rank."""
    world = dist.get_world_size(group)
    rows_all, row_counts = all_gather_ragged(rows_local, group)
    rec_local = torch.cat([lengths_local.reshape(-1, 1), meta_local.reshape(len(lengths_local), -1)], dim=1)
    rec_all, req_counts = all_gather_ragged(rec_local, group)
    m = rec_all.shape[1] - 1
    lengths = torch.zeros(n_total, dtype=torch.int64, device=rows_all.device)
    meta = torch.zeros((n_total, m), dtype=torch.int64, device=rows_all.device)
    src_start = torch.zeros(n_total, dtype=torch.int64, device=rows_all.device)
    row_base, req_base = 0, 0
    for r in range(world):
        owned = torch.arange(r, n_total, world, device=rows_all.device)            # round_robin(n_total, r, world)
        assert len(owned) == req_counts[r], (r, len(owned), req_counts[r])
        rec = rec_all[req_base:req_base + req_counts[r]]
        lengths[owned] = rec[:, 0]
        meta[owned] = rec[:, 1:]
        src_start[owned] = row_base + torch.cumsum(rec[:, 0], 0) - rec[:, 0]
        row_base += row_counts[r]
        req_base += req_counts[r]
    offsets = torch.zeros(n_total + 1, dtype=torch.int64, device=rows_all.device)
    offsets[1:] = torch.cumsum(lengths, 0)
    # permute the rows into request order with one gather
    total = int(offsets[-1].item())
    if total:
        req_of_row = torch.repeat_interleave(torch.arange(n_total, device=rows_all.device), lengths)
        within = torch.arange(total, device=rows_all.device) - offsets[:-1][req_of_row]
        rows = rows_all[src_start[req_of_row] + within]
    else:
        rows = rows_all[:0]
    return rows, offsets, meta


def bind_to_gpu_numa_node(pci_bus_id):
    """Pin this process (and therefore its later page allocations: Linux places pages on the node of the allocating
    CPU) to the CPUs of the NUMA node the GPU hangs off, so that pinned staging buffers are NUMA-local to the GPU's
    PCIe root.  Returns (node, n_cpus) or None when the topology is not exposed.  Call BEFORE allocating pinned memory."""
    import os
    try:
        dom = pci_bus_id.lower()
        if len(dom.split(":")[0]) == 8:                       # CUDA prints an 8-digit domain, sysfs uses 4
            dom = dom[4:]
        node = int(open(f"/sys/bus/pci/devices/{dom}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node, len(cpus)
    except (OSError, ValueError):
        return None
