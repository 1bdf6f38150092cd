"""GPU parity: device tree + nearest-neighbour scan vs the oracle's findNearestNeighbour.
Bar: indices exact except on distance ties within 1e-6 (BASELINE.json north_star)."""
import numpy as np
import pytest

import nao_wholebody_planning_b200 as nwb
import oracle_lib as OL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = nwb.Context()
    yield c
    c.close()


def check(idx, dist, o_idx, o_dist, nodes, q):
    if np.abs(dist - o_dist).max() >= 1e-12:
        bad = np.nonzero(np.abs(dist - o_dist) >= 1e-12)[0]
        print("NN mismatches at queries", bad[:20], "got", idx[bad[:20]], dist[bad[:20]], "want", o_idx[bad[:20]], o_dist[bad[:20]])
    assert np.abs(dist - o_dist).max() < 1e-12
    bad = np.nonzero(idx != o_idx)[0]
    for k in bad:       # only exact-distance ties may differ (none expected: both take the lowest index)
        d_a = np.sqrt(((q[k] - nodes[idx[k]]) ** 2).sum())
        d_b = np.sqrt(((q[k] - nodes[o_idx[k]]) ** 2).sum())
        assert abs(d_a - d_b) < 1e-6
    assert len(bad) == 0


@pytest.mark.parametrize("n", [1, 2, 33, 1000, 10000, 100000])
def test_nn_matches_oracle(ctx, oracle, n):
    nodes = OL.uniform_configs(n, seed=100 + n, lhyp_zero=True)
    db = OL.golden_db()
    q = db[np.random.default_rng(n).integers(0, len(db) - 1, 64)]
    t = nwb.Tree(ctx, 16)
    t.add(nodes[: n // 2])
    t.add(nodes[n // 2:])          # growth path
    assert len(t) == n
    idx, dist = t.findNearestNeighbour(q)
    o_idx, o_dist = oracle.nearest_neighbour(nodes, q, threads=oracle.threads)
    check(idx, dist, o_idx, o_dist, nodes, q)
    back, pred = t.get()
    assert (back == nodes).all() and (pred == -1).all()
    t.close()


def test_nn_ties_and_duplicates_pick_the_lowest_index(ctx, oracle):
    rng = np.random.default_rng(1)
    base = OL.uniform_configs(500, seed=7)
    nodes = np.concatenate([base, base[::-1], base])            # every node three times
    q = np.concatenate([base[:40] + 1e-9, base[100:140]])        # queries on / next to nodes
    t = nwb.Tree(ctx)
    t.add(nodes, predecessor=np.arange(len(nodes), dtype=np.int32) - 1)
    idx, dist = t.findNearestNeighbour(q)
    o_idx, o_dist = oracle.nearest_neighbour(nodes, q)
    assert (idx == o_idx).all() and np.abs(dist - o_dist).max() < 1e-15
    assert (idx[40:] == np.arange(100, 140)).all() and (dist[40:] == 0).all()
    _, pred = t.get(10, 5)
    assert (pred == np.arange(9, 14)).all()
    t.close()


def test_nn_near_ties_need_the_exact_pass(ctx, oracle):
    """Nodes on a sphere of almost equal radius around the query: fp32 cannot order them."""
    rng = np.random.default_rng(2)
    q = OL.golden_db()[5]
    dirs = rng.normal(size=(20000, 21))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    radius = 0.7 + 1e-9 * rng.random(20000)
    nodes = q + dirs * radius[:, None]
    t = nwb.Tree(ctx)
    t.add(nodes)
    idx, dist = t.findNearestNeighbour(q[None])
    o_idx, o_dist = oracle.nearest_neighbour(nodes, q[None])
    assert idx[0] == o_idx[0] and dist[0] == o_dist[0]
    t.close()


@pytest.mark.parametrize("nq", [1, 16])
@pytest.mark.parametrize("order", ["descending", "ascending", "middle_first"])
def test_nn_three_near_ties_in_one_scan_thread(ctx, oracle, nq, order):
    """Three nodes that one scan thread evaluates one after the other, closer to the query than anything else and
    within the fp32 band of each other (the band is 1.2e-5 in distance, the tie tolerance 1e-6): whatever order
    their fp32 distances arrive in, the exact fp64 winner (lowest index on a tie) must come out.  Round 1 dropped
    a live candidate when a thread's two-slot list was pushed from the front."""
    n = 400_000
    nodes = OL.uniform_configs(n, seed=77, lhyp_zero=True)
    q = np.tile(OL.golden_db()[9], (nq, 1))
    q[1:] += 1e-3 * np.random.default_rng(5).normal(size=(nq - 1, 21))
    t = nwb.Tree(ctx, n)
    t.add(nodes)
    # three nodes of one owner
    first = 1234
    own = t.scan_owner(nq, first)
    same = [first]
    k = first + 1
    while len(same) < 3 and k < min(n, first + 200000):
        if t.scan_owner(nq, k) == own:
            same.append(k)
        k += 1
    if len(same) < 3:                       # a scan form with one node per thread: any three nodes will do
        same = [first, first + 4321, first + 98765]
    t.close()
    d = np.random.default_rng(6).normal(size=21)
    d[15] = 0.0
    d /= np.linalg.norm(d)
    # radii: distinct in fp32 (2e-6 apart in distance -> 1.2e-6 in d^2 at r = 0.3, 40 ulp) but inside the band
    radii = {"descending": [0.300004, 0.300002, 0.300000], "ascending": [0.300000, 0.300002, 0.300004],
             "middle_first": [0.300002, 0.300004, 0.300000]}[order]
    for node, r in zip(same, radii):
        rot = np.roll(d, node % 5)
        rot[15] = 0.0
        rot /= np.linalg.norm(rot)
        nodes[node] = q[0] + r * rot
    t = nwb.Tree(ctx, n)
    t.add(nodes)
    idx, dist = t.findNearestNeighbour(q)
    o_idx, o_dist = oracle.nearest_neighbour(nodes, q, threads=oracle.threads)
    assert o_idx[0] in same
    check(idx, dist, o_idx, o_dist, nodes, q)
    # and an exact three-way tie inside one thread: lowest index wins
    for node in same:
        nodes[node] = nodes[same[2]]
    t.clear()
    t.add(nodes)
    idx, dist = t.findNearestNeighbour(q)
    assert idx[0] == same[0]
    t.close()


def _bf16(x):
    import torch
    return torch.from_numpy(np.asarray(x, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def test_tensor_core_accumulator_matches_its_specification_and_error_bound(ctx):
    """The raw tcgen05 accumulator of the batched scan (first 128 nodes x up to 128 queries) against a NumPy evaluation
    of the same bf16 hi/mid split, and its distance to the TRUE |n|^2 - 2 q.n against the bound E the filter uses
    (nn_tensor.cu tc_error_bound): the filter is only as good as that bound."""
    nodes = OL.uniform_configs(1000, seed=5, lhyp_zero=True)
    q = np.concatenate([OL.golden_db()[:60], OL.uniform_configs(40, seed=6)])
    t = nwb.Tree(ctx)
    t.add(nodes)
    tile = t.debug_tensor_tile(q)
    t.close()
    n128 = nodes[:128]
    qf, nf = q.astype(np.float32), n128.astype(np.float32)
    qh, nh = _bf16(qf), _bf16(nf)
    qm, nm = _bf16(qf - qh), _bf16(nf - nh)
    nn = (n128 ** 2).sum(1).astype(np.float32)
    n0 = _bf16(nn)
    n1 = _bf16(nn - n0)
    n2 = _bf16(nn - n0 - n1)
    spec = (n0 + n1 + n2)[None, :].astype(np.float64) - 2.0 * (qh.astype(np.float64) @ nh.T.astype(np.float64)
                                                                + qh.astype(np.float64) @ nm.T.astype(np.float64)
                                                                + qm.astype(np.float64) @ nh.T.astype(np.float64))
    got = tile[:len(q)].astype(np.float64)
    scale = (n128 ** 2).sum(1)[None, :] + 2 * np.linalg.norm(q, axis=1)[:, None] * np.linalg.norm(n128, axis=1)[None, :]
    assert np.abs(got - spec).max() < 1e-5 * scale.max(), np.abs(got - spec).max()      # fp32 accumulation of 66 terms
    true = (n128 ** 2).sum(1)[None, :] - 2.0 * q @ n128.T
    nmax = np.sqrt((nodes ** 2).sum(1).max())
    qn = np.linalg.norm(q, axis=1) * nmax
    bound = 1.25 * (2.4e-5 * qn + 1.6e-5 * (nmax * nmax + 2.02 * qn) + 1e-7)
    ratio = np.abs(got - true).max(1) / bound
    print(f"\ntensor-core filter: max |v - true| / E = {ratio.max():.3f} (must stay well below 1)")
    assert ratio.max() < 0.5
    assert np.abs(tile[len(q):]).max() < 1e3                         # padding rows: |n|^2 only (no query), finite


def test_tensor_core_path_flood_of_near_ties(ctx, oracle):
    """More near ties than the candidate list of a CTA holds: the flooded tiles are evaluated exactly, pair by pair."""
    rng = np.random.default_rng(12)
    q = OL.golden_db()[20:84]                                            # 64 queries -> tensor-core path
    dirs = rng.normal(size=(6000, 21))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    nodes = q[0] + dirs * (0.5 + 1e-9 * rng.random(6000))[:, None]       # a shell around query 0
    nodes[100:140] = nodes[3000:3040]                                    # and exact duplicates
    t = nwb.Tree(ctx)
    t.add(nodes)
    idx, dist = t.findNearestNeighbour(q)
    t.close()
    o_idx, o_dist = oracle.nearest_neighbour(nodes, q, threads=oracle.threads)
    check(idx, dist, o_idx, o_dist, nodes, q)


@pytest.mark.parametrize("nq", [32, 127, 128, 129, 300])
def test_tensor_core_path_ragged_query_counts(ctx, oracle, nq):
    nodes = OL.noisy_db_configs(5000 + nq, seed=nq)                      # clustered data: many candidates per query
    q = OL.noisy_db_configs(nq, seed=1000 + nq, sigma=0.02)
    t = nwb.Tree(ctx)
    t.add(nodes[:1000])
    t.add(nodes[1000:])
    idx, dist = t.findNearestNeighbour(q)
    o_idx, o_dist = oracle.nearest_neighbour(nodes, q, threads=oracle.threads)
    check(idx, dist, o_idx, o_dist, nodes, q)
    t.clear()                                                            # stale rows past the new end must not be seen
    t.add(nodes[:77])
    idx, dist = t.findNearestNeighbour(q)
    o_idx, o_dist = oracle.nearest_neighbour(nodes[:77], q)
    check(idx, dist, o_idx, o_dist, nodes[:77], q)
    t.close()


def test_nn_randomised_stress_all_forms(ctx, oracle):
    """Random tree sizes / query counts / data shapes through every form of the operator (exact small-tree kernel, fp32
    stream, CUDA-core groups, tensor cores), trees grown in random pieces: always the oracle's answer."""
    rng = np.random.default_rng(2026)
    for case in range(28):
        n = int(rng.choice([1, 5, 127, 128, 129, 255, 257, 1000, 4097, 20000, 70001, 131073]))
        nq = int(rng.choice([1, 2, 9, 31, 32, 33, 127, 128, 129, 200, 257, 600]))
        kind = case % 3
        if kind == 0:
            nodes = OL.uniform_configs(n, seed=3000 + case, lhyp_zero=True)
            q = OL.uniform_configs(nq, seed=4000 + case)
        elif kind == 1:
            nodes = OL.noisy_db_configs(n, seed=3000 + case, sigma=0.03)
            q = OL.noisy_db_configs(nq, seed=4000 + case, sigma=0.01)
        else:                                    # heavy duplication: exact ties everywhere
            base = OL.uniform_configs(max(1, n // 7), seed=3000 + case)
            nodes = base[rng.integers(0, len(base), n)]
            q = base[rng.integers(0, len(base), nq)] + (rng.random((nq, 1)) < 0.5) * 1e-7
        t = nwb.Tree(ctx, 16)
        cut = sorted(rng.integers(0, n + 1, 2))
        for a, b in ((0, cut[0]), (cut[0], cut[1]), (cut[1], n)):
            if b > a:
                t.add(nodes[a:b])
        idx, dist = t.findNearestNeighbour(q)
        t.close()
        o_idx, o_dist = oracle.nearest_neighbour(nodes, q, threads=oracle.threads)
        assert (idx == o_idx).all() and np.abs(dist - o_dist).max() < 1e-12, (case, n, nq, kind)


@pytest.mark.parametrize("kind", ["shell", "ties", "scaled", "beyond_fp16"])
def test_single_query_stream_over_the_fp16_filter_copy(ctx, oracle, kind):
    """One query over a large tree streams the fp16 filter copy and re-evaluates the candidates in fp64: the answer is
    the oracle's whatever the fp16 rounding does.  shell: thousands of nodes inside the fp16 error shell around the
    nearest one (differences of 1e-6 ... 5e-3, far below an fp16 ulp); ties: exact duplicates of the nearest node at
    higher and lower indices; scaled: coordinates up to 200 (the error bound follows the tree's largest |node|);
    beyond_fp16: one coordinate of 1e5 is not representable - every node becomes a candidate and the exact rescan runs."""
    rng = np.random.default_rng({"shell": 1, "ties": 2, "scaled": 3, "beyond_fp16": 4}[kind])
    n = 150_000
    nodes = OL.uniform_configs(n, seed=77, lhyp_zero=True)
    q = OL.uniform_configs(6, seed=78)
    if kind == "shell":
        for k in range(6):                                     # 3000 nodes at distance 0.05 + (1e-6 ... 5e-3) from query k
            u = rng.normal(size=(3000, 21))
            u /= np.linalg.norm(u, axis=1, keepdims=True)
            r = 0.05 + 10.0 ** rng.uniform(-6, -2.3, 3000)
            nodes[rng.choice(n, 3000, replace=False)] = q[k] + u * r[:, None]
    elif kind == "ties":
        for k in range(6):
            near = q[k] + 1e-3 * rng.normal(size=21)
            nodes[rng.choice(n, 40, replace=False)] = near     # 40 exact copies: the lowest index must win
    elif kind == "scaled":
        nodes *= 60.0
        q = q * 60.0
    else:
        nodes[12345, 3] = 1.0e5
    t = nwb.Tree(ctx, 16)
    t.add(nodes[:70_000])
    t.add(nodes[70_000:])
    o_idx, o_dist = oracle.nearest_neighbour(nodes, q, threads=oracle.threads)
    for k in range(6):
        idx, dist = t.findNearestNeighbour(q[k:k + 1])
        assert idx[0] == o_idx[k] and abs(dist[0] - o_dist[k]) < 1e-12, (kind, k, idx[0], o_idx[k])
    t.close()


def test_nn_empty_tree_and_far_nodes(ctx):
    t = nwb.Tree(ctx)
    idx, dist = t.findNearestNeighbour(np.zeros((3, 21)))
    assert (idx == -1).all() and (dist == 10000.0).all()          # RP:347 quirk C4
    t.add(np.full((4, 21), 5000.0))                               # farther than best_norm's initial value
    idx, dist = t.findNearestNeighbour(np.zeros((1, 21)))
    assert idx[0] == -1
    t.clear()
    assert len(t) == 0
    t.close()


def test_nn_full_size_properties(ctx):
    """N = 1e6 (BASELINE config 3): every node is its own nearest neighbour; batch-split invariance."""
    n = 1_000_000
    nodes = OL.uniform_configs(n, seed=20121001, lhyp_zero=True)
    t = nwb.Tree(ctx, n)
    t.add(nodes)
    pick = np.random.default_rng(3).integers(0, n, 256)
    idx, dist = t.findNearestNeighbour(nodes[pick])
    assert (idx == pick).all() and (dist == 0).all()
    q = OL.golden_db()[:96]
    i1, d1 = t.findNearestNeighbour(q)
    i2 = np.concatenate([t.findNearestNeighbour(q[a:a + 7])[0] for a in range(0, 96, 7)])
    assert (i1 == i2).all()
    # property: no node is strictly closer than the reported one
    for k in range(0, 96, 13):
        d = np.sqrt(((nodes - q[k]) ** 2).sum(1))
        assert abs(d.min() - d1[k]) < 1e-12 and np.argmin(d) == i1[k]     # numpy sums pairwise: last-ulp difference
    t.close()
