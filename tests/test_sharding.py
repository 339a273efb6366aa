"""Multi-rank host logic on CPU: world_size 2, gloo backend.  The per-rank evaluator is the CPU
oracle (allowed in tests); the thing under test is the partition / pad / all_gather / reassembly of
kiwi_b200.sharding, which is what the GPU ranks run over NCCL."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_block_partition():
    from kiwi_b200.sharding import block_partition
    assert block_partition(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert block_partition(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert block_partition(0, 2) == [(0, 0), (0, 0)]
    for n in (1, 7, 64, 1000):
        for w in (1, 2, 4, 8):
            p = block_partition(n, w)
            assert p[0][0] == 0 and p[-1][1] == n and all(a[1] == b[0] for a, b in zip(p, p[1:]))
            sizes = [e - b for b, e in p]
            assert max(sizes) - min(sizes) <= 1


def test_cyclic_partition():
    from kiwi_b200.sharding import cyclic_partition
    assert [list(i) for i in cyclic_partition(7, 3)] == [[0, 3, 6], [1, 4], [2, 5]]
    assert [list(i) for i in cyclic_partition(2, 4)] == [[0], [1], [], []]
    for n in (0, 1, 9, 64):
        for w in (1, 2, 8):
            parts = cyclic_partition(n, w)
            assert sorted(int(v) for i in parts for v in i) == list(range(n))
            assert max(len(i) for i in parts) - min(len(i) for i in parts) <= 1


def test_balanced_partition():
    from kiwi_b200.sharding import balanced_partition
    # a sweep whose fastest parameter alternates between a cheap and an expensive value: dealing out in turn would give
    # one of two ranks all the expensive ones
    costs = np.tile([30.0, 50.0], 8)
    parts = balanced_partition(costs, 2)
    assert [len(i) for i in parts] == [8, 8] and [float(costs[i].sum()) for i in parts] == [320.0, 320.0]
    rng = np.random.default_rng(1)
    for n, w in ((64, 8), (7, 3), (5, 8), (0, 2)):
        c = rng.uniform(1.0, 2.0, n)
        parts = balanced_partition(c, w)
        assert sorted(int(v) for i in parts for v in i) == list(range(n))
        assert max(len(i) for i in parts) - min(len(i) for i in parts) <= 1
        if n == 64:
            sums = [c[i].sum() for i in parts]
            assert max(sums) / min(sums) < 1.02


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import scenario as sc
    from oracle_lib import OracleEngine
    from kiwi_b200.sharding import eval_sources_sharded
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    comps = ["ned", "d", "ar"]
    lat, lon, dep = sc.small_receivers(3)
    o = OracleEngine(threads=1)
    sc.setup(o, sc.small_db(), lat, lon, dep, comps)
    o.eval_sources("bilateral", sc.BILAT_SMALL)
    sc.set_refs_from(o, [o], [3, 1, 2])
    p = np.tile(sc.BILAT_SMALL, (5, 1)); p[:, 5] += np.arange(5) * 7.0      # 5 candidates over 2 ranks: blocks of 3 and 2
    mis, st = eval_sources_sharded(o, "bilateral", p)
    misc, stc = eval_sources_sharded(o, "bilateral", p, partition="cyclic")      # ranks get candidates 0,2,4 and 1,3
    misb, stb = eval_sources_sharded(o, "bilateral", p, partition=[[4, 0], [1, 2, 3]])
    assert np.array_equal(misb, misc) and np.array_equal(stb, stc)
    ref, rst = o.eval_sources("bilateral", p)
    # fewer candidates than ranks: the receivers are partitioned by distance instead, the answers merged by receiver
    from kiwi_b200.sharding import ShardedEngine, receiver_partition
    se = ShardedEngine(o)
    one, ost = se.eval_sources("bilateral", p[:1])
    assert np.array_equal(one, ref[:1]) and np.array_equal(ost, rst[:1]) and o.enabled_receivers() == [True, True, True]
    shares = receiver_partition(o.get_distances()[0], [True] * 3, 2)
    assert sorted(int(i) for s_ in shares for i in s_) == [1, 2, 3] and [len(s_) for s_ in shares] == [2, 1]
    o.switch_receiver(2, False)                     # a disabled receiver stays out of the partition and of the answer
    two, _ = se.eval_sources("bilateral", p[:1])
    ref2, _ = o.eval_sources("bilateral", p[:1])
    assert two.shape == (1, 5, 2) and np.array_equal(two, ref2) and o.enabled_receivers() == [True, False, True]
    o.switch_receiver(2, True)
    lm1, _ = se.eval_sources_by_receivers("bilateral", p[:3], keep_sharded=True)      # optimiser loop: the partition stays
    lm2, _ = se.eval_sources_by_receivers("bilateral", p[:3], keep_sharded=True)
    se.unshard_receivers()
    assert np.array_equal(lm1, ref[:3]) and np.array_equal(lm2, ref[:3]) and o.enabled_receivers() == [True, True, True]
    q.put((rank, bool(np.array_equal(mis, ref) and np.array_equal(misc, ref)), bool(np.array_equal(st, rst) and np.array_equal(stc, rst)), mis.shape))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_evaluation_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same_m, same_s, shape in res:
        assert same_m and same_s and shape == (5, 6, 2), (rank, same_m, same_s, shape)
