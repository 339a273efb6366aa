"""Candidate sharding over the GPUs of one box.

The reference parallelises over receivers by running one `minimizer` process per group of receivers
and merging their text answers (python/tunguska/seismosizer.py:659-673, 785-827).  Here the
database, receivers and references are replicated in every GPU's HBM and the *candidates* of a
grid search are partitioned over the ranks (contiguous blocks, or dealt out in turn to even out the work); the only data that crosses NVLink is the small
per-candidate misfit block [ns_local, nmisfits, 2] (+ status), gathered with one all_gather.
"""
import numpy as np


def block_partition(n, world):
    """[(begin, end)] per rank: contiguous blocks, sizes differing by at most one."""
    base, rem = divmod(int(n), int(world))
    out, b = [], 0
    for r in range(world):
        e = b + base + (1 if r < rem else 0)
        out.append((b, e))
        b = e
    return out


def cyclic_partition(n, world):
    """index arrays per rank: candidate i goes to rank i mod world.  Grid searches list their candidates with the
    parameters varying in a fixed nesting, so cost (e.g. fault length -> number of sub-sources) varies slowly along the
    list; dealing the candidates out in turn evens the work out where contiguous blocks would not."""
    return [np.arange(r, int(n), int(world)) for r in range(int(world))]


def balanced_partition(costs, world):
    """index arrays per rank with (nearly) equal counts AND (nearly) equal summed cost: the candidates in order of decreasing
    cost each go to the rank with the smallest sum so far that still has room (longest-processing-time rule).  `costs`: any
    per-candidate proxy of the work, e.g. the fault area of a rupture (number of sub-sources)."""
    costs = np.asarray(costs, dtype=np.float64)
    n, world = costs.size, int(world)
    room = [(n + world - 1 - r) // world for r in range(world)]      # sizes differing by at most one
    total = np.zeros(world)
    out = [[] for _ in range(world)]
    for i in np.argsort(-costs, kind="stable"):
        r = min((r for r in range(world) if len(out[r]) < room[r]), key=lambda r: (total[r], r))
        out[r].append(int(i)); total[r] += costs[i]
    return [np.array(sorted(o), dtype=np.int64) for o in out]


def eval_sources_sharded(engine, sourcetype, params, group=None, device=None, partition="block"):
    """Evaluate params[ns, nparams] with the candidates split over the ranks of `group`
    (torch.distributed, NCCL on GPUs / gloo on CPU).  Every rank returns the full
    (misfits[ns, nmisfits, 2], status[ns]).  `engine` is this rank's Engine (one GPU)."""
    import torch
    import torch.distributed as dist
    p = np.ascontiguousarray(params, dtype=np.float32)
    if p.ndim == 1:
        p = p[None, :]
    ns = p.shape[0]
    if not (dist.is_available() and dist.is_initialized()):
        return engine.eval_sources(sourcetype, p)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if not isinstance(partition, str):      # explicit shares, e.g. from balanced_partition
        idx = [np.asarray(i, dtype=np.int64) for i in partition]
        if len(idx) != world or sorted(int(v) for i in idx for v in i) != list(range(ns)):
            raise ValueError("partition must give every candidate to exactly one of the %d ranks" % world)
    elif partition == "cyclic":
        idx = cyclic_partition(ns, world)
    elif partition == "block":
        idx = [np.arange(pb, pe) for pb, pe in block_partition(ns, world)]
    else:
        raise ValueError("partition must be 'block', 'cyclic' or a list of index arrays")
    mine = p[idx[rank]]
    b, e = 0, mine.shape[0]
    nm = engine.nmisfits
    width = max(len(i) for i in idx)                    # all_gather needs equal shapes: pad the short shares
    backend = dist.get_backend(group)
    dev = device if device is not None else ("cuda" if backend == "nccl" else "cpu")
    local = torch.zeros((width, nm * 2 + 1), dtype=torch.float32, device=dev)
    if e > b:
        if dev != "cpu" and hasattr(engine, "eval_sources_device"):
            block = torch.empty((e - b, nm, 2), dtype=torch.float32, device=dev)
            st = engine.eval_sources_device(sourcetype, mine, block.data_ptr())   # results never leave the GPU
            local[:e - b, :nm * 2] = block.reshape(e - b, nm * 2)
            local[:e - b, nm * 2] = torch.from_numpy(st.astype(np.float32)).to(dev)
        else:
            m, st = engine.eval_sources(sourcetype, mine)
            local[:e - b, :nm * 2] = torch.from_numpy(m.reshape(e - b, nm * 2)).to(dev)
            local[:e - b, nm * 2] = torch.from_numpy(st.astype(np.float32)).to(dev)
    gathered = torch.empty((world * width, nm * 2 + 1), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(gathered, local, group=group)
    g = gathered.cpu().numpy().reshape(world, width, nm * 2 + 1)
    mis = np.zeros((ns, nm * 2), np.float32)
    status = np.zeros(ns, np.int32)
    for r, i in enumerate(idx):
        mis[i] = g[r, :len(i), :nm * 2]
        status[i] = g[r, :len(i), nm * 2].astype(np.int32)
    return mis.reshape(ns, nm, 2), status
