"""Candidate sharding over the GPUs of one box.

The reference parallelises over receivers by running one `minimizer` process per group of receivers
and merging their text answers (python/tunguska/seismosizer.py:659-673, 785-827).  Here the
database, receivers and references are replicated in every GPU's HBM and the *candidates* of a
grid search are block-partitioned over the ranks; the only data that crosses NVLink is the small
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


def eval_sources_sharded(engine, sourcetype, params, group=None, device=None):
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
    parts = block_partition(ns, world)
    b, e = parts[rank]
    nm = engine.nmisfits
    width = max(pe - pb for pb, pe in parts)            # all_gather needs equal shapes: pad the short blocks
    backend = dist.get_backend(group)
    dev = device if device is not None else ("cuda" if backend == "nccl" else "cpu")
    local = torch.zeros((width, nm * 2 + 1), dtype=torch.float32, device=dev)
    if e > b:
        if dev != "cpu" and hasattr(engine, "eval_sources_device"):
            block = torch.empty((e - b, nm, 2), dtype=torch.float32, device=dev)
            st = engine.eval_sources_device(sourcetype, p[b:e], block.data_ptr())   # results never leave the GPU
            local[:e - b, :nm * 2] = block.reshape(e - b, nm * 2)
            local[:e - b, nm * 2] = torch.from_numpy(st.astype(np.float32)).to(dev)
        else:
            m, st = engine.eval_sources(sourcetype, p[b:e])
            local[:e - b, :nm * 2] = torch.from_numpy(m.reshape(e - b, nm * 2)).to(dev)
            local[:e - b, nm * 2] = torch.from_numpy(st.astype(np.float32)).to(dev)
    gathered = torch.empty((world * width, nm * 2 + 1), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(gathered, local, group=group)
    g = gathered.cpu().numpy().reshape(world, width, nm * 2 + 1)
    mis = np.concatenate([g[r, :pe - pb, :nm * 2] for r, (pb, pe) in enumerate(parts)], 0).reshape(ns, nm, 2)
    status = np.concatenate([g[r, :pe - pb, nm * 2] for r, (pb, pe) in enumerate(parts)], 0).astype(np.int32)
    return mis, status
