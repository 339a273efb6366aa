"""Sharding over the GPUs of one box.

The reference parallelises over receivers by running one `minimizer` process per group of receivers
and merging their text answers (python/tunguska/seismosizer.py:659-673, 785-827).  Here the
database, receivers and references are replicated in every GPU's HBM and

* the *candidates* of a grid search are partitioned over the ranks (contiguous blocks, dealt out in turn, or balanced by
  a cost proxy); the only data that crosses NVLink is the small per-candidate misfit block [ns_local, nmisfits, 2]
  (+ status), gathered with one all_gather;
* when there are fewer candidates than ranks (a single evaluation on a dense array, the n+1 sources of a
  Levenberg-Marquardt Jacobian) the *receivers* are partitioned instead, contiguously by epicentral distance so that every
  GPU touches a compact distance range of the database -- the reference's balance method `112233`
  (seismosizer.py:785-827); the per-receiver misfit blocks are merged by receiver as seismosizer.py:659-673 merges the
  answers of its processes.
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


def receiver_partition(distances, enabled, world):
    """1-based receiver numbers per rank: the enabled receivers in order of epicentral distance, cut into `world` contiguous
    runs of (nearly) equal length (balance method `112233`, seismosizer.py:785-827); each run in ascending receiver number."""
    distances = np.asarray(distances, dtype=np.float64)
    on = np.flatnonzero(np.asarray(enabled, dtype=bool))
    order = on[np.argsort(distances[on], kind="stable")]
    return [np.sort(order[b:e]) + 1 for b, e in block_partition(order.size, world)]


def _resolve_partition(partition, ns, world, costs=None):
    if not isinstance(partition, str):      # explicit shares, e.g. from balanced_partition
        idx = [np.asarray(i, dtype=np.int64) for i in partition]
        if len(idx) != world or sorted(int(v) for i in idx for v in i) != list(range(ns)):
            raise ValueError("partition must give every candidate to exactly one of the %d ranks" % world)
        return idx
    if partition == "cyclic":
        return cyclic_partition(ns, world)
    if partition == "block":
        return [np.arange(pb, pe) for pb, pe in block_partition(ns, world)]
    if partition == "balanced":
        return balanced_partition(np.ones(ns) if costs is None else costs, world)
    raise ValueError("partition must be 'block', 'cyclic', 'balanced' or a list of index arrays")


class ShardedEngine:
    """This rank's Engine (one GPU) as a member of a torch.distributed group (NCCL on GPUs / gloo on CPU).

    eval_sources() returns the full (misfits[ns, nmisfits, 2], status[ns]) on every rank; eval_sources_device() leaves the
    gathered block on the device (torch tensors), so that only what the caller reads afterwards crosses PCIe."""

    def __init__(self, engine, group=None, device=None):
        import torch.distributed as dist
        self.engine, self.group = engine, group
        self.active = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.active else 1
        self.rank = dist.get_rank(group) if self.active else 0
        backend = dist.get_backend(group) if self.active else "gloo"
        if device is not None:
            self.device = device
        elif backend == "nccl":     # the engine's own GPU, not torch's current device: the raw pointer handed to the engine lives there
            self.device = "cuda:%d" % getattr(engine, "device", 0)
        else:
            self.device = "cpu"
        self._shard = None          # receiver numbers of this rank while the receivers are partitioned

    # ---- candidates over the ranks -----------------------------------------------------------------------------
    def eval_sources_device(self, sourcetype, params, partition="block", costs=None):
        """-> (misfits [ns, nmisfits, 2], status [ns]) as torch tensors on self.device, identical on every rank"""
        import torch
        import torch.distributed as dist
        p = np.ascontiguousarray(params, dtype=np.float32)
        if p.ndim == 1:
            p = p[None, :]
        ns, nm = p.shape[0], self.engine.nmisfits
        idx = _resolve_partition(partition, ns, self.world, costs)
        mine = p[idx[self.rank]]
        n = mine.shape[0]
        width = max(len(i) for i in idx)                    # all_gather needs equal shapes: short shares are padded
        dev = self.device
        local = torch.zeros((width, nm * 2 + 1), dtype=torch.float32, device=dev)
        if n > 0:
            if dev != "cpu" and hasattr(self.engine, "eval_sources_device"):
                block = torch.empty((n, nm, 2), dtype=torch.float32, device=dev)
                st = self.engine.eval_sources_device(sourcetype, mine, block.data_ptr())   # results never leave the GPU
                local[:n, :nm * 2] = block.reshape(n, nm * 2)
            else:
                m, st = self.engine.eval_sources(sourcetype, mine)
                local[:n, :nm * 2] = torch.from_numpy(np.ascontiguousarray(m).reshape(n, nm * 2)).to(dev)
            local[:n, nm * 2] = torch.from_numpy(st.astype(np.float32)).to(dev)
        if self.world > 1:
            gathered = torch.empty((self.world * width, nm * 2 + 1), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(gathered, local, group=self.group)
        else:
            gathered = local
        g = gathered.reshape(self.world, width, nm * 2 + 1)
        out = torch.zeros((ns, nm * 2 + 1), dtype=torch.float32, device=dev)
        for r, i in enumerate(idx):
            if len(i):
                out[torch.from_numpy(np.asarray(i, dtype=np.int64)).to(dev)] = g[r, :len(i)]
        return out[:, :nm * 2].reshape(ns, nm, 2), out[:, nm * 2].to(torch.int32)

    def eval_sources(self, sourcetype, params, partition="auto", costs=None):
        """-> (misfits[ns, nmisfits, 2], status[ns]) numpy, identical on every rank.  partition 'auto': candidates in
        contiguous blocks, or -- fewer candidates than ranks -- the receivers by distance."""
        p = np.ascontiguousarray(params, dtype=np.float32)
        if p.ndim == 1:
            p = p[None, :]
        if not self.active or self.world == 1:
            return self.engine.eval_sources(sourcetype, p)
        if isinstance(partition, str) and partition == "auto":
            if p.shape[0] < self.world:
                return self.eval_sources_by_receivers(sourcetype, p)
            partition = "block"
        m, st = self.eval_sources_device(sourcetype, p, partition, costs)
        return m.cpu().numpy(), st.cpu().numpy()

    # ---- receivers over the ranks -------------------------------------------------------------------------------
    def shard_receivers(self):
        """switch off the receivers of the other ranks (until unshard_receivers); -> the layout needed to merge the answers"""
        e = self.engine
        dist_m = e.get_distances()[0]
        enabled = np.asarray(e.enabled_receivers(), dtype=bool)
        ncomp = np.asarray(e.components_per_receiver(), dtype=np.int64)
        shares = receiver_partition(dist_m, enabled, self.world)
        base = np.zeros(enabled.size, np.int64)                  # first misfit pair of every receiver in the complete answer
        base[enabled] = np.cumsum(ncomp[enabled]) - ncomp[enabled]
        self._shard = dict(shares=shares, base=base, ncomp=ncomp, enabled=enabled, nm=int(ncomp[enabled].sum()))
        mine = set(int(i) for i in shares[self.rank])
        for ir in np.flatnonzero(enabled) + 1:
            if int(ir) not in mine:
                e.switch_receiver(int(ir), False)
        return self._shard

    def unshard_receivers(self):
        if self._shard is None:
            return
        for ir in np.flatnonzero(self._shard["enabled"]) + 1:
            self.engine.switch_receiver(int(ir), True)
        self._shard = None

    def eval_sources_by_receivers(self, sourcetype, params, keep_sharded=False):
        """every rank evaluates ALL candidates on its share of the receivers; the misfit blocks are merged by receiver
        (seismosizer.py:659-673).  keep_sharded: leave the receivers partitioned for the next call (optimiser loops)."""
        import torch
        import torch.distributed as dist
        p = np.ascontiguousarray(params, dtype=np.float32)
        if p.ndim == 1:
            p = p[None, :]
        ns = p.shape[0]
        if not self.active or self.world == 1:
            return self.engine.eval_sources(sourcetype, p)
        sh = self._shard or self.shard_receivers()
        try:
            nm_local = self.engine.nmisfits
            width = max(int(sh["ncomp"][s - 1].sum()) for s in sh["shares"]) if len(sh["shares"]) else 0
            local = torch.zeros((ns, width * 2 + 1), dtype=torch.float32, device=self.device)
            if nm_local > 0:
                m, st = self.engine.eval_sources(sourcetype, p)
                local[:, :nm_local * 2] = torch.from_numpy(np.ascontiguousarray(m).reshape(ns, nm_local * 2)).to(self.device)
                local[:, width * 2] = torch.from_numpy(st.astype(np.float32)).to(self.device)
            gathered = torch.empty((self.world * ns, width * 2 + 1), dtype=torch.float32, device=self.device)
            dist.all_gather_into_tensor(gathered, local, group=self.group)
            g = gathered.cpu().numpy().reshape(self.world, ns, width * 2 + 1)
        finally:
            if not keep_sharded:
                self.unshard_receivers()
        mis = np.zeros((ns, sh["nm"], 2), np.float32)
        status = np.zeros(ns, np.int32)
        for r, share in enumerate(sh["shares"]):
            col = 0
            for ir in share:                                 # the rank's answer lists its receivers in ascending number
                nc = int(sh["ncomp"][ir - 1]); b = int(sh["base"][ir - 1])
                mis[:, b:b + nc, :] = g[r, :, 2 * col:2 * (col + nc)].reshape(ns, nc, 2)
                col += nc
            if len(share):
                status = np.maximum(status, g[r, :, width * 2].astype(np.int32))
        return mis, status


def eval_sources_sharded(engine, sourcetype, params, group=None, device=None, partition="block"):
    """Evaluate params[ns, nparams] with the candidates split over the ranks of `group`
    (torch.distributed, NCCL on GPUs / gloo on CPU).  Every rank returns the full
    (misfits[ns, nmisfits, 2], status[ns]).  `engine` is this rank's Engine (one GPU)."""
    return ShardedEngine(engine, group, device).eval_sources(sourcetype, params, partition)
