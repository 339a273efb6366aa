"""Brute-force grid search with built-in bootstrapping on the batched engine: the driver the reference has in
python/tunguska/gridsearch.py (`MisfitGrid`: compute / postprocess / stats, plotting left out), built anew on
`Engine.eval_sources_on_device` + `Engine.outer_misfits`.

Where the reference evaluates the grid one source at a time through the `minimizer` text pipe
(seismosizer.py:682-722) and re-reduces the whole misfit cube in numpy once per bootstrap iteration
(gridsearch.py:269-283), here the grid is one batched call (optionally sharded over ranks,
kiwi_b200.sharding) and all bootstrap realisations are reduced on the device in one call; only the
global misfits and the indices of the best sources come back.
"""
import numpy as np

from .engine import SOURCE_TYPES

# parameter names of the source types (psm_param_names_* of source_bilat.f90, source_circular.f90, source_point_lp.f90,
# source_eikonal.f90, source_mt_eikonal.f90, source_moment_tensor.f90), in parameter order
PARAM_NAMES = {
    "bilateral": ["time", "north-shift", "east-shift", "depth", "moment", "strike", "dip", "slip-rake", "rupture-rake", "length-a", "length-b",
                  "width", "rupture-velocity", "rise-time"],
    "circular": ["time", "north-shift", "east-shift", "depth", "moment", "strike", "dip", "slip-rake", "radius", "rupture-velocity", "rise-time"],
    "point_lp": ["time", "north-shift", "east-shift", "depth", "moment", "m_xx", "m_yy", "m_zz", "m_xy", "m_xz", "m_yz", "excitation-time",
                 "main-period"],
    "eikonal": ["time", "north-shift", "east-shift", "depth", "moment", "strike", "dip", "slip-rake", "bord-shift-x", "bord-shift-y", "bord-radius",
                "nukl-shift-x", "nukl-shift-y", "rel-rupture-velocity", "rise-time"],
    "mt_eikonal": ["time", "north-shift", "east-shift", "depth", "moment-factor", "strike", "dip", "bord-shift-x", "bord-shift-y", "bord-radius",
                   "nukl-shift-x", "nukl-shift-y", "rel-rupture-velocity", "mxx", "myy", "mzz", "mxy", "mxz", "myz", "rise-time"],
    "moment_tensor": ["time", "north-shift", "east-shift", "depth", "mxx", "myy", "mzz", "mxy", "mxz", "myz", "rise-time"],
}


def mimainc_to_gvals(mi, ma, inc):
    """grid values from (min, max, increment) as the reference's driver makes them (python/tunguska/gridsearch.py:18-22): the number of
    steps is rounded and the increment re-derived, so that both ends of the range are grid values"""
    lo, hi = float(mi), float(ma)
    nsteps = int(round((hi - lo) / float(inc)))
    if nsteps == 0:
        return np.array([lo])
    return lo + np.arange(nsteps + 1) * ((hi - lo) / nsteps)


def step_at(values, value):
    """width of the grid step that contains `value` (python/tunguska/gridsearch.py:24-27); 1 for a single grid value"""
    if len(values) < 2:
        return 1.
    k = min(max(int(np.searchsorted(values, value)), 1), len(values) - 1)
    return values[k] - values[k - 1]


def source_grid(sourcetype, base_params, param_values):
    """Source.grid (source.py:119-175): every combination of the listed values, the first parameter varying slowest."""
    names = PARAM_NAMES[sourcetype]
    base = np.asarray(base_params, dtype=np.float32)
    if base.size != len(names):
        raise ValueError("%s takes %d parameters" % (sourcetype, len(names)))
    cols = [names.index(p) for p, _ in param_values]
    axes = [np.asarray(v, dtype=np.float64) for _, v in param_values]
    if not axes:
        return base[None, :].copy()
    mesh = np.stack(np.meshgrid(*axes, indexing="ij"), -1).reshape(-1, len(axes))
    grid = np.tile(base, (mesh.shape[0], 1))
    for k, c in enumerate(cols):
        grid[:, c] = mesh[:, k].astype(np.float32)
    return grid


class MisfitGridStats:
    """gridsearch.py:44-109: best value and bootstrap distribution of one parameter; the 68 % interval is widened by half a grid step"""

    def __init__(self, paramname, best, distribution, tested_values=None):
        self.paramname, self.best = paramname, best
        self.distribution = np.asarray(distribution, dtype=float)
        self.tested_values = tested_values
        self.mean, self.std, self.median = np.mean(self.distribution), np.std(self.distribution), np.median(self.distribution)
        self.percentile16 = float(np.percentile(self.distribution, 16.))    # scipy.stats.scoreatpercentile: linear interpolation
        self.percentile84 = float(np.percentile(self.distribution, 84.))
        if tested_values is not None:
            self.percentile16 -= step_at(tested_values, self.percentile16) / 2.
            self.percentile84 += step_at(tested_values, self.percentile84) / 2.
            self.percentile16_warn = bool(self.percentile16 < np.min(tested_values))
            self.percentile84_warn = bool(self.percentile84 > np.max(tested_values))
        else:
            self.percentile16_warn = self.percentile84_warn = False

    def str_best_and_confidence(self, factor=1., unit=''):
        """one line in the wording of the reference's reports (python/tunguska/gridsearch.py:66-73)"""
        marks = (' (?)' if self.percentile16_warn else '', '(?) ' if self.percentile84_warn else '')
        head = '%s = %.3g %s' % (self.paramname.title(), self.best * factor, unit)
        interval = '[ %.3g%s, %.3g %s]' % (self.percentile16 * factor, marks[0], self.percentile84 * factor, marks[1])
        return head + '  (confidence interval 68%) = ' + interval + ' ' + unit


def bootstrap_weights(enabled, weights, iterations, rng):
    """seismosizer.py:853-877: per realisation, as many receivers as there are usable ones (enabled and with non-zero weight) are
    drawn from them with replacement; the weight of a receiver is the number of times it was drawn"""
    mask = np.asarray(enabled, dtype=bool)
    if weights is not None:
        mask = mask & (np.asarray(weights) != 0)
    idx = np.arange(mask.size)[mask]
    bw = np.zeros((iterations, mask.size))
    for b in range(iterations):
        if idx.size:
            bw[b] = np.bincount(idx[rng.integers(0, idx.size, idx.size)], minlength=mask.size)
    return bw


def global_misfit_of_one_source(block, components, enabled, receiver_weights=None, outer_norm="l2norm", anarchy=False):
    """make_global_misfits (seismosizer.py:879-920) for a single source from its misfit block [nmisfits][2] (enabled receivers only,
    receiver-major): used for the reference source, whose block is a few hundred numbers on the host."""
    nr = len(components)
    m, n = np.zeros(nr), np.zeros(nr)
    k = 0
    for ir in range(nr):
        if not enabled[ir]:
            continue
        nc = len(components[ir])
        mm, nn = block[k:k + nc, 0].astype(float), block[k:k + nc, 1].astype(float)
        k += nc
        if outer_norm == "l1norm":
            m[ir], n[ir] = mm.sum(), nn.sum()
        else:
            m[ir], n[ir] = np.sqrt((mm ** 2).sum()), np.sqrt((nn ** 2).sum())
    w = np.ones(nr) if receiver_weights is None else np.asarray(receiver_weights, dtype=float).copy()
    if anarchy:
        w = np.maximum(w / np.where(n != 0., n, -1.), 0.)
    if outer_norm == "l1norm":
        ms, ns = (m * w).sum(), (n * w).sum()
        return float(ms / ns) if ns > 0. else float("nan")
    ms, ns = ((m * w) ** 2).sum(), ((n * w) ** 2).sum()
    return float(np.sqrt(ms / ns)) if ns > 0. else float("nan")


def merge_best(parts):
    """parts[rank] = [[global candidate number or -1 per realisation], [its misfit or NaN]] -> (best, value) per realisation over all
    ranks: the smallest non-NaN misfit, the lowest candidate number among equals (what nanargmin over the whole grid returns)"""
    parts = np.asarray(parts, dtype=np.float64)
    idx, val = parts[:, 0, :], parts[:, 1, :]
    val = np.where((idx >= 0) & np.isfinite(val), val, np.inf)
    key_idx = np.where(np.isfinite(val), idx, np.inf)
    order = np.lexsort((key_idx, val), axis=0)[0]                       # per realisation: by misfit, then by candidate number
    cols = np.arange(idx.shape[1])
    best = np.where(np.isfinite(val[order, cols]), idx[order, cols], -1).astype(np.int64)
    bestv = np.where(best >= 0, val[order, cols], np.nan)
    return best, bestv


def _world(group):
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(group), dist.get_rank(group)
    except ImportError:
        pass
    return 1, 0


def _all_gather(arr, group):
    """[world, ...] of equally shaped float64 arrays (NCCL: through the GPU)"""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)).to(dev)
    world = dist.get_world_size(group)
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=torch.float64, device=dev)      # concatenated along the first axis
    dist.all_gather_into_tensor(out, t, group=group)
    return out.cpu().numpy().reshape((world,) + tuple(t.shape))


def _gather_rows(local, shares, ns, group):
    """local[k][n_local] per rank -> [k][ns] with every rank's columns at its candidates' places"""
    width = max(len(i) for i in shares)
    pad = np.full((local.shape[0], width), np.nan)
    pad[:, :local.shape[1]] = local
    g = _all_gather(pad, group)
    out = np.full((local.shape[0], ns), np.nan)
    for r, i in enumerate(shares):
        out[:, i] = g[r][:, :len(i)]
    return out


class MisfitGrid:
    """Brute force grid search minimizer with built-in bootstrapping (gridsearch.py:112-305).

        grid = MisfitGrid("bilateral", base_params, param_ranges=[("strike", 60., 120., 10.), ("depth", 2e3, 6e3, 1e3)])
        grid.compute(engine)                               # one batched evaluation, misfit cube stays on the GPU
        grid.postprocess(outer_norm="l2norm", bootstrap_iterations=1000)
        grid.best_source, grid.stats["strike"].str_best_and_confidence()
    """

    def __init__(self, sourcetype, base_params, param_ranges=None, param_values=None, ref_params=None):
        if sourcetype not in SOURCE_TYPES:
            raise ValueError("unknown source type name: %s" % sourcetype)
        self.sourcetype = sourcetype
        self.base_source = np.asarray(base_params, dtype=np.float32).copy()
        self.ref_source = self.base_source.copy() if ref_params is None else np.asarray(ref_params, dtype=np.float32).copy()
        if param_values is not None:
            self.param_values = [(p, np.asarray(v, dtype=float)) for p, v in param_values]
        else:
            self.param_values = [(p, mimainc_to_gvals(mi, ma, inc)) for p, mi, ma, inc in (param_ranges or [])]
        self.sources = source_grid(sourcetype, self.base_source, self.param_values)
        self.sourceparams = [p for p, _ in self.param_values]
        self.status = self.ref_misfit = self.best_source = self.misfits_by_s = self.bootstrap_sources = self.stats = None
        self._engine = None

    def compute(self, engine, group=None, costs=None):
        """Let the engine calculate the trace misfits (gridsearch.py:159-203): the whole grid in one batched call.  With an
        initialised torch.distributed process group (one rank per GPU) every rank evaluates its share of the grid
        (kiwi_b200.sharding.balanced_partition over `costs`, default equal costs); postprocess() then exchanges only per-candidate
        global misfits and per-realisation best candidates."""
        self._engine = engine
        ref_block, ref_status = engine.eval_sources(self.sourcetype, self.ref_source)
        self._ref_block = ref_block.astype(np.float64)
        self._group, self._share = group, None
        world, rank = _world(group)
        if world > 1:
            from .sharding import balanced_partition
            c = np.ones(self.sources.shape[0]) if costs is None else np.asarray(costs, dtype=float)
            self._shares = balanced_partition(c, world)
            self._share = self._shares[rank]
        mine = self.sources if self._share is None else np.ascontiguousarray(self.sources[self._share])
        status = engine.eval_sources_on_device(self.sourcetype, mine) if mine.shape[0] else np.zeros(0, np.int32)   # failed: status != 0, NaN misfits
        self._nlocal = mine.shape[0]
        self.status = status if self._share is None else _gather_rows(status.astype(np.float64)[None, :], self._shares, self.sources.shape[0],
                                                                   group)[0].astype(np.int32)
        self.best_source = self.misfits_by_s = self.bootstrap_sources = self.stats = None

    def postprocess(self, receiver_weights=None, outer_norm="l2norm", anarchy=False, bootstrap_iterations=1000, seed=0, enabled=None):
        """Combine trace misfits to global misfits, find the best source, make statistics (gridsearch.py:205-219).
        enabled: receiver mask (default: the receivers enabled in the engine).  seed: of the bootstrap draws (all ranks of a sharded
        search must use the same)."""
        e = self._engine
        if e is None:
            raise RuntimeError("compute() first")
        mask = np.asarray(e._enabled, dtype=bool) if enabled is None else (np.asarray(enabled, dtype=bool) & np.asarray(e._enabled, dtype=bool))
        w = None if receiver_weights is None else np.asarray(receiver_weights, dtype=np.float64)
        rng = np.random.default_rng(seed)
        nb = int(bootstrap_iterations)
        ns = self.sources.shape[0]
        kw = dict(receiver_weights=w, outer_norm=outer_norm, anarchy=anarchy)
        # two reductions of the cube on the device: the plain misfits of all candidates, and the best candidate of every bootstrap
        # realisation (the [realisations x candidates] matrix never leaves the GPU)
        if self._nlocal:
            plain, best0, bestv0 = e.outer_misfits(ns=self._nlocal, want_matrix=True, **kw)
            if nb > 0:
                _, bestb, bestvb = e.outer_misfits(ns=self._nlocal, bweights=bootstrap_weights(mask, w, nb, rng), want_matrix=False, **kw)
                best, bestv = np.concatenate([best0[:1], bestb[1:]]), np.concatenate([bestv0[:1], bestvb[1:]])
            else:
                best, bestv = best0[:1], bestv0[:1]
        else:
            plain, best, bestv = np.zeros((1, 0)), np.full(nb + 1, -1, np.int32), np.full(nb + 1, np.nan)
        if self._share is not None:      # sharded: local -> global candidate numbers, then the best over the ranks
            gbest = np.where(best >= 0, self._share[np.maximum(best, 0)] if self._nlocal else -1, -1).astype(np.float64)
            plain = _gather_rows(plain[:1], self._shares, ns, self._group)
            best, bestv = merge_best(_all_gather(np.stack([gbest, bestv]), self._group))
        self.ref_misfit = global_misfit_of_one_source(self._ref_block[0], e._components, [a and b for a, b in zip(e._enabled, mask)], **kw)
        self.misfits_by_s = plain[0]
        ibest = int(best[0]) if best[0] >= 0 else 0                      # gridsearch.py:255-257
        self.best_source = self.sources[ibest].copy()
        self.best_misfit = float(bestv[0])
        self.bootstrap_sources = self.sources[np.maximum(best[1:], 0)] if nb > 0 else self.sources[:0]
        names = PARAM_NAMES[self.sourcetype]
        self.stats = {}
        for p, gvalues in self.param_values:                             # gridsearch.py:285-294
            col = names.index(p)
            dist = self.bootstrap_sources[:, col].astype(float) if nb > 0 else np.array([self.best_source[col]], dtype=float)
            self.stats[p] = MisfitGridStats(p, float(self.best_source[col]), dist, tested_values=gvalues)
        return self.best_source

    def get_best_misfit(self):
        return float(np.nanmin(self.misfits_by_s))
