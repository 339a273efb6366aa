"""TEST INFRASTRUCTURE ONLY.  numpy restatement of make_global_misfits
(python/tunguska/seismosizer.py:843-922; the module itself is Python 2 + pyrocko and cannot be imported).
The bootstrap multiplicities are an input here (the reference draws them with num.random inside, :863-877)."""
import numpy as num


def cube_from_block(block, enabled, ncomps):
    """[ns, nmisfits, 2] block of get_misfits (enabled receivers only) -> misfits_by_src, norms_by_src
    [ns, nreceivers, ncomponents] as make_misfits_for_sources builds them (seismosizer.py:682-722):
    zero padded, zeros for disabled receivers, zeros for failed sources."""
    ns = block.shape[0]
    nr, nc = len(ncomps), max(ncomps)
    m = num.zeros((ns, nr, nc)); n = num.zeros((ns, nr, nc))
    k = 0
    for ir in range(nr):
        if not enabled[ir]:
            continue
        for ic in range(ncomps[ir]):
            m[:, ir, ic] = block[:, k, 0]; n[:, ir, ic] = block[:, k, 1]
            k += 1
    bad = ~num.isfinite(block.reshape(ns, -1)).all(axis=1)
    m[bad] = 0.; n[bad] = 0.
    return m, n


def make_global_misfits(misfits_by_src, norms_by_src, receiver_weights=1., outer_norm='l2norm', anarchy=False, bweights=None):
    if isinstance(receiver_weights, float):
        rweights = receiver_weights
    else:
        rweights = receiver_weights[num.newaxis, :].copy()
    if outer_norm == 'l1norm':                                       # :879-900
        misfits_by_sr = num.sum(misfits_by_src, 2)
        norms_by_sr = num.sum(norms_by_src, 2)
        if anarchy:
            xrweights = num.zeros(norms_by_sr.shape, dtype=float)
            xrweights[:, :] = rweights
            xrweights /= num.where(norms_by_sr != 0., norms_by_sr, -1.)
            rweights = num.maximum(xrweights, 0.)
        if bweights is not None:
            rweights = rweights * bweights
        misfits_by_sr = misfits_by_sr * rweights
        norms_by_sr = norms_by_sr * rweights
        ms = num.sum(misfits_by_sr, 1)
        ns = num.sum(norms_by_sr, 1)
        with num.errstate(all='ignore'):
            misfits_by_s = num.where(ns > 0., ms / ns, -1.)
        misfits_by_s = num.where(misfits_by_s < 0., num.nan, misfits_by_s)
    elif outer_norm == 'l2norm':                                     # :902-920
        misfits_by_sr = num.sqrt(num.sum(misfits_by_src ** 2, 2))
        norms_by_sr = num.sqrt(num.sum(norms_by_src ** 2, 2))
        if anarchy:
            rweights = rweights / num.where(norms_by_sr != 0., norms_by_sr, -1.)
            rweights = num.maximum(rweights, 0.)
        if bweights is not None:
            rweights = rweights * num.sqrt(bweights)
        misfits_by_sr = misfits_by_sr * rweights
        norms_by_sr = norms_by_sr * rweights
        ms = num.sum((misfits_by_sr) ** 2, 1)
        ns = num.sum((norms_by_sr) ** 2, 1)
        with num.errstate(all='ignore'):
            misfits_by_s = num.where(ns > 0., num.sqrt(ms / ns), -1.)
        misfits_by_s = num.where(misfits_by_s < 0, num.nan, misfits_by_s)
    else:
        raise Exception('unknown norm method: %s' % outer_norm)
    return misfits_by_s, misfits_by_sr
