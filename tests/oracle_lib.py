"""ctypes driver of the CPU parity oracle (oracle/_build/liboracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
WIDE_LIB_PATH = os.path.join(ROOT, "oracle", "_build", "liboracle_wide.so")   # strip arithmetic in double
KAT_PATH = os.path.join(ROOT, "oracle", "_build", "kat")

fp = C.POINTER(C.c_float)
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
lp = C.POINTER(C.c_longlong)

_libs = {}


def lib(wide=False):
    if wide not in _libs:
        L = C.CDLL(WIDE_LIB_PATH if wide else LIB_PATH)
        L.oracle_create.restype = C.c_void_p
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_last_error.argtypes = [C.c_void_p]
        L.oracle_get_global_misfit.restype = C.c_float
        L.oracle_time_eval.restype = C.c_double
        if hasattr(L, "oracle_enorm"):
            L.oracle_enorm.restype = C.c_float
        for name in ("oracle_destroy", "oracle_set_fresh", "oracle_set_database", "oracle_set_local_interpolation",
                     "oracle_set_spacial_undersampling", "oracle_set_receivers", "oracle_switch_receiver",
                     "oracle_set_source_location", "oracle_set_effective_dt", "oracle_set_ref_seismogram",
                     "oracle_set_misfit_method", "oracle_set_misfit_taper", "oracle_set_misfit_filter",
                     "oracle_set_synthetics_factor", "oracle_set_floating_shiftrange", "oracle_get_nmisfits",
                     "oracle_eval_sources", "oracle_get_global_misfit", "oracle_get_floating_shifts", "oracle_get_seismogram",
                     "oracle_get_probe_spans", "oracle_discretize_source", "oracle_record_indices", "oracle_get_indices",
                     "oracle_receiver_geometry", "oracle_trace_span", "oracle_time_eval", "oracle_get_strip_spans",
                     "oracle_set_source_params_mask", "oracle_set_source_subparams", "oracle_set_source_subparams_limits",
                     "oracle_get_source_subparams", "oracle_minimize_lm", "oracle_lmdif", "oracle_get_ground_motion",
                     "oracle_shift_ref_seismogram", "oracle_autoshift_ref_seismogram", "oracle_get_cross_correlations",
                     "oracle_get_distances", "oracle_get_source_crustal_thickness", "oracle_principal_axes", "oracle_get_probe"):
            if hasattr(L, name):
                getattr(L, name).argtypes = None
        _libs[wide] = L
    return _libs[wide]


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


SOURCE_TYPES = {"bilateral": 1, "circular": 2, "point_lp": 3, "eikonal": 4, "mt_eikonal": 5, "moment_tensor": 6}
NORMS = {"l2norm": 1, "l1norm": 2, "ampspec_l2norm": 3, "ampspec_l1norm": 4, "scalar_product": 5, "peak": 6,
         "floating_l2norm": 7, "floating_l1norm": 8}


class OracleError(RuntimeError):
    pass


class OracleEngine:
    """Same surface as kiwi_b200.Engine, computed by the CPU restatement of the Fortran."""

    def __init__(self, threads=None, wide=False):
        self.L = lib(wide)
        self.h = C.c_void_p(self.L.oracle_create())
        if threads:
            self.L.oracle_set_num_threads(C.c_int(threads))
        self._keep = None
        table = os.path.join(ROOT, "kiwi_b200", "data", "crust2x2.kcr")
        if os.path.exists(table):
            self._check(self.L.oracle_set_crust2x2(self.h, table.encode()))

    def set_source_constraints(self, points, normals):
        p, n = _f32(points).reshape(-1, 3), _f32(normals).reshape(-1, 3)
        self._check(self.L.oracle_set_source_constraints(self.h, C.c_int(p.shape[0]), p.ctypes.data_as(fp), n.ctypes.data_as(fp)))

    def set_source_crustal_thickness_limit(self, limit):
        self._check(self.L.oracle_set_source_crustal_thickness_limit(self.h, C.c_float(limit)))

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise OracleError(self.L.oracle_last_error(self.h).decode())

    def set_database(self, db):
        m = db.meta()
        span0, length, offset, data = db.view()
        self._keep = db
        self._check(self.L.oracle_set_database(self.h, C.c_int(m["nx"]), C.c_int(m["nz"]), C.c_int(m["ng"]), C.c_float(m["dt"]),
                                               C.c_float(m["dx"]), C.c_float(m["dz"]), C.c_float(m["firstx"]), C.c_float(m["firstz"]),
                                               span0.ctypes.data_as(ip), length.ctypes.data_as(ip), offset.ctypes.data_as(lp),
                                               data.ctypes.data_as(fp)))

    def set_local_interpolation(self, method):
        if isinstance(method, str):
            method = method == "bilinear"
        self._check(self.L.oracle_set_local_interpolation(self.h, C.c_int(int(bool(method)))))

    def set_spacial_undersampling(self, xu, zu):
        self._check(self.L.oracle_set_spacial_undersampling(self.h, C.c_int(xu), C.c_int(zu)))

    def set_receivers(self, lat_deg, lon_deg, depth=None, components=None):
        lat = np.ascontiguousarray(lat_deg, dtype=np.float64)
        lon = np.ascontiguousarray(lon_deg, dtype=np.float64)
        n = lat.size
        dep = _f32(np.zeros(n) if depth is None else depth)
        if components is None:
            components = ["ned"] * n
        if isinstance(components, str):
            components = [components] * n
        arr = (C.c_char_p * n)(*[c.encode() for c in components])
        self._check(self.L.oracle_set_receivers(self.h, C.c_int(n), lat.ctypes.data_as(dp), lon.ctypes.data_as(dp), dep.ctypes.data_as(fp), arr))
        self.nreceivers = n
        self._components = list(components)
        self._enabled = [True] * n

    def switch_receiver(self, irec, state):
        self._check(self.L.oracle_switch_receiver(self.h, C.c_int(irec), C.c_int(int(bool(state)))))
        self._enabled[irec - 1] = bool(state)

    def enabled_receivers(self):
        return list(self._enabled)

    def components_per_receiver(self):
        return [len(c) for c in self._components]

    def set_source_location(self, lat, lon, ref_time=0.0):
        self._check(self.L.oracle_set_source_location(self.h, C.c_float(lat), C.c_float(lon), C.c_double(ref_time)))

    def set_effective_dt(self, dt):
        self._check(self.L.oracle_set_effective_dt(self.h, C.c_float(dt)))

    def set_ref_seismogram(self, irec, icomp, tbegin, data):
        d = _f32(data)
        self._check(self.L.oracle_set_ref_seismogram(self.h, C.c_int(irec), C.c_int(icomp), C.c_float(tbegin), C.c_int(d.size), d.ctypes.data_as(fp)))

    def set_misfit_method(self, norm):
        if isinstance(norm, str):
            norm = NORMS[norm]
        self._check(self.L.oracle_set_misfit_method(self.h, C.c_int(norm)))

    def set_misfit_taper(self, irec, x, y):
        x, y = _f32(x), _f32(y)
        self._check(self.L.oracle_set_misfit_taper(self.h, C.c_int(irec), C.c_int(x.size), x.ctypes.data_as(fp), y.ctypes.data_as(fp)))

    def set_misfit_filter(self, x, y, irec=0):
        x, y = _f32(x), _f32(y)
        self._check(self.L.oracle_set_misfit_filter(self.h, C.c_int(irec), C.c_int(x.size), x.ctypes.data_as(fp), y.ctypes.data_as(fp)))

    def set_synthetics_factor(self, f):
        self._check(self.L.oracle_set_synthetics_factor(self.h, C.c_float(f)))

    def set_floating_shiftrange(self, lo, hi, irec=0):
        self._check(self.L.oracle_set_floating_shiftrange(self.h, C.c_int(irec), C.c_float(lo), C.c_float(hi)))

    @property
    def nmisfits(self):
        return self.L.oracle_get_nmisfits(self.h)

    def eval_sources(self, sourcetype, params):
        if isinstance(sourcetype, str):
            sourcetype = SOURCE_TYPES[sourcetype]
        p = _f32(params)
        if p.ndim == 1:
            p = p[None, :]
        ns, npar = p.shape
        nm = self.nmisfits
        out = np.zeros((ns, nm, 2), dtype=np.float32)
        status = np.zeros(ns, dtype=np.int32)
        self._check(self.L.oracle_eval_sources(self.h, C.c_int(sourcetype), C.c_int(ns), C.c_int(npar), p.ctypes.data_as(fp),
                                               out.ctypes.data_as(fp), status.ctypes.data_as(ip)))
        return out, status

    def set_source_params(self, sourcetype, params):
        """set_source_params + synthesis (the oracle evaluates eagerly; without references the status is 1)"""
        self.eval_sources(sourcetype, params)

    def time_eval(self, sourcetype, params):
        if isinstance(sourcetype, str):
            sourcetype = SOURCE_TYPES[sourcetype]
        p = _f32(params)
        if p.ndim == 1:
            p = p[None, :]
        return self.L.oracle_time_eval(self.h, C.c_int(sourcetype), C.c_int(p.shape[0]), C.c_int(p.shape[1]), p.ctypes.data_as(fp))

    def get_global_misfit(self):
        return float(self.L.oracle_get_global_misfit(self.h))

    def get_ground_motion(self, sourcetype, params, which):
        """which: 1 peak velocity, 2 peak acceleration, 3 Arias intensity -> values of the enabled receivers"""
        if isinstance(sourcetype, str):
            sourcetype = SOURCE_TYPES[sourcetype]
        p = _f32(params).ravel()
        out = np.zeros(4096, np.float32)
        n = self.L.oracle_get_ground_motion(self.h, C.c_int(sourcetype), p.ctypes.data_as(fp), C.c_int(p.size), C.c_int(which), out.ctypes.data_as(fp),
                                            C.c_int(out.size))
        if n < 0:
            self._check(1)
        return out[:n].copy()

    # ---- sub-parameters and Levenberg-Marquardt (sequential restatement) ----
    def set_source_params_mask(self, mask):
        m = np.ascontiguousarray(np.asarray(mask).astype(bool), dtype=np.int32)
        self._check(self.L.oracle_set_source_params_mask(self.h, m.ctypes.data_as(ip), C.c_int(m.size)))

    def set_source_subparams(self, sub):
        p = _f32(sub).ravel()
        self._check(self.L.oracle_set_source_subparams(self.h, p.ctypes.data_as(fp), C.c_int(p.size)))

    def set_source_subparams_limits(self, mins, maxs):
        a, b = _f32(mins).ravel(), _f32(maxs).ravel()
        self._check(self.L.oracle_set_source_subparams_limits(self.h, a.ctypes.data_as(fp), b.ctypes.data_as(fp), C.c_int(a.size)))

    def get_source_subparams(self):
        out = np.zeros(64, np.float32)
        n = self.L.oracle_get_source_subparams(self.h, out.ctypes.data_as(fp), C.c_int(out.size))
        return out[:n].copy()

    def minimize_lm(self):
        info, it, mis = C.c_int(), C.c_int(), C.c_float()
        self._check(self.L.oracle_minimize_lm(self.h, C.byref(info), C.byref(it), C.byref(mis)))
        return info.value, it.value, mis.value

    def get_floating_shifts(self):
        out = np.zeros(4096, np.int32)
        n = self.L.oracle_get_floating_shifts(self.h, out.ctypes.data_as(ip))
        return out[:n]

    def shift_ref_seismogram(self, irec, shift):
        self._check(self.L.oracle_shift_ref_seismogram(self.h, C.c_int(irec), C.c_float(shift)))

    def autoshift_ref_seismogram(self, irec, lo, hi):
        out = np.zeros(4096, np.float32)
        self._check(self.L.oracle_autoshift_ref_seismogram(self.h, C.c_int(irec), C.c_float(lo), C.c_float(hi), out.ctypes.data_as(fp)))
        return out[:(self.nreceivers if irec == 0 else 1)]

    def get_cross_correlations(self, irec, lo, hi):
        out = np.zeros(5 * 8192, np.float32)
        nc, ns = C.c_int(), C.c_int()
        self._check(self.L.oracle_get_cross_correlations(self.h, C.c_int(irec), C.c_float(lo), C.c_float(hi), out.ctypes.data_as(fp), C.byref(nc), C.byref(ns)))
        return out[:nc.value * ns.value].reshape(nc.value, ns.value).copy()

    def get_distances(self):
        d, a = np.zeros(4096), np.zeros(4096)
        n = self.L.oracle_get_distances(self.h, d.ctypes.data_as(dp), a.ctypes.data_as(dp))
        return d[:n], a[:n]

    def get_source_crustal_thickness(self):
        t = C.c_float()
        self._check(self.L.oracle_get_source_crustal_thickness(self.h, C.byref(t)))
        return t.value

    def principal_axes(self, strike, dip, rake):
        p, t = np.zeros(2, np.float32), np.zeros(2, np.float32)
        self.L.oracle_principal_axes(C.c_float(strike), C.c_float(dip), C.c_float(rake), p.ctypes.data_as(fp), t.ctypes.data_as(fp))
        return p, t

    def get_probe(self, irec, icomp, which_probe="synthetics", processing="plain", spectrum=False):
        """(first index, samples) of probe_get, or (df, amplitudes) of probe_get_amp_spectrum"""
        first, n, df = C.c_int(), C.c_int(), C.c_float()
        buf = np.empty(1 << 16, np.float32)
        self._check(self.L.oracle_get_probe(self.h, C.c_int(irec), C.c_int(icomp), C.c_int(["synthetics", "references"].index(which_probe)),
                                            C.c_int(["plain", "tapered", "filtered"].index(processing)), C.c_int(int(spectrum)), C.byref(first), C.byref(n),
                                            C.byref(df), buf.ctypes.data_as(fp), C.c_int(buf.size)))
        return (df.value if spectrum else first.value), buf[:n.value].copy()

    def get_seismogram(self, irec, icomp, which=0):
        first, n = C.c_int(), C.c_int()
        cap = 1 << 16
        buf = np.empty(cap, np.float32)
        self._check(self.L.oracle_get_seismogram(self.h, C.c_int(irec), C.c_int(icomp), C.c_int(which), C.byref(first), C.byref(n), buf.ctypes.data_as(fp), C.c_int(cap)))
        return first.value, buf[:n.value].copy()

    def discretize_source(self, sourcetype, params, cap=1 << 20):
        if isinstance(sourcetype, str):
            sourcetype = SOURCE_TYPES[sourcetype]
        p = _f32(params).ravel()
        table = np.empty((cap, 10), np.float32)
        grid = np.zeros(3, np.int32)
        n = self.L.oracle_discretize_source(self.h, C.c_int(sourcetype), C.c_int(p.size), p.ctypes.data_as(fp), table.ctypes.data_as(fp), C.c_int(cap), grid.ctypes.data_as(ip))
        if n < 0:
            raise OracleError(self.L.oracle_last_error(self.h).decode())
        return table[:n].copy(), grid, n

    def record_indices(self, on=True):
        self.L.oracle_record_indices(self.h, C.c_int(int(on)))

    def get_indices(self, irec, cap=1 << 20):
        ix = np.zeros(cap, np.int32); iz = np.zeros(cap, np.int32); its = np.zeros(cap, np.int32)
        dix = np.zeros(cap, np.float32); diz = np.zeros(cap, np.float32)
        dist = np.zeros(cap); azi = np.zeros(cap); bazi = np.zeros(cap)
        n = self.L.oracle_get_indices(self.h, C.c_int(irec), ix.ctypes.data_as(ip), iz.ctypes.data_as(ip), its.ctypes.data_as(ip),
                                      dix.ctypes.data_as(fp), diz.ctypes.data_as(fp), dist.ctypes.data_as(dp), azi.ctypes.data_as(dp),
                                      bazi.ctypes.data_as(dp), C.c_int(cap))
        return dict(ix=ix[:n], iz=iz[:n], its=its[:n], dix=dix[:n], diz=diz[:n], dist=dist[:n], azi=azi[:n], bazi=bazi[:n])

    def trace_span(self, ix, iz, ig):
        s = np.zeros(2, np.int32); ns = C.c_int()
        self._check(self.L.oracle_trace_span(self.h, C.c_int(ix), C.c_int(iz), C.c_int(ig), s.ctypes.data_as(ip), C.byref(ns)))
        return s, ns.value


ORACLE_LM_FCN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, fp, fp)


def oracle_lmdif(fcn, x0, m, ftol=None, xtol=None, gtol=0.0, maxfev=None, epsfcn=0.0, diag=None, mode=1, factor=100.0):
    """sequential lmdif of the oracle on a Python function fcn(x[n]) -> fvec[m] (None = failure). -> (x, fvec, info, nfev)"""
    L = lib()
    x = _f32(x0).ravel().copy()
    n = x.size
    tol = float(np.sqrt(np.float32(1.192091e-07)))
    ftol = tol if ftol is None else ftol
    xtol = tol if xtol is None else xtol
    maxfev = 200 * (n + 1) if maxfev is None else maxfev
    d = np.ones(n, np.float32) if diag is None else _f32(diag).ravel().copy()
    fvec = np.zeros(m, np.float32)

    def cb(user, n_, m_, xp, fpp):
        xa = np.ctypeslib.as_array(xp, shape=(n_,))
        out = fcn(xa)
        if out is None:
            return -2
        np.ctypeslib.as_array(fpp, shape=(m_,))[:] = np.asarray(out, np.float32)
        return 0

    cfn = ORACLE_LM_FCN(cb)
    info, nfev = C.c_int(), C.c_int()
    L.oracle_lmdif(cfn, None, C.c_int(m), C.c_int(n), x.ctypes.data_as(fp), fvec.ctypes.data_as(fp), C.c_float(ftol), C.c_float(xtol),
                   C.c_float(gtol), C.c_int(maxfev), C.c_float(epsfcn), d.ctypes.data_as(fp), C.c_int(mode), C.c_float(factor), C.byref(info),
                   C.byref(nfev))
    return x, fvec, info.value, nfev.value


def oracle_enorm(x):
    a = _f32(x).ravel()
    return float(lib().oracle_enorm(C.c_int(a.size), a.ctypes.data_as(fp)))


def gfdb_interpolate(db, nipx, nipz, wide=False):
    """Gulunay-interpolated database (gfdb.f90:1109-1310, interpolation.f90) through the oracle: dict (ix, iz, ig) -> (span0, samples)
    plus the meta of the interpolated grid."""
    L = lib(wide)
    m = db.meta()
    span0, length, offset, data = db.view()
    lp = C.POINTER(C.c_longlong)
    rc = L.oracle_gfdb_interpolate(C.c_int(m["nx"]), C.c_int(m["nz"]), C.c_int(m["ng"]), C.c_float(m["dt"]), C.c_float(m["dx"]), C.c_float(m["dz"]),
                                   C.c_float(m["firstx"]), C.c_float(m["firstz"]), span0.ctypes.data_as(ip), length.ctypes.data_as(ip),
                                   offset.ctypes.data_as(lp), data.ctypes.data_as(fp), C.c_int(nipx), C.c_int(nipz))
    assert rc == 0
    nx, nz, dx, dz = C.c_int(), C.c_int(), C.c_float(), C.c_float()
    L.oracle_interp_meta(C.byref(nx), C.byref(nz), C.byref(dx), C.byref(dz))
    out = {}
    buf = np.empty(1 << 16, np.float32)
    s0, n = C.c_int(), C.c_int()
    for ix in range(1, nx.value + 1):
        for iz in range(1, nz.value + 1):
            for ig in range(1, m["ng"] + 1):
                rc = L.oracle_interp_trace(C.c_int(ix), C.c_int(iz), C.c_int(ig), C.byref(s0), C.byref(n), buf.ctypes.data_as(fp), C.c_int(buf.size))
                assert rc == 0
                if n.value > 0:
                    out[(ix, iz, ig)] = (s0.value, buf[:n.value].copy())
    return dict(nx=nx.value, nz=nz.value, dx=dx.value, dz=dz.value), out


def gulunay(a, l1, l2, ntmargin, margin1, margin2, wide=False):
    """interpolation.f90 gulunay2d / gulunay3d on a field a[s2][s1][t] (C order; Fortran (t, s1, s2)); returns (tapered a, out[s2*l2][s1*l1][t])"""
    L = lib(wide)
    a = np.ascontiguousarray(a, np.float32).copy()
    s2, s1, t = a.shape
    out = np.zeros((s2 * l2, s1 * l1, t), np.float32)
    rc = L.oracle_gulunay(a.ctypes.data_as(fp), C.c_int(t), C.c_int(s1), C.c_int(s2), C.c_int(l1), C.c_int(l2), out.ctypes.data_as(fp),
                          C.c_int(ntmargin), C.c_int(margin1), C.c_int(margin2))
    assert rc == 0
    return a, out


def ahfull_node(rho, alpha, beta, stf, dt, x, z, nfflag=True, ffflag=True, cap=1 << 14):
    """the ten traces of gfdb_build_ahfull for one node, restated from the Fortran (oracle/ko_ahfull.hpp): [(span0, samples)] * 10"""
    L = lib()
    stf = _f32(stf).ravel()
    span0 = np.zeros(10, np.int32); length = np.zeros(10, np.int32); data = np.zeros((10, cap), np.float32)
    rc = L.oracle_ahfull_node(C.c_float(rho), C.c_float(alpha), C.c_float(beta), stf.ctypes.data_as(fp), C.c_int(stf.size), C.c_float(dt),
                              C.c_float(x), C.c_float(z), C.c_int(int(nfflag)), C.c_int(int(ffflag)), span0.ctypes.data_as(ip),
                              length.ctypes.data_as(ip), data.ctypes.data_as(fp), C.c_int(cap))
    if rc != 0:
        raise OracleError("oracle_ahfull_node: trace longer than %d samples" % cap)
    return [(int(span0[i]), data[i, :length[i]].copy()) for i in range(10)]
