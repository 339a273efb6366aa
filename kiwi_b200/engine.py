"""Host-side mirror of the reference's engine interface for the hot path.

Method names, argument meaning and error behaviour follow the commands of the `minimizer` program
(minimizer.f90:1729-1811) and the subroutines of minimizer_engine.f90 they call; a failing call
raises KiwiError carrying the engine's error string (the reference answers "<cmd>: nok" +
g_errstr, minimizer.f90:1676-1701).
"""
import ctypes as C
import os

import numpy as np

from ._lib import lib, c_float_p, c_double_p, c_int_p, c_ll_p

# source_all.f90:58-60 / parameterized_source.f90
SOURCE_TYPES = {"bilateral": 1, "circular": 2, "point_lp": 3, "eikonal": 4, "mt_eikonal": 5, "moment_tensor": 6}
# comparator.f90:33-42, names as accepted by set_misfit_method (minimizer.f90:842-873)
NORMS = {"l2norm": 1, "l1norm": 2, "ampspec_l2norm": 3, "ampspec_l1norm": 4, "scalar_product": 5, "peak": 6,
         "floating_l2norm": 7, "floating_l1norm": 8}
# benchmark/kiwibench.py:51-72: the 20-sample ramp source time function of the kiwibench database
KIWIBENCH_STF = np.array([0, 0, 0, 0, 0, 0, .1, .2, .3, .4, .5, .6, .7, .8, .9, 1, 1, 1, 1, 1], dtype=np.float32)


# the CRUST2.0 table converted by tools/make_crust2x2_table.py (the reference installs the text files under
# $(datadir)/kiwi/aux/crust2x2, Makefile:135-139)
CRUST2X2_TABLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "crust2x2.kcr")


class KiwiError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise KiwiError(lib.kiwi_last_error().decode("utf-8", "replace"))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(c_float_p)


def n_source_params(sourcetype):
    if isinstance(sourcetype, str):
        sourcetype = SOURCE_TYPES[sourcetype]
    return lib.kiwi_get_n_source_params(int(sourcetype))


def global_misfits(misfits):
    """sqrt(sum m^2)/sqrt(sum n^2) per candidate (minimizer_engine.f90:937-942)."""
    m = _f32(misfits)
    ns, nm = m.shape[0], m.shape[1]
    out = np.empty(ns, dtype=np.float32)
    _check(lib.kiwi_global_misfits(ns, nm, _fp(m), _fp(out)))
    return out


class Gfdb:
    """Green's function database on the host (gfdb.f90 t_gfdb; tools gfdb_build, gfdb_build_ahfull)."""

    def __init__(self, handle):
        if not handle:
            raise KiwiError(lib.kiwi_last_error().decode())
        self._h = C.c_void_p(handle)

    @classmethod
    def create(cls, nx, nz, ng, dt, dx, dz, firstx, firstz):
        """gfdb_build <db> nchunks nx nz ng dt dx dz firstx firstz (gfdb_build.f90)."""
        return cls(lib.kiwi_gfdb_create(nx, nz, ng, dt, dx, dz, firstx, firstz))

    @classmethod
    def read(cls, path):
        return cls(lib.kiwi_gfdb_read(str(path).encode()))

    @classmethod
    def read_hdf(cls, basepath):
        """Kiwi's own HDF5 database <basepath>.index + <basepath>.<i>.chunk (gfdb_io_hdf.f90), no libhdf5 needed."""
        return cls(lib.kiwi_gfdb_read_hdf(str(basepath).encode()))

    def write(self, path):
        _check(lib.kiwi_gfdb_write(self._h, str(path).encode()))

    def save_array(self, ix, iz, ig, span0, data):
        d = _f32(data)
        _check(lib.kiwi_gfdb_save_array(self._h, ix, iz, ig, span0, d.size, _fp(d)))

    def build_ahfull(self, rho, alpha, beta, stf=KIWIBENCH_STF, nfflag=True, ffflag=True, nthreads=0):
        """gfdb_build_ahfull for every grid node (gfdb_build_ahfull.f90:70-216)."""
        s = _f32(stf)
        _check(lib.kiwi_gfdb_build_ahfull(self._h, rho, alpha, beta, _fp(s), s.size, int(nfflag), int(ffflag), nthreads))
        return self

    def interpolate(self, nipx, nipz, device=0):
        """The database `set_database dbpath nipx nipz` works on: Gulunay f-k interpolation on the GPU (gfdb.f90:1109-1310)."""
        return Gfdb(lib.kiwi_gfdb_interpolate(self._h, nipx, nipz, device))

    def meta(self):
        nx, nz, ng = C.c_int(), C.c_int(), C.c_int()
        dt, dx, dz, fx, fz = C.c_float(), C.c_float(), C.c_float(), C.c_float(), C.c_float()
        nt, ns = C.c_longlong(), C.c_longlong()
        _check(lib.kiwi_gfdb_meta(self._h, nx, nz, ng, dt, dx, dz, fx, fz, nt, ns))
        return dict(nx=nx.value, nz=nz.value, ng=ng.value, dt=dt.value, dx=dx.value, dz=dz.value, firstx=fx.value,
                    firstz=fz.value, ntraces=nt.value, nsamples=ns.value)

    def view(self):
        """Borrowed flat arrays (span0, len, offset, data); index ((ix-1)*nz + (iz-1))*ng + (ig-1)."""
        m = self.meta()
        n = m["nx"] * m["nz"] * m["ng"]
        ps0, pl, po, pd = c_int_p(), c_int_p(), c_ll_p(), c_float_p()
        _check(lib.kiwi_gfdb_view(self._h, C.byref(ps0), C.byref(pl), C.byref(po), C.byref(pd)))
        span0 = np.ctypeslib.as_array(ps0, shape=(n,))
        length = np.ctypeslib.as_array(pl, shape=(n,))
        offset = np.ctypeslib.as_array(po, shape=(n,))
        data = np.ctypeslib.as_array(pd, shape=(max(int(m["nsamples"]), 1),))
        return span0, length, offset, data

    def close(self):
        if self._h:
            lib.kiwi_gfdb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """One engine context = the module state of minimizer_engine.f90, resident on one B200."""

    def __init__(self, device=0):
        h = lib.kiwi_create(device)
        if not h:
            raise KiwiError(lib.kiwi_last_error().decode())
        self._h = C.c_void_p(h)
        self._db = None
        self._device = device
        if os.path.exists(CRUST2X2_TABLE):          # minimizer loads crust2x2 at start-up (minimizer.f90:1669-1674)
            self.set_crust2x2(CRUST2X2_TABLE)

    def close(self):
        if getattr(self, "_pin_ptr", None):
            lib.kiwi_host_free(self._pin_ptr)
            self._pin_ptr, self._pin_bytes = None, 0
        if self._h:
            lib.kiwi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setters (one per reference command) --------------------------------------------------------
    def set_database(self, db, nipx=1, nipz=1):
        """set_database dbpath [nipx nipz] (minimizer.f90:89-135): nipx, nipz > 1 turn on Gulunay's interpolation of the database."""
        if nipx != 1 or nipz != 1:
            db = db.interpolate(nipx, nipz, self._device)
        _check(lib.kiwi_set_database(self._h, db._h))
        self._db = db

    def set_local_interpolation(self, method):
        if isinstance(method, str):
            if method not in ("nearest_neighbor", "bilinear"):
                raise KiwiError("unknown interpolation method: " + method)   # minimizer.f90:170-176
            method = method == "bilinear"
        _check(lib.kiwi_set_local_interpolation(self._h, int(bool(method))))

    def set_spacial_undersampling(self, xunder, zunder):
        _check(lib.kiwi_set_spacial_undersampling(self._h, xunder, zunder))

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
        _check(lib.kiwi_set_receivers(self._h, n, lat.ctypes.data_as(c_double_p), lon.ctypes.data_as(c_double_p), _fp(dep), arr))
        self._nreceivers = n
        self._components = list(components)
        self._enabled = [True] * n

    def set_receivers_file(self, path, has_depth=False):
        """set_receivers <file> [has_depth]: lat lon [depth] [components] per line (minimizer_engine.f90:165-286)."""
        lat, lon, dep, comps = [], [], [], []
        with open(path) as f:
            for line in f:
                line = line.split("#")[0].split()
                if not line:
                    continue
                lat.append(float(line[0])); lon.append(float(line[1]))
                k = 2
                if has_depth:
                    dep.append(float(line[2])); k = 3
                else:
                    dep.append(0.0)
                comps.append(line[k] if len(line) > k else "ned")
        self.set_receivers(lat, lon, dep, comps)
        return len(lat)

    def switch_receiver(self, ireceiver, state):
        _check(lib.kiwi_switch_receiver(self._h, ireceiver, int(bool(state))))
        self._enabled[ireceiver - 1] = bool(state)

    @property
    def device(self):
        return self._device

    def enabled_receivers(self):
        """[bool] per receiver (switch_receiver)"""
        return list(self._enabled)

    def components_per_receiver(self):
        """[number of components] per receiver: the misfit pairs it contributes to get_misfits while enabled"""
        return [len(c) for c in self._components]

    def set_source_location(self, lat_deg, lon_deg, ref_time=0.0):
        _check(lib.kiwi_set_source_location(self._h, lat_deg, lon_deg, ref_time))

    def set_crust2x2(self, path):
        _check(lib.kiwi_set_crust2x2(self._h, str(path).encode()))

    def set_source_constraints(self, points, normals):
        p, n = _f32(points).reshape(-1, 3), _f32(normals).reshape(-1, 3)
        _check(lib.kiwi_set_source_constraints(self._h, p.shape[0], _fp(p), _fp(n)))

    def set_source_crustal_thickness_limit(self, limit):
        _check(lib.kiwi_set_source_crustal_thickness_limit(self._h, limit))

    def set_effective_dt(self, dt):
        _check(lib.kiwi_set_effective_dt(self._h, dt))

    def set_ref_seismogram(self, ireceiver, icomponent, tbegin, data):
        d = _f32(data)
        _check(lib.kiwi_set_ref_seismogram(self._h, ireceiver, icomponent, tbegin, d.size, _fp(d)))

    def set_misfit_method(self, norm):
        if isinstance(norm, str):
            if norm not in NORMS:
                raise KiwiError("unknown norm method: " + norm)
            norm = NORMS[norm]
        _check(lib.kiwi_set_misfit_method(self._h, norm))

    def set_misfit_taper(self, ireceiver, x, y):
        x, y = _f32(x), _f32(y)
        _check(lib.kiwi_set_misfit_taper(self._h, ireceiver, x.size, _fp(x), _fp(y)))

    def set_misfit_filter(self, x, y, ireceiver=0):
        x, y = _f32(x), _f32(y)
        _check(lib.kiwi_set_misfit_filter(self._h, ireceiver, x.size, _fp(x), _fp(y)))

    def set_synthetics_factor(self, factor):
        _check(lib.kiwi_set_synthetics_factor(self._h, factor))

    def set_share_syntheses(self, enabled):
        """candidates differing only in the moment share one synthesis (default on)"""
        _check(lib.kiwi_set_share_syntheses(self._h, int(bool(enabled))))

    def set_mt_grid(self, enabled):
        """tensor-core path for point moment-tensor grid searches on (default) / off; 2 = on, synthesis not fused into the contraction"""
        _check(lib.kiwi_set_mt_grid(self._h, 2 if (not isinstance(enabled, bool) and enabled == 2) else int(bool(enabled))))

    def set_accumulation(self, reference_order):
        """synthesis in the reference's order of floating-point operations (slow; kiwi_set_accumulation) on / off (default)"""
        _check(lib.kiwi_set_accumulation(self._h, int(bool(reference_order))))

    def set_eikonal_device(self, min_batch):
        """fast-marching solves of eikonal batches: -1 shared between host threads and device (default), 0 host only, k > 0 all on the device from k candidates on"""
        _check(lib.kiwi_set_eikonal_device(self._h, int(min_batch)))

    def set_floating_shiftrange(self, lo, hi, ireceiver=0):
        _check(lib.kiwi_set_floating_shiftrange(self._h, ireceiver, lo, hi))

    # ---- evaluation -------------------------------------------------------------------------------
    @property
    def nmisfits(self):
        return lib.kiwi_get_nmisfits(self._h)

    def eval_sources(self, sourcetype, params, pinned=False):
        """Batched set_source_params + get_misfits: returns (misfits[ns, nmisfits, 2], status[ns]).
        pinned=True returns the block in page-locked memory owned by the engine (valid until the next pinned call):
        the device-to-host copy of a large block then runs at bus speed."""
        if isinstance(sourcetype, str):
            sourcetype = SOURCE_TYPES[sourcetype]
        p = _f32(params)
        if p.ndim == 1:
            p = p[None, :]
        ns, nparams = p.shape
        nm = self.nmisfits
        out = self._pinned_out(ns, nm) if pinned else np.empty((ns, nm, 2), dtype=np.float32)
        status = np.zeros(ns, dtype=np.int32)
        _check(lib.kiwi_eval_sources(self._h, sourcetype, ns, nparams, _fp(p), _fp(out), status.ctypes.data_as(c_int_p)))
        self._last_ns = ns
        return out, status

    def _pinned_out(self, ns, nm):
        """misfit block in page-locked memory owned by the engine (kiwi_host_alloc), reused by the next pinned call"""
        need = max(ns * nm * 2, 1) * 4
        if getattr(self, "_pin_bytes", 0) < need:
            if getattr(self, "_pin_ptr", None):
                lib.kiwi_host_free(self._pin_ptr)
            self._pin_ptr = lib.kiwi_host_alloc(need)
            if not self._pin_ptr:
                raise KiwiError(lib.kiwi_last_error().decode())
            self._pin_bytes = need
        buf = (C.c_float * (ns * nm * 2)).from_address(self._pin_ptr)
        return np.frombuffer(buf, dtype=np.float32).reshape(ns, nm, 2)

    def eval_sources_on_device(self, sourcetype, params):
        """Evaluate and leave the misfit cube on the GPU (for outer_misfits); returns status[ns]."""
        if isinstance(sourcetype, str):
            sourcetype = SOURCE_TYPES[sourcetype]
        p = _f32(params)
        if p.ndim == 1:
            p = p[None, :]
        status = np.zeros(p.shape[0], dtype=np.int32)
        _check(lib.kiwi_eval_sources(self._h, sourcetype, p.shape[0], p.shape[1], _fp(p), None, status.ctypes.data_as(c_int_p)))
        self._last_ns = p.shape[0]
        return status

    def outer_misfits(self, ns=None, receiver_weights=None, outer_norm="l2norm", anarchy=False, bweights=None, d_misfits_ptr=None,
                      want_matrix=True):
        """make_global_misfits + best source (seismosizer.py:843-922, gridsearch.py:250-266) on the device:
        (misfits_by_s[nboot+1, ns] or None, best[nboot+1], best_value[nboot+1])."""
        if ns is None:
            ns = getattr(self, "_last_ns", 0)
        rw = None if receiver_weights is None else np.ascontiguousarray(receiver_weights, dtype=np.float64)
        bw = None if bweights is None else np.ascontiguousarray(bweights, dtype=np.float64)
        nboot = 0 if bw is None else bw.shape[0]
        out = np.empty((nboot + 1, ns), dtype=np.float64) if want_matrix else None
        best = np.zeros(nboot + 1, dtype=np.int32); bestv = np.zeros(nboot + 1, dtype=np.float64)
        norm = NORMS[outer_norm] if isinstance(outer_norm, str) else outer_norm
        _check(lib.kiwi_outer_misfits(self._h, ns, C.c_void_p(d_misfits_ptr) if d_misfits_ptr else None,
                                      rw.ctypes.data_as(c_double_p) if rw is not None else None, norm, int(bool(anarchy)), nboot,
                                      bw.ctypes.data_as(c_double_p) if bw is not None else None,
                                      out.ctypes.data_as(c_double_p) if out is not None else None, best.ctypes.data_as(c_int_p),
                                      bestv.ctypes.data_as(c_double_p)))
        return out, best, bestv

    def eval_sources_device(self, sourcetype, params, d_misfits_ptr):
        """Same, results left at the device address d_misfits_ptr ([ns][nmisfits][2] fp32)."""
        if isinstance(sourcetype, str):
            sourcetype = SOURCE_TYPES[sourcetype]
        p = _f32(params)
        if p.ndim == 1:
            p = p[None, :]
        ns, nparams = p.shape
        status = np.zeros(ns, dtype=np.int32)
        _check(lib.kiwi_eval_sources_device(self._h, sourcetype, ns, nparams, _fp(p), C.c_void_p(d_misfits_ptr), status.ctypes.data_as(c_int_p)))
        return status

    def set_source_params(self, sourcetype, params):
        if isinstance(sourcetype, str):
            sourcetype = SOURCE_TYPES[sourcetype]
        p = _f32(params).ravel()
        _check(lib.kiwi_set_source_params(self._h, sourcetype, p.size, _fp(p)))

    def get_misfits(self):
        """(misfit, norm factor) pairs of the enabled receivers, receiver-major (minimizer_engine.f90:1130-1172)."""
        nm = self.nmisfits
        out = np.empty((max(nm, 1), 2), dtype=np.float32)
        n = C.c_int()
        _check(lib.kiwi_get_misfits(self._h, _fp(out), nm, n))
        return out[:n.value]

    def get_global_misfit(self):
        v = C.c_float()
        _check(lib.kiwi_get_global_misfit(self._h, v))
        return v.value

    # ---- ground-motion diagnostics (minimizer_engine.f90:1174-1245) ---------------------------------------
    def get_peak_amplitudes(self, differentiate):
        out = np.zeros(8192, dtype=np.float32)
        n = C.c_int()
        _check(lib.kiwi_get_peak_amplitudes(self._h, differentiate, _fp(out), out.size, n))
        return out[:n.value].copy()

    def get_arias_intensities(self):
        out = np.zeros(8192, dtype=np.float32)
        n = C.c_int()
        _check(lib.kiwi_get_arias_intensities(self._h, _fp(out), out.size, n))
        return out[:n.value].copy()

    def eval_ground_motion(self, sourcetype, params, nenabled):
        """-> (values[ns, nenabled, 3] = peak velocity, peak acceleration, Arias intensity; status[ns])"""
        if isinstance(sourcetype, str):
            sourcetype = SOURCE_TYPES[sourcetype]
        p = _f32(params)
        if p.ndim == 1:
            p = p[None, :]
        out = np.zeros((p.shape[0], nenabled, 3), dtype=np.float32)
        status = np.zeros(p.shape[0], dtype=np.int32)
        _check(lib.kiwi_eval_ground_motion(self._h, sourcetype, p.shape[0], p.shape[1], _fp(p), _fp(out), status.ctypes.data_as(c_int_p)))
        return out, status

    # ---- sub-parameters and Levenberg-Marquardt (minimizer_engine.f90:525-610, 729-874) ------------------
    def set_source_params_mask(self, mask):
        m = np.ascontiguousarray(np.asarray(mask).astype(bool), dtype=np.int32)
        _check(lib.kiwi_set_source_params_mask(self._h, m.ctypes.data_as(c_int_p), m.size))

    def set_source_subparams(self, subparams):
        p = _f32(subparams).ravel()
        _check(lib.kiwi_set_source_subparams(self._h, _fp(p), p.size))

    def set_source_subparams_limits(self, mins, maxs):
        a, b = _f32(mins).ravel(), _f32(maxs).ravel()
        if a.size != b.size:
            raise KiwiError("wrong number of subparam_maxs")
        _check(lib.kiwi_set_source_subparams_limits(self._h, _fp(a), _fp(b), a.size))

    def get_source_subparams(self):
        out = np.zeros(64, dtype=np.float32)
        n = C.c_int()
        _check(lib.kiwi_get_source_subparams(self._h, _fp(out), out.size, n))
        return out[:n.value].copy()

    def minimize_lm(self):
        """-> (info, iterations, misfit); the source is left at the last model evaluated, as in the reference."""
        info, it, mis = C.c_int(), C.c_int(), C.c_float()
        _check(lib.kiwi_minimize_lm(self._h, info, it, mis))
        return info.value, it.value, mis.value

    def get_floating_shifts(self):
        nr = 4096
        out = np.zeros(nr, dtype=np.int32)
        n = C.c_int()
        _check(lib.kiwi_get_floating_shifts(self._h, out.ctypes.data_as(c_int_p), nr, n))
        return out[:n.value]

    def shift_ref_seismogram(self, ireceiver, shift):
        """shift_ref_seismogram (minimizer_engine.f90:354-378): seconds."""
        _check(lib.kiwi_shift_ref_seismogram(self._h, ireceiver, shift))

    def autoshift_ref_seismogram(self, ireceiver, shift_lo, shift_hi):
        """autoshift_ref_seismogram (minimizer_engine.f90:380-416): the applied shifts in seconds (ireceiver 0 = all)."""
        out = np.zeros(max(1, getattr(self, "_nreceivers", 4096) if ireceiver == 0 else 1), dtype=np.float32)
        n = C.c_int()
        _check(lib.kiwi_autoshift_ref_seismogram(self._h, ireceiver, shift_lo, shift_hi, out.ctypes.data_as(c_float_p), out.size, n))
        return out[:n.value]

    def get_cross_correlations(self, ireceiver, shift_lo, shift_hi):
        """In-memory replacement of output_cross_correlations: array [component][shift]."""
        out = np.zeros(5 * 8192, dtype=np.float32)
        nc, ns = C.c_int(), C.c_int()
        _check(lib.kiwi_get_cross_correlations(self._h, ireceiver, shift_lo, shift_hi, out.ctypes.data_as(c_float_p), out.size, nc, ns))
        return out[:nc.value * ns.value].reshape(nc.value, ns.value).copy()

    def get_distances(self):
        """get_distances (minimizer_engine.f90:1260-1281): (distances [m], azimuths [rad]) of all receivers."""
        nr = getattr(self, "_nreceivers", 4096)
        d, a = np.zeros(nr), np.zeros(nr)
        n = C.c_int()
        _check(lib.kiwi_get_distances(self._h, d.ctypes.data_as(c_double_p), a.ctypes.data_as(c_double_p), nr, n))
        return d[:n.value], a[:n.value]

    def get_source_crustal_thickness(self):
        t = C.c_float()
        _check(lib.kiwi_get_source_crustal_thickness(self._h, t))
        return t.value

    def get_principal_axes(self):
        """(pax, tax), each (azimuth, polar angle) in degrees."""
        p, t = np.zeros(2, np.float32), np.zeros(2, np.float32)
        _check(lib.kiwi_get_principal_axes(self._h, _fp(p), _fp(t)))
        return p, t

    def get_probe(self, ireceiver, icomponent, which_probe="synthetics", processing="plain", spectrum=False):
        """output_seismograms / output_seismogram_spectra in memory: (first index, samples) or (df, amplitudes)."""
        wp, pr = ["synthetics", "references"].index(which_probe), ["plain", "tapered", "filtered"].index(processing)
        buf = np.empty(1 << 15, dtype=np.float32)
        n = C.c_int()
        if spectrum:
            df = C.c_float()
            _check(lib.kiwi_get_probe_spectrum(self._h, ireceiver, icomponent, wp, pr, df, n, _fp(buf), buf.size))
            return df.value, buf[:n.value].copy()
        first = C.c_int()
        _check(lib.kiwi_get_probe(self._h, ireceiver, icomponent, wp, pr, first, n, _fp(buf), buf.size))
        return first.value, buf[:n.value].copy()

    def set_synthetic_reference(self, scale=1.0):
        """Calculate seismograms of the current source and use these as reference (seismosizer.py:523-527)."""
        for ir, comps in enumerate(self._components, start=1):
            if not self._enabled[ir - 1]:
                continue
            for ic in range(1, len(comps) + 1):
                first, data = self.get_probe(ir, ic, "synthetics", "plain")
                self.set_ref_seismogram(ir, ic, (first - 1) * self._dt(), data * np.float32(scale))

    def get_receivers_snapshot(self, which_seismograms=("syn", "ref"), which_spectra=("syn", "ref"), which_processing="filtered"):
        """Seismograms and amplitude spectra of all enabled receivers as the Python driver of the reference collects them through
        output_seismograms / output_seismogram_spectra (seismosizer.py:541-610): a list (one dict per receiver, None for disabled
        ones) with keys 'syn_seismograms', 'ref_seismograms' -> [(t0, dt, samples) per component] and 'syn_spectra',
        'ref_spectra' -> [(df, amplitudes) per component]."""
        dt = self._dt()
        out = []
        for ir, comps in enumerate(self._components, start=1):
            if not self._enabled[ir - 1]:
                out.append(None)
                continue
            rec = {}
            for key, probe in (("syn", "synthetics"), ("ref", "references")):
                if key in which_seismograms:
                    rec[key + "_seismograms"] = []
                    for ic in range(1, len(comps) + 1):
                        first, data = self.get_probe(ir, ic, probe, which_processing)
                        rec[key + "_seismograms"].append(((first - 1) * dt, dt, data))
                if key in which_spectra:
                    rec[key + "_spectra"] = [self.get_probe(ir, ic, probe, which_processing, spectrum=True) for ic in range(1, len(comps) + 1)]
            out.append(rec)
        return out

    def _dt(self):
        return self._db.meta()["dt"]

    def get_seismogram(self, ireceiver, icomponent, which=0):
        """In-memory replacement of output_seismograms: (first_index, samples)."""
        first, n = C.c_int(), C.c_int()
        cap = 1 << 16
        buf = np.empty(cap, dtype=np.float32)
        _check(lib.kiwi_get_seismogram(self._h, ireceiver, icomponent, which, first, n, _fp(buf), cap))
        return first.value, buf[:min(n.value, cap)].copy()

    # ---- inspection (bit-exact integer contract) --------------------------------------------------------
    def discretize_source(self, sourcetype, params, cap=1 << 20):
        if isinstance(sourcetype, str):
            sourcetype = SOURCE_TYPES[sourcetype]
        p = _f32(params).ravel()
        table = np.empty((cap, 10), dtype=np.float32)
        n = C.c_int()
        grid = np.zeros(3, dtype=np.int32)
        _check(lib.kiwi_discretize_source(self._h, sourcetype, p.size, _fp(p), _fp(table), cap, n, grid.ctypes.data_as(c_int_p)))
        return table[:min(n.value, cap)].copy(), grid, n.value

    def get_indices(self, ireceiver, cap=1 << 20):
        ix = np.zeros(cap, np.int32); iz = np.zeros(cap, np.int32); its = np.zeros(cap, np.int32)
        dix = np.zeros(cap, np.float32); diz = np.zeros(cap, np.float32); near = np.zeros(cap, np.int32)
        n = C.c_int()
        _check(lib.kiwi_get_indices(self._h, ireceiver, ix.ctypes.data_as(c_int_p), iz.ctypes.data_as(c_int_p), its.ctypes.data_as(c_int_p),
                                    _fp(dix), _fp(diz), near.ctypes.data_as(c_int_p), cap, n))
        k = min(n.value, cap)
        return dict(ix=ix[:k], iz=iz[:k], its=its[:k], dix=dix[:k], diz=diz[:k], near=near[:k])

    def get_spans(self, ireceiver):
        s = np.zeros(6, np.int32)
        _check(lib.kiwi_get_spans(self._h, ireceiver, s.ctypes.data_as(c_int_p)))
        return s

    def trace_span(self, ix, iz, ig):
        s = np.zeros(2, np.int32)
        _check(lib.kiwi_trace_span(self._h, ix, iz, ig, s.ctypes.data_as(c_int_p)))
        return s

    # ---- measurement ---------------------------------------------------------------------------------
    def last_batch_bytes(self, max_candidates=4):
        """(B_alg, B_log) bytes per evaluation, candidates sampled, centroids skipped."""
        a, b, n, k = C.c_double(), C.c_double(), C.c_int(), C.c_longlong()
        _check(lib.kiwi_last_batch_bytes(self._h, max_candidates, a, b, n, k))
        return a.value, b.value, n.value, k.value

    def last_timing(self):
        ms = np.zeros(5, np.float32); ln = np.zeros(4, np.int32)
        _check(lib.kiwi_last_timing(self._h, _fp(ms), ln.ctypes.data_as(c_int_p)))
        return dict(discretise_ms=float(ms[0]), geometry_ms=float(ms[1]), synthesis_ms=float(ms[2]), misfit_ms=float(ms[3]),
                    total_ms=float(ms[4]), launches=[int(v) for v in ln])


def gulunay(a, l1, l2, ntmargin, margin1, margin2, device=0):
    """gulunay2d / gulunay3d (interpolation.f90) on fields a[batch][s2][s1][t]: returns (tapered a, out[batch][s2*l2][s1*l1][t])."""
    a = np.ascontiguousarray(a, dtype=np.float32).copy()
    batch, s2, s1, t = a.shape
    out = np.zeros((batch, s2 * l2, s1 * l1, t), dtype=np.float32)
    _check(lib.kiwi_gulunay(device, _fp(a), batch, t, s1, s2, l1, l2, ntmargin, margin1, margin2, _fp(out)))
    return a, out


def eikonal_fmm(speed, origin, delta, initialpoint):
    """eikonal_solver_fmm (eikonal.f90:29-199) on the host: speed[ny][nx] -> times[ny][nx]."""
    sp = np.ascontiguousarray(speed, dtype=np.float32)
    ny, nx = sp.shape
    times = np.zeros_like(sp)
    o, d, p = _f32(origin), _f32(delta), _f32(initialpoint)
    _check(lib.kiwi_eikonal_fmm(nx, ny, _fp(sp), _fp(o), _fp(d), _fp(p), _fp(times)))
    return times


def eikonal_fmm_device(speeds, origins, deltas, initialpoints):
    """the same on the device, a batch of grids at a time (kiwi_eikonal_fmm_device): list of speed[ny][nx] -> (list of times, kernel ms)"""
    sps = [np.ascontiguousarray(s, dtype=np.float32) for s in speeds]
    n = len(sps)
    nx = np.array([s.shape[1] for s in sps], dtype=np.int32); ny = np.array([s.shape[0] for s in sps], dtype=np.int32)
    times = [np.zeros_like(s) for s in sps]
    o, d, p = (np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(n, 2)) for v in (origins, deltas, initialpoints))
    sp_ptrs = (c_float_p * n)(*[_fp(s) for s in sps]); t_ptrs = (c_float_p * n)(*[_fp(t) for t in times])
    ms = C.c_float(0)
    _check(lib.kiwi_eikonal_fmm_device(n, nx.ctypes.data_as(c_int_p), ny.ctypes.data_as(c_int_p), sp_ptrs, _fp(o), _fp(d), _fp(p), t_ptrs, C.byref(ms)))
    return times, float(ms.value)


def lmdif_batched(fcn, x0, m, ftol=None, xtol=None, gtol=0.0, maxfev=None, epsfcn=0.0, diag=None, mode=1, factor=100.0):
    """MINPACK lmdif (single precision) with the Jacobian columns evaluated as one batch (kiwi_lmdif_batched).
    fcn(xs[ncols, n]) -> fvecs[ncols, m] (rows may be returned short to signal a failure at that column).
    Returns (x, fvec, info, nfev)."""
    from ._lib import LM_FCN
    x = _f32(x0).ravel().copy()
    n = x.size
    tol = float(np.sqrt(np.float32(1.192091e-07)))
    ftol = tol if ftol is None else ftol
    xtol = tol if xtol is None else xtol
    maxfev = 200 * (n + 1) if maxfev is None else maxfev
    d = np.ones(n, dtype=np.float32) if diag is None else _f32(diag).ravel().copy()
    fvec = np.zeros(m, dtype=np.float32)

    def cb(user, ncols, n_, m_, xs, fs):
        xa = np.ctypeslib.as_array(xs, shape=(ncols, n_))
        fa = np.ctypeslib.as_array(fs, shape=(ncols, m_))
        out = fcn(xa)
        if out is None:
            return 0
        out = np.asarray(out, dtype=np.float32).reshape(-1, m_)
        fa[:out.shape[0]] = out
        return int(out.shape[0])

    cfn = LM_FCN(cb)
    info, nfev = C.c_int(), C.c_int()
    _check(lib.kiwi_lmdif_batched(C.cast(cfn, C.c_void_p), None, m, n, _fp(x), _fp(fvec), ftol, xtol, gtol, maxfev, epsfcn, _fp(d), mode, factor,
                                  info, nfev))
    return x, fvec, info.value, nfev.value


def h5_root_members(path):
    """names of the members of the root group of an HDF5 file (kiwi_h5_read_root_dataset with name = NULL)"""
    n = C.c_longlong()
    _check(lib.kiwi_h5_read_root_dataset(str(path).encode(), None, None, None, None, None, None, 0, n, None))
    buf = C.create_string_buffer(max(int(n.value), 1))
    _check(lib.kiwi_h5_read_root_dataset(str(path).encode(), None, None, None, None, None, buf, n.value, n, None))
    return [s.decode() for s in buf.raw[:n.value].split(b"\0") if s]


def h5_read_root_dataset(path, name):
    """-> (array, nattrs) of a dataset in the root group of an HDF5 file, through the library's minimal parser"""
    cls, sz, rank, na = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    dims = (C.c_longlong * 8)()
    nb = C.c_longlong()
    _check(lib.kiwi_h5_read_root_dataset(str(path).encode(), name.encode(), cls, sz, rank, dims, None, 0, nb, na))
    raw = np.zeros(max(int(nb.value), 1), dtype=np.uint8)
    _check(lib.kiwi_h5_read_root_dataset(str(path).encode(), name.encode(), cls, sz, rank, dims, raw.ctypes.data_as(C.c_void_p), raw.size, nb, na))
    kinds = {(0, 4): np.int32, (0, 8): np.int64, (0, 2): np.int16, (0, 1): np.int8, (1, 4): np.float32, (1, 8): np.float64, (7, 8): np.uint64}
    dt = kinds.get((cls.value, sz.value))
    if dt is None:
        raise KiwiError("unsupported element type: class %d, %d bytes" % (cls.value, sz.value))
    shape = tuple(int(dims[i]) for i in range(rank.value))
    return raw[:nb.value].view(dt).reshape(shape), na.value
