// TEST INFRASTRUCTURE ONLY (see ko_base.hpp).  C ABI around the oracle so that tests/ (ctypes),
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs can drive it.
// The entry points deliberately mirror include/kiwi_b200.h one to one.
#include "ko_engine.hpp"
#include "ko_interp.hpp"
#include "ko_ahfull.hpp"
#include <chrono>

using namespace ko;

extern "C" {

void* oracle_create() { return new Engine(); }
void oracle_destroy(void* h) { delete (Engine*)h; }
const char* oracle_last_error(void* h) { return ((Engine*)h)->errstr.c_str(); }
void oracle_set_fresh(void* h, int fresh) { ((Engine*)h)->fresh = fresh != 0; }
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int oracle_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// flat array form of a GFDB, see include/kiwi_b200.h kiwi_gfdb_view()
int oracle_set_database(void* h, int nx, int nz, int ng, float dt, float dx, float dz, float firstx, float firstz,
                        const int* span0, const int* len, const long long* offset, const float* data) {
    Engine& e = *(Engine*)h;
    gfdb_from_arrays(e.db, nx, nz, ng, dt, dx, dz, firstx, firstz, span0, len, offset, data);
    e.database_inited = true;
    return 0;
}
int oracle_set_local_interpolation(void* h, int bilinear) { ((Engine*)h)->interpolate = bilinear != 0; return 0; }
int oracle_set_spacial_undersampling(void* h, int xu, int zu) {
    Engine& e = *(Engine*)h;
    if (xu < 1 || zu < 1) { e.errstr = "invalid undersampling value"; return 1; }
    e.xundersample = xu; e.zundersample = zu; return 0;
}
int oracle_set_receivers(void* h, int n, const double* lat_deg, const double* lon_deg, const float* depth,
                         const char* const* comps) {
    return set_receivers(*(Engine*)h, n, lat_deg, lon_deg, depth, comps) ? 0 : 1;
}
int oracle_switch_receiver(void* h, int irec, int state) {
    Engine& e = *(Engine*)h;
    if (irec < 1 || irec > (int)e.receivers.size()) { e.errstr = "receiver index out of range"; return 1; }
    receiver_set_enabled(e.receivers[irec - 1], state != 0); return 0;
}
int oracle_set_source_location(void* h, float lat_deg, float lon_deg, double ref_time) {
    set_source_location(*(Engine*)h, lat_deg, lon_deg, ref_time); return 0;
}
int oracle_set_effective_dt(void* h, float dt) { ((Engine*)h)->effective_dt = dt; return 0; }
int oracle_set_ref_seismogram(void* h, int irec, int icomp, float tbegin, int n, const float* data) {
    Engine& e = *(Engine*)h;
    if (!set_ref_seismogram(e, irec, icomp, data, n, tbegin)) return 1;
    finish_ref_seismograms(e);
    return 0;
}
int oracle_set_misfit_method(void* h, int id) { ((Engine*)h)->misfit_method = id; return 0; }
int oracle_set_misfit_taper(void* h, int irec, int n, const float* x, const float* y) { return set_misfit_taper(*(Engine*)h, irec, x, y, n) ? 0 : 1; }
int oracle_set_misfit_filter(void* h, int irec, int n, const float* x, const float* y) { return set_misfit_filter(*(Engine*)h, irec, x, y, n) ? 0 : 1; }
int oracle_set_synthetics_factor(void* h, float f) { set_synthetics_factor(*(Engine*)h, f); return 0; }
int oracle_set_floating_shiftrange(void* h, int irec, float lo, float hi) {  // minimizer_engine.f90:418-451
    Engine& e = *(Engine*)h;
    int r[2] = {f_nint(lo / e.db.dt), f_nint(hi / e.db.dt)};
    if (irec == 0) { for (auto& rc : e.receivers) { rc.floating_shiftrange[0] = r[0]; rc.floating_shiftrange[1] = r[1]; } }
    else if (irec < 1 || irec > (int)e.receivers.size()) { e.errstr = "receiver index out of range"; return 1; }
    else { e.receivers[irec - 1].floating_shiftrange[0] = r[0]; e.receivers[irec - 1].floating_shiftrange[1] = r[1]; }
    return 0;
}
// Under the fresh-state semantics a shifted reference is a reference set anew with its first sample moved by ishift samples
static void reinit_shifted_ref(Engine& e, int ir, int ishift) {
    Receiver& r = e.receivers[ir];
    for (int c = 0; c < r.ncomponents; c++) {
        const Probe& old = e.ref_probes_initial[ir][c];
        Probe np; probe_init(np, r.dt);
        np.taper = old.taper; np.filter = old.filter; np.factor = old.factor;
        if (old.array.alloc) {
            Strip strip;
            strip_init(old.dataspan[0] + ishift, old.dataspan[1] + ishift, &old.array.d[old.dataspan[0] - old.array.lo], slen(old.dataspan), strip);
            probe_set_array(np, strip);
        }
        e.ref_probes_initial[ir][c] = np;
        r.ref_probes[c] = np;
    }
}
// shift_ref_seismogram (minimizer_engine.f90:354-378)
int oracle_shift_ref_seismogram(void* h, int irec, float shift) {
    Engine& e = *(Engine*)h;
    if (!e.ref_probes_inited) { e.errstr = "no reference seismograms set"; return 1; }
    if (irec < 1 || irec > (int)e.receivers.size()) { e.errstr = "receiver index out of range"; return 1; }
    reinit_shifted_ref(e, irec - 1, f_nint(shift / e.db.dt));
    return 0;
}
// autoshift_ref_seismogram (minimizer_engine.f90:380-416): update_misfits of the current source (as the first evaluation of a fresh
// process), then receiver_autoshift_ref_seismogram; shifts[] in seconds, one per receiver (irec = 0) or one
int oracle_autoshift_ref_seismogram(void* h, int irec, float lo, float hi, float* shifts) {
    Engine& e = *(Engine*)h;
    if (e.cur_params.empty()) { e.errstr = "no source parameters set"; return 1; }
    std::vector<float> params = e.cur_params;
    if (evaluate(e, e.cur_type, params.data(), (int)params.size(), nullptr, 0) < 0) return 1;
    const int r[2] = {f_nint(lo / e.db.dt), f_nint(hi / e.db.dt)};
    if (irec != 0 && (irec < 1 || irec > (int)e.receivers.size())) { e.errstr = "receiver index out of range"; return 1; }
    const int i0 = irec == 0 ? 0 : irec - 1, i1 = irec == 0 ? (int)e.receivers.size() : irec;
    for (int i = i0; i < i1; i++) {
        const int ishift = receiver_autoshift_ref_seismogram(e.receivers[i], r);
        shifts[i - i0] = ishift * e.db.dt;
        reinit_shifted_ref(e, i, ishift);
    }
    return 0;
}
// output_cross_correlations (minimizer_engine.f90:1283-1306) in memory: update_syn_probes of the current source in a fresh process,
// then receiver_calculate_cross_correlations; cc[component][shift]
int oracle_get_cross_correlations(void* h, int irec, float lo, float hi, float* cc, int* ncomp, int* nshift) {
    Engine& e = *(Engine*)h;
    if (e.cur_params.empty()) { e.errstr = "no source parameters set"; return 1; }
    if (irec < 1 || irec > (int)e.receivers.size()) { e.errstr = "receiver index out of range"; return 1; }
    std::vector<float> params = e.cur_params;
    if (!set_source_params(e, e.cur_type, params.data(), (int)params.size())) return 1;
    if (!calculate_seismograms(e)) return 1;
    if (!scale_seismograms(e)) return 1;
    const int r[2] = {f_nint(lo / e.db.dt), f_nint(hi / e.db.dt)};
    Receiver& rc = e.receivers[irec - 1];
    std::vector<float> v;
    if (rc.enabled) receiver_calculate_cross_correlations(rc, r, v);
    for (size_t i = 0; i < v.size(); i++) cc[i] = v[i];
    *ncomp = rc.enabled ? rc.ncomponents : 0; *nshift = slen(r);
    return 0;
}
// ---- Gulunay interpolation of a database (gfdb.f90:1109-1310, interpolation.f90), every block eagerly ----------------
static InterpGfdb g_interp;
int oracle_gfdb_interpolate(int nx, int nz, int ng, float dt, float dx, float dz, float firstx, float firstz, const int* span0, const int* len,
                            const long long* offset, const float* data, int nipx, int nipz) {
    Gfdb src;
    gfdb_from_arrays(src, nx, nz, ng, dt, dx, dz, firstx, firstz, span0, len, offset, data);
    interp_gfdb_init(g_interp, src, nipx, nipz);
    gfdb_interpolate_all(g_interp);
    return 0;
}
int oracle_interp_meta(int* nx, int* nz, float* dx, float* dz) {
    *nx = g_interp.db.nx; *nz = g_interp.db.nz; *dx = g_interp.db.dx; *dz = g_interp.db.dz; return 0;
}
// dense samples of trace (ix, iz, ig) of the interpolated database over its span (gaps between strips are zeros)
int oracle_interp_trace(int ix, int iz, int ig, int* span0, int* len, float* buf, int cap) {
    Trace* t = gfdb_get_trace(g_interp.db, ix, iz, ig);
    if (!t) { *span0 = 0; *len = 0; return 0; }
    *span0 = t->span[0]; *len = t->span[1] - t->span[0] + 1;
    if (*len > cap) return 1;
    for (int i = 0; i < *len; i++) buf[i] = 0.f;
    for (const Strip& st : t->strips) for (int i = st.lo; i <= st.hi(); i++) buf[i - t->span[0]] = (float)st.at(i);
    return 0;
}
// gulunay on a caller-provided field (t, s1, s2) -> (t, s1*l1, s2*l2); a is tapered in place
int oracle_gulunay(float* a, int t, int s1, int s2, int l1, int l2, float* out, int ntmargin, int margin1, int margin2) {
    std::vector<float> A(a, a + (size_t)t * s1 * s2), I;
    gulunay(A, t, s1, s2, l1, l2, I, ntmargin, margin1, margin2);
    for (size_t i = 0; i < A.size(); i++) a[i] = A[i];
    for (size_t i = 0; i < I.size(); i++) out[i] = I[i];
    return 0;
}
// get_distances (minimizer_engine.f90:1260-1281)
int oracle_get_distances(void* h, double* distances, double* azimuths) {
    Engine& e = *(Engine*)h;
    for (size_t i = 0; i < e.receivers.size(); i++) {
        azimuths[i] = azimuth(e.psm.origin, e.receivers[i].origin);
        distances[i] = distance_accurate50m(e.psm.origin, e.receivers[i].origin);
    }
    return (int)e.receivers.size();
}
// psm_get_crustal_thickness (parameterized_source.f90:207-221)
int oracle_get_source_crustal_thickness(void* h, float* thickness) {
    Engine& e = *(Engine*)h;
    if (!e.crust.loaded) { e.errstr = "crust2x2 model not loaded"; return 1; }
    Crust1dProfile profile = crust2x2_get_profile(e.crust, r2d_tgc(e.psme.origin));
    float vp, vs, vrho;
    crust2x2_get_profile_averages(profile, vp, vs, vrho, *thickness);
    if (e.psme.crustal_thickness_limit > 0) *thickness = std::min(e.psme.crustal_thickness_limit, *thickness);
    return 0;
}
// p- and t-axis of a shear source: source_bilat.f90:232-237 with polar / domeshot / wrap (:565-594)
static inline float wrap_f(float x, float mi, float ma) { return x - std::floor((x - mi) / (ma - mi)) * (ma - mi); }
int oracle_principal_axes(float strike_deg, float dip_deg, float rake_deg, float* pax, float* tax) {
    float rot[3][3];
    init_euler(d2r_r(dip_deg), d2r_r(strike_deg), -d2r_r(rake_deg), rot);
    const float sq = std::sqrt(2.f);
    for (int which = 0; which < 2; which++) {
        const float v[3] = {which == 0 ? sq : -sq, 0.f, -sq};
        float x[3];
        for (int i = 0; i < 3; i++) { float a = 0.f; for (int j = 0; j < 3; j++) a = a + rot[i][j] * v[j]; x[i] = a; }
        float pol[3] = {std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]), std::atan2(x[1], x[0]), 0.f};
        pol[2] = std::acos(x[2] / pol[0]);
        float d[3] = {pol[0], wrap_f(pol[1], pi, -pi), wrap_f(pol[2], pi, -pi)};
        if (d[2] > pi / 2.f) { d[1] = wrap_f(d[1] + pi, -pi, pi); d[2] = pi - d[2]; }
        float* out = which == 0 ? pax : tax;
        out[0] = r2d_r(d[1]); out[1] = r2d_r(d[2]);
    }
    return 0;
}
// output_seismograms / output_seismogram_spectra (minimizer_engine.f90:947-1067) in memory: update_syn_probes of the current source in a
// fresh process, then probe_get / probe_get_amp_spectrum of the synthetic (which_probe 0) or reference (1) probe
int oracle_get_probe(void* h, int irec, int icomp, int which_probe, int which_processing, int spectrum, int* first, int* n, float* df, float* buf, int cap) {
    Engine& e = *(Engine*)h;
    if (e.cur_params.empty()) { e.errstr = "no source parameters set"; return 1; }
    if (irec < 1 || irec > (int)e.receivers.size()) { e.errstr = "receiver index out of range"; return 1; }
    std::vector<float> params = e.cur_params;
    if (!set_source_params(e, e.cur_type, params.data(), (int)params.size())) return 1;
    if (!calculate_seismograms(e)) return 1;
    if (!scale_seismograms(e)) return 1;
    Receiver& rc = e.receivers[irec - 1];
    if (icomp < 1 || icomp > rc.ncomponents) { e.errstr = "component index out of range"; return 1; }
    Probe& p = which_probe == 0 ? rc.syn_probes[icomp - 1] : rc.ref_probes[icomp - 1];
    std::vector<float> v;
    *first = 1; *df = 0.f;
    if (spectrum) probe_get_amp_spectrum(p, which_processing, *df, v); else probe_get(p, which_processing, *first, v);
    *n = (int)v.size();
    for (int i = 0; i < std::min(*n, cap); i++) buf[i] = v[i];
    return 0;
}
int oracle_set_crust2x2(void* h, const char* path) {
    Engine& e = *(Engine*)h;
    if (!crust2x2_load(path, e.crust)) { e.errstr = "can't load crust2x2 table"; return 1; }
    if (e.source_location_inited) psm_set_default_constraints(e.psme, e.crust);
    return 0;
}
int oracle_set_source_constraints(void* h, int n, const float* points, const float* normals) {   // parameterized_source.f90:147-166
    Engine& e = *(Engine*)h;
    e.psme.constraints.assign(n, HalfSpace());
    for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) { e.psme.constraints[i].point[k] = points[3 * i + k]; e.psme.constraints[i].normal[k] = normals[3 * i + k]; }
    return 0;
}
int oracle_set_source_crustal_thickness_limit(void* h, float limit) {
    Engine& e = *(Engine*)h;
    e.psme.crustal_thickness_limit = limit;
    if (e.crust.loaded && e.source_location_inited) psm_set_default_constraints(e.psme, e.crust);
    return 0;
}
// eikonal solver on a caller-provided speed field (test_eikonal.f90): column-major speed(nx,ny)
int oracle_eikonal_fmm(int nx, int ny, const float* speed, const float* origin, const float* delta, const float* initialpoint, float* times) {
    Field sp, tm; sp.alloc(nx, ny);
    for (int i = 0; i < nx * ny; i++) sp.a[i + 1] = speed[i];
    eikonal_solver_fmm(sp, origin, delta, initialpoint, tm);
    for (int i = 0; i < nx * ny; i++) times[i] = tm.a[i + 1];
    return 0;
}
int oracle_get_nmisfits(void* h) {
    Engine& e = *(Engine*)h; int n = 0;
    for (auto& r : e.receivers) if (r.enabled) n += r.ncomponents;
    return n;
}
// batched in form, sequential in execution (this IS the reference's loop, seismosizer.py:703-716)
int oracle_eval_sources(void* h, int sourcetype, int ns, int nparams, const float* params, float* misfits, int* status) {
    Engine& e = *(Engine*)h;
    int nm = oracle_get_nmisfits(h);
    for (int s = 0; s < ns; s++) {
        int n = evaluate(e, sourcetype, params + (size_t)s * nparams, nparams, misfits ? misfits + (size_t)s * nm * 2 : nullptr, nm);
        if (status) status[s] = (n == nm) ? 0 : 1;
        if (n < 0 && !status) return 1;
    }
    return 0;
}
float oracle_get_global_misfit(void* h) { return ((Engine*)h)->misfit; }
int oracle_get_floating_shifts(void* h, int* shifts) {
    Engine& e = *(Engine*)h; int n = 0;
    for (auto& r : e.receivers) if (r.enabled) shifts[n++] = r.floating_shift;
    return n;
}
// synthetic displacement of the last evaluated source (receiver%displacement, before scaling)
// which: 0 = displacement strip, 1 = syn probe array over dataspan (scaled by moment)
int oracle_get_seismogram(void* h, int irec, int icomp, int which, int* first_index, int* n, float* buf, int cap) {
    Engine& e = *(Engine*)h;
    if (irec < 1 || irec > (int)e.receivers.size()) return 1;
    Receiver& r = e.receivers[irec - 1];
    if (icomp < 1 || icomp > r.ncomponents) return 1;
    if (which == 0) {
        const Strip& s = r.displacement[icomp - 1];
        if (!s.alloc) { e.errstr = "no synthetic seismogram available"; return 1; }
        *first_index = s.lo; *n = strip_length(s);
        for (int i = 0; i < std::min(*n, cap); i++) buf[i] = s.d[i];
    } else {
        const Probe& p = r.syn_probes[icomp - 1];
        if (!p.array.alloc) { e.errstr = "no synthetic seismogram available"; return 1; }
        *first_index = p.dataspan[0]; *n = slen(p.dataspan);
        for (int i = 0; i < std::min(*n, cap); i++) buf[i] = p.array.at(p.dataspan[0] + i);
    }
    return 0;
}
int oracle_get_probe_spans(void* h, int irec, int icomp, int* out8) {
    Engine& e = *(Engine*)h;
    Receiver& r = e.receivers[irec - 1];
    const Probe& a = r.ref_probes[icomp - 1]; const Probe& b = r.syn_probes[icomp - 1];
    out8[0] = a.span[0]; out8[1] = a.span[1]; out8[2] = a.dataspan[0]; out8[3] = a.dataspan[1];
    out8[4] = b.span[0]; out8[5] = b.span[1]; out8[6] = b.dataspan[0]; out8[7] = b.dataspan[1];
    return 0;
}
// discretisation only: centroid table in the reference's AoS order (10 floats per centroid:
// north east depth time mxx myy mzz mxy mxz myz).  Returns ncentroids (or -1); grid = nx,ny,nt.
int oracle_discretize_source(void* h, int sourcetype, int nparams, const float* params, float* table, int cap, int* grid3) {
    Engine& e = *(Engine*)h;
    if (!set_source_params(e, sourcetype, params, nparams)) return -1;
    bool ok;
    if (sourcetype == PSM_EIKONAL || sourcetype == PSM_MT_EIKONAL) {
        ok = psm_to_tdsm_eikonal(e.psme, e.crust, e.tdsm, e.effective_dt, e.errstr);
        e.psm.grid_size = {e.psme.grid_size[0], e.psme.grid_size[1]};
    } else psm_to_tdsm(e.psm, e.tdsm, e.effective_dt, ok);
    if (!ok) return -1;
    int n = (int)e.tdsm.centroids.size();
    for (int i = 0; i < std::min(n, cap); i++) {
        const Centroid& c = e.tdsm.centroids[i];
        float* t = table + (size_t)i * 10;
        t[0] = c.north; t[1] = c.east; t[2] = c.depth; t[3] = c.time;
        for (int k = 0; k < 6; k++) t[4 + k] = c.m[k];
    }
    if (grid3) for (size_t i = 0; i < 3; i++) grid3[i] = i < e.psm.grid_size.size() ? e.psm.grid_size[i] : 1;
    return n;
}
// per-centroid integer arrays of the last evaluation for receiver irec (needs record on)
void oracle_record_indices(void* h, int on) { ((Engine*)h)->record_indices = on != 0; }
int oracle_get_indices(void* h, int irec, int* ix, int* iz, int* its, float* dix, float* diz, double* dist, double* azi, double* bazi, int cap) {
    Engine& e = *(Engine*)h;
    if (irec < 1 || irec > (int)e.index_records.size()) return -1;
    auto& v = e.index_records[irec - 1];
    int n = (int)v.size();
    for (int i = 0; i < std::min(n, cap); i++) {
        ix[i] = v[i].ix0; iz[i] = v[i].iz0; its[i] = v[i].its; dix[i] = v[i].dix; diz[i] = v[i].diz;
        if (dist) dist[i] = v[i].dist; if (azi) azi[i] = v[i].azi; if (bazi) bazi[i] = v[i].bazi;
    }
    return n;
}
// per-receiver base geometry (seismogram.f90:99-100)
int oracle_receiver_geometry(void* h, int irec, double* azi, double* bazi, double* dist) {
    Engine& e = *(Engine*)h;
    Receiver& r = e.receivers[irec - 1];
    azibazi(e.psm.origin, r.origin, *azi, *bazi);
    *dist = distance_accurate50m(e.psm.origin, r.origin);
    return 0;
}
// stored span of one GF trace after trace_pack (for span parity of the slab packer)
int oracle_trace_span(void* h, int ix, int iz, int ig, int* span2, int* nstrips) {
    Engine& e = *(Engine*)h;
    Trace* t = gfdb_get_trace(e.db, ix, iz, ig);
    if (!t) return 1;
    span2[0] = t->span[0]; span2[1] = t->span[1]; *nstrips = t->nstrips;
    return 0;
}
// time `ns` evaluations; returns seconds of wall time
// ---- ground-motion diagnostics of the current source (minimizer_engine.f90:1174-1245), enabled receivers in order ----
// which: 1 peak velocity, 2 peak acceleration, 3 Arias intensity.  update_syn_probes of the given source in a fresh state
// (no misfit calculation before: that would widen the probe spans, comparator.f90:464-486), then the receivers' values
int oracle_get_ground_motion(void* h, int sourcetype, const float* params, int nparams, int which, float* out, int cap) {
    Engine& e = *(Engine*)h;
    if (!set_source_params(e, sourcetype, params, nparams)) return -1;
    if (!calculate_seismograms(e)) return -1;
    if (!scale_seismograms(e)) return -1;
    int n = 0;
    for (auto& r : e.receivers) {
        if (!r.enabled) continue;
        const float v = which == 3 ? receiver_get_arias_intensity(r) : receiver_get_maxabs(r, which);
        if (n < cap) out[n] = v;
        n++;
    }
    return n;
}

// ---- sub-parameters and Levenberg-Marquardt ----------------------------------------------------------------------
int oracle_set_source_params_mask(void* h, const int* mask, int n) {
    Engine& e = *(Engine*)h;
    if (!e.source_inited) { e.errstr = "no source parameters set"; return 1; }
    if (n != (int)e.cur_params.size()) { e.errstr = "wrong number of elements in source params mask"; return 1; }
    e.params_mask.assign(n, 0);
    for (int i = 0; i < n; i++) e.params_mask[i] = mask[i] ? 1 : 0;
    e.sub_mins.clear(); e.sub_maxs.clear();
    return 0;
}
int oracle_set_source_subparams(void* h, const float* sub, int n) {
    Engine& e = *(Engine*)h;
    if (!e.source_inited) { e.errstr = "no source parameters set"; return 1; }
    if (n != lm_count_mask(e)) { e.errstr = "wrong number of subparams"; return 1; }
    return set_subparams(e, sub, false) ? 0 : 1;
}
int oracle_set_source_subparams_limits(void* h, const float* mins, const float* maxs, int n) {
    Engine& e = *(Engine*)h;
    if (!e.source_inited) { e.errstr = "no source parameters set"; return 1; }
    if (n != lm_count_mask(e)) { e.errstr = "wrong number of subparam_mins"; return 1; }
    e.sub_mins.assign(mins, mins + n); e.sub_maxs.assign(maxs, maxs + n);
    return 0;
}
int oracle_get_source_subparams(void* h, float* sub, int cap) {
    Engine& e = *(Engine*)h;
    int k = 0;
    for (size_t i = 0; i < e.cur_params.size(); i++) if (lm_masked(e, i)) { if (k < cap) sub[k] = e.cur_params[i]; k++; }
    return k;
}
int oracle_minimize_lm(void* h, int* info, int* iterations, float* misfit) {
    Engine& e = *(Engine*)h;
    return minimize_lm(e, *info, *iterations, *misfit) ? 0 : 1;
}
// lmdif on a caller-supplied function (sequential): fcn(user, n, m, x, fvec) returns iflag
typedef int (*oracle_lm_fcn)(void* user, int n, int m, float* x, float* fvec);
void oracle_lmdif(oracle_lm_fcn fcn, void* user, int m, int n, float* x, float* fvec, float ftol, float xtol, float gtol, int maxfev, float epsfcn,
                  float* diag, int mode, float factor, int* info, int* nfev) {
    lm_lmdif([&](int m_, int n_, float* x_, float* f_) { return fcn(user, n_, m_, x_, f_); }, m, n, x, fvec, ftol, xtol, gtol, maxfev, epsfcn, diag, mode,
             factor, *info, *nfev);
}
float oracle_enorm(int n, const float* x) { return lm_enorm(n, x); }

double oracle_time_eval(void* h, int sourcetype, int ns, int nparams, const float* params) {
    auto t0 = std::chrono::steady_clock::now();
    oracle_eval_sources(h, sourcetype, ns, nparams, params, nullptr, nullptr);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}


// the ten packed traces gfdb_build_ahfull stores for the input line "x z nfflag ffflag" (ko_ahfull.hpp), each unpacked onto its
// span: span0[ig], len[ig], data[ig*cap ..]
int oracle_ahfull_node(float rho, float alpha, float beta, const float* stf, int nstf, float dt, float x, float z, int nfflag, int ffflag,
                       int* span0, int* len, float* data, int cap) {
    const ahfull::Medium m = ahfull::make_medium(rho, alpha, beta, stf, nstf, dt);
    Trace tr[10];
    ahfull::addentry(m, dt, x, z, nfflag != 0, ffflag != 0, tr);
    for (int ig = 0; ig < 10; ig++) {
        Strip s;
        trace_unpack(tr[ig], s);
        span0[ig] = tr[ig].span[0]; len[ig] = s.size();
        if (s.size() > cap) return 1;
        for (int i = 0; i < s.size(); i++) data[(size_t)ig * cap + i] = (float)s.d[i];
    }
    return 0;
}
}  // extern "C"
