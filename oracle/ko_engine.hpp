// TEST INFRASTRUCTURE ONLY (see ko_base.hpp).  Restates the hot-path part of
// minimizer_engine.f90: state singletons (:78-108), setters, calculate_seismograms /
// scale_seismograms / calculate_misfits (:885-945), get_misfits (:1130-1172).
//
// Evaluation semantics: the reference never shrinks receiver%displacement strips
// (seismogram.f90:102-106) nor probe spans (comparator.f90:245-249), so untapered and
// amplitude-spectrum misfits depend on which sources were evaluated earlier.  A batched evaluator
// cannot reproduce an evaluation order, so both the oracle and the CUDA path define FRESH-STATE
// semantics: every candidate is evaluated as if it were the first one after set_ref_seismograms in
// a new process.  `fresh=false` keeps the reference's history-dependent behaviour for study.
#pragma once
#include "ko_receiver.hpp"
#include "ko_eikonal.hpp"
#include "ko_lm.hpp"
#ifdef _OPENMP
#include <omp.h>
#endif

namespace ko {

struct Engine {
    Psm psm;
    PsmE psme;                         // eikonal / mt_eikonal parameters (source_eikonal.f90)
    CrustModel crust;                  // crust2x2 model (minimizer.f90:1669-1674)
    float effective_dt = 1.f;          // minimizer_engine.f90:79
    Tdsm tdsm;
    std::vector<Receiver> receivers;
    std::vector<std::vector<Probe>> ref_probes_initial;  // state right after set_ref_seismograms
    Gfdb db;
    int misfit_method = L2NORM;
    float misfit = 0.f;
    bool interpolate = false;
    int xundersample = 1, zundersample = 1;
    bool database_inited = false, receivers_inited = false, source_location_inited = false, source_inited = false,
         ref_probes_inited = false;
    bool fresh = true;
    std::string errstr;
    std::vector<Trace> scratch;  // per-thread bilinear scratch traces
    std::vector<std::vector<IndexRecord>> index_records;  // filled when record_indices
    bool record_indices = false;
    // parameter vector as set, sub-parameter mask and limits (psm%params, psm%params_mask, g_subparam_mins/maxs)
    int cur_type = 0;
    std::vector<float> cur_params;
    std::vector<char> params_mask;              // empty = all true (source_all.f90:251)
    std::vector<float> sub_mins, sub_maxs;      // empty = none
    int iterations = 0;                         // minimizer_engine.f90:104
};

// minimizer_engine.f90:165-286; coordinates in degrees as in the receivers file
static inline bool set_receivers(Engine& e, int n, const double* lat_deg, const double* lon_deg, const float* depth,
                                 const char* const* comps) {
    if (!e.database_inited) { e.errstr = "no database set"; return false; }
    e.receivers.assign(n, Receiver());
    for (int i = 0; i < n; i++) {
        GeoCoords o; o.lat = lat_deg[i]; o.lon = lon_deg[i];
        if (!receiver_init(e.receivers[i], d2r_tgc(o), depth[i], comps[i], e.db.dt)) {
            e.errstr = "initializing receiver failed";
            e.receivers.clear(); e.receivers_inited = false; return false;
        }
    }
    e.receivers_inited = true; e.ref_probes_inited = false;
    return true;
}
// minimizer.f90:485-519 + minimizer_engine.f90:453-467: degrees are converted in default real
static inline void set_source_location(Engine& e, float lat_deg, float lon_deg, double ref_time) {
    e.psm.origin.lat = (double)d2r_r(lat_deg);
    e.psm.origin.lon = (double)d2r_r(lon_deg);
    e.psm.ref_time = ref_time;
    e.psme.origin = e.psm.origin;
    if (e.crust.loaded) psm_set_default_constraints(e.psme, e.crust);   // psm_set_origin_and_time, parameterized_source.f90:183-196
    e.source_location_inited = true;
}
// receiver.f90:746-801 with the file reading stripped: data starts at time `tbegin` rel. to ref time
static inline bool set_ref_seismogram(Engine& e, int irec, int icomp, const float* data, int n, float tbegin) {
    if (irec < 1 || irec > (int)e.receivers.size()) { e.errstr = "receiver index out of range"; return false; }
    Receiver& r = e.receivers[irec - 1];
    if (icomp < 1 || icomp > r.ncomponents) { e.errstr = "component index out of range"; return false; }
    Strip strip;
    seismogram_to_strip(data, n, tbegin, r.dt, strip);
    probe_set_array(r.ref_probes[icomp - 1], strip);
    return true;
}
static inline void finish_ref_seismograms(Engine& e) {
    e.ref_probes_initial.clear();
    for (auto& r : e.receivers) e.ref_probes_initial.push_back(r.ref_probes);
    e.ref_probes_inited = true;
}
static inline void sync_initial_probe_settings(Engine& e) {
    if (!e.ref_probes_inited) return;
    for (size_t i = 0; i < e.receivers.size(); i++)
        for (int c = 0; c < e.receivers[i].ncomponents; c++) {
            Probe& p = e.ref_probes_initial[i][c];
            p.taper = e.receivers[i].ref_probes[c].taper;
            p.filter = e.receivers[i].ref_probes[c].filter;
            dirtyfy_array_tapered(p);
        }
}
// minimizer_engine.f90:668-698: ireceiver 1..n only
static inline bool set_misfit_taper(Engine& e, int irec, const float* x, const float* y, int n) {
    if (!e.receivers_inited) { e.errstr = "no receivers set"; return false; }
    if (irec < 1 || irec > (int)e.receivers.size()) { e.errstr = "receiver index out of range"; return false; }
    Plf t; plf_make(t, std::vector<float>(x, x + n), std::vector<float>(y, y + n));
    receiver_set_taper(e.receivers[irec - 1], t);
    sync_initial_probe_settings(e);
    return true;
}
// minimizer_engine.f90:632-666: ireceiver 0 = all receivers
static inline bool set_misfit_filter(Engine& e, int irec, const float* x, const float* y, int n) {
    if (!e.receivers_inited) { e.errstr = "no receivers set"; return false; }
    if (irec < 0 || irec > (int)e.receivers.size()) { e.errstr = "receiver index out of range"; return false; }
    Plf f; plf_make(f, std::vector<float>(x, x + n), std::vector<float>(y, y + n));
    if (irec == 0) { for (auto& r : e.receivers) receiver_set_filter(r, f); }
    else receiver_set_filter(e.receivers[irec - 1], f);
    sync_initial_probe_settings(e);
    return true;
}
static inline void set_synthetics_factor(Engine& e, float f) { for (auto& r : e.receivers) receiver_set_synthetics_factor(r, f); }

// minimizer_engine.f90:500-523 (change detection is a speed-up only; always re-evaluated here)
static inline bool set_source_params(Engine& e, int sourcetype, const float* params, int nparams) {
    if (!e.source_location_inited) { e.errstr = "no source location set"; return false; }
    bool omc;
    if (e.cur_type != sourcetype) e.params_mask.clear();   // a new source type starts with every parameter selected
    e.cur_type = sourcetype; e.cur_params.assign(params, params + nparams);
    if (sourcetype == PSM_EIKONAL || sourcetype == PSM_MT_EIKONAL) {
        if (nparams != (sourcetype == PSM_EIKONAL ? 15 : 20)) { e.errstr = "wrong number of source parameters"; return false; }
        if (!e.crust.loaded) { e.errstr = "crust2x2 model not loaded"; return false; }
        psm_set_eikonal(e.psme, params, sourcetype == PSM_MT_EIKONAL);
        e.psm.sourcetype = sourcetype; e.psm.moment = e.psme.moment; e.psm.risetime = e.psme.risetime;
        e.source_inited = true;
        return true;
    }
    if (!psm_set(e.psm, sourcetype, params, nparams, omc)) { e.errstr = "unknown source type or wrong number of parameters"; return false; }
    e.source_inited = true;
    return true;
}

static inline void reset_to_fresh_state(Engine& e) {
    for (size_t i = 0; i < e.receivers.size(); i++) {
        Receiver& r = e.receivers[i];
        for (int c = 0; c < r.ncomponents; c++) {
            r.displacement[c] = Strip();
            Probe np; probe_init(np, r.dt);
            np.taper = r.syn_probes[c].taper; np.filter = r.syn_probes[c].filter; np.factor = r.syn_probes[c].factor;
            r.syn_probes[c] = np;
            if (e.ref_probes_inited) r.ref_probes[c] = e.ref_probes_initial[i][c];
        }
    }
}

// minimizer_engine.f90:876-907 discretize_source + calculate_seismograms
static inline bool calculate_seismograms(Engine& e) {
    if (!e.database_inited) { e.errstr = "no database set"; return false; }
    if (!e.receivers_inited) { e.errstr = "no receivers set"; return false; }
    if (!e.source_location_inited) { e.errstr = "no source location set"; return false; }
    if (!e.source_inited) { e.errstr = "no source parameters set"; return false; }
    if (e.fresh) reset_to_fresh_state(e);
    bool ok;
    if (e.psm.sourcetype == PSM_EIKONAL || e.psm.sourcetype == PSM_MT_EIKONAL) {
        ok = psm_to_tdsm_eikonal(e.psme, e.crust, e.tdsm, e.effective_dt, e.errstr);
        e.tdsm.origin = e.psm.origin; e.tdsm.ref_time = e.psm.ref_time;
        e.psm.grid_size = {e.psme.grid_size[0], e.psme.grid_size[1]};
    } else psm_to_tdsm(e.psm, e.tdsm, e.effective_dt, ok);
    if (!ok) return false;
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    if ((int)e.scratch.size() < nthreads) e.scratch.resize(nthreads);
    int nr = (int)e.receivers.size();
    if (e.record_indices) e.index_records.assign(nr, {});
#pragma omp parallel for schedule(dynamic)
    for (int ir = 0; ir < nr; ir++) {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        if (e.receivers[ir].enabled)
            make_seismogram(e.tdsm, e.receivers[ir], e.db, e.interpolate, e.xundersample, e.zundersample, e.scratch[tid],
                            e.record_indices ? &e.index_records[ir] : nullptr);
    }
    return true;
}
// minimizer_engine.f90:909-921
// An enabled receiver none of whose centroids found its Green's functions has no displacement strip; the reference then takes
// strip_span of an unallocated array (undefined).  Reported as a failed evaluation here, as the CUDA engine does.
static inline bool scale_seismograms(Engine& e) {
    for (auto& r : e.receivers)
        if (r.enabled)
            for (int c = 0; c < r.ncomponents; c++)
                if (!r.displacement[c].alloc) { e.errstr = "no synthetic seismogram: every centroid is outside the database"; return false; }
    for (auto& r : e.receivers) receiver_scaled_seismograms_to_probes(r, e.psm.risetime, e.psm.moment);
    return true;
}
// minimizer_engine.f90:924-945
static inline bool calculate_misfits(Engine& e) {
    if (!e.ref_probes_inited) { e.errstr = "no reference seismograms set"; return false; }
    float misfit = 0.f, nf = 0.f;
    for (auto& r : e.receivers) {
        receiver_calculate_misfits(r, e.misfit_method);
        float s = 0.f; for (float m : r.misfits) s = s + m * m;
        misfit = misfit + s;
        s = 0.f; for (float m : r.misfits_norm_factors) s = s + m * m;
        nf = nf + s;
    }
    e.misfit = sqrtf(misfit) / sqrtf(nf);
    return true;
}
// one full evaluation: set_source_params + get_misfits; out = (misfit, norm_factor) pairs of the
// enabled receivers, receiver-major (minimizer_engine.f90:1130-1172).  Returns nmisfits or -1.
static inline int evaluate(Engine& e, int sourcetype, const float* params, int nparams, float* out, int cap) {
    if (!set_source_params(e, sourcetype, params, nparams)) return -1;
    if (!calculate_seismograms(e)) return -1;
    if (!scale_seismograms(e)) return -1;
    if (!calculate_misfits(e)) return -1;
    int n = 0;
    for (auto& r : e.receivers) {
        if (!r.enabled) continue;
        for (int c = 0; c < r.ncomponents; c++) {
            if (out && n < cap) { out[2 * n] = r.misfits[c]; out[2 * n + 1] = r.misfits_norm_factors[c]; }
            n++;
        }
    }
    return n;
}


// ---- sub-parameters and minimize_lm (minimizer_engine.f90:525-610, 729-874; source_all.f90:377-428) ----------------
static inline const std::vector<float>& psm_params_norm(int sourcetype) {
    static const std::vector<float> bilat = {1.f, 10000.f, 10000.f, 10000.f, 7e18f, 360.f, 90.f, 360.f, 360.f, 10000.f, 10000.f, 10000.f, 3000.f, 1.f};
    static const std::vector<float> circular = {1.f, 10000.f, 10000.f, 10000.f, 7e18f, 360.f, 90.f, 360.f, 10000.f, 3000.f, 1.f};
    static const std::vector<float> point_lp = {1.f, 10000.f, 10000.f, 10000.f, 7e18f, 1.f, 0.f, -1.f, 1.f, 1.f, 1.f, 20.f, 1.f};
    static const std::vector<float> eikonal = {1.f, 10000.f, 10000.f, 10000.f, 7e18f, 360.f, 90.f, 360.f, 10000.f, 10000.f, 10000.f, 360.f, 10000.f, 1.f, 1.f};
    static const std::vector<float> mt_eikonal = {1.f, 10000.f, 10000.f, 10000.f, 7e18f, 360.f, 90.f, 10000.f, 10000.f, 10000.f, 360.f, 10000.f, 1.f, 7e18f,
                                                  7e18f, 7e18f, 7e18f, 7e18f, 7e18f, 1.f};
    static const std::vector<float> mt = {1.f, 10000.f, 10000.f, 10000.f, 7e18f, 7e18f, 7e18f, 7e18f, 7e18f, 7e18f, 1.f};
    static const std::vector<float> none;
    switch (sourcetype) {
        case PSM_BILAT: return bilat;
        case PSM_CIRCULAR: return circular;
        case PSM_POINT_LP: return point_lp;
        case PSM_EIKONAL: return eikonal;
        case PSM_MT_EIKONAL: return mt_eikonal;
        case PSM_MOMENT_TENSOR: return mt;
    }
    return none;
}
static inline bool lm_masked(const Engine& e, size_t i) { return e.params_mask.empty() || e.params_mask[i]; }
static inline int lm_count_mask(const Engine& e) {
    int k = 0;
    for (size_t i = 0; i < e.cur_params.size(); i++) k += lm_masked(e, i) ? 1 : 0;
    return k;
}
// psm_set_subparams: psm_get_params(normalized) -> replace the masked ones -> psm_set_*(normalized)
static inline bool set_subparams(Engine& e, const float* sub, bool normalized) {
    const std::vector<float>& norm = psm_params_norm(e.cur_type);
    std::vector<float> copy(e.cur_params.size());
    for (size_t i = 0; i < copy.size(); i++) copy[i] = normalized ? e.cur_params[i] / norm[i] : e.cur_params[i];
    size_t isub = 0;
    for (size_t i = 0; i < copy.size(); i++) if (lm_masked(e, i)) copy[i] = sub[isub++];
    if (normalized) for (size_t i = 0; i < copy.size(); i++) copy[i] = copy[i] * norm[i];
    return set_source_params(e, e.cur_type, copy.data(), (int)copy.size());
}
// update_misfits of the current source (fresh-state semantics)
static inline bool update_misfits(Engine& e) {
    if (!calculate_seismograms(e)) return false;
    if (!scale_seismograms(e)) return false;
    return calculate_misfits(e);
}
static inline bool minimize_lm(Engine& e, int& info, int& iterations_, float& misfit_) {
    if (!e.source_inited) { e.errstr = "no source parameters set"; return false; }
    if (!update_misfits(e)) return false;
    const int nsubparams = lm_count_mask(e);
    int nmisfits = 0;
    for (auto& r : e.receivers) nmisfits += r.ncomponents;
    if (nsubparams <= 0 || nmisfits < nsubparams) { e.errstr = "something went wrong in minimize_lm"; return false; }
    const std::vector<float>& norm = psm_params_norm(e.cur_type);
    std::vector<float> subparams, subparams_norm;
    for (size_t i = 0; i < e.cur_params.size(); i++)
        if (lm_masked(e, i)) { subparams.push_back(e.cur_params[i] / norm[i]); subparams_norm.push_back(norm[i]); }
    std::vector<float> misfits(nmisfits), diag(nsubparams, 1.f);
    const float tol = sqrtf(lm_epsmch);
    e.iterations = 0;
    // lm_forward_step, minimizer_engine.f90:808-874
    LmFcn forward = [&](int m, int n, float* x, float* fvec) -> int {
        float penalty = 0.0f;
        if (!e.sub_mins.empty() && !e.sub_maxs.empty())
            for (int i = 0; i < n; i++) {
                if (x[i] * subparams_norm[i] < e.sub_mins[i]) {
                    penalty = penalty + fabsf(x[i] * subparams_norm[i] - e.sub_mins[i]) / fabsf(e.sub_maxs[i] - e.sub_mins[i]);
                    x[i] = e.sub_mins[i] / subparams_norm[i];
                }
                if (x[i] * subparams_norm[i] > e.sub_maxs[i]) {
                    penalty = penalty + fabsf(x[i] * subparams_norm[i] - e.sub_maxs[i]) / fabsf(e.sub_maxs[i] - e.sub_mins[i]);
                    x[i] = e.sub_maxs[i] / subparams_norm[i];
                }
            }
        if (!set_subparams(e, x, true)) return -2;
        if (!update_misfits(e)) return -2;
        int im = 0;
        for (auto& r : e.receivers)
            for (int c = 0; c < r.ncomponents; c++) {
                const float v = r.enabled ? r.misfits[c] : 0.f;
                if (!std::isfinite(v)) return -2;
                fvec[im++] = v * (1.0f + penalty);
            }
        (void)m;
        e.iterations = e.iterations + 1;
        return 0;
    };
    int nfev = 0;
    lm_lmdif(forward, nmisfits, nsubparams, subparams.data(), misfits.data(), tol, tol, 0.f, 500 * (nsubparams + 1), 0.f, diag.data(), 2, 0.01f, info, nfev);
    if (info == 8) info = 4;
    iterations_ = e.iterations;
    misfit_ = e.misfit;
    return true;
}

}  // namespace ko
