// TEST INFRASTRUCTURE ONLY (see ko_base.hpp).  Restates receiver.f90 (hot-path part) and
// seismogram.f90.
#pragma once
#include "ko_comparator.hpp"
#include "ko_gfdb.hpp"
#include "ko_source.hpp"

namespace ko {

// receiver.f90:35-48
enum Comp { C_AWAY = 1, C_COMING = -1, C_RIGHT = 2, C_LEFT = -2, C_DOWN = 3, C_UP = -3, C_NORTH = 4, C_SOUTH = -4,
            C_EAST = 5, C_WEST = -5 };
// receiver.f90:56  component_names(-5:5) = w s u l c ? a r d n e
static inline int character_to_id(char ch) {  // receiver.f90:297-313
    static const char names[11] = {'w', 's', 'u', 'l', 'c', '?', 'a', 'r', 'd', 'n', 'e'};
    for (int i = -5; i <= 5; i++) if (ch == names[i + 5]) return i;
    return 0;
}

struct Receiver {  // receiver.f90:58-99
    bool enabled = false;
    float dt = 0.f;
    GeoCoords origin;
    float depth = 0.f;
    int ncomponents = 0;
    std::vector<int> components;
    std::vector<Strip> displacement;
    std::vector<float> misfits, misfits_norm_factors;
    std::vector<Probe> ref_probes, syn_probes;
    int floating_shiftrange[2] = {0, 0};
    int floating_shift = 0;
};

// receiver.f90:136-209
static inline bool receiver_init(Receiver& self, const GeoCoords& origin, float depth, const std::string& components_str, float dt) {
    self = Receiver();
    self.enabled = true;
    self.dt = dt;
    int nc = (int)components_str.size();
    self.components.assign(nc, 0);
    if (nc == 0) self.enabled = false;
    for (int i = 0; i < nc; i++) {
        int id = character_to_id(components_str[i]);
        if (id == 0) { self.components.clear(); return false; }
        for (int j = 0; j < i; j++) if (abs(self.components[j]) == abs(id)) { self.components.clear(); return false; }
        self.components[i] = id;
    }
    self.ncomponents = nc;
    self.origin = origin;
    self.depth = depth;
    self.displacement.assign(nc, Strip());
    self.ref_probes.assign(nc, Probe());
    self.syn_probes.assign(nc, Probe());
    self.misfits.assign(nc, 0.f);
    self.misfits_norm_factors.assign(nc, 0.f);
    for (int i = 0; i < nc; i++) { probe_init(self.ref_probes[i], dt); probe_init(self.syn_probes[i], dt); }
    return true;
}
// receiver.f90:315-335 (1-based index, 0 = not available)
static inline int receiver_component_index(const Receiver& r, int component) {
    for (size_t i = 0; i < r.components.size(); i++) if (abs(r.components[i]) == abs(component)) return (int)i + 1;
    return 0;
}
// receiver.f90:337-357
static inline float receiver_component_sign(const Receiver& r, int component) {
    for (size_t i = 0; i < r.components.size(); i++)
        if (abs(r.components[i]) == abs(component)) return r.components[i] < 0 ? -1.f : 1.f;
    return 0.f;
}
// receiver.f90:275-295
static inline void receiver_set_enabled(Receiver& r, bool newstate) {
    if (!newstate) for (auto& s : r.displacement) strip_nullify(s);
    r.enabled = newstate;
}
static inline void receiver_set_filter(Receiver& r, const Plf& f) { for (int i = 0; i < r.ncomponents; i++) { probe_set_filter(r.ref_probes[i], f); probe_set_filter(r.syn_probes[i], f); } }
static inline void receiver_set_taper(Receiver& r, const Plf& t) { for (int i = 0; i < r.ncomponents; i++) { probe_set_taper(r.ref_probes[i], t); probe_set_taper(r.syn_probes[i], t); } }
static inline void receiver_set_synthetics_factor(Receiver& r, float f) { for (int i = 0; i < r.ncomponents; i++) probe_set_factor(r.syn_probes[i], f); }

// receiver.f90:439-510
static inline void receiver_calculate_floating_misfits(Receiver& self, int misfit_method, const int shiftrange[2]) {
    int method = (misfit_method == FLOATING_L1NORM) ? L1NORM : L2NORM;
    if (self.ncomponents == 0) return;
    if (!self.enabled) {
        for (int i = 0; i < self.ncomponents; i++) { self.misfits[i] = 0.f; self.misfits_norm_factors[i] = 0.f; }
        return;
    }
    int ns = slen(shiftrange), nc = self.ncomponents;
    std::vector<float> misfits((size_t)nc * ns), norms((size_t)nc * ns);
    int ishift = shiftrange[0];
    for (int i = 0; i < ns; i++) {
        for (int ic = 0; ic < nc; ic++) {
            probe_shift(self.ref_probes[ic], ishift);
            misfits[(size_t)i * nc + ic] = probes_norm(self.ref_probes[ic], self.syn_probes[ic], method);
            norms[(size_t)i * nc + ic] = probe_norm(self.ref_probes[ic], method);
        }
        ishift = 1;
    }
    int iloc = 0; float best = 0.f;
    for (int i = 0; i < ns; i++) {  // minloc(sum(misfits[**2], 1), 1): first minimum
        float s = 0.f;
        for (int ic = 0; ic < nc; ic++) { float m = misfits[(size_t)i * nc + ic]; s = s + (method == L1NORM ? m : m * m); }
        if (i == 0 || s < best) { best = s; iloc = i; }
    }
    self.floating_shift = shiftrange[0] + iloc;
    for (int ic = 0; ic < nc; ic++) {
        self.misfits[ic] = misfits[(size_t)iloc * nc + ic];
        float s = 0.f;
        for (int i = 0; i < ns; i++) s = s + norms[(size_t)i * nc + ic];
        self.misfits_norm_factors[ic] = s / (float)ns;
    }
    for (int ic = 0; ic < nc; ic++) probe_shift(self.ref_probes[ic], -shiftrange[1]);
}
// receiver.f90:407-437
// receiver.f90:505-542: horizontal components preferably a/c and r/l, else n/s and e/w; none if incomplete
static inline void get_component_ids(const Receiver& self, int& iver, int& ihor1, int& ihor2) {
    iver = ihor1 = ihor2 = 0;
    for (int ic = 1; ic <= self.ncomponents; ic++) {
        const int ict = abs(self.components[ic - 1]);
        if (ict == 1) ihor1 = ic;
        if (ict == 2) ihor2 = ic;
        if (ict == 3) iver = ic;
    }
    if (ihor1 == 0 || ihor2 == 0)
        for (int ic = 1; ic <= self.ncomponents; ic++) {
            const int ict = abs(self.components[ic - 1]);
            if (ict == 4) ihor1 = ic;
            if (ict == 5) ihor2 = ic;
        }
    if (ihor1 == 0 || ihor2 == 0) { ihor1 = 0; ihor2 = 0; }
}
// receiver.f90:544-576: peak of the vector norm of the velocity (differentiate 1) or acceleration (2) of the synthetics
static inline float receiver_get_maxabs(Receiver& self, int differentiate) {
    if (!self.enabled) return 0.f;
    int ic[3];
    get_component_ids(self, ic[0], ic[1], ic[2]);
    Probe* p[3]; int n = 0;
    for (int i = 0; i < 3; i++) if (ic[i] != 0) p[n++] = &self.syn_probes[ic[i] - 1];
    if (n == 0) return 0.f;
    return probes_ground_motion(p, n, differentiate);
}
// receiver.f90:578-596
static inline float receiver_get_arias_intensity(Receiver& self) {
    if (!self.enabled) return 0.f;
    int iver, ihor1, ihor2;
    get_component_ids(self, iver, ihor1, ihor2);
    Probe* p[3];
    if (iver != 0 && ihor1 != 0 && ihor2 != 0) { p[0] = &self.syn_probes[iver - 1]; p[1] = &self.syn_probes[ihor1 - 1]; p[2] = &self.syn_probes[ihor2 - 1]; return probes_ground_motion(p, 3, 3); }
    if (ihor1 != 0 && ihor2 != 0) { p[0] = &self.syn_probes[ihor1 - 1]; p[1] = &self.syn_probes[ihor2 - 1]; return probes_ground_motion(p, 2, 3); }
    if (iver != 0) { p[0] = &self.syn_probes[iver - 1]; return probes_ground_motion(p, 1, 3); }
    return 0.f;
}

// receiver.f90:597-616
static inline void receiver_calculate_cross_correlations(Receiver& self, const int shiftrange[2], std::vector<float>& cross_corr /* [ncomp][nshift] */) {
    const int ns = slen(shiftrange);
    cross_corr.assign((size_t)self.ncomponents * ns, 0.f);
    for (int ic = 0; ic < self.ncomponents; ic++)
        probes_windowed_cross_corr(self.syn_probes[ic], self.ref_probes[ic], shiftrange, &cross_corr[(size_t)ic * ns]);
}
// receiver.f90:802-814
static inline void receiver_shift_ref_seismogram(Receiver& self, int ishift) {
    for (int ic = 0; ic < self.ncomponents; ic++) probe_shift(self.ref_probes[ic], ishift);
}
// receiver.f90:816-832: shift of the references that maximises the summed squared positive normalised cross-correlation
static inline int receiver_autoshift_ref_seismogram(Receiver& self, const int ishiftrange[2]) {
    int ishift = 0;
    if (!self.enabled) return ishift;
    std::vector<float> cc;
    receiver_calculate_cross_correlations(self, ishiftrange, cc);
    const int ns = slen(ishiftrange);
    float mx = -std::numeric_limits<float>::max();
    for (float v : cc) mx = std::max(mx, v);
    const float den = std::max(1.f, mx);
    int imax = 0; float best = 0.f;
    for (int i = 0; i < ns; i++) {   // maxloc(sum(max(cc/den,0.)**2, 2), 1): first maximum
        float sum = 0.f;
        for (int ic = 0; ic < self.ncomponents; ic++) { const float v = std::max(cc[(size_t)ic * ns + i] / den, 0.f); sum = sum + v * v; }
        if (i == 0 || sum > best) { best = sum; imax = i; }
    }
    ishift = imax + ishiftrange[0];
    receiver_shift_ref_seismogram(self, ishift);
    return ishift;
}

static inline void receiver_calculate_misfits(Receiver& self, int misfit_method) {
    if (misfit_method == FLOATING_L1NORM || misfit_method == FLOATING_L2NORM) {
        receiver_calculate_floating_misfits(self, misfit_method, self.floating_shiftrange);
        return;
    }
    for (int ic = 0; ic < self.ncomponents; ic++) {
        if (self.enabled) {
            self.misfits[ic] = probes_norm(self.ref_probes[ic], self.syn_probes[ic], misfit_method);
            self.misfits_norm_factors[ic] = probe_norm(self.ref_probes[ic], misfit_method);
        } else { self.misfits[ic] = 0.f; self.misfits_norm_factors[ic] = 0.f; }
    }
}
// receiver.f90:834-851: sample i of a file trace sits at index nint(tbegin/dt)+i
static inline void seismogram_to_strip(const float* seis, int n, float tbegin, float deltat, Strip& strip) {
    int ibeg = f_nint(tbegin / deltat);
    strip_init(ibeg + 1, ibeg + n, seis, n, strip);
}
// receiver.f90:853-904
static inline void receiver_scaled_seismograms_to_probes(Receiver& rc, float risetime, float moment) {
    std::vector<float> weights, shifts;
    if (rc.enabled) {
        if (risetime > 0.0f) {
            float rrise[2] = {-risetime / 2.f, +risetime / 2.f};
            int nshifts = 1 + 2 * f_nint(0.5f * risetime / rc.dt);
            weights.assign(nshifts, 0.f); shifts.assign(nshifts, 0.f);
            for (int is = 1; is <= nshifts; is++) {
                float ts = ((float)(is - 1) - 0.5f * (float)(nshifts - 1)) * rc.dt;
                float rsamp[2] = {ts - rc.dt / 2.f, ts + rc.dt / 2.f};
                float rover[2] = {std::max(rrise[0], rsamp[0]), std::min(rrise[1], rsamp[1])};
                weights[is - 1] = std::max(0.f, rover[1] - rover[0]);
                shifts[is - 1] = ts / rc.dt;
            }
            float sum = 0.f; for (float w : weights) sum = sum + w;  // sum(weights), sequential
            for (float& w : weights) w = w / sum;
        }
        for (int ic = 0; ic < rc.ncomponents; ic++) {
            Strip tmp;
            strip_copy(rc.displacement[ic], tmp);
            if (risetime > 0.0f) strip_fold(tmp, shifts, weights);
            probe_set_array(rc.syn_probes[ic], tmp, false, moment);
        }
    }
}

// ---- seismogram.f90 ---------------------------------------------------------------------------
// :316-336
static inline void make_weights(float azimuth_, const float m[6], float f[6]) {
    float sa = sinf(azimuth_), ca = cosf(azimuth_), s2a = sinf(2.f * azimuth_), c2a = cosf(2.f * azimuth_);
    f[0] = m[0] * (ca * ca) + m[1] * (sa * sa) + m[3] * s2a;
    f[1] = m[4] * ca + m[5] * sa;
    f[2] = m[2];
    f[3] = 0.5f * (m[1] - m[0]) * s2a + m[3] * c2a;
    f[4] = m[5] * ca - m[4] * sa;
    f[5] = m[0] * (sa * sa) + m[1] * (ca * ca) - m[3] * s2a;
}

// per-(centroid, receiver) integers, recorded for the bit-exact index parity tests
struct IndexRecord { int ix0, iz0, its; float dix, diz; double dist, azi, bazi; };

// :36-301.  `scratch` plays the role of greensf%interpolated_traces(thread).
static inline void make_seismogram(const Tdsm& source, Receiver& receiver, Gfdb& greensf, bool interpolate, int xundersample,
                                   int zundersample, Trace& scratch, std::vector<IndexRecord>* rec = nullptr) {
    int ja = receiver_component_index(receiver, C_AWAY), jr = receiver_component_index(receiver, C_RIGHT),
        jd = receiver_component_index(receiver, C_DOWN), jn = receiver_component_index(receiver, C_NORTH),
        je = receiver_component_index(receiver, C_EAST);
    float sa = receiver_component_sign(receiver, C_AWAY), sr = receiver_component_sign(receiver, C_RIGHT),
          sd = receiver_component_sign(receiver, C_DOWN), sn = receiver_component_sign(receiver, C_NORTH),
          se = receiver_component_sign(receiver, C_EAST);
    bool need_horizontal = ja != 0 || jr != 0 || jn != 0 || je != 0;
    double azi_orig, bazi_orig;
    azibazi(source.origin, receiver.origin, azi_orig, bazi_orig);
    double dist_orig = distance_accurate50m(source.origin, receiver.origin);
    for (int i = 0; i < receiver.ncomponents; i++) strip_nullify(receiver.displacement[i]);
    Strip temp[2], ar[2];
    if (need_horizontal) {
        if (ja) strip_extend_to_same_span({&receiver.displacement[ja - 1], &ar[0], &temp[0], &ar[1], &temp[1]});
        if (jr) strip_extend_to_same_span({&receiver.displacement[jr - 1], &ar[0], &temp[0], &ar[1], &temp[1]});
        if (jn) strip_extend_to_same_span({&receiver.displacement[jn - 1], &ar[0], &temp[0], &ar[1], &temp[1]});
        if (je) strip_extend_to_same_span({&receiver.displacement[je - 1], &ar[0], &temp[0], &ar[1], &temp[1]});
        for (int i = 0; i < 2; i++) strip_nullify(ar[i]);
    }
    if (rec) rec->clear();
    for (size_t ic = 0; ic < source.centroids.size(); ic++) {
        const Centroid& c = source.centroids[ic];
        float dnorth = c.north, deast = c.east, depth = c.depth, time = c.time;
        float f[6];
        float rshift = time / greensf.dt;
        double azi, bazi, dist;
        approx_differential_azidist(dnorth, deast, azi_orig, bazi_orig, dist_orig, azi, bazi, dist);
        make_weights((float)azi, c.m, f);
        int ix[2], iz[2]; float dix, diz;
        if (interpolate) {
            gfdb_get_indices_bilin(greensf, (float)dist, depth - receiver.depth, xundersample, zundersample, ix, iz, dix, diz);
        } else {
            gfdb_get_indices(greensf, (float)dist, depth - receiver.depth, ix[0], iz[0]);
            ix[1] = ix[0] + 1; iz[1] = iz[0] + 1; dix = 0.f; diz = 0.f;
        }
        if (rec) rec->push_back(IndexRecord{ix[0], iz[0], f_floor(rshift), dix, diz, dist, azi, bazi});
        Trace* tp;
#define KO_FETCH(ig) tp = gfdb_get_trace_bilin(greensf, ix, iz, (ig), dix, diz, scratch); if (!tp) continue;
        if (need_horizontal) {
            double lambda = bazi - bazi_orig;
            if (lambda != 0.) {
                float cl = (float)cos(lambda), sl = (float)sin(lambda);
                strip_nullify(temp[0]);
                KO_FETCH(1) trace_multiply_add(*tp, temp[0], f[0], SHIFT_REAL, 0, rshift);
                KO_FETCH(2) trace_multiply_add(*tp, temp[0], f[1], SHIFT_REAL, 0, rshift);
                KO_FETCH(3) trace_multiply_add(*tp, temp[0], f[2], SHIFT_REAL, 0, rshift);
                if (greensf.ng == 10) { KO_FETCH(9) trace_multiply_add(*tp, temp[0], f[5], SHIFT_REAL, 0, rshift); }
                strip_nullify(temp[1]);
                KO_FETCH(4) trace_multiply_add(*tp, temp[1], f[3], SHIFT_REAL, 0, rshift);
                KO_FETCH(5) trace_multiply_add(*tp, temp[1], f[4], SHIFT_REAL, 0, rshift);
                strip_extend_to_same_span({&temp[0], &temp[1], &ar[0], &ar[1]});
                int n = ar[0].size();
                for (int i = 0; i < n; i++) ar[0].d[i] = ar[0].d[i] + cl * temp[0].d[i] - sl * temp[1].d[i];
                for (int i = 0; i < n; i++) ar[1].d[i] = ar[1].d[i] + cl * temp[1].d[i] + sl * temp[0].d[i];
            } else {
                KO_FETCH(1) trace_multiply_add(*tp, ar[0], f[0], SHIFT_REAL, 0, rshift);
                KO_FETCH(2) trace_multiply_add(*tp, ar[0], f[1], SHIFT_REAL, 0, rshift);
                KO_FETCH(3) trace_multiply_add(*tp, ar[0], f[2], SHIFT_REAL, 0, rshift);
                if (greensf.ng == 10) { KO_FETCH(9) trace_multiply_add(*tp, ar[0], f[5], SHIFT_REAL, 0, rshift); }
                KO_FETCH(4) trace_multiply_add(*tp, ar[1], f[3], SHIFT_REAL, 0, rshift);
                KO_FETCH(5) trace_multiply_add(*tp, ar[1], f[4], SHIFT_REAL, 0, rshift);
            }
        }
        if (jd != 0) {
            Strip& dz = receiver.displacement[jd - 1];
            KO_FETCH(6) trace_multiply_add(*tp, dz, f[0] * sd, SHIFT_REAL, 0, rshift);
            KO_FETCH(7) trace_multiply_add(*tp, dz, f[1] * sd, SHIFT_REAL, 0, rshift);
            KO_FETCH(8) trace_multiply_add(*tp, dz, f[2] * sd, SHIFT_REAL, 0, rshift);
            if (greensf.ng == 10) { KO_FETCH(10) trace_multiply_add(*tp, dz, f[5] * sd, SHIFT_REAL, 0, rshift); }
        }
#undef KO_FETCH
    }
    if (need_horizontal) {
        if (ja) {
            Strip& d = receiver.displacement[ja - 1];
            strip_extend_to_same_span({&d, &ar[0]});
            for (int i = 0; i < d.size(); i++) d.d[i] = ar[0].d[i] * sa;
        }
        if (jr) {
            Strip& d = receiver.displacement[jr - 1];
            strip_extend_to_same_span({&d, &ar[1]});
            for (int i = 0; i < d.size(); i++) d.d[i] = ar[1].d[i] * sr;
        }
        if (jn || je) {
            float cl = (float)cos(bazi_orig + (double)pi);
            float sl = (float)sin(bazi_orig + (double)pi);
            strip_extend_to_same_span({&ar[0], &ar[1]});
            for (int i = 0; i < ar[0].size(); i++) {  // seismogram.f90:303-314 rotate
                sreal a = ar[0].d[i], b = ar[1].d[i];
                sreal aa = cl * a - sl * b;
                b = cl * b + sl * a;
                ar[0].d[i] = aa; ar[1].d[i] = b;
            }
            if (jn) {
                Strip& d = receiver.displacement[jn - 1];
                strip_extend_to_same_span({&d, &ar[0]});
                for (int i = 0; i < d.size(); i++) d.d[i] = ar[0].d[i] * sn;
            }
            if (je) {
                Strip& d = receiver.displacement[je - 1];
                strip_extend_to_same_span({&d, &ar[1]});
                for (int i = 0; i < d.size(); i++) d.d[i] = ar[1].d[i] * se;
            }
        }
    }
}

}  // namespace ko
