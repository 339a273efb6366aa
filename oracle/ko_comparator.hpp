// TEST INFRASTRUCTURE ONLY (see ko_base.hpp).  Restates comparator.f90.
// FFTW3 single precision (sfftw_plan_dft_r2c_1d / c2r, comparator.f90:1201-1210, 1244-1249) is a
// third-party system library with no pinned version (Makefile:49-50 -lfftw3f) and is absent
// from this image; it is replaced by a radix-2 fp32 transform with the same conventions
// (unnormalised forward r2c with exp(-i...), unnormalised inverse c2r).  Results agree with any
// correct fp32 FFT to ~1e-7 relative; pinned by test_comparator.f90:89-94 only.
#pragma once
#include "ko_trace.hpp"
#include <complex>

namespace ko {

enum Norm { L2NORM = 1, L1NORM = 2, AMPSPEC_L2NORM = 3, AMPSPEC_L1NORM = 4, SCALAR_PRODUCT = 5, PEAK = 6,
            FLOATING_L2NORM = 7, FLOATING_L1NORM = 8 };  // comparator.f90:33-42

typedef std::complex<float> cfloat;

// in-place radix-2 complex FFT, fp32 data, twiddles rounded from double; sign=-1 forward
static inline void fft_c(std::vector<cfloat>& a, int sign) {
    size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; i++) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        double ang = sign * 2.0 * M_PI / (double)len;
        for (size_t i = 0; i < n; i += len) {
            for (size_t k = 0; k < len / 2; k++) {
                cfloat w((float)cos(ang * (double)k), (float)sin(ang * (double)k));
                cfloat u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
        }
    }
}
static inline void fft_r2c(const sreal* in, int n, std::vector<cfloat>& out) {
    std::vector<cfloat> a(n);
    for (int i = 0; i < n; i++) a[i] = cfloat((float)in[i], 0.f);
    if (n > 1) fft_c(a, -1);
    out.assign(a.begin(), a.begin() + n / 2 + 1);
}
static inline void fft_c2r(const std::vector<cfloat>& in, int n, sreal* out) {
    std::vector<cfloat> a(n);
    for (int k = 0; k <= n / 2; k++) a[k] = in[k];
    for (int k = n / 2 + 1; k < n; k++) a[k] = std::conj(in[n - k]);
    a[0] = cfloat(a[0].real(), 0.f);
    if (n > 1) { a[n / 2] = cfloat(a[n / 2].real(), 0.f); fft_c(a, +1); }
    for (int i = 0; i < n; i++) out[i] = a[i].real();
}

struct Probe {  // comparator.f90:55-80
    float dt = 0.f, df = 0.f;
    int span[2] = {0, 0};
    int dataspan[2] = {0, 0};
    Strip array, array_tapered, array_filtered;  // arrays with arbitrary lower bound
    std::vector<cfloat> spectrum, spectrum_filtered;
    std::vector<float> amp_spectrum, amp_spectrum_filtered;
    bool array_dirty = true, array_tapered_dirty = true, spectrum_dirty = true, spectrum_filtered_dirty = true,
         array_filtered_dirty = true;
    float paddingfactor = 2.f;
    Plf taper, filter;
    float factor = 1.f;
};

static inline int slen(const int s[2]) { return s[1] - s[0] + 1; }                      // :1139-1143
static inline void span_union(const int a[2], const int b[2], int c[2]) { c[0] = std::min(a[0], b[0]); c[1] = std::max(a[1], b[1]); }
static inline void span_intersection(const int a[2], const int b[2], int c[2]) { c[0] = std::max(a[0], b[0]); c[1] = std::min(a[1], b[1]); }
static inline bool containing(const int outer[2], const int inner[2]) {                 // :1120-1128
    int is[2]; span_intersection(outer, inner, is); return is[0] == inner[0] && is[1] == inner[1];
}
static inline int next_power_of_two(int n) {  // :1111-1118 (default-real log)
    return 1 << f_ceiling(logf((float)n) / logf(2.f));
}
static inline void allowed_span(const int span[2], int minlength, int newspan[2]) {  // :1092-1109
    newspan[0] = span[0]; newspan[1] = span[1];
    int length = slen(newspan);
    if (length < minlength) length = minlength;
    int lengthp = next_power_of_two(length);
    newspan[0] = newspan[0] - f_floor((float)(lengthp - slen(span)) / 2.f);
    newspan[1] = newspan[0] + lengthp - 1;
}
static inline void discrete_plf_span(const Plf& plf, float dt, int out[2]) {  // :1145-1157
    float r0 = 0.f, r1 = -1.f;  // plf_span of an undefined plf, piecewise_linear_function.f90:124-135
    if (plf.defined) { r0 = plf.x[0]; r1 = plf.x[plf.n() - 1]; }
    out[0] = f_ceiling(r0 / dt);
    out[1] = f_floor(r1 / dt);
}

static inline void dirtyfy_array_filtered(Probe& p) { p.array_filtered_dirty = true; }
static inline void dirtyfy_spectrum_filtered(Probe& p) { p.spectrum_filtered_dirty = true; dirtyfy_array_filtered(p); }
static inline void dirtyfy_spectrum(Probe& p) { p.spectrum_dirty = true; dirtyfy_spectrum_filtered(p); }
static inline void dirtyfy_array_tapered(Probe& p) { p.array_tapered_dirty = true; dirtyfy_spectrum(p); }
static inline void dirtyfy_array(Probe& p) { p.array_dirty = true; dirtyfy_array_tapered(p); }

static inline void probe_init(Probe& p, float dt) {  // :186-200
    p = Probe();
    p.dt = dt;
}
static inline void probe_set_taper(Probe& p, const Plf& plf) { p.taper = plf; dirtyfy_array_tapered(p); }   // :436-444
static inline void probe_set_filter(Probe& p, const Plf& plf) { p.filter = plf; dirtyfy_spectrum_filtered(p); }  // :446-454
static inline void probe_set_factor(Probe& p, float f) { p.factor = f; }  // :456-462

// :222-271
static inline void probe_set_array(Probe& self, const Strip& strip, bool allow_shrink = false, float factor = 1.f) {
    self.dataspan[0] = strip.lo; self.dataspan[1] = strip.hi();
    int newspan[2];
    int sspan[2] = {strip.lo, strip.hi()};
    if (allow_shrink || !self.array.alloc) { newspan[0] = sspan[0]; newspan[1] = sspan[1]; }
    else span_union(sspan, self.span, newspan);
    int datalength = slen(self.dataspan);
    int tmp[2] = {newspan[0], newspan[1]};
    allowed_span(tmp, f_ceiling((float)datalength * self.paddingfactor), newspan);
    if (newspan[0] != self.span[0] || newspan[1] != self.span[1]) {
        resize(self.array, newspan[0], slen(newspan));
        resize(self.array_tapered, newspan[0], slen(newspan));
    }
    self.span[0] = newspan[0]; self.span[1] = newspan[1];
    if (self.span[0] <= self.dataspan[0] - 1) for (int i = self.span[0]; i <= self.dataspan[0] - 1; i++) self.array.at(i) = 0.f;
    for (int i = self.dataspan[0]; i <= self.dataspan[1]; i++) self.array.at(i) = strip.at(i) * factor;
    if (self.dataspan[1] + 1 <= self.span[1])
        for (int i = self.dataspan[1] + 1; i <= self.span[1]; i++) self.array.at(i) = self.array.at(self.dataspan[1]);
    dirtyfy_array(self);
}
// :273-288
static inline void probe_shift(Probe& self, int ishift) {
    if (!self.array.alloc) return;
    Strip strip;
    int nds[2] = {self.dataspan[0] + ishift, self.dataspan[1] + ishift};
    strip_init(nds[0], nds[1], &self.array.d[self.dataspan[0] - self.array.lo], slen(self.dataspan), strip);
    probe_set_array(self, strip);
}
// :291-330
static inline void probe_extend_span(Probe& self, const int span[2]) {
    int newspan[2];
    if (!self.array.alloc) {
        allowed_span(span, 0, newspan);
        resize(self.array, span[0], slen(span));
        std::fill(self.array.d.begin(), self.array.d.end(), 0.f);
        resize(self.array_tapered, span[0], slen(span));
        return;
    }
    int u[2];
    span_union(span, self.dataspan, u);
    allowed_span(u, 0, newspan);
    if (self.span[0] == newspan[0] && self.span[1] == newspan[1]) return;
    Strip temp;
    strip_init(self.dataspan[0], self.dataspan[1], &self.array.d[self.dataspan[0] - self.array.lo], slen(self.dataspan), temp);
    resize(self.array, newspan[0], slen(newspan));
    resize(self.array_tapered, newspan[0], slen(newspan));
    self.span[0] = newspan[0]; self.span[1] = newspan[1];
    if (self.span[0] <= self.dataspan[0] - 1) for (int i = self.span[0]; i <= self.dataspan[0] - 1; i++) self.array.at(i) = 0.f;
    for (int i = self.dataspan[0]; i <= self.dataspan[1]; i++) self.array.at(i) = temp.at(i);
    if (self.dataspan[1] + 1 <= self.span[1])
        for (int i = self.dataspan[1] + 1; i <= self.span[1]; i++) self.array.at(i) = self.array.at(self.dataspan[1]);
    dirtyfy_array(self);
}
// :464-486
static inline void probes_adjust_spans(Probe& a, Probe& b) {
    if (a.dt != b.dt) { fprintf(stderr, "probes_adjust_spans(): both probes must have same dt\n"); abort(); }
    int u[2], newspan[2];
    span_union(a.dataspan, b.dataspan, u);
    int minlength = std::max(f_ceiling((float)slen(a.dataspan) * a.paddingfactor), f_ceiling((float)slen(b.dataspan) * b.paddingfactor));
    allowed_span(u, minlength, newspan);
    if (a.span[0] == b.span[0] && a.span[1] == b.span[1] && slen(a.span) == slen(newspan) && containing(a.span, b.dataspan) &&
        containing(b.span, a.dataspan))
        return;
    probe_extend_span(a, newspan);
    probe_extend_span(b, newspan);
}

// :1173-1184
static inline void make_array_tapered(Probe& self) {
    if (self.taper.defined) {
        self.array_tapered.d = self.array.d;
        Strip& at = self.array_tapered;
        plf_taper_generic(self.taper, self.dataspan[0], self.span[1], self.dt, ip_cos,
                          [&](int j, float f) { at.at(j) = at.at(j) * f; }, [&](int j) { at.at(j) = 0.f; });
    }
}
// :1186-1215
static inline void make_spectrum(Probe& self) {
    int ntrans = self.array.size();
    if (self.taper.defined) fft_r2c(self.array_tapered.d.data(), ntrans, self.spectrum);
    else fft_r2c(self.array.d.data(), ntrans, self.spectrum);
    self.spectrum_filtered.resize(ntrans / 2 + 1);
    self.amp_spectrum.resize(ntrans / 2 + 1);
    self.amp_spectrum_filtered.resize(ntrans / 2 + 1);
    for (int k = 0; k <= ntrans / 2; k++) self.amp_spectrum[k] = std::abs(self.spectrum[k]);
    self.df = 1.f / ((float)ntrans * self.dt);
}
// :1217-1231
static inline void make_spectrum_filtered(Probe& self) {
    if (self.filter.defined) {
        self.amp_spectrum_filtered = self.amp_spectrum;
        self.spectrum_filtered = self.spectrum;
        int n = (int)self.spectrum_filtered.size();
        plf_taper_generic(self.filter, 0, n - 1, self.df, ip_cos,
                          [&](int j, float f) { self.spectrum_filtered[j] = self.spectrum_filtered[j] * f; },
                          [&](int j) { self.spectrum_filtered[j] = cfloat(0.f, 0.f); });
        plf_taper_generic(self.filter, 0, n - 1, self.df, ip_cos,
                          [&](int j, float f) { self.amp_spectrum_filtered[j] = self.amp_spectrum_filtered[j] * f; },
                          [&](int j) { self.amp_spectrum_filtered[j] = 0.f; });
    }
}
// :1233-1263
static inline void make_array_filtered(Probe& self) {
    if (self.filter.defined) {
        int ntrans = self.array.size();
        resize(self.array_filtered, self.array.lo, ntrans);
        fft_c2r(self.spectrum_filtered, ntrans, self.array_filtered.d.data());
        for (int i = 0; i < ntrans; i++) self.array_filtered.d[i] = self.array_filtered.d[i] / (float)ntrans;
        if (self.taper.defined) {
            Strip& af = self.array_filtered;
            plf_taper_generic(self.taper, self.span[0], self.span[1], self.dt, ip_zero_one,
                              [&](int j, float f) { af.at(j) = af.at(j) * f; }, [&](int j) { af.at(j) = 0.f; });
        }
    }
}
// :1267-1306 dataflow
static inline void update_array(Probe& p) { p.array_dirty = false; }
static inline void update_array_tapered(Probe& p) { update_array(p); if (p.array_tapered_dirty) make_array_tapered(p); p.array_tapered_dirty = false; }
static inline void update_spectrum(Probe& p) { update_array_tapered(p); if (p.spectrum_dirty) make_spectrum(p); p.spectrum_dirty = false; }
static inline void update_spectrum_filtered(Probe& p) { update_spectrum(p); if (p.spectrum_filtered_dirty) make_spectrum_filtered(p); p.spectrum_filtered_dirty = false; }
static inline void update_array_filtered(Probe& p) { update_spectrum_filtered(p); if (p.array_filtered_dirty) make_array_filtered(p); p.array_filtered_dirty = false; }

// probe_get_plain / _tapered / _filtered (:356-433) and probe_get_amp_spectrum (:332-354): the samples over the span the reference
// hands out; which_processing 0 plain, 1 tapered, 2 filtered
static inline void probe_get(Probe& self, int which_processing, int& first, std::vector<float>& out) {
    out.clear(); first = 1;
    if (!self.array.alloc) { out.push_back(0.f); return; }
    int span[2] = {self.dataspan[0], self.dataspan[1]};
    const Strip* src = &self.array;
    if (which_processing == 2 && self.filter.defined) {
        update_array_filtered(self);
        if (self.taper.defined) {
            int d[2]; discrete_plf_span(self.taper, self.dt, d); span_intersection(d, self.span, span);
            if (span[0] > span[1]) { span[0] = self.dataspan[0]; span[1] = self.dataspan[1]; }
        }
        src = &self.array_filtered;
    } else if (which_processing >= 1 && self.taper.defined) {
        update_array_tapered(self);
        int d[2]; discrete_plf_span(self.taper, self.dt, d); span_intersection(d, self.dataspan, span);
        if (span[0] > span[1]) { span[0] = self.dataspan[0]; span[1] = self.dataspan[1]; }
        src = &self.array_tapered;
    }
    first = span[0];
    for (int i = span[0]; i <= span[1]; i++) out.push_back((float)src->at(i));
}
static inline void probe_get_amp_spectrum(Probe& self, int which_processing, float& df, std::vector<float>& out) {
    out.clear(); df = 0.f;
    if (!self.array.alloc) { out.push_back(0.f); return; }
    update_spectrum_filtered(self);
    df = self.df;
    out = (self.filter.defined && which_processing == 2) ? self.amp_spectrum_filtered : self.amp_spectrum;
}

// norm functions, :627-697: element arithmetic in fp32, sum in fp64, result fp32
struct Norm2 { virtual float f(const float* a, const float* b, int n, float dt, float fa, float fb) const = 0; virtual ~Norm2() {} };
template <class T>
static inline float scalar_product_2(const T* a, const T* b, int n, float dt, float fa, float fb) {
    (void)dt; double s = 0.;
    if (fa == 1.f && fb == 1.f) for (int i = 0; i < n; i++) s += (double)(a[i] * b[i]);
    else for (int i = 0; i < n; i++) s += (double)(a[i] * fa * b[i] * fb);
    return (float)s;
}
template <class T>
static inline float l1norm_func(const T* a, const T* b, int n, float dt, float fa, float fb) {
    double s = 0.;
    if (fa == 1.f && fb == 1.f) for (int i = 0; i < n; i++) s += (double)std::abs(a[i] - b[i]);
    else for (int i = 0; i < n; i++) s += (double)std::abs(fa * a[i] - fb * b[i]);
    return (float)((double)dt * s);
}
template <class T>
static inline float l2norm_func(const T* a, const T* b, int n, float dt, float fa, float fb) {
    double s = 0.;
    if (fa == 1.f && fb == 1.f) for (int i = 0; i < n; i++) { double d = (double)(a[i] - b[i]); s += d * d; }
    else for (int i = 0; i < n; i++) { double d = (double)(fa * a[i] - fb * b[i]); s += d * d; }
    return (float)sqrt((double)dt * s);
}
template <class T>
static inline float maxabs_func(const T* a, const T* b, int n, float dt, float fa, float fb) {
    (void)dt; double m = -std::numeric_limits<double>::max();
    for (int i = 0; i < n; i++) { double x = (double)(fa * a[i]), y = (double)(fb * b[i]); m = std::max(m, sqrt(x * x + y * y)); }
    return (float)m;
}
template <class T>
static inline float scalar_product_1(const T* a, int n, float dt, float fa) {
    (void)dt; double s = 0.; for (int i = 0; i < n; i++) s += (double)(a[i] * a[i]);
    return fa * fa * (float)s;
}
template <class T>
static inline float l1norm_func_1(const T* a, int n, float dt, float fa) {
    double s = 0.; for (int i = 0; i < n; i++) s += (double)std::abs(a[i]);
    return fa * (float)((double)dt * s);
}
template <class T>
static inline float l2norm_func_1(const T* a, int n, float dt, float fa) {
    double s = 0.; for (int i = 0; i < n; i++) { double d = (double)a[i]; s += d * d; }
    return fa * (float)sqrt((double)dt * s);
}
template <class T>
static inline float maxabs_func_1(const T* a, int n, float dt, float fa) {
    (void)dt; float m = -std::numeric_limits<float>::max(); for (int i = 0; i < n; i++) m = std::max(m, (float)std::abs(a[i]));
    return fa * m;
}
typedef float (*normfn2)(const sreal*, const sreal*, int, float, float, float);
typedef float (*normfn1)(const sreal*, int, float, float);

static long g_warn_empty_region = 0;

// :770-822
static inline float probes_norm_timedomain(Probe& a, Probe& b, normfn2 fn) {
    probes_adjust_spans(a, b);
    int span[2], at[2], bt[2];
    if (a.taper.defined && b.taper.defined) {
        int da[2], db[2];
        discrete_plf_span(a.taper, a.dt, da); span_intersection(da, a.span, at);
        discrete_plf_span(b.taper, b.dt, db); span_intersection(db, b.span, bt);
        if (at[0] > at[1]) { span[0] = bt[0]; span[1] = bt[1]; }
        else if (bt[0] > bt[1]) { span[0] = at[0]; span[1] = at[1]; }
        else span_union(at, bt, span);
    } else {
        probes_adjust_spans(a, b);
        span_union(a.dataspan, b.dataspan, span);
    }
    if (span[0] > span[1]) { g_warn_empty_region++; return 0.f; }
    int n = slen(span);
    if (a.filter.defined && b.filter.defined) {
        update_array_filtered(a); update_array_filtered(b);
        return fn(&a.array_filtered.d[span[0] - a.array_filtered.lo], &b.array_filtered.d[span[0] - b.array_filtered.lo], n, a.dt, a.factor, b.factor);
    } else if (a.taper.defined && b.taper.defined) {
        update_array_tapered(a); update_array_tapered(b);
        return fn(&a.array_tapered.d[span[0] - a.array_tapered.lo], &b.array_tapered.d[span[0] - b.array_tapered.lo], n, a.dt, a.factor, b.factor);
    }
    return fn(&a.array.d[span[0] - a.array.lo], &b.array.d[span[0] - b.array.lo], n, a.dt, a.factor, b.factor);
}
// :824-859
static inline float probe_norm_timedomain(Probe& a, normfn1 fn) {
    int span[2];
    if (a.taper.defined) { int da[2]; discrete_plf_span(a.taper, a.dt, da); span_intersection(da, a.span, span); }
    else { span[0] = a.dataspan[0]; span[1] = a.dataspan[1]; }
    int n = slen(span);
    if (n < 0) n = 0;
    if (a.filter.defined) { update_array_filtered(a); return fn(&a.array_filtered.d[span[0] - a.array_filtered.lo], n, a.dt, a.factor); }
    else if (a.taper.defined) { update_array_tapered(a); return fn(&a.array_tapered.d[span[0] - a.array_tapered.lo], n, a.dt, a.factor); }
    return fn(&a.array.d[span[0] - a.array.lo], n, a.dt, a.factor);
}
// ---- ground-motion diagnostics (comparator.f90:488-517 probes_adjust_spans_3, :519-625 norm functions, :700-765) ----
static inline void probes_adjust_spans_3(Probe& a, Probe& b, Probe& c) {
    int t[2], newspan[2];
    span_union(a.dataspan, b.dataspan, t);
    int u[2]; span_union(t, c.dataspan, u);
    const int minlength = std::max(std::max(f_ceiling((float)slen(a.dataspan) * a.paddingfactor), f_ceiling((float)slen(b.dataspan) * b.paddingfactor)),
                                   f_ceiling((float)slen(c.dataspan) * c.paddingfactor));
    allowed_span(u, minlength, newspan);
    const bool same = a.span[0] == b.span[0] && a.span[1] == b.span[1] && a.span[0] == c.span[0] && a.span[1] == c.span[1] &&
                      slen(a.span) == slen(newspan) && containing(a.span, b.dataspan) && containing(b.span, a.dataspan) &&
                      containing(a.span, c.dataspan) && containing(c.span, a.dataspan) && containing(b.span, c.dataspan) && containing(c.span, b.dataspan);
    if (same) return;
    probe_extend_span(a, newspan); probe_extend_span(b, newspan); probe_extend_span(c, newspan);
}
// which == 1: max_vecnorm_d1_*, 2: max_vecnorm_d2_*, 3: arias_intensity_* over nc = 1..3 arrays of n samples
static inline float ground_motion_func(int which, const sreal* const* arr, const float* fac, int nc, int n, float dt) {
    if (which == 1) {
        double m = -std::numeric_limits<double>::max();
        for (int i = 0; i + 1 < n; i++) {
            double s = 0.;
            for (int c = 0; c < nc; c++) { const double d = (double)(arr[c][i] - arr[c][i + 1]); s += (double)(fac[c] * fac[c]) * (d * d); }
            m = std::max(m, s);
        }
        if (n < 2) return 0.f;
        return (float)(sqrt(m) / (double)dt);
    }
    double acc = which == 2 ? -std::numeric_limits<double>::max() : 0.;
    for (int i = 0; i + 2 < n; i++) {
        double s = 0.;
        for (int c = 0; c < nc; c++) {
            const double d = (double)(arr[c][i] - (sreal)2.0f * arr[c][i + 1] + arr[c][i + 2]);
            s += (double)(fac[c] * fac[c]) * (d * d);
        }
        if (which == 2) acc = std::max(acc, s); else acc += s;
    }
    if (n < 3) return 0.f;
    if (which == 2) return (float)(sqrt(acc) / (double)(dt * dt));
    return (float)((double)(pi / (2.f * 9.81f) * dt) * acc / (double)(dt * dt));
}
// probe_norm_timedomain / probes_norm_timedomain / probes_norm_timedomain_3 with one of the functions above
static inline float probes_ground_motion(Probe** p, int nc, int which) {
    if (nc == 2) probes_adjust_spans(*p[0], *p[1]);
    if (nc == 3) probes_adjust_spans_3(*p[0], *p[1], *p[2]);
    bool tapered = true, filtered = true;
    for (int c = 0; c < nc; c++) { tapered = tapered && p[c]->taper.defined; filtered = filtered && p[c]->filter.defined; }
    int span[2] = {0, -1};
    if (tapered) {
        int ts[3][2];
        for (int c = 0; c < nc; c++) { int d[2]; discrete_plf_span(p[c]->taper, p[c]->dt, d); span_intersection(d, p[c]->span, ts[c]); }
        if (nc == 1) { span[0] = ts[0][0]; span[1] = ts[0][1]; }
        else if (nc == 2) {
            if (ts[0][0] > ts[0][1]) { span[0] = ts[1][0]; span[1] = ts[1][1]; }
            else if (ts[1][0] > ts[1][1]) { span[0] = ts[0][0]; span[1] = ts[0][1]; }
            else span_union(ts[0], ts[1], span);
        } else {
            if (ts[0][0] > ts[0][1]) { span[0] = ts[1][0]; span[1] = ts[1][1]; }
            else if (ts[1][0] > ts[1][1]) { span[0] = ts[0][0]; span[1] = ts[0][1]; }
            else if (ts[2][0] > ts[2][1]) { span[0] = ts[2][0]; span[1] = ts[2][1]; }
            else { int t[2]; span_union(ts[0], ts[1], t); span_union(t, ts[2], span); }
        }
    } else {
        span[0] = p[0]->dataspan[0]; span[1] = p[0]->dataspan[1];
        for (int c = 1; c < nc; c++) { int t[2] = {span[0], span[1]}; span_union(t, p[c]->dataspan, span); }
    }
    if (span[0] > span[1]) { g_warn_empty_region++; return 0.f; }
    const int n = slen(span);
    const sreal* arr[3]; float fac[3];
    for (int c = 0; c < nc; c++) {
        fac[c] = p[c]->factor;
        if (filtered) { update_array_filtered(*p[c]); arr[c] = &p[c]->array_filtered.d[span[0] - p[c]->array_filtered.lo]; }
        else if (tapered) { update_array_tapered(*p[c]); arr[c] = &p[c]->array_tapered.d[span[0] - p[c]->array_tapered.lo]; }
        else arr[c] = &p[c]->array.d[span[0] - p[c]->array.lo];
    }
    return ground_motion_func(which, arr, fac, nc, n, p[0]->dt);
}

// :861-886
typedef float (*normfn2f)(const float*, const float*, int, float, float, float);
typedef float (*normfn1f)(const float*, int, float, float);
static inline float probes_norm_frequencydomain(Probe& a, Probe& b, normfn2f fn) {
    probes_adjust_spans(a, b);
    if (a.filter.defined && b.filter.defined) {
        update_spectrum_filtered(a); update_spectrum_filtered(b);
        return fn(a.amp_spectrum_filtered.data(), b.amp_spectrum_filtered.data(), (int)a.amp_spectrum_filtered.size(), a.df, a.factor, b.factor);
    }
    update_spectrum(a); update_spectrum(b);
    return fn(a.amp_spectrum.data(), b.amp_spectrum.data(), (int)a.amp_spectrum.size(), a.df, a.factor, b.factor);
}
// :888-909
static inline float probe_norm_frequencydomain(Probe& a, normfn1f fn) {
    if (a.filter.defined) { update_spectrum_filtered(a); return fn(a.amp_spectrum_filtered.data(), (int)a.amp_spectrum_filtered.size(), a.df, a.factor); }
    update_spectrum(a);
    return fn(a.amp_spectrum.data(), (int)a.amp_spectrum.size(), a.df, a.factor);
}
// :911-953
static inline float probes_norm(Probe& a, Probe& b, int method = L2NORM) {
    switch (method) {
        case L2NORM: return probes_norm_timedomain(a, b, l2norm_func<sreal>);
        case L1NORM: return probes_norm_timedomain(a, b, l1norm_func<sreal>);
        case SCALAR_PRODUCT: return probes_norm_timedomain(a, b, scalar_product_2<sreal>);
        case AMPSPEC_L2NORM: return probes_norm_frequencydomain(a, b, l2norm_func<float>);
        case AMPSPEC_L1NORM: return probes_norm_frequencydomain(a, b, l1norm_func<float>);
        case PEAK: return probes_norm_timedomain(a, b, maxabs_func<sreal>);
    }
    fprintf(stderr, "probes_norm(): unknown norm method\n"); abort();
}
// :955-996
static inline float probe_norm(Probe& a, int method = L2NORM) {
    switch (method) {
        case L2NORM: return probe_norm_timedomain(a, l2norm_func_1<sreal>);
        case L1NORM: return probe_norm_timedomain(a, l1norm_func_1<sreal>);
        case SCALAR_PRODUCT: return probe_norm_timedomain(a, scalar_product_1<sreal>);
        case AMPSPEC_L2NORM: return probe_norm_frequencydomain(a, l2norm_func_1<float>);
        case AMPSPEC_L1NORM: return probe_norm_frequencydomain(a, l1norm_func_1<float>);
        case PEAK: return probe_norm_timedomain(a, maxabs_func_1<sreal>);
    }
    fprintf(stderr, "probe_norm(): unknown norm method\n"); abort();
}
// :1060-1090
static inline void probes_windowed_cross_corr(Probe& a, Probe& b, const int shiftrange[2], float* cross_corr) {
    int ishift = shiftrange[0];
    for (int i = 0; i < slen(shiftrange); i++) {
        probe_shift(b, ishift);
        ishift = 1;
        cross_corr[i] = probes_norm_timedomain(a, b, scalar_product_2<sreal>);
    }
    probe_shift(b, -shiftrange[1]);
}

}  // namespace ko
