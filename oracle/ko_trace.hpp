// TEST INFRASTRUCTURE ONLY (see ko_base.hpp).  Restates sparse_trace.f90.
#pragma once
#include "ko_base.hpp"

namespace ko {

static const int maxgap = 5;  // sparse_trace.f90:24

// Sample type of strips.  The reference has `real` (fp32) everywhere and that is what the oracle
// uses.  -DKO_WIDE builds a second library whose strip arithmetic (trace sums, probes, norms) runs
// in double on the same fp32 inputs, indices and weights: it measures how far the fp32 reference
// path itself is from the exactly accumulated result (accuracy evidence only, never a parity bar).
#ifdef KO_WIDE
typedef double sreal;
#else
typedef float sreal;
#endif

// util.f90:339-357 `resize`: no preservation, no initialisation, length 0 deallocates
struct Strip {  // sparse_trace.f90:28-32
    int lo = 0;
    std::vector<sreal> d;
    bool alloc = false;
    int hi() const { return lo + (int)d.size() - 1; }
    int size() const { return (int)d.size(); }
    sreal& at(int i) { return d[i - lo]; }
    sreal at(int i) const { return d[i - lo]; }
};
static inline void resize(Strip& s, int offset, int length) {
    if (!s.alloc) {
        if (length == 0) return;
        s.lo = offset; s.d.assign(length, 0.f); s.alloc = true;  // contents unspecified in Fortran
        return;
    }
    if (s.size() != length || s.lo != offset) {
        s.d.clear(); s.alloc = false;
        if (length != 0) { s.lo = offset; s.d.assign(length, 0.f); s.alloc = true; }
    }
}

struct Trace {  // sparse_trace.f90:34-50
    int nstrips = 0;
    int span[2] = {0, 0};
    std::vector<Strip> strips;
    bool alloc = false;  // allocated(trace%strips)
};

template <class T>
static inline void strip_init(int s0, int s1, const T* data, int ndata, Strip& strip) {  // :72-88
    int length = s1 - s0 + 1;
    if (ndata != length) { fprintf(stderr, "strip_init(): length of data does not match span\n"); abort(); }
    resize(strip, s0, length);
    for (int i = 0; i < length; i++) strip.d[i] = data[i];
}
static inline int strip_length(const Strip& s) { return s.alloc ? s.size() : 0; }  // :90-99
static inline void strip_destroy(Strip& s) { resize(s, 1, 0); }                    // :101-104
static inline void strip_nullify(Strip& s) { if (s.alloc) std::fill(s.d.begin(), s.d.end(), 0.f); }  // :106-111
static inline void strip_copy(const Strip& src, Strip& dst) {  // :201-211
    resize(dst, src.lo, src.size());
    if (src.alloc) dst.d = src.d;
}
static inline void trace_destroy(Trace& t) {  // :896-913
    t.strips.clear(); t.alloc = false; t.nstrips = 0; t.span[0] = 0; t.span[1] = 0;
}
static inline bool trace_is_empty(const Trace& t) { return !t.alloc; }  // :915-921

// :316-345
static inline void strip_extend(Strip& s, int n0, int n1) {
    std::vector<sreal> temp; int r0 = 0, r1 = -1; bool had = false;
    if (s.alloc) { r0 = s.lo; r1 = s.hi(); temp = s.d; had = true; }
    resize(s, n0, n1 - n0 + 1);
    if (had) {
        if (n0 < r0) for (int i = n0; i <= r0 - 1; i++) s.at(i) = 0.f;
        if (n1 > r1) for (int i = r1 + 1; i <= n1; i++) s.at(i) = temp[r1 - r0];
        for (int i = r0; i <= r1; i++) s.at(i) = temp[i - r0];
    } else {
        std::fill(s.d.begin(), s.d.end(), 0.f);
    }
}
// :213-314 strip_extend_to_same_span_{5,4,2}
static inline void strip_extend_to_same_span(std::initializer_list<Strip*> ss) {
    int s0 = std::numeric_limits<int>::max(), s1 = -std::numeric_limits<int>::max();
    for (Strip* p : ss) if (p->alloc) { s0 = std::min(p->lo, s0); s1 = std::max(p->hi(), s1); }
    if (s0 < s1) for (Strip* p : ss) strip_extend(*p, s0, s1);
}
// :347-377
static inline void strip_dataspan(const Strip& s, int out[2]) {
    if (strip_length(s) == 0) { out[0] = 0; out[1] = -1; return; }
    int s0 = s.lo, s1 = s.hi();
    out[0] = s0; out[1] = s1;
    sreal firstvalue = 0.f;
    for (int i = s0; i <= s1; i++) { out[0] = i; if (s.at(i) != firstvalue) break; }
    sreal lastvalue = s.at(s1);
    for (int i = s1; i >= s0; i--) { if (s.at(i) != lastvalue) break; out[1] = i; }
}
// :404-418
template <class T>
static inline void trace_create_simple(Trace& t, const T* data, int s0, int s1) {
    trace_destroy(t);
    t.strips.resize(1); t.alloc = true; t.nstrips = 1;
    resize(t.strips[0], s0, s1 - s0 + 1);
    for (int i = 0; i <= s1 - s0; i++) t.strips[0].d[i] = data[i];
    t.span[0] = s0; t.span[1] = s1;
}
// :420-432
static inline void trace_create_simple_nodata(Trace& t, int s0, int s1) {
    trace_destroy(t);
    t.strips.resize(1); t.alloc = true; t.nstrips = 1;
    resize(t.strips[0], s0, s1 - s0 + 1);
    t.span[0] = s0; t.span[1] = s1;
}
static inline bool span_contains(const int span[2], const int sub[2]) {  // :434-441
    return sub[0] <= sub[1] && span[0] <= sub[0] && sub[1] <= span[1];
}

// :443-555
static inline void trace_pack(const Strip& strip, Trace& trace, const int* last_span_as_hint = nullptr) {
    int gap = 0; bool interest = false; int istrip = 0;
    for (int i = strip.lo; i <= strip.hi(); i++) {
        if (strip.at(i) != 0.f) {
            if (!interest) { interest = true; istrip++; }
            gap = 0;
        } else if (interest) {
            gap++;
            if (gap > maxgap) interest = false;
        }
    }
    int nstrips = istrip;
    trace_destroy(trace);
    int ibeg = 0, iend = 0;
    if (nstrips == 0) {
        trace.strips.resize(1); trace.alloc = true;
        int span[2] = {strip.lo, strip.hi()};
        ibeg = span[0]; iend = ibeg;
        if (last_span_as_hint && span_contains(span, last_span_as_hint)) { ibeg = last_span_as_hint[0]; iend = ibeg; }
        resize(trace.strips[0], ibeg, iend - ibeg + 1);
        std::fill(trace.strips[0].d.begin(), trace.strips[0].d.end(), 0.f);
        trace.nstrips = 1; trace.span[0] = ibeg; trace.span[1] = iend;
        return;
    }
    trace.strips.resize(nstrips); trace.alloc = true; trace.nstrips = nstrips;
    gap = 0; interest = false; istrip = 0;
    auto emit = [&](int is, int b, int e) {  // copy strip%data(b:e) into strips(is)
        resize(trace.strips[is - 1], b, e - b + 1);
        for (int i = b; i <= e; i++) trace.strips[is - 1].at(i) = strip.at(i);
    };
    for (int i = strip.lo; i <= strip.hi(); i++) {
        if (strip.at(i) != 0.f) {
            if (!interest) { interest = true; ibeg = i; istrip++; }
            gap = 0; iend = i;
        } else if (interest) {
            gap++;
            if (gap > maxgap) { emit(istrip, ibeg, iend + 1); interest = false; }  // add one of the zeros
        }
    }
    if (interest) {
        if (gap > 0) emit(istrip, ibeg, iend + 1); else emit(istrip, ibeg, iend);
    }
    trace.span[0] = trace.strips[0].lo;
    trace.span[1] = trace.strips[nstrips - 1].hi();
}
// :557-579
static inline void trace_unpack(const Trace& trace, Strip& strip) {
    int length = trace.span[1] - trace.span[0] + 1;
    resize(strip, trace.span[0], length);
    std::fill(strip.d.begin(), strip.d.end(), 0.f);
    for (int is = 0; is < trace.nstrips; is++)
        for (int i = trace.strips[is].lo; i <= trace.strips[is].hi(); i++) strip.at(i) = trace.strips[is].at(i);
}
// :156-171
static inline void trace_copy(const Trace& src, Trace& dst) {
    trace_destroy(dst);
    if (src.alloc) { dst.strips = src.strips; dst.alloc = true; dst.span[0] = src.span[0]; dst.span[1] = src.span[1]; dst.nstrips = src.nstrips; }
}
// :122-154
static inline void trace_join(const Trace& a, const Trace& b, Trace& c) {
    trace_destroy(c);
    if (!a.alloc) { trace_copy(b, c); return; }
    if (!b.alloc) { trace_copy(a, c); return; }
    if (a.span[1] >= b.span[0]) { fprintf(stderr, "trace_join(): span overlap detected\n"); abort(); }
    c.nstrips = a.nstrips + b.nstrips; c.span[0] = a.span[0]; c.span[1] = b.span[1];
    c.strips = a.strips; c.strips.insert(c.strips.end(), b.strips.begin(), b.strips.end()); c.alloc = true;
}

enum ShiftKind { SHIFT_NONE = 0, SHIFT_INT = 1, SHIFT_REAL = 2 };

// :597-707  strip(x) += factor * trace(x - shift); output grows; last sample repeats to the right
static inline void trace_multiply_add(const Trace& trace, Strip& strip, float factor = 1.f, ShiftKind kind = SHIFT_NONE,
                                      int itraceshift_ = 0, float rtraceshift_ = 0.f) {
    int itraceshift = 0;
    float weight_right = 0.f, weight_left = 0.f;
    bool rpresent = (kind == SHIFT_REAL);
    if (kind == SHIFT_INT) itraceshift = itraceshift_;
    if (rpresent) {
        itraceshift = f_floor(rtraceshift_);
        weight_right = rtraceshift_ - (float)itraceshift;
        weight_left = 1.f - weight_right;
        weight_right = weight_right * factor;
        weight_left = weight_left * factor;
    }
    int span[2] = {trace.span[0] + itraceshift, trace.span[1] + itraceshift};
    int need[2] = {span[0], span[1]};
    if (rpresent) need[1] = need[1] + 1;
    if (strip.alloc) {
        int c0 = std::min(need[0], strip.lo), c1 = std::max(need[1], strip.hi());
        if (c0 != strip.lo || c1 != strip.hi()) strip_extend(strip, c0, c1);
    } else {
        resize(strip, need[0], need[1] - need[0] + 1);
        std::fill(strip.d.begin(), strip.d.end(), 0.f);
    }
    for (int is = 1; is <= trace.nstrips; is++) {
        const Strip& ts = trace.strips[is - 1];
        int ss0 = ts.lo + itraceshift, ss1 = ts.hi() + itraceshift;
        if (ss1 < span[0]) continue;
        if (ss0 > span[1]) break;
        int r0 = std::max(ss0, span[0]), r1 = std::min(ss1, span[1]);
        if (!rpresent) {
            for (int x = r0; x <= r1; x++) strip.at(x) = strip.at(x) + factor * ts.at(x - itraceshift);
        } else {
            for (int x = r0; x <= r1; x++) strip.at(x) = strip.at(x) + weight_left * ts.at(x - itraceshift);
            if (is == trace.nstrips) {  // last point covered by repeat e.p. below
                for (int x = r0 + 1; x <= r1; x++) strip.at(x) = strip.at(x) + weight_right * ts.at(x - 1 - itraceshift);
            } else {
                for (int x = r0 + 1; x <= r1 + 1; x++) strip.at(x) = strip.at(x) + weight_right * ts.at(x - 1 - itraceshift);
            }
        }
        if (is == trace.nstrips && r1 + 1 <= strip.hi()) {
            sreal lastval = ts.at(ts.hi());
            if (lastval != 0.f) for (int x = r1 + 1; x <= strip.hi(); x++) strip.at(x) = strip.at(x) + factor * lastval;
        }
    }
}

// :710-792 fixed window variant; array covers [a0, a1]
static inline void trace_multiply_add_nogrow(const Trace& trace, sreal* array, int a0, int a1, float factor = 1.f,
                                             ShiftKind kind = SHIFT_NONE, int itraceshift_ = 0, float rtraceshift_ = 0.f) {
    auto A = [&](int x) -> sreal& { return array[x - a0]; };
    int itraceshift = 0;
    float weight_right = 0.f, weight_left = 0.f;
    bool rpresent = (kind == SHIFT_REAL);
    if (kind == SHIFT_INT) itraceshift = itraceshift_;
    if (rpresent) {
        itraceshift = f_floor(rtraceshift_);
        weight_right = rtraceshift_ - (float)itraceshift;
        weight_left = 1.f - weight_right;
        weight_right = weight_right * factor;
        weight_left = weight_left * factor;
    }
    int sh[2] = {trace.span[0] + itraceshift, trace.span[1] + itraceshift};
    int span[2] = {std::max(a0, sh[0]), std::min(a1, sh[1])};
    if (span[1] < span[0]) return;
    for (int is = 1; is <= trace.nstrips; is++) {
        const Strip& ts = trace.strips[is - 1];
        int ss0 = ts.lo + itraceshift, ss1 = ts.hi() + itraceshift;
        if (ss1 < span[0]) continue;
        if (ss0 > span[1]) break;
        int r0 = std::max(ss0, span[0]), r1 = std::min(ss1, span[1]);
        if (!rpresent) {
            for (int x = r0; x <= r1; x++) A(x) = A(x) + factor * ts.at(x - itraceshift);
        } else {
            for (int x = r0; x <= r1; x++) A(x) = A(x) + weight_left * ts.at(x - itraceshift);
            if (is == trace.nstrips || r1 + 1 > a1) {
                for (int x = r0 + 1; x <= r1; x++) A(x) = A(x) + weight_right * ts.at(x - 1 - itraceshift);
            } else {
                for (int x = r0 + 1; x <= r1 + 1; x++) A(x) = A(x) + weight_right * ts.at(x - 1 - itraceshift);
            }
        }
        if (is == trace.nstrips && r1 + 1 <= a1) {
            sreal lastval = ts.at(ts.hi());
            if (lastval != 0.f) for (int x = r1 + 1; x <= a1; x++) A(x) = A(x) + factor * lastval;
        }
    }
}

// :379-402
static inline void strip_fold(Strip& s, const std::vector<float>& shifts, const std::vector<float>& amplitudes) {
    int ds[2];
    strip_dataspan(s, ds);
    if (ds[1] < ds[0]) return;
    Trace t;
    trace_create_simple(t, &s.d[ds[0] - s.lo], ds[0], ds[1]);
    std::fill(s.d.begin(), s.d.end(), 0.f);
    for (size_t i = 0; i < shifts.size(); i++) trace_multiply_add(t, s, amplitudes[i], SHIFT_REAL, 0, shifts[i]);
    trace_destroy(t);
}

}  // namespace ko
