// TEST INFRASTRUCTURE ONLY -- CPU parity oracle for kiwi_b200.
//
// This directory is a line-by-line CPU restatement of the emolch/kiwi Fortran hot path
// (reference mounted at /root/reference; citations are file:line in that tree).  It is the
// checker for the CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may build, link, import or execute anything in oracle/.  The product
// (kiwi_b200/) never does.
//
// Parity status: the reference cannot be compiled in this image (no Fortran compiler, no HDF5,
// no FFTW), so the oracle is pinned against the reference's own known-answer tests
// (test_sparse_trace.f90, test_comparator.f90, test_piecewise_linear_function.f90,
// test_source_bilat.f90, test_orthodrome.f90, test_euler.f90 -> oracle/kat_main.cpp).
// make_seismogram, the bilinear GF fetch, component rotations, scaling and the global misfit
// formula have no reference test: for those "parity unpinned by reference tests"; they rest on
// this restatement following the cited lines.
//
// Arithmetic rules: `real` -> float, `real*8`/`double precision` -> double, exactly where the
// Fortran has them; build with -ffp-contract=off (gfortran on baseline x86-64 emits no FMA).
//
// ko_base.hpp: constants.f90, orthodrome.f90, euler.f90, piecewise_linear_function.f90
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <limits>

namespace ko {

// ---- Fortran intrinsics -------------------------------------------------------------------
static inline int f_nint(float x) { return (int)lroundf(x); }    // NINT: half away from zero
static inline int f_nint(double x) { return (int)lround(x); }
static inline int f_floor(float x) { return (int)floorf(x); }
static inline int f_floor(double x) { return (int)floor(x); }
static inline int f_ceiling(float x) { return (int)ceilf(x); }

// ---- constants.f90:21-25 ------------------------------------------------------------------
// All initialisers are default-real expressions, also for the real*8 parameters.
static const float pi = 3.14159265358979f;
static const double pi_ = (double)3.14159265358979f;
static const float earthradius = 6371.f * 1000.f;
static const float earthradius_equator = 6378.14f * 1000.f;
static const double earth_oblateness = (double)(1.f / 298.257223563f);

// ---- orthodrome.f90 -----------------------------------------------------------------------
struct GeoCoords { double lat = 0., lon = 0.; };  // orthodrome.f90:32-34

// orthodrome.f90:296-350 -- `2./360.*pi` is a default-real constant expression
static inline float d2r_r(float deg) { return 2.f / 360.f * pi * deg; }
static inline float r2d_r(float rad) { return 360.f / 2.f / pi * rad; }
static inline double d2r_d(double deg) { return (double)(2.f / 360.f * pi) * deg; }
static inline double r2d_d(double rad) { return (double)(360.f / 2.f / pi) * rad; }
static inline GeoCoords d2r_tgc(const GeoCoords& deg) {
    GeoCoords r; r.lat = (double)(2.f / 360.f * pi) * deg.lat; r.lon = (double)(2.f / 360.f * pi) * deg.lon; return r;
}
static inline GeoCoords r2d_tgc(const GeoCoords& rad) {
    GeoCoords r; r.lat = (double)(360.f / 2.f / pi) * rad.lat; r.lon = (double)(360.f / 2.f / pi) * rad.lon; return r;
}

// orthodrome.f90:158-170
static inline double clip(double x, double mi, double ma) { return std::min(std::max(mi, x), ma); }
static inline double wrap(double x, double mi, double ma) { return x - floor((x - mi) / (ma - mi)) * (ma - mi); }

// orthodrome.f90:285-294
static inline double cosdelta(const GeoCoords& a, const GeoCoords& b) {
    return sin(a.lat) * sin(b.lat) + cos(a.lat) * cos(b.lat) * cos(b.lon - a.lon);
}
// orthodrome.f90:172-181
static inline double arcdistance(const GeoCoords& a, const GeoCoords& b) { return acos(cosdelta(a, b)); }
// orthodrome.f90:231-243
static inline double azimuth(const GeoCoords& a, const GeoCoords& b) {
    return atan2(cos(a.lat) * cos(b.lat) * sin(b.lon - a.lon), sin(b.lat) - sin(a.lat) * cosdelta(a, b));
}
// orthodrome.f90:245-265
static inline void azibazi(const GeoCoords& a, const GeoCoords& b, double& azi, double& bazi) {
    double t = cos(a.lat) * cos(b.lat) * sin(b.lon - a.lon);
    double sb = sin(b.lat);
    double sa = sin(a.lat);
    double cd = cosdelta(a, b);
    azi = atan2(t, sb - sa * cd);
    bazi = atan2(-t, sa - sb * cd);
}
// orthodrome.f90:193-229
static inline double distance_accurate50m(const GeoCoords& a, const GeoCoords& b) {
    double f = (a.lat + b.lat) / 2.;
    double g = (a.lat - b.lat) / 2.;
    double l = (a.lon - b.lon) / 2.;
    double sg = sin(g), cl = cos(l), cf = cos(f), sl = sin(l), cg = cos(g), sf = sin(f);
    double s = sg * sg * (cl * cl) + cf * cf * (sl * sl);
    double c = cg * cg * (cl * cl) + sf * sf * (sl * sl);
    double w = atan(sqrt(s / c));
    double r = sqrt(s * c) / w;
    double d = 2. * w * (double)earthradius_equator;
    double h1 = (3. * r - 1.) / (2. * c);
    double h2 = (3. * r + 1.) / (2. * s);
    return d * (1. + earth_oblateness * h1 * (sf * sf) * (cg * cg) - earth_oblateness * h2 * (cf * cf) * (sg * sg));
}

// orthodrome.f90:67,72 -- both approximations are switched off by constants
static const double max_distance_flat_approx = -1.;
static const double min_ratio_const_azimuth_approx = std::numeric_limits<double>::max();

// orthodrome.f90:77-156
static inline void approx_differential_azidist(float delta_x, float delta_y, double azimuth_, double backazimuth,
                                               double dist, double& new_azimuth, double& new_backazimuth,
                                               double& new_dist) {
    if (dist < max_distance_flat_approx) {
        double ndx = dist * cos(azimuth_) - (double)delta_x;
        double ndy = dist * sin(azimuth_) - (double)delta_y;
        new_azimuth = atan2(ndy, ndx);
        new_backazimuth = backazimuth + (new_azimuth - azimuth_);
        new_dist = sqrt(ndx * ndx + ndy * ndy);
    } else {
        // r = sqrt(delta_x**2 + delta_y**2): a default-real expression assigned to real*8
        double r = (double)sqrtf(delta_x * delta_x + delta_y * delta_y);
        if (dist / r > min_ratio_const_azimuth_approx) {  // r == 0 -> +Inf > huge
            new_azimuth = azimuth_;
            new_backazimuth = backazimuth;
            new_dist = dist - ((double)delta_x * cos(azimuth_) + (double)delta_y * sin(azimuth_));
        } else {
            double a = r / (double)earthradius;
            double b = dist / (double)earthradius;
            double lambda = (double)atan2f(delta_y, delta_x);  // default-real atan2
            double gamma = azimuth_ - lambda;
            double c = acos(clip(cos(a) * cos(b) + sin(a) * sin(b) * cos(gamma), -1., 1.));
            double alpha = asin(clip(sin(a) * sin(gamma) / sin(c), -1., 1.));
            double beta = asin(clip(sin(b) * sin(gamma) / sin(c), -1., 1.));
            if (cos(a) - cos(b) * cos(c) < 0) {
                if (alpha > 0) alpha = pi_ - alpha; else alpha = -pi_ - alpha;
            }
            if (cos(b) - cos(a) * cos(c) < 0) {
                if (beta > 0) beta = pi_ - beta; else beta = -pi_ - beta;
            }
            new_dist = c * (double)earthradius;
            new_backazimuth = wrap(backazimuth + alpha, -pi_, pi_);
            new_azimuth = wrap(lambda - pi_ - beta, -pi_, pi_);
        }
    }
}

// ---- euler.f90:28-67 ----------------------------------------------------------------------
// mat is indexed mat[row][col] == Fortran mat(row+1, col+1)
static inline void init_euler(float alpha, float beta, float gamma, float mat[3][3]) {
    float ca = cosf(alpha), cb = cosf(beta), cg = cosf(gamma);
    float sa = sinf(alpha), sb = sinf(beta), sg = sinf(gamma);
    mat[0][0] = cb * cg - ca * sb * sg;
    mat[1][0] = sb * cg + ca * cb * sg;
    mat[2][0] = sa * sg;
    mat[0][1] = -cb * sg - ca * sb * cg;
    mat[1][1] = -sb * sg + ca * cb * cg;
    mat[2][1] = sa * cg;
    mat[0][2] = sa * sb;
    mat[1][2] = -sa * cb;
    mat[2][2] = ca;
}
// Fortran matmul(A, v) for 3x3 * 3: sum over the contracted index in ascending order
static inline void matvec3(const float a[3][3], const float v[3], float out[3]) {
    for (int i = 0; i < 3; i++) {
        float s = 0.f;
        for (int j = 0; j < 3; j++) s = s + a[i][j] * v[j];
        out[i] = s;
    }
}
static inline void matmul3(const float a[3][3], const float b[3][3], float out[3][3]) {
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) {
            float s = 0.f;
            for (int j = 0; j < 3; j++) s = s + a[i][j] * b[j][k];
            out[i][k] = s;
        }
}

// ---- piecewise_linear_function.f90 -----------------------------------------------------------
struct Plf {  // :27-35
    std::vector<float> x, y;
    bool defined = false;  // allocated(s%f)
    int n() const { return (int)x.size(); }
};
static inline void plf_make(Plf& s, const std::vector<float>& x, const std::vector<float>& y) {  // :81-95
    size_t n = std::min(x.size(), y.size());
    s.x.assign(x.begin(), x.begin() + n);
    s.y.assign(y.begin(), y.begin() + n);
    s.defined = true;
}
static inline void plf_destroy(Plf& s) { s.x.clear(); s.y.clear(); s.defined = false; }
static inline float trapezoid_centroid(float x0, float y0, float x1, float y1) {  // :285-294
    if (y0 + y1 == 0.f) return (x0 + x1) / 2.f;
    return (x0 * (2.f * y0 + y1) + x1 * (y0 + 2.f * y1)) / (3.f * (y0 + y1));
}
static inline float trapezoid_area(float x0, float y0, float x1, float y1) { return (y0 + y1) * (x1 - x0) / 2.f; }  // :296-300
static inline float ip_linear(float x0, float y0, float x1, float y1, float xi) {  // :302-306
    return y0 + (y1 - y0) / (x1 - x0) * (xi - x0);
}
static inline float ip_cos(float x0, float y0, float x1, float y1, float xi) {  // :308-316
    if (y1 != y0) return y0 + (y1 - y0) * (0.5f - 0.5f * cosf((xi - x0) / (x1 - x0) * pi));
    return y0;
}
static inline float ip_zero_one(float x0, float y0, float x1, float y1, float xi) {  // :318-327
    if (y0 == 0.f && y1 == 0.f) return 0.f + 0.f * (x0 + x1 + xi);
    return 1.f;
}
// :137-161
static inline float plf_integrate(const Plf& s, float a, float b) {
    float area = 0.f;
    if (!s.defined) return area;
    int n = s.n();
    if (b <= s.x[0]) return area;
    if (a >= s.x[n - 1]) return area;
    for (int i = 0; i < n - 1; i++) {
        if (a >= s.x[i + 1]) continue;
        if (b <= s.x[i]) return area;
        float x0 = std::max(a, s.x[i]);
        float x1 = std::min(b, s.x[i + 1]);
        float y0 = s.y[i];
        if (x0 != s.x[i]) y0 = ip_linear(s.x[i], s.y[i], s.x[i + 1], s.y[i + 1], a);
        float y1 = s.y[i + 1];
        if (x1 != s.x[i + 1]) y1 = ip_linear(s.x[i], s.y[i], s.x[i + 1], s.y[i + 1], b);
        area = area + trapezoid_area(x0, y0, x1, y1);
    }
    return area;
}
// :163-193
static inline void plf_integrate_and_centroid(const Plf& s, float a, float b, float& area, float& centroid) {
    area = 0.f;
    centroid = (a + b) / 2.f;
    float c = 0.f;
    if (!s.defined) return;
    int n = s.n();
    if (b <= s.x[0]) return;
    if (a >= s.x[n - 1]) return;
    for (int i = 0; i < n - 1; i++) {
        if (a >= s.x[i + 1]) continue;
        if (b <= s.x[i]) break;
        float x0 = std::max(a, s.x[i]);
        float x1 = std::min(b, s.x[i + 1]);
        float y0 = s.y[i];
        if (x0 != s.x[i]) y0 = ip_linear(s.x[i], s.y[i], s.x[i + 1], s.y[i + 1], a);
        float y1 = s.y[i + 1];
        if (x1 != s.x[i + 1]) y1 = ip_linear(s.x[i], s.y[i], s.x[i + 1], s.y[i + 1], b);
        float areathis = trapezoid_area(x0, y0, x1, y1);
        c = c + areathis * trapezoid_centroid(x0, y0, x1, y1);
        area = area + areathis;
    }
    centroid = c / area;
}

// :195-237 (real) and :239-282 (complex): array(j) for j in [span0, span1]; `at(j)` returns a
// reference-like accessor.  MulFn(j, factor) multiplies element j, ZeroFn(j) zeroes it.
template <class MulFn, class ZeroFn, class IpFn>
static inline void plf_taper_generic(const Plf& s, int span0, int span1, float dx, IpFn ip, MulFn mul, ZeroFn zero) {
    int n = s.n();
    int ibeg = f_floor(s.x[0] / dx);
    if (span0 <= ibeg) {
        for (int j = span0; j <= std::min(ibeg, span1); j++) zero(j);
    }
    int ibegatleast = span0;
    for (int i = 0; i < n - 1; i++) {
        ibeg = std::max(std::max(f_floor(s.x[i] / dx) + 1, span0), ibegatleast);
        int iend = std::min(f_floor(s.x[i + 1] / dx), span1);
        if (ibeg <= iend) {
            for (int j = ibeg; j <= iend; j++) mul(j, ip(s.x[i], s.y[i], s.x[i + 1], s.y[i + 1], (float)j * dx));
        }
        ibegatleast = iend + 1;
    }
    int iend = f_floor(s.x[n - 1] / dx) + 1;
    if (span1 >= iend) {
        for (int j = std::max(iend, span0); j <= span1; j++) zero(j);
    }
}

}  // namespace ko
