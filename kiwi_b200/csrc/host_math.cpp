// Host-side arithmetic of the engine: everything that is evaluated once per receiver or once
// per candidate and either needs libm (so it is done with the host's glibc, like the reference)
// or is tiny and sequential.  fp32/fp64 exactly where the Fortran has real/real*8; built with
// -ffp-contract=off.  Citations are file:line of /root/reference.
#include "host_math.hpp"
#include <cmath>
#include <algorithm>

namespace kh {

// constants.f90:21-25: default-real initialisers also for the real*8 parameters
static const float pi = 3.14159265358979f;
static const float earthradius_equator = 6378.14f * 1000.f;
static const double earth_oblateness = (double)(1.f / 298.257223563f);

float d2r_r(float deg) { return 2.f / 360.f * pi * deg; }                 // orthodrome.f90:316-323
double d2r_d(double deg) { return (double)(2.f / 360.f * pi) * deg; }     // orthodrome.f90:296-305, 334-341

static double cosdelta(double alat, double alon, double blat, double blon) {  // orthodrome.f90:285-294
    return sin(alat) * sin(blat) + cos(alat) * cos(blat) * cos(blon - alon);
}
void azibazi(double alat, double alon, double blat, double blon, double* azi, double* bazi) {  // orthodrome.f90:245-265
    double t = cos(alat) * cos(blat) * sin(blon - alon);
    double sb = sin(blat), sa = sin(alat);
    double cd = cosdelta(alat, alon, blat, blon);
    *azi = atan2(t, sb - sa * cd);
    *bazi = atan2(-t, sa - sb * cd);
}
double distance_accurate50m(double alat, double alon, double blat, double blon) {  // orthodrome.f90:193-229
    double f = (alat + blat) / 2., g = (alat - blat) / 2., l = (alon - blon) / 2.;
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
void final_rotation(double bazi0, float* cl0, float* sl0) {  // seismogram.f90:270-271: real(cos(bazi_orig+pi))
    *cl0 = (float)cos(bazi0 + (double)pi);
    *sl0 = (float)sin(bazi0 + (double)pi);
}

// euler.f90:28-67; row-major mat[i*3+j] = mat(i+1,j+1)
void init_euler(float alpha, float beta, float gamma, float* mat) {
    float ca = cosf(alpha), cb = cosf(beta), cg = cosf(gamma), sa = sinf(alpha), sb = sinf(beta), sg = sinf(gamma);
    mat[0 * 3 + 0] = cb * cg - ca * sb * sg;
    mat[1 * 3 + 0] = sb * cg + ca * cb * sg;
    mat[2 * 3 + 0] = sa * sg;
    mat[0 * 3 + 1] = -cb * sg - ca * sb * cg;
    mat[1 * 3 + 1] = -sb * sg + ca * cb * cg;
    mat[2 * 3 + 1] = sa * cg;
    mat[0 * 3 + 2] = sa * sb;
    mat[1 * 3 + 2] = -sa * cb;
    mat[2 * 3 + 2] = ca;
}

// ---- piecewise_linear_function.f90 ------------------------------------------------------------
static float trapezoid_centroid(float x0, float y0, float x1, float y1) {  // :285-294
    if (y0 + y1 == 0.f) return (x0 + x1) / 2.f;
    return (x0 * (2.f * y0 + y1) + x1 * (y0 + 2.f * y1)) / (3.f * (y0 + y1));
}
static float trapezoid_area(float x0, float y0, float x1, float y1) { return (y0 + y1) * (x1 - x0) / 2.f; }
static float ip_linear(float x0, float y0, float x1, float y1, float xi) { return y0 + (y1 - y0) / (x1 - x0) * (xi - x0); }
static float ip_cos(float x0, float y0, float x1, float y1, float xi) {  // :308-316
    if (y1 != y0) return y0 + (y1 - y0) * (0.5f - 0.5f * cosf((xi - x0) / (x1 - x0) * pi));
    return y0;
}
// :163-193
void plf_integrate_and_centroid(const float* px, const float* py, int n, float a, float b, float* area_, float* centroid_) {
    float area = 0.f, centroid = (a + b) / 2.f, c = 0.f;
    *area_ = area; *centroid_ = centroid;
    if (n <= 0) return;
    if (b <= px[0]) return;
    if (a >= px[n - 1]) return;
    for (int i = 0; i < n - 1; i++) {
        if (a >= px[i + 1]) continue;
        if (b <= px[i]) break;
        float x0 = std::max(a, px[i]), x1 = std::min(b, px[i + 1]);
        float y0 = py[i];
        if (x0 != px[i]) y0 = ip_linear(px[i], py[i], px[i + 1], py[i + 1], a);
        float y1 = py[i + 1];
        if (x1 != px[i + 1]) y1 = ip_linear(px[i], py[i], px[i + 1], py[i + 1], b);
        float areathis = trapezoid_area(x0, y0, x1, y1);
        c = c + areathis * trapezoid_centroid(x0, y0, x1, y1);
        area = area + areathis;
    }
    *area_ = area; *centroid_ = c / area;
}

// Pointwise form of plf_taper_array_r with ip_cos (:195-237): multiplier of sample j for
// j in [tp0, tp1] = [floor(x1/dt)+1, floor(xn/dt)]; every sample outside that range is zeroed.
void taper_table(const std::vector<float>& x, const std::vector<float>& y, float dt, int* tp0, int* tp1, std::vector<float>* tab) {
    int n = (int)x.size();
    *tp0 = (int)floorf(x[0] / dt) + 1;
    *tp1 = (int)floorf(x[n - 1] / dt);
    tab->clear();
    if (*tp1 < *tp0) return;
    tab->assign((size_t)(*tp1 - *tp0 + 1), 0.f);
    int ibegatleast = *tp0;
    for (int i = 0; i < n - 1; i++) {
        int ibeg = std::max((int)floorf(x[i] / dt) + 1, ibegatleast);
        int iend = (int)floorf(x[i + 1] / dt);
        for (int j = ibeg; j <= iend; j++) (*tab)[j - *tp0] = ip_cos(x[i], y[i], x[i + 1], y[i + 1], (float)j * dt);
        ibegatleast = iend + 1;
    }
}
void discrete_plf_span(const std::vector<float>& x, float dt, int* s0, int* s1) {  // comparator.f90:1145-1157
    *s0 = (int)ceilf(x.front() / dt);
    *s1 = (int)floorf(x.back() / dt);
}

// comparator.f90:1092-1118
static int next_power_of_two(int n) { return 1 << (int)ceilf(logf((float)n) / logf(2.f)); }
void allowed_span(int s0, int s1, int minlength, int* n0, int* n1) {
    int slen = s1 - s0 + 1;
    int length = std::max(slen, minlength);
    int lengthp = next_power_of_two(length);
    *n0 = s0 - (int)floorf((float)(lengthp - slen) / 2.f);
    *n1 = *n0 + lengthp - 1;
}
// probe_set_array on a freshly initialised probe (comparator.f90:222-271), paddingfactor 2
void initial_probe_span(int ds0, int ds1, int* s0, int* s1) {
    int datalength = ds1 - ds0 + 1;
    allowed_span(ds0, ds1, (int)ceilf((float)datalength * 2.f), s0, s1);
}

// ---- source_bilat.f90 ------------------------------------------------------------------------------
bool prep_bilateral(const float* p, float shortest_doi, SourcePrep* out) {
    SourcePrep& o = *out;
    o = SourcePrep();
    // psm_update_dep_params_bilat :216-239
    float strike = d2r_r(p[5]), dip = d2r_r(p[6]), rake = d2r_r(p[7]), rupdir = d2r_r(p[8]);
    float rot_slip[9];
    init_euler(dip, strike, -rupdir, o.rot_rup);
    init_euler(dip, strike, -rake, rot_slip);
    o.moment = p[4]; o.risetime = 0.f;   // :207-209, parameterized_source.f90:121-125
    // psm_to_tdsm_bilat :241-271
    float rupvel = p[12];
    float maxdt = shortest_doi, maxdx = 0.5f * shortest_doi * rupvel, maxdy = shortest_doi * rupvel;
    // psm_to_tdsm_size_bilat :274-315
    float length_a = p[9], length_b = p[10], width = p[11], risetime = p[13];
    float length = length_a + length_b;
    float fx = length / maxdx, fy = width / maxdy;
    if (!(fabsf(fx) < 1e6f) || !(fabsf(fy) < 1e6f) || !(rupvel > 0.f) || !std::isfinite(risetime)) return false;
    int nx = (int)floorf(fx) + 1;
    if (nx <= 1) nx = 2;
    if (length == 0.f) nx = 1;
    int ny = (int)floorf(fy) + 1;
    if (ny <= 1) ny = 2;
    if (width == 0.f) ny = 1;
    float dursf = length / (float)nx / rupvel;
    float durfull = risetime + dursf;
    float ft = durfull / maxdt;
    if (!(fabsf(ft) < 1e6f)) return false;
    int nt = (int)floorf(ft) + 1;
    if (nt <= 1) nt = 2;
    o.nx = nx; o.ny = ny; o.nt = nt; o.ngroups = nx * ny;
    // STF :379-411
    dursf = length / (float)nx / rupvel;
    float sx[4], sy[4];
    if (risetime < dursf) {
        sx[0] = (-dursf - risetime) / 2.f; sx[1] = (-dursf + risetime) / 2.f; sx[2] = (dursf - risetime) / 2.f; sx[3] = (dursf + risetime) / 2.f;
        sy[0] = 0.f; sy[1] = 1.f / dursf; sy[2] = 1.f / dursf; sy[3] = 0.f;
    } else {
        sx[0] = (-risetime - dursf) / 2.f; sx[1] = (-risetime + dursf) / 2.f; sx[2] = (risetime - dursf) / 2.f; sx[3] = (risetime + dursf) / 2.f;
        sy[0] = 0.f; sy[1] = 1.f / risetime; sy[2] = 1.f / risetime; sy[3] = 0.f;
    }
    durfull = dursf + risetime;
    float tbeg = sx[0];
    float dt = durfull / (float)nt;
    o.toff.resize(nt); o.wt.resize(nt);
    for (int it = 1; it <= nt; it++) {
        float ta = tbeg + dt * (float)(it - 1);
        float tb = tbeg + dt * (float)it;
        plf_integrate_and_centroid(sx, sy, 4, ta, tb, &o.wt[it - 1], &o.toff[it - 1]);
    }
    // m_rot = matmul(rotmat_slip, matmul(m_unrot, transpose(rotmat_slip))) / np  :426-428
    const float m_unrot[9] = {0, 0, -1, 0, 0, 0, -1, 0, 0};
    float trot[9], tmp[9], m_rot[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) trot[i * 3 + j] = rot_slip[j * 3 + i];
    auto matmul3 = [](const float* a, const float* b, float* c) {
        for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) {
            float s = 0.f;
            for (int j = 0; j < 3; j++) s = s + a[i * 3 + j] * b[j * 3 + k];
            c[i * 3 + k] = s;
        }
    };
    matmul3(m_unrot, trot, tmp);
    matmul3(rot_slip, tmp, m_rot);
    int np = nx * ny;
    for (int i = 0; i < 9; i++) m_rot[i] = m_rot[i] / (float)np;
    o.mhat[0] = m_rot[0]; o.mhat[1] = m_rot[4]; o.mhat[2] = m_rot[8]; o.mhat[3] = m_rot[1]; o.mhat[4] = m_rot[2]; o.mhat[5] = m_rot[5];
    for (int i = 0; i < 16 && i < 14; i++) o.p[i] = p[i];
    return true;
}

// ---- source_moment_tensor.f90:205-267 ------------------------------------------------------------
bool prep_moment_tensor(const float* p, float shortest_doi, SourcePrep* out) {
    SourcePrep& o = *out;
    o = SourcePrep();
    float risetime = p[10], time = p[0];
    float ft = risetime / shortest_doi;
    if (!(fabsf(ft) < 1e6f)) return false;
    int nt = (int)floorf(ft) + 1;
    if (nt <= 1) nt = 2;
    float sx[4] = {(-risetime) / 2.f, (-risetime) / 2.f, (risetime) / 2.f, (risetime) / 2.f};
    float sy[4] = {0.f, 1.f / risetime, 1.f / risetime, 0.f};
    float tbeg = sx[0];
    float dt = risetime / (float)nt;
    o.toff.resize(nt); o.wt.resize(nt);
    for (int it = 1; it <= nt; it++) {
        float ta = tbeg + dt * (float)(it - 1);
        float tb = tbeg + dt * (float)it;
        plf_integrate_and_centroid(sx, sy, 4, ta, tb, &o.wt[it - 1], &o.toff[it - 1]);
    }
    o.nx = 1; o.ny = 1; o.nt = nt; o.ngroups = 1;
    o.moment = 1.f; o.risetime = 0.f;   // source_moment_tensor.f90:199
    for (int i = 0; i < 6; i++) o.mhat[i] = p[4 + i];
    for (int i = 0; i < 11; i++) o.p[i] = p[i];
    o.point[0] = p[1]; o.point[1] = p[2]; o.point[2] = p[3]; o.time = time;
    return true;
}

// ---- source_circular.f90: circular rupture with constant rupture velocity ---------------------------
// 11 parameters: time north east depth moment strike dip rake radius rupture-velocity rise-time
bool prep_circular(const float* p, float shortest_doi, SourcePrep* out) {
    SourcePrep& o = *out;
    o = SourcePrep();
    for (int i = 0; i < 11; i++) if (!std::isfinite(p[i])) return false;
    // psm_update_dep_params_circular :209-232 -- note that it takes params(9), the radius, as rupture direction
    const float strike = d2r_r(p[5]), dip = d2r_r(p[6]), rake = d2r_r(p[7]), rupdir = d2r_r(p[8]);
    float rot_slip[9];
    init_euler(dip, strike, -rupdir, o.rot_rup);
    init_euler(dip, strike, -rake, rot_slip);
    o.moment = p[4]; o.risetime = 0.f;
    const float time = p[0], north = p[1], east = p[2], depth = p[3], radius = p[8], rupvel = p[9], risetime = p[10];
    if (!(rupvel > 0.f)) return false;
    // psm_to_tdsm_size_circular :267-302
    const float maxdt = shortest_doi, maxdx = 0.5f * shortest_doi * rupvel;
    const float length = radius * 2.f;
    const float fx = length / maxdx;
    if (!(fabsf(fx) < 3000.f)) return false;
    int nx = (int)floorf(fx) + 1;
    if (nx <= 1) nx = 2;
    if (length == 0.f) nx = 1;
    const int ny = nx;
    float dursf = length / (float)nx / rupvel;
    float durfull = risetime + dursf;
    const float ft = durfull / maxdt;
    if (!(fabsf(ft) < 1e6f)) return false;
    int nt = (int)floorf(ft) + 1;
    if (nt <= 1) nt = 2;
    // psm_to_tdsm_table_circular :305-444
    for (int ix = 1; ix <= nx; ix++)
        for (int iy = 1; iy <= ny; iy++) {
            const float x = (2.f * ((float)ix - 1.f) - (float)nx + 1.f) / (2.f * (float)nx) * length;
            const float y = (2.f * ((float)iy - 1.f) - (float)ny + 1.f) / (2.f * (float)ny) * length;
            const float r = sqrtf(x * x + y * y);
            if (r <= radius) {
                const float g[3] = {x, y, 0.f};
                float q[3];
                for (int i = 0; i < 3; i++) { float a = 0.f; for (int j = 0; j < 3; j++) a = a + o.rot_rup[i * 3 + j] * g[j]; q[i] = a; }
                o.g_north.push_back(q[0] + north); o.g_east.push_back(q[1] + east); o.g_depth.push_back(q[2] + depth);
                o.g_tbase.push_back(r / rupvel + time);
                o.g_gw.push_back(1.f); o.g_tap_begin.push_back(0); o.g_tap_count.push_back(nt);
            }
        }
    const int np = (int)o.g_north.size();
    if (np == 0) return false;
    dursf = length / (float)nx / rupvel;
    float sx[4], sy[4];
    if (risetime < dursf) {
        sx[0] = (-dursf - risetime) / 2.f; sx[1] = (-dursf + risetime) / 2.f; sx[2] = (dursf - risetime) / 2.f; sx[3] = (dursf + risetime) / 2.f;
        sy[0] = 0.f; sy[1] = 1.f / dursf; sy[2] = 1.f / dursf; sy[3] = 0.f;
    } else {
        sx[0] = (-risetime - dursf) / 2.f; sx[1] = (-risetime + dursf) / 2.f; sx[2] = (risetime - dursf) / 2.f; sx[3] = (risetime + dursf) / 2.f;
        sy[0] = 0.f; sy[1] = 1.f / risetime; sy[2] = 1.f / risetime; sy[3] = 0.f;
    }
    durfull = dursf + risetime;
    const float tbeg = sx[0], dt = durfull / (float)nt;
    o.toff.resize(nt); o.wt.resize(nt);
    for (int it = 1; it <= nt; it++) plf_integrate_and_centroid(sx, sy, 4, tbeg + dt * (float)(it - 1), tbeg + dt * (float)it, &o.wt[it - 1], &o.toff[it - 1]);
    const float m_unrot[9] = {0, 0, -1, 0, 0, 0, -1, 0, 0};
    float trot[9], tmp[9], m_rot[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) trot[i * 3 + j] = rot_slip[j * 3 + i];
    auto matmul3 = [](const float* a, const float* b, float* c) {
        for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) { float v = 0.f; for (int j = 0; j < 3; j++) v = v + a[i * 3 + j] * b[j * 3 + k]; c[i * 3 + k] = v; }
    };
    matmul3(m_unrot, trot, tmp);
    matmul3(rot_slip, tmp, m_rot);
    for (int i = 0; i < 9; i++) m_rot[i] = m_rot[i] / (float)np;
    o.mhat[0] = m_rot[0]; o.mhat[1] = m_rot[4]; o.mhat[2] = m_rot[8]; o.mhat[3] = m_rot[1]; o.mhat[4] = m_rot[2]; o.mhat[5] = m_rot[5];
    o.explicit_groups = true;
    o.nx = nx; o.ny = ny; o.nt = nt; o.ngroups = np;
    return true;
}

// ---- source_point_lp.f90: long-period point source with an analytic source time function ---------------
// 13 parameters: time north east depth moment mxx myy mzz mxy mxz myz duration-of-excitation period
static float stf_point_lp(float reltime, float prd, float dur_exc) {   // :412-421
    const float t1 = 2.f;
    const float t2 = t1 + dur_exc - 5.f;
    const float t3 = t2 / 4.f;
    return expf(-((reltime - t3) * (reltime - t3)) / (2.f * pi * dur_exc)) * 1.f / (1.f + expf(-2.f * (reltime - t1))) * 1.f /
           (1.f + expf(0.5f * (reltime - t2))) * sinf(2.f * pi / prd * reltime);
}
bool prep_point_lp(const float* p, float shortest_doi, SourcePrep* out) {
    SourcePrep& o = *out;
    o = SourcePrep();
    for (int i = 0; i < 13; i++) if (!std::isfinite(p[i])) return false;
    const float maxdt = shortest_doi, dur_exc = p[11], prd = p[12];
    const float ft = dur_exc / maxdt;
    if (!(fabsf(ft) < 4096.f)) return false;
    int nt = (int)floorf(ft) + 1;   // :238-241
    if (nt <= 1) nt = 2;
    o.moment = p[4]; o.risetime = 0.f;
    for (int i = 0; i < 6; i++) o.mhat[i] = p[5 + i];
    o.toff.resize(nt); o.wt.resize(nt);
    for (int it = 1; it <= nt; it++) {   // :303-316
        const float rel_time = (float)(it - 1) * maxdt;
        o.wt[it - 1] = stf_point_lp(rel_time, prd, dur_exc);
        o.toff[it - 1] = p[0] + (float)it * maxdt;
    }
    // one position; at most 32 time centroids per device group
    for (int t0 = 0; t0 < nt; t0 += 32) {
        o.g_north.push_back(p[1]); o.g_east.push_back(p[2]); o.g_depth.push_back(p[3]); o.g_tbase.push_back(0.f); o.g_gw.push_back(1.f);
        o.g_tap_begin.push_back(t0); o.g_tap_count.push_back(std::min(32, nt - t0));
    }
    o.explicit_groups = true;
    o.nx = 1; o.ny = 1; o.nt = std::min(nt, 32); o.ngroups = (int)o.g_north.size();
    return true;
}


// source_bilat.f90:565-594
static void polar3(const float xyz[3], float pol[3]) {
    pol[0] = sqrtf(xyz[0] * xyz[0] + xyz[1] * xyz[1] + xyz[2] * xyz[2]);
    pol[1] = atan2f(xyz[1], xyz[0]);
    pol[2] = acosf(xyz[2] / pol[0]);
}
static float wrap_r(float x, float mi, float ma) { return x - floorf((x - mi) / (ma - mi)) * (ma - mi); }
static void domeshot3(const float pol[3], float out[3]) {
    const float pi = 3.14159265358979f;
    out[0] = pol[0];
    out[1] = wrap_r(pol[1], pi, -pi); out[2] = wrap_r(pol[2], pi, -pi);
    if (out[2] > pi / 2.f) { out[1] = wrap_r(out[1] + pi, -pi, pi); out[2] = pi - out[2]; }
}
void principal_axes(float strike_deg, float dip_deg, float rake_deg, float pax[2], float tax[2]) {
    const float pi = 3.14159265358979f;
    float rot[9];
    init_euler(d2r_r(dip_deg), d2r_r(strike_deg), -d2r_r(rake_deg), rot);
    const float r2 = sqrtf(2.f);
    const float vp[3] = {r2, 0.f, -r2}, vt[3] = {-r2, 0.f, -r2};
    for (int which = 0; which < 2; which++) {
        const float* v = which == 0 ? vp : vt;
        float x[3], pol[3], ds[3];
        for (int i = 0; i < 3; i++) { float a = 0.f; for (int j = 0; j < 3; j++) a = a + rot[i * 3 + j] * v[j]; x[i] = a; }
        polar3(x, pol); domeshot3(pol, ds);
        float* out = which == 0 ? pax : tax;
        out[0] = 360.f / 2.f / pi * ds[1]; out[1] = 360.f / 2.f / pi * ds[2];   // r2d
    }
}

}  // namespace kh
