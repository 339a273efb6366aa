// TEST INFRASTRUCTURE ONLY (see ko_base.hpp).  Restates the analytical full-space fixture generator: gfdb_build_ahfull.f90
// (addentry, gfdb_save_array), elseis.f90 (elseis_mt, factors_mt, radpat_mt, material_factors_mt, make_direction_cosine,
// make_istfs), elseis_oo.f90 (set_stf), integration.f90 (antiderivate), differentiation.f90 (differentiate).
// Written from the Fortran, independently of the product's builder (kiwi_b200/csrc/gfdb_host.cpp), so that
// tests/test_gfdb_host.py can hold the two against each other trace by trace: a database both sides consume is only as good as
// its one builder otherwise.
#pragma once
#include "ko_trace.hpp"

namespace ko {
namespace ahfull {

// integration.f90:27-59
static inline void antiderivate(float dt, const std::vector<float>& f, std::vector<float>& ff) {
    const int n = (int)std::min(f.size(), ff.size());
    if ((int)f.size() < 2) { for (float& v : ff) v = 0.0f; return; }
    ff[0] = 0.0f;
    for (int i = 1; i <= n - 1; i++) ff[i] = ff[i - 1] + (f[i] + f[i - 1]) / 2 * dt;
}

// differentiation.f90:27-70
static inline void differentiate(float dt, const std::vector<float>& f, std::vector<float>& df) {
    const int n = (int)f.size();
    for (int i = 2; i <= n - 1; i++) df[i - 1] = (f[i] - f[i - 2]) / (dt * 2);
    df[0] = (f[1] - f[0]) / dt;
    df[n - 1] = (f[n - 1] - f[n - 2]) / dt;
}

struct Medium {   // elseis_oo.f90 elseis_t, the parts addentry uses
    float rho, alpha, beta, dt;
    std::vector<float> stf, dstf, istf, istftau;
    float material_factor_mt[5];
};

// elseis_oo.f90:127-157 set_stf + :74-84 set_material (elseis.f90:434-452 make_istfs, :321-337 material_factors_mt)
static inline Medium make_medium(float rho, float alpha, float beta, const float* stf, int lstf, float dt) {
    Medium m;
    m.rho = rho; m.alpha = alpha; m.beta = beta; m.dt = dt;
    m.stf.assign(stf, stf + lstf);
    m.dstf.assign(lstf, 0.f); m.istf.assign(lstf, 0.f); m.istftau.assign(lstf, 0.f);
    std::vector<float> stftau(lstf);
    for (int i = 1; i <= lstf; i++) stftau[i - 1] = m.stf[i - 1] * (i - 1) * dt;
    antiderivate(dt, m.stf, m.istf);
    antiderivate(dt, stftau, m.istftau);
    differentiate(dt, m.stf, m.dstf);
    const float PI = pi;   // constants.f90:21
    m.material_factor_mt[0] = 1.0f / (4.0f * PI * rho);
    m.material_factor_mt[1] = 1.0f / (4.0f * PI * rho * (alpha * alpha));            // alpha**2
    m.material_factor_mt[2] = 1.0f / (4.0f * PI * rho * (beta * beta));
    m.material_factor_mt[3] = 1.0f / (4.0f * PI * rho * ((alpha * alpha) * alpha));    // alpha**3 as gfortran expands it
    m.material_factor_mt[4] = 1.0f / (4.0f * PI * rho * ((beta * beta) * beta));
    return m;
}

static inline float kron(int a, int b) { return a == b ? 1.f : 0.f; }

// elseis.f90:343-373
static inline void radpat_mt(const float gamma[3], int n, int p, int q, float rpc[5]) {
    const float gn = gamma[n - 1], gp = gamma[p - 1], gq = gamma[q - 1];
    rpc[0] = (15 * gn * gp * gq) - (3 * gn * kron(p, q)) - (3 * gp * kron(n, q)) - (3 * gq * kron(n, p));
    rpc[1] = (6 * gn * gp * gq) - (gn * kron(p, q)) - (gp * kron(n, q)) - (gq * kron(n, p));
    rpc[2] = -((6 * gn * gp * gq) - (gn * kron(p, q)) - (gp * kron(n, q)) - (2 * gq * kron(n, p)));
    rpc[3] = gn * gp * gq;
    rpc[4] = -(gn * gp - kron(n, p)) * gq;
}

// elseis.f90:293-305 (integer powers as gfortran expands them: r**2 = r*r, r**4 = (r*r)*(r*r))
static inline void factors_mt(const float matfac[5], const float radpat[5], float r, float factors[5]) {
    const float r2 = r * r;
    factors[0] = matfac[0] * radpat[0] / (r2 * r2);
    factors[1] = matfac[1] * radpat[1] / r2;
    factors[2] = matfac[2] * radpat[2] / r2;
    factors[3] = matfac[3] * radpat[3] / r;
    factors[4] = matfac[4] * radpat[4] / r;
}

// elseis.f90:133-209, the `addweight` form: elseism(1:npt) += term * addweight
static inline void elseis_mt_add(const float factors[5], float r, const Medium& m, float toffset, bool nfflag, bool ffflag, float* elseism, int npt,
                                 float addweight) {
    const float dt = m.dt, alpha = m.alpha, beta = m.beta;
    const int lstf = (int)m.stf.size();
    const int ita_delta = f_nint(toffset / dt - r / alpha / dt);
    const int itb_delta = f_nint(toffset / dt - r / beta / dt);
    for (int it = 1; it <= npt; it++) {
        const float t = toffset + (it - 1) * dt;
        const float ta = t - r / alpha;
        const float tb = t - r / beta;
        int ita = ita_delta + (it - 1);
        int itb = itb_delta + (it - 1);
        if (ita < 0) ita = 0;                 // to_bounds( 0, lstf-1, . )
        if (lstf - 1 < ita) ita = lstf - 1;
        if (itb < 0) itb = 0;
        if (lstf - 1 < itb) itb = lstf - 1;
        float ta_delta = 0.f, tb_delta = 0.f;
        if (nfflag) { ta_delta = ta - ita * dt; tb_delta = tb - itb * dt; }
        ita = ita + 1;                        // 1-based from here on, as in the Fortran
        itb = itb + 1;
        const float sa = m.stf[ita - 1], sb = m.stf[itb - 1];
        float term = 0.0f;
        if (nfflag) {
            const float integral_term =
                t * (m.istf[ita - 1] - m.istf[itb - 1] + ta_delta * sa - tb_delta * sb) -
                (m.istftau[ita - 1] + ta_delta * sa * (ita - 1) * dt + 0.5f * sa * (ta_delta * ta_delta) - m.istftau[itb - 1] -
                 tb_delta * sb * (itb - 1) * dt - 0.5f * sb * (tb_delta * tb_delta));
            term = term + factors[0] * integral_term;
            term = term + factors[1] * sa;
            term = term + factors[2] * sb;
        }
        if (ffflag) {
            term = term + factors[3] * m.dstf[ita - 1];
            term = term + factors[4] * m.dstf[itb - 1];
        }
        elseism[it - 1] = elseism[it - 1] + term * addweight;
    }
}

// gfdb_build_ahfull.f90:34-37: reshape((/.../),(/3,3/)) fills column by column, source(p,q) = list((q-1)*3 + p)
static inline float source_weight(int which, int p, int q) {
    static const float lists[4][9] = {{1, 1, 0, 1, 0, 0, 0, 0, 0}, {0, 0, 1, 0, 0, 1, 1, 1, 0}, {0, 0, 0, 0, 0, 0, 0, 0, 1}, {0, 0, 0, 0, 1, 0, 0, 0, 0}};
    return lists[which][(q - 1) * 3 + (p - 1)];
}

// gfdb_build_ahfull.f90:70-191 addentry for the line "x z nfflag ffflag": the ten packed traces of the node, as
// gfdb_save_array (:193-216) hands them to gfdb_save_trace.  traces[ig-1], ig = 1..10.
static inline void addentry(const Medium& es, float dbdt, float x, float z, bool nfflag, bool ffflag, Trace traces[10]) {
    const float s_location[3] = {0.f, 0.f, z}, r_location[3] = {x, 0.f, 0.f};
    float rel_location[3];
    for (int i = 0; i < 3; i++) rel_location[i] = r_location[i] - s_location[i];
    const float d = sqrtf((s_location[0] - r_location[0]) * (s_location[0] - r_location[0]) + (s_location[1] - r_location[1]) * (s_location[1] - r_location[1]) +
                          (s_location[2] - r_location[2]) * (s_location[2] - r_location[2]));
    const float tstf = ((int)es.stf.size() - 1) * es.dt;
    auto snapdown = [](float t, float dt) { return f_floor(t / dt) * dt; };
    auto snapup = [](float t, float dt) { return (float)(int)ceilf(t / dt) * dt; };
    const float firstarrival_p = snapdown(d / es.alpha, dbdt);
    const float lastarrival_p = snapup(d / es.alpha + tstf, dbdt);
    const float firstarrival_s = snapdown(d / es.beta, dbdt);
    const float lastarrival_s = snapup(d / es.beta + tstf, dbdt) + dbdt * 2;   // add 2 samples of zero/static at the end
    const float tbegin_total = firstarrival_p, tend_total = lastarrival_s;
    int nwindows;
    float tbegin[2], tend[2];
    if (lastarrival_p >= firstarrival_s || nfflag) {
        nwindows = 1; tbegin[0] = firstarrival_p; tend[0] = lastarrival_s;
    } else {
        nwindows = 2; tbegin[0] = firstarrival_p; tend[0] = lastarrival_p; tbegin[1] = firstarrival_s; tend[1] = lastarrival_s;
    }
    const int nsamples = f_nint((tend_total - tbegin_total) / es.dt + 1);
    std::vector<float> seismograms((size_t)12 * nsamples, 0.f);   // seismograms(12,nsamples), row i at [(i-1)*nsamples]
    // set_coords: elseis.f90:399-414 make_direction_cosine
    const float r = sqrtf(rel_location[0] * rel_location[0] + rel_location[1] * rel_location[1] + rel_location[2] * rel_location[2]);
    const float gamma[3] = {rel_location[0] / r, rel_location[1] / r, rel_location[2] / r};
    for (int n = 1; n <= 3; n++)
        for (int p = 1; p <= 3; p++)
            for (int q = 1; q <= 3; q++) {
                float radiation_factor[5], factor[5];     // set_npq -> update_radiation_factors -> update_factors
                radpat_mt(gamma, n, p, q, radiation_factor);
                factors_mt(es.material_factor_mt, radiation_factor, r, factor);
                for (int iwindow = 0; iwindow < nwindows; iwindow++) {
                    const int itbegin = f_nint((tbegin[iwindow] - tbegin_total) / es.dt) + 1;
                    const int itend = f_nint((tend[iwindow] - tbegin_total) / es.dt) + 1;
                    for (int which = 0; which < 4; which++)   // source_a .. source_d -> rows n, n+3, n+6, n+9
                        elseis_mt_add(factor, r, es, tbegin[iwindow], nfflag, ffflag, &seismograms[(size_t)(n + 3 * which - 1) * nsamples + (itbegin - 1)],
                                      itend - itbegin + 1, source_weight(which, p, q));
                }
            }
    static const int row_of_component[10] = {1, 4, 7, 2, 5, 3, 6, 9, 10, 12};   // gfdb_build_ahfull.f90:166-175
    for (int ig = 1; ig <= 10; ig++) {   // gfdb_save_array: span(1) = nint(tbegin/db%dt), strip_init, trace_pack
        const int span1 = f_nint(tbegin_total / dbdt);
        Strip conti;
        strip_init(span1, span1 + nsamples - 1, &seismograms[(size_t)(row_of_component[ig - 1] - 1) * nsamples], nsamples, conti);
        trace_pack(conti, traces[ig - 1]);
    }
}

}  // namespace ahfull
}  // namespace ko
