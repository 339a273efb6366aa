// TEST INFRASTRUCTURE ONLY (see ko_base.hpp).  Restates discrete_source.f90,
// parameterized_source.f90 (fields used on the hot path), source_bilat.f90,
// source_moment_tensor.f90 and the dispatch of source_all.f90.
#pragma once
#include "ko_base.hpp"

namespace ko {

struct Centroid {  // discrete_source.f90:27-30
    float north, east, depth, time;
    float m[6];  // mxx myy mzz mxy mxz myz
};
struct Tdsm {  // discrete_source.f90:32-45
    double ref_time = 0.;
    GeoCoords origin;
    std::vector<Centroid> centroids;
};

// source type ids as in parameterized_source.f90 (psm_bilat ... psm_moment_tensor)
enum SourceType { PSM_NONE = 0, PSM_BILAT = 1, PSM_CIRCULAR = 2, PSM_POINT_LP = 3, PSM_EIKONAL = 4,
                  PSM_MT_EIKONAL = 5, PSM_MOMENT_TENSOR = 6 };

struct Psm {  // parameterized_source.f90:63-94
    int sourcetype = 0;
    double ref_time = 0.;
    GeoCoords origin;
    float moment = 1.0f;
    float risetime = 0.0f;
    std::vector<float> params;
    float rotmat_rup[3][3], rotmat_slip[3][3];
    std::vector<int> grid_size;
};

static const int n_source_params_bilat = 14;          // source_bilat.f90:32
static const int n_source_params_moment_tensor = 11;  // source_moment_tensor.f90

// source_bilat.f90:216-239 (p-/t-axis bookkeeping is diagnostics only and omitted)
static inline void psm_update_dep_params_bilat(Psm& psm) {
    float strike = d2r_r(psm.params[5]);
    float dip = d2r_r(psm.params[6]);
    float rake = d2r_r(psm.params[7]);
    float rupdir = d2r_r(psm.params[8]);
    init_euler(dip, strike, -rupdir, psm.rotmat_rup);
    init_euler(dip, strike, -rake, psm.rotmat_slip);
}
// source_bilat.f90:173-214
static inline void psm_set_bilat(Psm& psm, const float* params, bool& only_moment_changed) {
    std::vector<float> np(params, params + n_source_params_bilat);
    bool same_type = (psm.sourcetype == PSM_BILAT) && (int)psm.params.size() == n_source_params_bilat;
    int ndiff = 0;
    if (same_type) { for (int i = 0; i < n_source_params_bilat; i++) if (np[i] != psm.params[i]) ndiff++; }
    else ndiff = n_source_params_bilat;  // psm%params freshly resized: contents undefined in Fortran
    only_moment_changed = same_type && (ndiff <= 1 && np[4] != psm.params[4]);
    psm.params = np;
    psm.sourcetype = PSM_BILAT;
    psm.moment = psm.params[4];
    psm.risetime = 0.0f;  // psm_reset_dependents, parameterized_source.f90:121-125
    psm_update_dep_params_bilat(psm);
}
// source_bilat.f90:274-315
static inline void psm_to_tdsm_size_bilat(const Psm& in, float maxdx, float maxdy, float maxdt, int& nx, int& ny, int& nt) {
    float length_a = in.params[9], length_b = in.params[10], width = in.params[11], rupvel = in.params[12],
          risetime = in.params[13];
    float length = length_a + length_b;
    nx = f_floor(length / maxdx) + 1;
    if (nx <= 1) nx = 2;
    if (length == 0.f) nx = 1;
    ny = f_floor(width / maxdy) + 1;
    if (ny <= 1) ny = 2;
    if (width == 0.f) ny = 1;
    float dursf = length / (float)nx / rupvel;
    float durfull = risetime + dursf;
    nt = f_floor(durfull / maxdt) + 1;
    if (nt <= 1) nt = 2;
}
// source_bilat.f90:318-459
static inline void psm_to_tdsm_table_bilat(Psm& psm, Tdsm& out, int nx, int ny, int nt) {
    float north = psm.params[1], east = psm.params[2], depth = psm.params[3];
    float length_a = psm.params[9], length_b = psm.params[10], width = psm.params[11], rupvel = psm.params[12],
          risetime = psm.params[13];
    float length = length_a + length_b;
    int np = nx * ny;
    std::vector<float> tshift(np), grid(3 * (size_t)np), wt(nt), toff(nt);
    int ip = 0;
    for (int ix = 1; ix <= nx; ix++) {
        for (int iy = 1; iy <= ny; iy++) {
            float g[3];
            g[0] = (2.f * ((float)ix - 1.f) - (float)nx + 1.f) / (2.f * (float)nx) * length;
            g[1] = (2.f * ((float)iy - 1.f) - (float)ny + 1.f) / (2.f * (float)ny) * width;
            g[2] = 0.f;
            tshift[ip] = fabsf(length / 2.f - length_b + g[0]) / rupvel + psm.params[0] -
                         std::max(length_a, length_b) / 2.f / rupvel;
            float p[3];
            matvec3(psm.rotmat_rup, g, p);
            grid[3 * ip + 0] = p[0] + north; grid[3 * ip + 1] = p[1] + east; grid[3 * ip + 2] = p[2] + depth;
            ip++;
        }
    }
    float dursf = length / (float)nx / rupvel;
    Plf stf;
    if (risetime < dursf) {
        plf_make(stf, {(-dursf - risetime) / 2.f, (-dursf + risetime) / 2.f, (dursf - risetime) / 2.f, (dursf + risetime) / 2.f},
                 {0.f, 1.f / dursf, 1.f / dursf, 0.f});
    } else {
        plf_make(stf, {(-risetime - dursf) / 2.f, (-risetime + dursf) / 2.f, (risetime - dursf) / 2.f, (risetime + dursf) / 2.f},
                 {0.f, 1.f / risetime, 1.f / risetime, 0.f});
    }
    float durfull = dursf + risetime;
    float tbeg = stf.x[0];
    float dt = durfull / (float)nt;
    for (int it = 1; it <= nt; it++) {
        float ta = tbeg + dt * (float)(it - 1);
        float tb = tbeg + dt * (float)it;
        plf_integrate_and_centroid(stf, ta, tb, wt[it - 1], toff[it - 1]);
    }
    int nd = np * nt;
    out.centroids.assign(nd, Centroid());
    float m_unrot[3][3] = {{0, 0, -1}, {0, 0, 0}, {-1, 0, 0}};  // reshape((/0,0,-1,0,0,0,-1,0,0/),(/3,3/)) is symmetric
    float trotmat[3][3], tmp[3][3], m_rot[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) trotmat[i][j] = psm.rotmat_slip[j][i];
    matmul3(m_unrot, trotmat, tmp);
    matmul3(psm.rotmat_slip, tmp, m_rot);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m_rot[i][j] = m_rot[i][j] / (float)np;
    int id = 0;
    for (ip = 0; ip < np; ip++) {
        for (int it = 0; it < nt; it++) {
            Centroid& c = out.centroids[id];
            c.north = grid[3 * ip + 0]; c.east = grid[3 * ip + 1]; c.depth = grid[3 * ip + 2];
            c.time = tshift[ip] + toff[it];
            c.m[0] = m_rot[0][0] * wt[it];
            c.m[1] = m_rot[1][1] * wt[it];
            c.m[2] = m_rot[2][2] * wt[it];
            c.m[3] = m_rot[0][1] * wt[it];
            c.m[4] = m_rot[0][2] * wt[it];
            c.m[5] = m_rot[1][2] * wt[it];
            id++;
        }
    }
}
// source_bilat.f90:241-271
static inline void psm_to_tdsm_bilat(Psm& psm, Tdsm& tdsm, float shortest_doi, bool& ok) {
    ok = true;
    float rupvel = psm.params[12];
    float maxdt = shortest_doi;
    float maxdx = 0.5f * shortest_doi * rupvel;
    float maxdy = shortest_doi * rupvel;
    int nx, ny, nt;
    psm_to_tdsm_size_bilat(psm, maxdx, maxdy, maxdt, nx, ny, nt);
    psm_to_tdsm_table_bilat(psm, tdsm, nx, ny, nt);
    psm.grid_size = {nx, ny, nt};
}

// source_moment_tensor.f90:164-203
static inline void psm_set_moment_tensor(Psm& psm, const float* params, bool& only_moment_changed) {
    only_moment_changed = false;
    psm.params.assign(params, params + n_source_params_moment_tensor);
    psm.sourcetype = PSM_MOMENT_TENSOR;
    psm.moment = 1.f;
    psm.risetime = 0.0f;
}
// source_moment_tensor.f90:205-267
static inline void psm_to_tdsm_moment_tensor(Psm& psm, Tdsm& tdsm, float shortest_doi, bool& ok) {
    ok = true;
    float point[3] = {psm.params[1], psm.params[2], psm.params[3]};
    float m[6]; for (int i = 0; i < 6; i++) m[i] = psm.params[4 + i];
    float risetime = psm.params[10];
    float time = psm.params[0];
    float maxdt = shortest_doi;
    int nt = f_floor(risetime / maxdt) + 1;
    if (nt <= 1) nt = 2;
    psm.grid_size = {nt};
    Plf stf;
    plf_make(stf, {(-risetime) / 2.f, (-risetime) / 2.f, (risetime) / 2.f, (risetime) / 2.f},
             {0.f, 1.f / risetime, 1.f / risetime, 0.f});
    float tbeg = stf.x[0];
    float dt = risetime / (float)nt;
    std::vector<float> wt(nt), toff(nt);
    for (int it = 1; it <= nt; it++) {
        float ta = tbeg + dt * (float)(it - 1);
        float tb = tbeg + dt * (float)it;
        plf_integrate_and_centroid(stf, ta, tb, wt[it - 1], toff[it - 1]);
    }
    tdsm.centroids.assign(nt, Centroid());
    for (int it = 0; it < nt; it++) {
        Centroid& c = tdsm.centroids[it];
        c.north = point[0]; c.east = point[1]; c.depth = point[2];
        c.time = toff[it] + time;
        for (int i = 0; i < 6; i++) c.m[i] = m[i] * wt[it];
    }
}

// ---- source_circular.f90 ---------------------------------------------------------------------------------
static const int n_source_params_circular = 11;   // :33
// :166-232 (note: psm_update_dep_params_circular reads params(9), the radius, as the rupture direction)
static inline void psm_set_circular(Psm& psm, const float* params, bool& only_moment_changed) {
    only_moment_changed = false;
    psm.params.assign(params, params + n_source_params_circular);
    psm.sourcetype = PSM_CIRCULAR;
    psm.moment = psm.params[4];
    psm.risetime = 0.0f;
    float strike = d2r_r(psm.params[5]), dip = d2r_r(psm.params[6]), rake = d2r_r(psm.params[7]), rupdir = d2r_r(psm.params[8]);
    init_euler(dip, strike, -rupdir, psm.rotmat_rup);
    init_euler(dip, strike, -rake, psm.rotmat_slip);
}
// :235-444
static inline void psm_to_tdsm_circular(Psm& psm, Tdsm& out, float shortest_doi, bool& ok) {
    ok = true;
    float time = psm.params[0], north = psm.params[1], east = psm.params[2], depth = psm.params[3];
    float radius = psm.params[8], rupvel = psm.params[9], risetime = psm.params[10];
    float maxdt = shortest_doi, maxdx = 0.5f * shortest_doi * rupvel;
    float length = radius * 2;
    int nx = f_floor(length / maxdx) + 1;
    if (nx <= 1) nx = 2;
    if (length == 0.f) nx = 1;
    int ny = nx;
    float dursf = length / (float)nx / rupvel;
    float durfull = risetime + dursf;
    int nt = f_floor(durfull / maxdt) + 1;
    if (nt <= 1) nt = 2;
    std::vector<float> tshift, grid;
    for (int ix = 1; ix <= nx; ix++)
        for (int iy = 1; iy <= ny; iy++) {
            float x = (2.f * ((float)ix - 1.f) - (float)nx + 1.f) / (2.f * (float)nx) * length;
            float y = (2.f * ((float)iy - 1.f) - (float)ny + 1.f) / (2.f * (float)ny) * length;
            float r = sqrtf(x * x + y * y);
            float g[3] = {x, y, 0.f}, p[3];
            matvec3(psm.rotmat_rup, g, p);
            if (r <= radius) {
                grid.push_back(p[0] + north); grid.push_back(p[1] + east); grid.push_back(p[2] + depth);
                tshift.push_back(r / rupvel + time);
            }
        }
    int np = (int)tshift.size();
    if (np == 0) { ok = false; return; }
    dursf = length / (float)nx / rupvel;
    Plf stf;
    if (risetime < dursf) plf_make(stf, {(-dursf - risetime) / 2.f, (-dursf + risetime) / 2.f, (dursf - risetime) / 2.f, (dursf + risetime) / 2.f}, {0.f, 1.f / dursf, 1.f / dursf, 0.f});
    else plf_make(stf, {(-risetime - dursf) / 2.f, (-risetime + dursf) / 2.f, (risetime - dursf) / 2.f, (risetime + dursf) / 2.f}, {0.f, 1.f / risetime, 1.f / risetime, 0.f});
    durfull = dursf + risetime;
    float tbeg = stf.x[0], dt = durfull / (float)nt;
    std::vector<float> wt(nt), toff(nt);
    for (int it = 1; it <= nt; it++) plf_integrate_and_centroid(stf, tbeg + dt * (float)(it - 1), tbeg + dt * (float)it, wt[it - 1], toff[it - 1]);
    float m_unrot[3][3] = {{0, 0, -1}, {0, 0, 0}, {-1, 0, 0}}, trotmat[3][3], tmp[3][3], m_rot[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) trotmat[i][j] = psm.rotmat_slip[j][i];
    matmul3(m_unrot, trotmat, tmp);
    matmul3(psm.rotmat_slip, tmp, m_rot);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m_rot[i][j] = m_rot[i][j] / (float)np;
    out.centroids.assign((size_t)np * nt, Centroid());
    int id = 0;
    for (int ip = 0; ip < np; ip++)
        for (int it = 0; it < nt; it++) {
            Centroid& c = out.centroids[id++];
            c.north = grid[3 * ip]; c.east = grid[3 * ip + 1]; c.depth = grid[3 * ip + 2];
            c.time = tshift[ip] + toff[it];
            c.m[0] = m_rot[0][0] * wt[it]; c.m[1] = m_rot[1][1] * wt[it]; c.m[2] = m_rot[2][2] * wt[it];
            c.m[3] = m_rot[0][1] * wt[it]; c.m[4] = m_rot[0][2] * wt[it]; c.m[5] = m_rot[1][2] * wt[it];
        }
    psm.grid_size = {nx, ny, nt};
}

// ---- source_point_lp.f90 ---------------------------------------------------------------------------------
static const int n_source_params_point_lp = 13;   // :14
static inline void psm_set_point_lp(Psm& psm, const float* params, bool& only_moment_changed) {   // :184-215
    only_moment_changed = false;
    psm.params.assign(params, params + n_source_params_point_lp);
    psm.sourcetype = PSM_POINT_LP;
    psm.moment = psm.params[4];
    psm.risetime = 0.0f;
}
static inline float stf_point_lp(float reltime, float prd, float dur_exc) {   // :412-421
    float t1 = 2.f;
    float t2 = t1 + dur_exc - 5;
    float t3 = t2 / 4.f;
    return expf(-((reltime - t3) * (reltime - t3)) / (2 * pi * dur_exc)) * 1.f / (1.f + expf(-2.f * (reltime - t1))) * 1.f /
           (1.f + expf(0.5f * (reltime - t2))) * sinf(2.f * pi / prd * reltime);
}
static inline void psm_to_tdsm_point_lp(Psm& psm, Tdsm& out, float shortest_doi, bool& ok) {   // :217-337
    ok = true;
    float maxdt = shortest_doi, dur_exc = psm.params[11], prd = psm.params[12];
    int nt = f_floor(dur_exc / maxdt) + 1;
    if (nt <= 1) nt = 2;
    float timestepsize = maxdt;
    out.centroids.assign(nt, Centroid());
    for (int it = 1; it <= nt; it++) {
        float rel_time = (float)(it - 1) * timestepsize;
        float tfactor = stf_point_lp(rel_time, prd, dur_exc);
        Centroid& c = out.centroids[it - 1];
        c.north = psm.params[1]; c.east = psm.params[2]; c.depth = psm.params[3];
        c.time = psm.params[0] + (float)it * timestepsize;
        for (int k = 0; k < 6; k++) c.m[k] = psm.params[5 + k] * tfactor;
    }
    psm.grid_size = {1, 1, nt};
}

// source_all.f90:216-261 / :431-465 (dispatch; only the source types in scope)
static inline bool psm_set(Psm& psm, int sourcetype, const float* params, int nparams, bool& only_moment_changed) {
    only_moment_changed = false;
    if (sourcetype == PSM_BILAT) {
        if (nparams != n_source_params_bilat) return false;
        psm_set_bilat(psm, params, only_moment_changed);
        return true;
    }
    if (sourcetype == PSM_MOMENT_TENSOR) {
        if (nparams != n_source_params_moment_tensor) return false;
        psm_set_moment_tensor(psm, params, only_moment_changed);
        return true;
    }
    if (sourcetype == PSM_CIRCULAR) {
        if (nparams != n_source_params_circular) return false;
        psm_set_circular(psm, params, only_moment_changed);
        return true;
    }
    if (sourcetype == PSM_POINT_LP) {
        if (nparams != n_source_params_point_lp) return false;
        psm_set_point_lp(psm, params, only_moment_changed);
        return true;
    }
    return false;
}
static inline void psm_to_tdsm(Psm& psm, Tdsm& tdsm, float shortest_doi, bool& ok) {
    tdsm.centroids.clear();
    ok = false;
    if (psm.sourcetype == PSM_BILAT) psm_to_tdsm_bilat(psm, tdsm, shortest_doi, ok);
    else if (psm.sourcetype == PSM_MOMENT_TENSOR) psm_to_tdsm_moment_tensor(psm, tdsm, shortest_doi, ok);
    else if (psm.sourcetype == PSM_CIRCULAR) psm_to_tdsm_circular(psm, tdsm, shortest_doi, ok);
    else if (psm.sourcetype == PSM_POINT_LP) psm_to_tdsm_point_lp(psm, tdsm, shortest_doi, ok);
    tdsm.origin = psm.origin;
    tdsm.ref_time = psm.ref_time;
}

}  // namespace ko
