// TEST INFRASTRUCTURE ONLY (see ko_base.hpp).  Restates geometry.f90, heap.f90, eikonal.f90,
// crust2x2.f90 (lookup part; the text tables are read from the converted binary table, see
// tools/make_crust2x2_table.py), the constraint part of parameterized_source.f90 and
// source_eikonal.f90 / source_mt_eikonal.f90.
#pragma once
#include "ko_source.hpp"
#include <cstdint>

namespace ko {

// ---- geometry.f90 ---------------------------------------------------------------------------
struct Vec3 { float v[3]; float& operator[](int i) { return v[i]; } float operator[](int i) const { return v[i]; } };
static inline Vec3 vsub(const Vec3& a, const Vec3& b) { return Vec3{{a[0] - b[0], a[1] - b[1], a[2] - b[2]}}; }
static inline float dot_product(const Vec3& a, const Vec3& b) { float s = 0.f; for (int i = 0; i < 3; i++) s = s + a[i] * b[i]; return s; }
struct HalfSpace { Vec3 point, normal; };                 // :25-28
struct Circle { Vec3 center; float transform[3][3]; };     // :30-34
typedef std::vector<Vec3> PolygonPts;                      // :36-38

static inline bool point_in_halfspace(const Vec3& point, const HalfSpace& hs) {   // :57-71
    return dot_product(hs.normal, vsub(hs.point, point)) >= 0.0f;
}
// :73-125
static inline void get_piercingpoint(const Vec3& a, const Vec3& b, const HalfSpace& hs, Vec3& pp, bool& between_ab, bool& parallel,
                                     bool* a_inside_ = nullptr, bool* b_inside_ = nullptr) {
    Vec3 ab = vsub(b, a);
    float lambda_a = dot_product(hs.normal, vsub(hs.point, a));
    float lambda_b = dot_product(hs.normal, vsub(hs.point, b));
    float lambda_ab = dot_product(hs.normal, ab);
    bool a_inside = lambda_a >= 0.f, b_inside = lambda_b >= 0.f;
    if (a_inside_) *a_inside_ = a_inside;
    if (b_inside_) *b_inside_ = b_inside;
    between_ab = (a_inside && !b_inside) || (b_inside && !a_inside);
    parallel = lambda_ab * lambda_ab < dot_product(ab, ab) / 16777216.f;   // 2**digits(lambda_ab), digits = 24
    if (parallel && between_ab) { pp = (fabsf(lambda_a) <= fabsf(lambda_b)) ? a : b; return; }
    if (parallel && !between_ab) { pp = Vec3{{0.f, 0.f, 0.f}}; return; }
    for (int i = 0; i < 3; i++) pp[i] = a[i] + ab[i] * lambda_a / lambda_ab;
}
// :191-211
static inline void circle_to_polygon(const Circle& c, int npoints, PolygonPts& poly) {
    poly.assign(npoints, Vec3{{0, 0, 0}});
    for (int i = 1; i <= npoints; i++) {
        float u[3] = {cosf((float)i * 2.f * pi / (float)npoints), sinf((float)i * 2.f * pi / (float)npoints), 0.f};
        float r[3];
        matvec3(c.transform, u, r);
        for (int k = 0; k < 3; k++) poly[i - 1][k] = r[k] + c.center[k];
    }
}
// :213-257
static inline void trim_polygon(const PolygonPts& poly, const HalfSpace& hs, PolygonPts& trimmed) {
    int npoints = (int)poly.size();
    std::vector<Vec3> piercing(npoints);
    std::vector<char> does_pierce(npoints), point_inside(npoints);
    for (int ip = 1; ip <= npoints; ip++) {
        int jp = ip % npoints + 1;
        bool dp, par, ai, bi;
        get_piercingpoint(poly[ip - 1], poly[jp - 1], hs, piercing[ip - 1], dp, par, &ai, &bi);
        does_pierce[ip - 1] = dp; point_inside[ip - 1] = ai;
    }
    trimmed.clear();
    for (int ip = 1; ip <= npoints; ip++) {
        if (point_inside[ip - 1]) trimmed.push_back(poly[ip - 1]);
        if (does_pierce[ip - 1]) trimmed.push_back(piercing[ip - 1]);
    }
}
// :259-276
static inline void trim_polygon(const PolygonPts& poly, const std::vector<HalfSpace>& hss, PolygonPts& trimmed) {
    PolygonPts temp = poly;
    trimmed = poly;
    for (size_t icon = 1; icon <= hss.size(); icon++) {
        if (icon != 1) temp = trimmed;
        PolygonPts t2;
        trim_polygon(temp, hss[icon - 1], t2);
        trimmed = t2;
    }
}
// :290-322
static inline float polygon_area(const PolygonPts& p) {
    int np = (int)p.size();
    float axy = 0.f, ayz = 0.f, azx = 0.f;
    if (np <= 2) return 0.f;
    for (int ip = 1; ip <= np; ip++) {
        int jp = ip % np + 1;
        const Vec3 &a = p[ip - 1], &b = p[jp - 1];
        axy = axy + (a[0] - b[0]) * (a[1] + b[1]) * 0.5f;
        ayz = ayz + (a[1] - b[1]) * (a[2] + b[2]) * 0.5f;
        azx = azx + (a[2] - b[2]) * (a[0] + b[0]) * 0.5f;
    }
    return sqrtf(axy * axy + ayz * ayz + azx * azx);
}

// ---- heap.f90: everything 1-based, iheap(0) unused ------------------------------------------------
struct IndexHeapO { std::vector<int> iheap; int n = 0; };
static inline void initheap(IndexHeapO& h, int maxsize) { h.iheap.assign(maxsize + 1, 0); h.n = 0; }
static inline void h_upheap(IndexHeapO& h, int element, const float* keys, int* back) {   // :210-232
    int v = element;
    while (v > 1) {
        int u = (v - 2) / 2 + 1;
        if (keys[h.iheap[u]] <= keys[h.iheap[v]]) return;
        std::swap(h.iheap[u], h.iheap[v]);
        if (back) std::swap(back[h.iheap[u]], back[h.iheap[v]]);
        v = u;
    }
}
static inline void h_downheap(IndexHeapO& h, int element, const float* keys, int* back) {   // :176-208
    int v = element, w = 2 * (v - 1) + 2;
    while (w <= h.n) {
        if (w + 1 <= h.n && keys[h.iheap[w + 1]] < keys[h.iheap[w]]) w = w + 1;
        if (keys[h.iheap[v]] <= keys[h.iheap[w]]) return;
        std::swap(h.iheap[v], h.iheap[w]);
        if (back) std::swap(back[h.iheap[v]], back[h.iheap[w]]);
        v = w; w = 2 * (v - 1) + 2;
    }
}
static inline void pushheap(IndexHeapO& h, int keyindex, const float* keys, int* back) {   // :70-93
    if (h.n + 1 > (int)h.iheap.size() - 1) return;
    h.n = h.n + 1;
    h.iheap[h.n] = keyindex;
    if (back) back[keyindex] = h.n;
    h_upheap(h, h.n, keys, back);
}
static inline void popheap(IndexHeapO& h, int& keyindex, const float* keys, int* back) {   // :95-124
    if (h.n == 0) { keyindex = 0; return; }
    std::swap(h.iheap[1], h.iheap[h.n]);
    if (back) { std::swap(back[h.iheap[1]], back[h.iheap[h.n]]); back[h.iheap[h.n]] = 0; }
    keyindex = h.iheap[h.n];
    h.n = h.n - 1;
    h_downheap(h, 1, keys, back);
}
static inline void updateheap(IndexHeapO& h, int keyindex, float newkey, float* keys, int* back) {   // :126-150
    float oldkey = keys[keyindex];
    keys[keyindex] = newkey;
    if (newkey < oldkey) h_upheap(h, back[keyindex], keys, back);
    if (newkey > oldkey) h_downheap(h, back[keyindex], keys, back);
}

// ---- eikonal.f90:29-199 ---------------------------------------------------------------------------------
// Field(ix,iy), 1-based, column-major like the Fortran arrays; linear index (iy-1)*nx+ix is also the heap key index
struct Field {
    int nx = 0, ny = 0; std::vector<float> a;   // a[0] unused
    void alloc(int nx_, int ny_, float v = 0.f) { nx = nx_; ny = ny_; a.assign((size_t)nx * ny + 1, v); }
    float& operator()(int ix, int iy) { return a[(size_t)(iy - 1) * nx + ix]; }
    float operator()(int ix, int iy) const { return a[(size_t)(iy - 1) * nx + ix]; }
};
static inline void eikonal_solver_fmm(const Field& speed, const float origin[2], const float delta[2], const float initialpoint[2],
                                      Field& times) {
    const int FARAWAY = -1, ALIVE = 0;
    float infinity = std::numeric_limits<float>::max() * 0.1f;
    int nx = speed.nx, ny = speed.ny;
    float dx = delta[0], dy = delta[1];
    std::vector<int> backpointers((size_t)nx * ny + 1, FARAWAY);
    auto ind = [nx](int ixl, int iyl) { return (iyl - 1) * nx + ixl; };
    int ix = (int)((initialpoint[0] - origin[0]) / dx) + 1;
    int iy = (int)((initialpoint[1] - origin[1]) / dy) + 1;
    if (ix < 1) ix = 1;
    if (nx < ix) ix = nx;
    if (iy < 1) iy = 1;
    if (ny < iy) iy = ny;
    times.alloc(nx, ny, infinity);
    times(ix, iy) = 0.0f;
    if (nx == 1 && ny == 1) return;
    backpointers[ind(ix, iy)] = ALIVE;
    int nalive = 1;
    IndexHeapO heap;
    initheap(heap, nx * ny);
    float* keys = times.a.data();
    int* back = backpointers.data();
    if (1 < ix) times(ix - 1, iy) = dx / speed(ix - 1, iy);
    if (ix < nx) times(ix + 1, iy) = dx / speed(ix + 1, iy);
    if (1 < iy) times(ix, iy - 1) = dy / speed(ix, iy - 1);
    if (iy < ny) times(ix, iy + 1) = dy / speed(ix, iy + 1);
    if (1 < ix) pushheap(heap, ind(ix - 1, iy), keys, back);
    if (ix < nx) pushheap(heap, ind(ix + 1, iy), keys, back);
    if (1 < iy) pushheap(heap, ind(ix, iy - 1), keys, back);
    if (iy < ny) pushheap(heap, ind(ix, iy + 1), keys, back);
    auto update_neighbor = [&](int ux, int uy) {
        int i = (uy - 1) * nx + ux;
        if (backpointers[i] == ALIVE) return;
        if (backpointers[i] == FARAWAY) pushheap(heap, i, keys, back);
        float a = infinity, b = infinity, c = infinity, d = infinity;
        float told = times(ux, uy);
        if (1 < ux) a = times(ux - 1, uy);
        if (ux < nx) b = times(ux + 1, uy);
        if (1 < uy) c = times(ux, uy - 1);
        if (uy < ny) d = times(ux, uy + 1);
        float t = 0.f;
        float aa = std::min(a, b), cc = std::min(c, d);
        float sp = speed(ux, uy);
        if (std::max(aa, cc) != infinity) {
            float e = (aa - cc) * sp;
            float s = dx * dx * (dy * dy) * (dx * dx + dy * dy - e * e);
            if (s >= 0.f) t = std::max(t, ((aa * (dy * dy) + cc * (dx * dx)) * sp + sqrtf(s)) / (sp * (dx * dx + dy * dy)));
        }
        if (std::min(c, d) == infinity) {
            if (a < infinity) t = std::max(t, a + dx / sp);
            if (b < infinity) t = std::max(t, b + dx / sp);
        }
        if (std::min(a, b) == infinity) {
            if (c < infinity) t = std::max(t, c + dy / sp);
            if (d < infinity) t = std::max(t, d + dy / sp);
        }
        if (t == 0.f) {
            t = infinity;
            if (a < infinity) t = std::min(t, a + dx / sp);
            if (b < infinity) t = std::min(t, b + dx / sp);
            if (c < infinity) t = std::min(t, c + dy / sp);
            if (d < infinity) t = std::min(t, d + dy / sp);
        }
        if (t != 0.f && told != t) updateheap(heap, ind(ux, uy), t, keys, back);
    };
    while (nalive <= nx * ny) {
        int imin;
        popheap(heap, imin, keys, back);
        if (imin == 0) break;
        ix = (imin - 1) % nx + 1;
        iy = (imin - 1) / nx + 1;
        backpointers[ind(ix, iy)] = ALIVE;
        nalive = nalive + 1;
        if (1 < ix) update_neighbor(ix - 1, iy);
        if (ix < nx) update_neighbor(ix + 1, iy);
        if (1 < iy) update_neighbor(ix, iy - 1);
        if (iy < ny) update_neighbor(ix, iy + 1);
    }
}

// ---- crust2x2.f90 --------------------------------------------------------------------------------------
struct Crust1dProfile { float vp[8], vs[8], rho[8], thickness[7], elevation; };   // :39-44
struct CrustModel {
    bool loaded = false; int ntypes = 0, nlo = 0, nla = 0;
    std::vector<Crust1dProfile> model;   // (ilon, ilat), ilon fastest
};
static inline bool crust2x2_load(const char* path, CrustModel& m) {   // :68-74 + :215-341 on the converted table
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    char magic[4]; int32_t hdr[3];
    if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "KCR1", 4) || fread(hdr, 4, 3, f) != 3) { fclose(f); return false; }
    m.ntypes = hdr[0]; m.nlo = hdr[1]; m.nla = hdr[2];
    std::vector<float> raw((size_t)m.ntypes * 31), elev((size_t)m.nlo * m.nla);
    std::vector<int16_t> tmap((size_t)m.nlo * m.nla);
    bool ok = fread(raw.data(), 4, raw.size(), f) == raw.size() && fread(tmap.data(), 2, tmap.size(), f) == tmap.size() &&
              fread(elev.data(), 4, elev.size(), f) == elev.size();
    fclose(f);
    if (!ok) return false;
    std::vector<Crust1dProfile> ctypes(m.ntypes);
    for (int i = 0; i < m.ntypes; i++) {
        Crust1dProfile& c = ctypes[i];
        for (int l = 0; l < 8; l++) { c.vp[l] = raw[i * 31 + l] * 1000.f; c.vs[l] = raw[i * 31 + 8 + l] * 1000.f; c.rho[l] = raw[i * 31 + 16 + l] * 1000.f; }
        for (int l = 0; l < 7; l++) c.thickness[l] = raw[i * 31 + 24 + l] * 1000.f;
        std::swap(c.vp[0], c.vp[1]); std::swap(c.vs[0], c.vs[1]); std::swap(c.rho[0], c.rho[1]); std::swap(c.thickness[0], c.thickness[1]);
        c.elevation = 0.f;
    }
    m.model.resize((size_t)m.nlo * m.nla);
    for (int j = 0; j < m.nla; j++)
        for (int i = 0; i < m.nlo; i++) {
            Crust1dProfile p = ctypes[tmap[(size_t)j * m.nlo + i]];
            p.elevation = elev[(size_t)j * m.nlo + i];
            if (p.elevation < 0.f && p.thickness[0] != 0.f) p.thickness[0] = -p.elevation;
            m.model[(size_t)j * m.nlo + i] = p;
        }
    m.loaded = true;
    return true;
}
static inline float c2_wrap(float x, float mi, float ma) { if (mi <= x && x <= ma) return x; return x - floorf((x - mi) / (ma - mi)) * (ma - mi); }
static inline Crust1dProfile crust2x2_get_profile(const CrustModel& m, const GeoCoords& location) {   // :76-92, :197-213
    float flat = std::min(std::max(-90.f, (float)location.lat), 90.f);
    float flon = c2_wrap((float)location.lon, -180.f, 180.f);
    float dx = 360.f / (float)m.nlo;
    float cola = 90.f - flat;
    int ilat = (int)(cola / dx) + 1, ilon = (int)((flon + 180.f) / dx) + 1;
    ilat = std::min(std::max(ilat, 1), m.nla); ilon = std::min(std::max(ilon, 1), m.nlo);
    return m.model[(size_t)(ilat - 1) * m.nlo + (ilon - 1)];
}
static inline void crust2x2_get_profile_averages(const Crust1dProfile& p, float& vvp, float& vvs, float& vrho, float& vthi) {   // :129-158
    vthi = 0.f; vvp = 0.f; vvs = 0.f; vrho = 0.f;
    for (int i = 2; i <= 7; i++) {
        vthi = vthi + p.thickness[i - 1];
        vvp = vvp + p.thickness[i - 1] / p.vp[i - 1];
        vvs = vvs + p.thickness[i - 1] / p.vs[i - 1];
        vrho = vrho + p.thickness[i - 1] * p.rho[i - 1];
    }
    vvp = vthi / vvp; vvs = vthi / vvs; vrho = vrho / vthi;
}
static inline void crust2x2_get_at_depth(const Crust1dProfile& p, float depth, float& vp, float& vs, float& rho) {   // :160-193
    float d = 0.f;
    for (int i = 3; i <= 7; i++) {
        d = d + p.thickness[i - 1];
        if (d >= depth) { vp = p.vp[i - 1]; vs = p.vs[i - 1]; rho = p.rho[i - 1]; return; }
    }
    vp = p.vp[7]; vs = p.vs[7]; rho = p.rho[7];
}

// ---- parameterized_source.f90:127-223 ------------------------------------------------------------------
struct EikonalGrid {   // :52-61
    float first[2], last[2], delta[2], initialpoint[2]; int ndims[2]; float minspeed;
    Field speed, times, durations, weights; std::vector<Vec3> points;   // points(:,ix,iy) at [(iy-1)*nx + ix-1]
};
struct PsmE {   // the fields of t_psm the eikonal sources use
    std::vector<float> params; GeoCoords origin; std::vector<HalfSpace> constraints; float crustal_thickness_limit = 0.f;
    float rotmat_rup[3][3], rotmat_slip[3][3]; float moment = 1.f, risetime = 0.f; int grid_size[2] = {1, 1};
    bool mt = false;
    int o(int i) const { return (mt && i >= 9) ? i - 1 : i; }   // 1-based index of params 9..15 (eikonal) in the mt_eikonal layout
    float par(int i) const { return params[o(i) - 1]; }
};
static inline void psm_set_default_constraints(PsmE& self, const CrustModel& cm) {   // :127-145, :209-223
    Crust1dProfile profile = crust2x2_get_profile(cm, r2d_tgc(self.origin));
    float vp, vs, vrho, thickness;
    crust2x2_get_profile_averages(profile, vp, vs, vrho, thickness);
    if (self.crustal_thickness_limit > 0) thickness = std::min(self.crustal_thickness_limit, thickness);
    self.constraints.assign(2, HalfSpace());
    self.constraints[0].point = Vec3{{0.f, 0.f, 1500.f}}; self.constraints[0].normal = Vec3{{0.f, 0.f, -1.f}};
    self.constraints[1].point = Vec3{{0.f, 0.f, thickness}}; self.constraints[1].normal = Vec3{{0.f, 0.f, 1.f}};
}
static inline bool psm_point_in_constraints(const PsmE& self, const Vec3& point) {   // :168-181
    for (const HalfSpace& h : self.constraints) if (!point_in_halfspace(point, h)) return false;
    return true;
}

// ---- source_eikonal.f90 ----------------------------------------------------------------------------------
static inline Vec3 psm_rc_to_ned(const PsmE& psm, const Vec3& rc) {   // :612-617
    float r[3]; matvec3(psm.rotmat_rup, rc.v, r);
    return Vec3{{r[0] + psm.params[1], r[1] + psm.params[2], r[2] + psm.params[3]}};
}
static inline Vec3 psm_ned_to_rc(const PsmE& psm, const Vec3& pt) {   // :605-610
    float t[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t[i][j] = psm.rotmat_rup[j][i];
    float d[3] = {pt[0] - psm.params[1], pt[1] - psm.params[2], pt[2] - psm.params[3]}, r[3];
    matvec3(t, d, r);
    return Vec3{{r[0], r[1], r[2]}};
}
static inline float veclen(const Vec3& a) { return sqrtf(dot_product(a, a)); }   // :888-892

// :185-257 (eikonal) / source_mt_eikonal.f90:185-262
static inline void psm_set_eikonal(PsmE& psm, const float* params, bool mt) {
    psm.mt = mt;
    psm.params.assign(params, params + (mt ? 20 : 15));
    psm.moment = psm.params[4];
    psm.risetime = mt ? psm.params[19] : psm.params[14];
    float strike = d2r_r(psm.params[5]), dip = d2r_r(psm.params[6]);
    if (!mt) { float rake = d2r_r(psm.params[7]); init_euler(dip, strike, -rake, psm.rotmat_slip); }
    init_euler(dip, strike, 0.f, psm.rotmat_rup);
}
// :714-764
static inline void discretize_subfault_time(float duration_subfault, float risetime, float maxdt, std::vector<float>& tweights,
                                            std::vector<float>& toffsets, int& nt) {
    float dursf = duration_subfault;
    float durfull = dursf + risetime;
    nt = f_floor(durfull / maxdt) + 1;
    if ((int)tweights.size() < nt) tweights.resize(nt);
    if ((int)toffsets.size() < nt) toffsets.resize(nt);
    if (nt == 1) { tweights[0] = 1.f; toffsets[0] = 0.f; return; }
    Plf stf;
    if (risetime < dursf) plf_make(stf, {(-dursf - risetime) / 2.f, (-dursf + risetime) / 2.f, (dursf - risetime) / 2.f, (dursf + risetime) / 2.f}, {0.f, 1.f / dursf, 1.f / dursf, 0.f});
    else plf_make(stf, {(-risetime - dursf) / 2.f, (-risetime + dursf) / 2.f, (risetime - dursf) / 2.f, (risetime + dursf) / 2.f}, {0.f, 1.f / risetime, 1.f / risetime, 0.f});
    float tbeg = stf.x[0];
    float dt = durfull / (float)nt;
    for (int it = 1; it <= nt; it++) plf_integrate_and_centroid(stf, tbeg + dt * (float)(it - 1), tbeg + dt * (float)it, tweights[it - 1], toffsets[it - 1]);
}
// :259-316 and everything it calls; returns false with `err` set where the reference sets ok=.false.
static inline bool psm_to_tdsm_eikonal(PsmE& psm, const CrustModel& cm, Tdsm& tdsm, float shortest_doi, std::string& err) {
    float bord_shift_x = psm.par(9), bord_shift_y = psm.par(10), bord_radius = psm.par(11);
    // psm_borderline_eikonal :318-348
    Circle circle;
    circle.center = psm_rc_to_ned(psm, Vec3{{bord_shift_x, bord_shift_y, 0.f}});
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) circle.transform[i][j] = -psm.rotmat_rup[i][j] * bord_radius;
    int n_initial_points = 180;
    if (bord_radius == 0.f) n_initial_points = 1;
    PolygonPts circle_poly, rupture_poly, rupture_poly_rc;
    circle_to_polygon(circle, n_initial_points, circle_poly);
    trim_polygon(circle_poly, psm.constraints, rupture_poly);
    rupture_poly_rc = rupture_poly;
    for (size_t ip = 0; ip < rupture_poly_rc.size(); ip++) rupture_poly_rc[ip] = psm_ned_to_rc(psm, rupture_poly[ip]);
    if (rupture_poly.size() == 0) { err = "Empty rupture area"; return false; }
    // polygon_box geometry.f90:278-286
    Vec3 min_rc = rupture_poly_rc[0], max_rc = rupture_poly_rc[0];
    for (const Vec3& q : rupture_poly_rc) for (int k = 0; k < 3; k++) { min_rc[k] = std::min(min_rc[k], q[k]); max_rc[k] = std::max(max_rc[k], q[k]); }
    float deltagrid = std::min(100.f * shortest_doi / 2.f, 4000.f);
    // psm_make_eikonal_grid :435-517
    EikonalGrid grid;
    float rel_rupture_velocity = psm.par(14);
    for (int k = 0; k < 2; k++) { grid.first[k] = min_rc[k]; grid.last[k] = max_rc[k]; }
    float dims[2] = {grid.last[0] - grid.first[0], grid.last[1] - grid.first[1]};
    for (int k = 0; k < 2; k++) { grid.ndims[k] = f_ceiling(dims[k] / deltagrid); if (grid.ndims[k] == 0) grid.ndims[k] = 1; }
    for (int k = 0; k < 2; k++) grid.delta[k] = dims[k] / (float)grid.ndims[k];
    int nxf = grid.ndims[0], nyf = grid.ndims[1];
    grid.speed.alloc(nxf, nyf); grid.points.assign((size_t)nxf * nyf, Vec3{{0, 0, 0}});
    Crust1dProfile profile = crust2x2_get_profile(cm, psm.origin);   // origin in radians: reference quirk (:472)
    Vec3 circle_center = psm_rc_to_ned(psm, Vec3{{bord_shift_x, bord_shift_y, 0.f}});
    {   // psm_initial_point_intolerant_rc :401-432
        float nukl_shift_x = psm.par(12), nukl_shift_y = psm.par(13);
        float nukl_shift = sqrtf(nukl_shift_x * nukl_shift_x + nukl_shift_y * nukl_shift_y);
        Vec3 ned = psm_rc_to_ned(psm, Vec3{{nukl_shift_x, nukl_shift_y, 0.f}});
        if (!psm_point_in_constraints(psm, ned) || nukl_shift > bord_radius) { err = "position of nucleation point is outside of rupture region"; return false; }
        grid.initialpoint[0] = nukl_shift_x; grid.initialpoint[1] = nukl_shift_y;
    }
    float minspeed = std::numeric_limits<float>::max();
    for (int iy = 1; iy <= nyf; iy++)
        for (int ix = 1; ix <= nxf; ix++) {
            Vec3 point_rc{{grid.first[0] + ((float)ix - 0.5f) * grid.delta[0], grid.first[1] + ((float)iy - 0.5f) * grid.delta[1], 0.f}};
            Vec3 point = psm_rc_to_ned(psm, point_rc);
            grid.points[(size_t)(iy - 1) * nxf + ix - 1] = point;
            if (veclen(vsub(point, circle_center)) > bord_radius || !psm_point_in_constraints(psm, point)) grid.speed(ix, iy) = 0.f;
            else {
                float vp, vs, rho;
                crust2x2_get_at_depth(profile, point[2], vp, vs, rho);
                grid.speed(ix, iy) = vs * rel_rupture_velocity;
                minspeed = std::min(grid.speed(ix, iy), minspeed);
            }
        }
    grid.minspeed = minspeed;
    if (minspeed == std::numeric_limits<float>::max()) { err = "no valid point in the rupture area"; return false; }
    float invalid_speed = minspeed * 0.5f;
    for (int iy = 1; iy <= nyf; iy++) for (int ix = 1; ix <= nxf; ix++) if (grid.speed(ix, iy) == 0.f) grid.speed(ix, iy) = invalid_speed;
    eikonal_solver_fmm(grid.speed, grid.first, grid.delta, grid.initialpoint, grid.times);
    for (int iy = 1; iy <= nyf; iy++) for (int ix = 1; ix <= nxf; ix++) if (grid.speed(ix, iy) == invalid_speed) grid.times(ix, iy) = -1.f;
    // optimal grid size :270-277, 617-638
    float maxdt = shortest_doi, maxdx = 0.5f * shortest_doi * grid.minspeed, maxdy = 0.5f * shortest_doi * grid.minspeed;
    float sizex = grid.last[0] - grid.first[0], sizey = grid.last[1] - grid.first[1];
    int nx = f_floor(sizex / maxdx) + 1;
    if (nx <= 1) nx = 2;
    if (sizex == 0.f) nx = 1;
    int ny = f_floor(sizey / maxdy) + 1;
    if (ny <= 1) ny = 2;
    if (sizey == 0.f) ny = 1;
    // psm_downsample_grid :519-601
    EikonalGrid cg;
    for (int k = 0; k < 2; k++) { cg.first[k] = grid.first[k]; cg.last[k] = grid.last[k]; }
    cg.ndims[0] = nx; cg.ndims[1] = ny;
    for (int k = 0; k < 2; k++) { cg.delta[k] = (cg.last[k] - cg.first[k]) / (float)cg.ndims[k]; if (cg.delta[k] == 0.f || cg.ndims[k] == 0) cg.delta[k] = 1.f; }
    Field ntimes; ntimes.alloc(nx, ny, 0.f);
    cg.times.alloc(nx, ny, -1.f); cg.speed.alloc(nx, ny, 0.f); cg.durations.alloc(nx, ny, 0.f); cg.weights.alloc(nx, ny, 0.f);
    cg.points.assign((size_t)nx * ny, Vec3{{0, 0, 0}});
    int npf = 0;
    for (int iyf = 1; iyf <= nyf; iyf++)
        for (int ixf = 1; ixf <= nxf; ixf++) {
            if (grid.times(ixf, iyf) < 0.f) continue;
            Vec3 prc = psm_ned_to_rc(psm, grid.points[(size_t)(iyf - 1) * nxf + ixf - 1]);
            int ixc = f_floor((prc[0] - cg.first[0]) / cg.delta[0]) + 1, iyc = f_floor((prc[1] - cg.first[1]) / cg.delta[1]) + 1;
            if (ixc < 1 || iyc < 1 || ixc > nx || iyc > ny) continue;   // "orphaned point in fine grid"
            ntimes(ixc, iyc) = ntimes(ixc, iyc) + 1.f;
            if (cg.times(ixc, iyc) == -1.f) cg.times(ixc, iyc) = 0.f;
            cg.times(ixc, iyc) = cg.times(ixc, iyc) + grid.times(ixf, iyf);
            cg.speed(ixc, iyc) = cg.speed(ixc, iyc) + 1.f / grid.speed(ixf, iyf);
            Vec3& cp = cg.points[(size_t)(iyc - 1) * nx + ixc - 1];
            const Vec3& fp = grid.points[(size_t)(iyf - 1) * nxf + ixf - 1];
            for (int k = 0; k < 3; k++) cp[k] = cp[k] + fp[k];
            npf = npf + 1;
        }
    for (int iy = 1; iy <= ny; iy++)
        for (int ix = 1; ix <= nx; ix++)
            if (ntimes(ix, iy) > 0.f) {
                cg.times(ix, iy) = 1.f / ntimes(ix, iy) * cg.times(ix, iy);
                cg.speed(ix, iy) = 1.f / (1.f / ntimes(ix, iy) * cg.speed(ix, iy));
                Vec3& cp = cg.points[(size_t)(iy - 1) * nx + ix - 1];
                for (int k = 0; k < 3; k++) cp[k] = 1.f / ntimes(ix, iy) * cp[k];
            }
    for (int iy = 1; iy <= ny; iy++) for (int ix = 1; ix <= nx; ix++) cg.weights(ix, iy) = ntimes(ix, iy) / (float)npf;
    for (int iyf = 1; iyf <= nyf; iyf++)
        for (int ixf = 1; ixf <= nxf; ixf++) {
            if (grid.times(ixf, iyf) < 0.f) continue;
            Vec3 prc = psm_ned_to_rc(psm, grid.points[(size_t)(iyf - 1) * nxf + ixf - 1]);
            int ixc = f_floor((prc[0] - cg.first[0]) / cg.delta[0]) + 1, iyc = f_floor((prc[1] - cg.first[1]) / cg.delta[1]) + 1;
            if (ixc < 1 || iyc < 1 || ixc > nx || iyc > ny) continue;
            cg.durations(ixc, iyc) = cg.durations(ixc, iyc) + fabsf(grid.times(ixf, iyf) - cg.times(ixc, iyc));
        }
    for (int iy = 1; iy <= ny; iy++) for (int ix = 1; ix <= nx; ix++) if (ntimes(ix, iy) > 0.f) cg.durations(ix, iy) = 4.f / ntimes(ix, iy) * cg.durations(ix, iy);
    // psm_to_tdsm_table_eikonal :640-712
    float origin_time = psm.params[0];
    float centertime = 0.f;
    for (int iy = 1; iy <= ny; iy++) for (int ix = 1; ix <= nx; ix++) if (cg.times(ix, iy) >= 0.f) centertime = centertime + cg.times(ix, iy) * cg.weights(ix, iy);
    float m6[6];
    if (!psm.mt) {
        float m_unrot[3][3] = {{0, 0, -1}, {0, 0, 0}, {-1, 0, 0}}, trotmat[3][3], tmp[3][3], m_rot[3][3];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) trotmat[i][j] = psm.rotmat_slip[j][i];
        matmul3(m_unrot, trotmat, tmp);
        matmul3(psm.rotmat_slip, tmp, m_rot);
        m6[0] = m_rot[0][0]; m6[1] = m_rot[1][1]; m6[2] = m_rot[2][2]; m6[3] = m_rot[0][1]; m6[4] = m_rot[0][2]; m6[5] = m_rot[1][2];
    } else for (int k = 0; k < 6; k++) m6[k] = psm.params[13 + k];   // source_mt_eikonal.f90:697-702
    tdsm.centroids.clear();
    std::vector<float> tweights, toffsets;
    for (int iy = 1; iy <= ny; iy++)
        for (int ix = 1; ix <= nx; ix++) {
            if (cg.times(ix, iy) < 0.f) continue;
            int nt;
            discretize_subfault_time(cg.durations(ix, iy), 0.f, maxdt, tweights, toffsets, nt);
            const Vec3& cp = cg.points[(size_t)(iy - 1) * nx + ix - 1];
            for (int it = 1; it <= nt; it++) {
                Centroid c;
                c.north = cp[0]; c.east = cp[1]; c.depth = cp[2];
                c.time = cg.times(ix, iy) + toffsets[it - 1] + origin_time - centertime;
                for (int k = 0; k < 6; k++) c.m[k] = m6[k] * tweights[it - 1] * cg.weights(ix, iy);
                tdsm.centroids.push_back(c);
            }
        }
    psm.grid_size[0] = nx; psm.grid_size[1] = ny;
    if (tdsm.centroids.empty()) { err = "Empty rupture area"; return false; }
    return true;
}

}  // namespace ko
