// Host-side discretisation of the eikonal / mt_eikonal rupture sources.
//
// source_eikonal.f90 / source_mt_eikonal.f90 turn 15 (20) parameters into sub-fault centroids:
// circular border clipped by constraint half-spaces (geometry.f90) -> fine grid in rupture
// coordinates with rupture speed = crustal S velocity (crust2x2.f90) x relative rupture velocity ->
// first-arrival times by fast marching (eikonal.f90 + heap.f90) -> coarse sub-fault grid with mean
// time, weight and duration -> per sub-fault time centroids.  The marching is inherently
// sequential (heap order decides ties) and per-candidate tiny, so it runs here on host threads,
// one candidate per task; the result feeds the same device SoA of groups and taps as the other
// source types.  fp32 throughout, statement order as in the reference, built -ffp-contract=off.
// Citations are file:line of /root/reference.
#include "host_math.hpp"
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>

namespace kh {

static const float pi_f = 3.14159265358979f;   // constants.f90:21

// ---- crust2x2.f90 -----------------------------------------------------------------------------------
static const int NLAYERS = 7, LWATER = 0 /* 1-based 1 */, LBELOWCRUST = 7 /* 1-based 8 */;

bool crust2x2_load(const char* path, Crust2x2* c, std::string* err) {
    FILE* f = fopen(path, "rb");
    if (!f) { *err = std::string("can't open file: ") + path; return false; }
    char magic[4]; int32_t hdr[3];
    bool ok = fread(magic, 1, 4, f) == 4 && memcmp(magic, "KCR1", 4) == 0 && fread(hdr, 4, 3, f) == 3;
    if (ok) {
        c->ntypes = hdr[0]; c->nlo = hdr[1]; c->nla = hdr[2];
        std::vector<float> raw((size_t)c->ntypes * 31);
        c->map.resize((size_t)c->nlo * c->nla); c->elev.resize((size_t)c->nlo * c->nla);
        ok = fread(raw.data(), 4, raw.size(), f) == raw.size() && fread(c->map.data(), 2, c->map.size(), f) == c->map.size() &&
             fread(c->elev.data(), 4, c->elev.size(), f) == c->elev.size();
        if (ok) {
            c->types.assign(c->ntypes, CrustProfile());
            for (int i = 0; i < c->ntypes; i++) {   // crust2x2.f90:274-297
                CrustProfile& p = c->types[i];
                const float* r = &raw[(size_t)i * 31];
                for (int l = 0; l < 8; l++) { p.vp[l] = r[l] * 1000.f; p.vs[l] = r[8 + l] * 1000.f; p.rho[l] = r[16 + l] * 1000.f; }
                for (int l = 0; l < 7; l++) p.thickness[l] = r[24 + l] * 1000.f;
                std::swap(p.vp[0], p.vp[1]); std::swap(p.vs[0], p.vs[1]); std::swap(p.rho[0], p.rho[1]);   // flip ice and water layers
                std::swap(p.thickness[0], p.thickness[1]);
                p.elevation = 0.f;
            }
        }
    }
    fclose(f);
    if (!ok) { *err = std::string("not a crust2x2 table: ") + path; return false; }
    c->loaded = true;
    return true;
}

// crust2x2.f90:76-92 + latlon2indices :197-213; lat/lon are used as given (see prep_eikonal about units)
CrustProfile crust2x2_get_profile(const Crust2x2& c, float lat, float lon) {
    float flat = std::min(std::max(-90.f, lat), 90.f);
    float flon = lon;
    if (!(-180.f <= flon && flon <= 180.f)) flon = flon - floorf((flon - (-180.f)) / (180.f - (-180.f))) * (180.f - (-180.f));
    const float dx = 360.f / (float)c.nlo;
    const float cola = 90.f - flat;
    int ilat = (int)(cola / dx) + 1;
    int ilon = (int)((flon + 180.f) / dx) + 1;
    ilat = std::min(std::max(ilat, 1), c.nla);   // the reference would index out of bounds at the south pole / date line
    ilon = std::min(std::max(ilon, 1), c.nlo);
    const size_t cell = (size_t)(ilat - 1) * c.nlo + (ilon - 1);
    CrustProfile p = c.types[c.map[cell]];
    p.elevation = c.elev[cell];
    if (p.elevation < 0.f && p.thickness[LWATER] != 0.f) p.thickness[LWATER] = -p.elevation;   // :336-340
    return p;
}
// crust2x2.f90:129-158
void crust2x2_get_profile_averages(const CrustProfile& p, float* vvp, float* vvs, float* vrho, float* vthi) {
    float thi = 0.f, vp = 0.f, vs = 0.f, rho = 0.f;
    for (int i = 1; i < NLAYERS; i++) {
        thi = thi + p.thickness[i];
        vp = vp + p.thickness[i] / p.vp[i];
        vs = vs + p.thickness[i] / p.vs[i];
        rho = rho + p.thickness[i] * p.rho[i];
    }
    *vvp = thi / vp; *vvs = thi / vs; *vrho = rho / thi; *vthi = thi;
}
// crust2x2.f90:160-193
static void crust2x2_get_at_depth(const CrustProfile& p, float depth, float* vp, float* vs, float* rho) {
    float d = 0.f;
    for (int i = 2; i < NLAYERS; i++) {
        d = d + p.thickness[i];
        if (d >= depth) { *vp = p.vp[i]; *vs = p.vs[i]; *rho = p.rho[i]; return; }
    }
    *vp = p.vp[LBELOWCRUST]; *vs = p.vs[LBELOWCRUST]; *rho = p.rho[LBELOWCRUST];
}

// parameterized_source.f90:127-145, 209-223: plane z >= 1500 m and z <= crustal thickness at the origin
// (this one converts the origin to degrees, unlike the velocity lookup of the eikonal grid)
void default_constraints(const Crust2x2& c, double olat_rad, double olon_rad, float thickness_limit, std::vector<Halfspace>* out) {
    const double r2d = (double)(360.f / 2.f / pi_f);   // orthodrome.f90:343-350
    CrustProfile p = crust2x2_get_profile(c, (float)(r2d * olat_rad), (float)(r2d * olon_rad));
    float vp, vs, rho, thickness;
    crust2x2_get_profile_averages(p, &vp, &vs, &rho, &thickness);
    if (thickness_limit > 0.f) thickness = std::min(thickness_limit, thickness);
    out->assign(2, Halfspace());
    (*out)[0] = Halfspace{{0.f, 0.f, 1500.f}, {0.f, 0.f, -1.f}};
    (*out)[1] = Halfspace{{0.f, 0.f, thickness}, {0.f, 0.f, 1.f}};
}

// ---- geometry.f90 -------------------------------------------------------------------------------------
typedef float V3[3];
static inline float dot3(const float* a, const float* b) { float s = 0.f; for (int i = 0; i < 3; i++) s = s + a[i] * b[i]; return s; }

static bool point_in_halfspace(const float* p, const Halfspace& h) {   // :57-71
    float d[3] = {h.point[0] - p[0], h.point[1] - p[1], h.point[2] - p[2]};
    return dot3(h.normal, d) >= 0.0f;
}
static bool point_in_constraints(const std::vector<Halfspace>& cs, const float* p) {   // parameterized_source.f90:168-181
    for (const Halfspace& h : cs) if (!point_in_halfspace(p, h)) return false;
    return true;
}
// :73-125
static void get_piercingpoint(const float* a, const float* b, const Halfspace& h, float* pp, bool* between_ab, bool* parallel, bool* a_inside_,
                              bool* b_inside_) {
    float ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    float da[3] = {h.point[0] - a[0], h.point[1] - a[1], h.point[2] - a[2]};
    float db[3] = {h.point[0] - b[0], h.point[1] - b[1], h.point[2] - b[2]};
    const float lambda_a = dot3(h.normal, da), lambda_b = dot3(h.normal, db), lambda_ab = dot3(h.normal, ab);
    const bool a_inside = lambda_a >= 0.f, b_inside = lambda_b >= 0.f;
    if (a_inside_) *a_inside_ = a_inside;
    if (b_inside_) *b_inside_ = b_inside;
    *between_ab = (a_inside && !b_inside) || (b_inside && !a_inside);
    *parallel = lambda_ab * lambda_ab < dot3(ab, ab) / 16777216.f;   // 2**digits(real) = 2**24
    if (*parallel && *between_ab) {
        const float* s = fabsf(lambda_a) <= fabsf(lambda_b) ? a : b;
        pp[0] = s[0]; pp[1] = s[1]; pp[2] = s[2];
        return;
    }
    if (*parallel && !*between_ab) { pp[0] = 0.f; pp[1] = 0.f; pp[2] = 0.f; return; }
    for (int i = 0; i < 3; i++) pp[i] = a[i] + ab[i] * lambda_a / lambda_ab;
}
typedef std::vector<float> Polygon;   // 3 floats per point
// :213-257
static void trim_polygon_one(const Polygon& poly, const Halfspace& h, Polygon* out) {
    const int n = (int)poly.size() / 3;
    std::vector<float> pierce((size_t)3 * std::max(n, 1));
    std::vector<char> does(n), inside(n);
    for (int i = 0; i < n; i++) {
        const int j = (i + 1) % n;
        bool d, par, ai, bi;
        get_piercingpoint(&poly[3 * i], &poly[3 * j], h, &pierce[3 * i], &d, &par, &ai, &bi);
        does[i] = d; inside[i] = ai;
    }
    out->clear();
    for (int i = 0; i < n; i++) {
        if (inside[i]) out->insert(out->end(), &poly[3 * i], &poly[3 * i] + 3);
        if (does[i]) out->insert(out->end(), &pierce[3 * i], &pierce[3 * i] + 3);
    }
}
// :259-276
static void trim_polygon_more(const Polygon& poly, const std::vector<Halfspace>& hs, Polygon* out) {
    Polygon temp = poly;
    for (size_t k = 0; k < hs.size(); k++) {
        if (k != 0) temp = *out;
        trim_polygon_one(temp, hs[k], out);
    }
    if (hs.empty()) *out = poly;   // zero-trip loop leaves the result unallocated in the reference; an untrimmed polygon is what is meant
}

// ---- heap.f90: index heap with back-pointers (1-based indices as in the reference) ----------------------
// The heap holds (key, index) pairs instead of indices into the key array: every comparison of heap.f90 reads the same values, so the
// pop order -- which the solver's result depends on, see below -- is the reference's, but the sift loops stay in one array.
struct HeapItem { float key; int idx; };
struct IndexHeap {
    std::vector<HeapItem> item;   // item[1..n]
    int n = 0, cap = 0;
    void init(int maxsize) { item.assign((size_t)maxsize + 1, HeapItem{0.f, 0}); n = 0; cap = maxsize; }
};
// (the sift loops move a hole instead of swapping: the same comparisons and the same final arrangement as heap.f90's swaps, half the writes)
static inline void upheap(IndexHeap& h, int element, int* bp) {   // :210-232
    HeapItem* a = h.item.data();
    int v = element;
    const HeapItem x = a[v];
    bool moved = false;
    while (v > 1) {
        const int u = (v - 2) / 2 + 1;
        if (a[u].key <= x.key) break;
        a[v] = a[u]; bp[a[v].idx] = v;
        v = u; moved = true;
    }
    if (moved) { a[v] = x; bp[x.idx] = v; }
}
static inline void downheap(IndexHeap& h, int element, int* bp) {   // :176-208
    HeapItem* a = h.item.data();
    int v = element;
    const HeapItem x = a[v];
    bool moved = false;
    int w = 2 * (v - 1) + 2;
    while (w <= h.n) {
        if (w + 1 <= h.n && a[w + 1].key < a[w].key) w = w + 1;
        if (x.key <= a[w].key) break;
        a[v] = a[w]; bp[a[v].idx] = v;
        v = w; moved = true;
        w = 2 * (v - 1) + 2;
    }
    if (moved) { a[v] = x; bp[x.idx] = v; }
}
static inline void pushheap(IndexHeap& h, int keyindex, const float* keys, int* bp) {   // :70-93
    if (h.n + 1 > h.cap) return;
    h.n = h.n + 1;
    h.item[h.n] = HeapItem{keys[keyindex], keyindex};
    bp[keyindex] = h.n;
    upheap(h, h.n, bp);
}
static inline int popheap(IndexHeap& h, int* bp) {   // :95-124
    if (h.n == 0) return 0;
    std::swap(h.item[1], h.item[h.n]);
    bp[h.item[1].idx] = 1;
    const int keyindex = h.item[h.n].idx;
    bp[keyindex] = 0;
    h.n = h.n - 1;
    downheap(h, 1, bp);
    return keyindex;
}
static inline void updateheap(IndexHeap& h, int keyindex, float newkey, float* keys, int* bp) {   // :126-150
    const float oldkey = keys[keyindex];
    keys[keyindex] = newkey;
    h.item[bp[keyindex]].key = newkey;
    if (newkey < oldkey) upheap(h, bp[keyindex], bp);
    if (newkey > oldkey) downheap(h, bp[keyindex], bp);
}

// ---- eikonal.f90:29-199: fast marching; arrays are (ix,iy) column-major, 1-based linear index i=(iy-1)*nx+ix ----
// Neighbour times enter an update whether they are final or still tentative (:151-155), so the result depends on the order in which the
// heap hands out equal and near-equal keys: the solver cannot be reordered or parallelised without changing the sub-source table, and
// runs per candidate on the host.  The loop-invariant products of :163-166 are formed once; they are the same operations on the same
// operands.
void eikonal_solver_fmm(const float* speed, int nx, int ny, const float origin[2], const float delta[2], const float initialpoint[2],
                        float* times) {
    const int FARAWAY = -1, ALIVE = 0;
    const float infinity = std::numeric_limits<float>::max() * 0.1f;
    const float dx = delta[0], dy = delta[1];
    const float dx2 = dx * dx, dy2 = dy * dy, dx2dy2 = dx2 * dy2, dx2pdy2 = dx2 + dy2;
    std::vector<int> bpv((size_t)nx * ny + 1, FARAWAY);
    int* bp = bpv.data();               // bp[1..nx*ny]
    float* T = times - 1;               // T[1..nx*ny]
    const float* S = speed - 1;
    auto ind = [nx](int ix, int iy) { return (iy - 1) * nx + ix; };
    int ix = (int)((initialpoint[0] - origin[0]) / dx) + 1;
    int iy = (int)((initialpoint[1] - origin[1]) / dy) + 1;
    if (ix < 1) ix = 1;
    if (nx < ix) ix = nx;
    if (iy < 1) iy = 1;
    if (ny < iy) iy = ny;
    for (int i = 1; i <= nx * ny; i++) T[i] = infinity;
    T[ind(ix, iy)] = 0.0f;
    if (nx == 1 && ny == 1) return;
    bp[ind(ix, iy)] = ALIVE;
    int nalive = 1;
    IndexHeap heap;
    heap.init(nx * ny);
    if (1 < ix) T[ind(ix - 1, iy)] = dx / S[ind(ix - 1, iy)];
    if (ix < nx) T[ind(ix + 1, iy)] = dx / S[ind(ix + 1, iy)];
    if (1 < iy) T[ind(ix, iy - 1)] = dy / S[ind(ix, iy - 1)];
    if (iy < ny) T[ind(ix, iy + 1)] = dy / S[ind(ix, iy + 1)];
    if (1 < ix) pushheap(heap, ind(ix - 1, iy), T, bp);
    if (ix < nx) pushheap(heap, ind(ix + 1, iy), T, bp);
    if (1 < iy) pushheap(heap, ind(ix, iy - 1), T, bp);
    if (iy < ny) pushheap(heap, ind(ix, iy + 1), T, bp);

    auto update_neighbor = [&](int jx, int jy, int i) {   // :134-190, i = ind(jx, jy)
        if (bp[i] == ALIVE) return;
        if (bp[i] == FARAWAY) pushheap(heap, i, T, bp);
        float a = infinity, b = infinity, c = infinity, d = infinity;
        const float told = T[i];
        if (1 < jx) a = T[i - 1];
        if (jx < nx) b = T[i + 1];
        if (1 < jy) c = T[i - nx];
        if (jy < ny) d = T[i + nx];
        float t = 0.f;
        const float aa = std::min(a, b), cc = std::min(c, d);
        const float sp = S[i];
        if (std::max(aa, cc) != infinity) {
            const float q = (aa - cc) * sp;
            const float s = dx2dy2 * (dx2pdy2 - q * q);
            if (s >= 0.f) t = std::max(t, ((aa * dy2 + cc * dx2) * sp + sqrtf(s)) / (sp * dx2pdy2));
        }
        if (cc == infinity) {
            if (a < infinity) t = std::max(t, a + dx / sp);
            if (b < infinity) t = std::max(t, b + dx / sp);
        }
        if (aa == infinity) {
            if (c < infinity) t = std::max(t, c + dy / sp);
            if (d < infinity) t = std::max(t, d + dy / sp);
        }
        if (t == 0.f) {   // fallback condition
            t = infinity;
            if (a < infinity) t = std::min(t, a + dx / sp);
            if (b < infinity) t = std::min(t, b + dx / sp);
            if (c < infinity) t = std::min(t, c + dy / sp);
            if (d < infinity) t = std::min(t, d + dy / sp);
        }
        if (t != 0.f && told != t) updateheap(heap, i, t, T, bp);
    };
    while (nalive <= nx * ny) {
        const int imin = popheap(heap, bp);
        if (imin == 0) break;
        iy = (imin - 1) / nx + 1;
        ix = imin - (iy - 1) * nx;
        bp[imin] = ALIVE;
        nalive = nalive + 1;
        if (1 < ix) update_neighbor(ix - 1, iy, imin - 1);
        if (ix < nx) update_neighbor(ix + 1, iy, imin + 1);
        if (1 < iy) update_neighbor(ix, iy - 1, imin - nx);
        if (iy < ny) update_neighbor(ix, iy + 1, imin + nx);
    }
}

// ---- source_eikonal.f90 ---------------------------------------------------------------------------------
namespace {
struct Psm {
    const float* p;          // parameters, 0-based
    float rot_rup[9];        // row-major
    int i_bsx, i_bsy, i_brad, i_nsx, i_nsy, i_relv;
};
inline void rc_to_ned(const Psm& s, const float* rc, float* out) {   // :612-617
    for (int i = 0; i < 3; i++) {
        float a = 0.f;
        for (int j = 0; j < 3; j++) a = a + s.rot_rup[i * 3 + j] * rc[j];
        out[i] = a + s.p[1 + i];
    }
}
inline void ned_to_rc(const Psm& s, const float* pt, float* out) {   // :605-610
    const float d[3] = {pt[0] - s.p[1], pt[1] - s.p[2], pt[2] - s.p[3]};
    for (int i = 0; i < 3; i++) {
        float a = 0.f;
        for (int j = 0; j < 3; j++) a = a + s.rot_rup[j * 3 + i] * d[j];
        out[i] = a;
    }
}
// discretize_subfault_time :714-764
void discretize_subfault_time(float duration_subfault, float risetime, float maxdt, std::vector<float>& tw, std::vector<float>& toff, int* nt_) {
    const float dursf = duration_subfault;
    const float durfull = dursf + risetime;
    const int nt = (int)floorf(durfull / maxdt) + 1;
    *nt_ = nt;
    if ((int)tw.size() < nt) tw.resize(nt);
    if ((int)toff.size() < nt) toff.resize(nt);
    if (nt == 1) { tw[0] = 1.f; toff[0] = 0.f; return; }
    float sx[4], sy[4];
    if (risetime < dursf) {
        sx[0] = (-dursf - risetime) / 2.f; sx[1] = (-dursf + risetime) / 2.f; sx[2] = (dursf - risetime) / 2.f; sx[3] = (dursf + risetime) / 2.f;
        sy[0] = 0.f; sy[1] = 1.f / dursf; sy[2] = 1.f / dursf; sy[3] = 0.f;
    } else {
        sx[0] = (-risetime - dursf) / 2.f; sx[1] = (-risetime + dursf) / 2.f; sx[2] = (risetime - dursf) / 2.f; sx[3] = (risetime + dursf) / 2.f;
        sy[0] = 0.f; sy[1] = 1.f / risetime; sy[2] = 1.f / risetime; sy[3] = 0.f;
    }
    const float tbeg = sx[0];
    const float dt = durfull / (float)nt;
    for (int it = 1; it <= nt; it++) {
        const float ta = tbeg + dt * (float)(it - 1);
        const float tb = tbeg + dt * (float)it;
        plf_integrate_and_centroid(sx, sy, 4, ta, tb, &tw[it - 1], &toff[it - 1]);
    }
}
}  // namespace

// The discretiser in three steps, so that the fast-marching solves of a large batch can run on the device (csrc/eikonal.cu) between the
// two host parts: prep_eikonal_begin (geometry, speed field) -> times of the fine grid -> prep_eikonal_finish (down-sampling, table).
static Psm psm_of(const EikonalWork& w) {
    Psm s;
    s.p = w.p; memcpy(s.rot_rup, w.rot_rup, sizeof s.rot_rup);
    s.i_bsx = w.idx[0]; s.i_bsy = w.idx[1]; s.i_brad = w.idx[2]; s.i_nsx = w.idx[3]; s.i_nsy = w.idx[4]; s.i_relv = w.idx[5];
    return s;
}

bool prep_eikonal_setup(const float* p, bool mt_variant, float shortest_doi, double olat_rad, double olon_rad, const Crust2x2& crust,
                        const std::vector<Halfspace>& constraints, EikonalWork* work, EikonalPrep* out) {
    EikonalPrep& o = *out;
    o = EikonalPrep();
    EikonalWork& w = *work;
    w = EikonalWork();
    if (!crust.loaded) { o.err = "crust2x2 model not loaded"; return false; }
    Psm s;
    s.p = p;
    // parameter layout: eikonal (15) has the rake at index 7, mt_eikonal (20) carries the tensor instead
    if (!mt_variant) { s.i_bsx = 8; s.i_bsy = 9; s.i_brad = 10; s.i_nsx = 11; s.i_nsy = 12; s.i_relv = 13; o.risetime = p[14]; }
    else { s.i_bsx = 7; s.i_bsy = 8; s.i_brad = 9; s.i_nsx = 10; s.i_nsy = 11; s.i_relv = 12; o.risetime = p[19]; }
    o.moment = p[4];
    for (int i = 0; i < (mt_variant ? 20 : 15); i++) if (!std::isfinite(p[i])) { o.err = "non-finite source parameter"; return false; }
    // psm_update_dep_params :233-257
    const float strike = d2r_r(p[5]), dip = d2r_r(p[6]);
    float* rot_slip = w.rot_slip;
    init_euler(dip, strike, 0.f, s.rot_rup);
    if (!mt_variant) { const float rake = d2r_r(p[7]); init_euler(dip, strike, -rake, rot_slip); }
    const float bord_shift_x = p[s.i_bsx], bord_shift_y = p[s.i_bsy], bord_radius = p[s.i_brad];
    // ---- psm_borderline :318-348 ---------------------------------------------------------------------
    float center[3];
    { const float rc[3] = {bord_shift_x, bord_shift_y, 0.f}; rc_to_ned(s, rc, center); }
    float transform[9];
    for (int i = 0; i < 9; i++) transform[i] = -s.rot_rup[i] * bord_radius;
    int npts = 180;
    if (bord_radius == 0.f) npts = 1;
    Polygon circle((size_t)3 * npts);
    for (int i = 1; i <= npts; i++) {   // circle_to_polygon geometry.f90:191-211
        const float ang = (float)i * 2.f * pi_f / (float)npts;
        const float v[3] = {cosf(ang), sinf(ang), 0.f};
        for (int r = 0; r < 3; r++) {
            float a = 0.f;
            for (int j = 0; j < 3; j++) a = a + transform[r * 3 + j] * v[j];
            circle[3 * (i - 1) + r] = a + center[r];
        }
    }
    Polygon rupture_poly;
    trim_polygon_more(circle, constraints, &rupture_poly);
    const int np = (int)rupture_poly.size() / 3;
    if (np == 0) { o.err = "Empty rupture area"; return false; }   // :285-289
    // polygon_box of the polygon in rupture coordinates
    float min_rc[3] = {0, 0, 0}, max_rc[3] = {0, 0, 0};
    for (int i = 0; i < np; i++) {
        float rc[3];
        ned_to_rc(s, &rupture_poly[3 * i], rc);
        for (int k = 0; k < 3; k++) {
            if (i == 0 || rc[k] < min_rc[k]) min_rc[k] = rc[k];
            if (i == 0 || rc[k] > max_rc[k]) max_rc[k] = rc[k];
        }
    }
    const float deltagrid = std::min(100.f * shortest_doi / 2.f, 4000.f);
    // ---- psm_make_eikonal_grid :435-517 ------------------------------------------------------------------
    const float rel_rupture_velocity = p[s.i_relv];
    float first[2] = {min_rc[0], min_rc[1]}, last[2] = {max_rc[0], max_rc[1]};
    float dims[2] = {last[0] - first[0], last[1] - first[1]};
    int nd[2] = {(int)ceilf(dims[0] / deltagrid), (int)ceilf(dims[1] / deltagrid)};
    if (nd[0] == 0) nd[0] = 1;
    if (nd[1] == 0) nd[1] = 1;
    if (nd[0] < 0 || nd[1] < 0 || (long long)nd[0] * nd[1] > 4000000LL) { o.err = "eikonal grid too large"; return false; }
    float delta[2] = {dims[0] / (float)nd[0], dims[1] / (float)nd[1]};
    const int fnx = nd[0], fny = nd[1];
    // crust2x2_get_profile(psm%origin): the origin is in RADIANS here (source_eikonal.f90:472), kept as is
    const CrustProfile profile = crust2x2_get_profile(crust, (float)olat_rad, (float)olon_rad);
    // psm_initial_point_intolerant_rc :401-432
    const float nukl_shift_x = p[s.i_nsx], nukl_shift_y = p[s.i_nsy];
    const float nukl_shift = sqrtf(nukl_shift_x * nukl_shift_x + nukl_shift_y * nukl_shift_y);
    {
        const float rc[3] = {nukl_shift_x, nukl_shift_y, 0.f};
        float ned[3];
        rc_to_ned(s, rc, ned);
        if (!point_in_constraints(constraints, ned) || nukl_shift > bord_radius) {
            o.err = "position of nucleation point is outside of rupture region";
            return false;
        }
    }
    const float initialpoint[2] = {nukl_shift_x, nukl_shift_y};
    w.p = p; memcpy(w.rot_rup, s.rot_rup, sizeof w.rot_rup);
    w.idx[0] = s.i_bsx; w.idx[1] = s.i_bsy; w.idx[2] = s.i_brad; w.idx[3] = s.i_nsx; w.idx[4] = s.i_nsy; w.idx[5] = s.i_relv;
    w.mt_variant = mt_variant; w.shortest_doi = shortest_doi;
    for (int k = 0; k < 2; k++) { w.first[k] = first[k]; w.last[k] = last[k]; w.delta[k] = delta[k]; w.initialpoint[k] = initialpoint[k]; }
    w.fnx = fnx; w.fny = fny;
    memcpy(w.center, center, sizeof w.center); w.bord_radius = bord_radius; w.relv = rel_rupture_velocity;
    w.profile = profile; w.constraints = &constraints;
    return true;
}

// psm_make_eikonal_grid :435-517, the loop over the fine grid: positions, rupture speed, slow rim
bool prep_eikonal_speed_host(EikonalWork* work, EikonalPrep* out) {
    EikonalPrep& o = *out;
    EikonalWork& w = *work;
    const Psm s = psm_of(w);
    const int fnx = w.fnx, fny = w.fny;
    const float* first = w.first; const float* delta = w.delta; const float* center = w.center;
    const float bord_radius = w.bord_radius, rel_rupture_velocity = w.relv;
    const std::vector<Halfspace>& constraints = *w.constraints;
    const CrustProfile& profile = w.profile;
    std::vector<float>&speed = w.speed, &points = w.points;
    speed.assign((size_t)fnx * fny, 0.f); points.assign((size_t)3 * fnx * fny, 0.f);
    float minspeed = std::numeric_limits<float>::max();
    for (int iy = 1; iy <= fny; iy++)
        for (int ix = 1; ix <= fnx; ix++) {
            const float rc[3] = {first[0] + ((float)ix - 0.5f) * delta[0], first[1] + ((float)iy - 0.5f) * delta[1], 0.f};
            float pt[3];
            rc_to_ned(s, rc, pt);
            const size_t c = (size_t)(iy - 1) * fnx + (ix - 1);
            points[3 * c] = pt[0]; points[3 * c + 1] = pt[1]; points[3 * c + 2] = pt[2];
            const float d[3] = {pt[0] - center[0], pt[1] - center[1], pt[2] - center[2]};
            if (sqrtf(dot3(d, d)) > bord_radius || !point_in_constraints(constraints, pt)) speed[c] = 0.f;
            else {
                float vp, vs, rho;
                crust2x2_get_at_depth(profile, pt[2], &vp, &vs, &rho);
                speed[c] = vs * rel_rupture_velocity;
                minspeed = std::min(speed[c], minspeed);
            }
        }
    const float invalid_speed = minspeed * 0.5f;
    for (float& v : speed) if (v == 0.f) v = invalid_speed;
    if (!(minspeed > 0.f) || minspeed == std::numeric_limits<float>::max()) { o.err = "no valid point in the rupture area"; return false; }
    w.minspeed = minspeed; w.invalid_speed = invalid_speed;
    return true;
}

bool prep_eikonal_begin(const float* p, bool mt_variant, float shortest_doi, double olat_rad, double olon_rad, const Crust2x2& crust,
                        const std::vector<Halfspace>& constraints, EikonalWork* work, EikonalPrep* out) {
    return prep_eikonal_setup(p, mt_variant, shortest_doi, olat_rad, olon_rad, crust, constraints, work, out) && prep_eikonal_speed_host(work, out);
}

// crust2x2_get_at_depth as a table for the device (csrc/eikonal.cu): cumulative thickness of the layers it walks and their vs
void eikonal_layer_table(const CrustProfile& p, float thr[5], float vs[6]) {
    float d = 0.f;
    for (int i = 2; i < NLAYERS; i++) { d = d + p.thickness[i]; thr[i - 2] = d; vs[i - 2] = p.vs[i]; }
    vs[5] = p.vs[LBELOWCRUST];
}

void prep_eikonal_solve_host(EikonalWork* work) {
    EikonalWork& w = *work;
    w.times.assign(w.speed.size(), 0.f);
    eikonal_solver_fmm(w.speed.data(), w.fnx, w.fny, w.first, w.delta, w.initialpoint, w.times.data());
}

// coarse grid size source_eikonal.f90:274-277, 617-638
bool prep_eikonal_coarse_dims(const EikonalWork& w, EikonalCoarse* cg, std::string* err) {
    const float* first = w.first; const float* last = w.last;
    const float maxdx = 0.5f * w.shortest_doi * w.minspeed, maxdy = 0.5f * w.shortest_doi * w.minspeed;
    const float sizex = last[0] - first[0], sizey = last[1] - first[1];
    const float fx = sizex / maxdx, fy = sizey / maxdy;
    if (!(fabsf(fx) < 1e5f) || !(fabsf(fy) < 1e5f)) { *err = "sub-fault grid too large"; return false; }
    int nxc = (int)floorf(fx) + 1;
    if (nxc <= 1) nxc = 2;
    if (sizex == 0.f) nxc = 1;
    int nyc = (int)floorf(fy) + 1;
    if (nyc <= 1) nyc = 2;
    if (sizey == 0.f) nyc = 1;
    cg->nxc = nxc; cg->nyc = nyc;
    cg->cdelta[0] = (last[0] - first[0]) / (float)nxc; cg->cdelta[1] = (last[1] - first[1]) / (float)nyc;
    if (cg->cdelta[0] == 0.f || nxc == 0) cg->cdelta[0] = 1.f;
    if (cg->cdelta[1] == 0.f || nyc == 0) cg->cdelta[1] = 1.f;
    return true;
}

// psm_downsample_grid :519-601 (sums per coarse cell in fine-grid order); times < 0 = outside the rupture area
void prep_eikonal_downsample_host(EikonalWork* work, EikonalCoarse* cgp) {
    EikonalWork& w = *work;
    EikonalCoarse& cg = *cgp;
    const Psm s = psm_of(w);
    const float* first = w.first;
    const float* cdelta = cg.cdelta;
    const int nxc = cg.nxc, nyc = cg.nyc;
    std::vector<float>&speed = w.speed, &times = w.times, &points = w.points;
    for (size_t c = 0; c < speed.size(); c++) if (speed[c] == w.invalid_speed) times[c] = -1.f;
    const size_t nc = (size_t)nxc * nyc;
    std::vector<float>&ntimes = cg.ntimes, &ctimes = cg.ctimes, &cpoints = cg.cpoints, &cdur = cg.cdur;
    ntimes.assign(nc, 0.f); ctimes.assign(nc, -1.f); cpoints.assign(3 * nc, 0.f); cdur.assign(nc, 0.f);
    std::vector<float> cspeed(nc, 0.f);
    auto coarse_cell = [&](size_t c, int* ixc, int* iyc) {
        float rc[3];
        ned_to_rc(s, &points[3 * c], rc);
        *ixc = (int)floorf((rc[0] - first[0]) / cdelta[0]) + 1;
        *iyc = (int)floorf((rc[1] - first[1]) / cdelta[1]) + 1;
        return !(*ixc < 1 || *iyc < 1 || *ixc > nxc || *iyc > nyc);   // else: "orphaned point in fine grid"
    };
    std::vector<int> cell_of(speed.size(), -1);   // coarse cell of every valid fine point (the second pass below needs it again)
    for (size_t c = 0; c < speed.size(); c++) {   // iyf outer, ixf inner = linear order
        if (times[c] < 0.f) continue;
        int ixc, iyc;
        if (!coarse_cell(c, &ixc, &iyc)) continue;
        const size_t k = (size_t)(iyc - 1) * nxc + (ixc - 1);
        cell_of[c] = (int)k;
        ntimes[k] = ntimes[k] + 1.f;
        if (ctimes[k] == -1.f) ctimes[k] = 0.f;
        ctimes[k] = ctimes[k] + times[c];
        cspeed[k] = cspeed[k] + 1.f / speed[c];
        for (int q = 0; q < 3; q++) cpoints[3 * k + q] = cpoints[3 * k + q] + points[3 * c + q];
    }
    for (size_t k = 0; k < nc; k++)
        if (ntimes[k] > 0.f) {
            ctimes[k] = 1.f / ntimes[k] * ctimes[k];
            cspeed[k] = 1.f / (1.f / ntimes[k] * cspeed[k]);
            for (int q = 0; q < 3; q++) cpoints[3 * k + q] = 1.f / ntimes[k] * cpoints[3 * k + q];
        }
    for (size_t c = 0; c < speed.size(); c++) {
        if (cell_of[c] < 0) continue;
        const size_t k = (size_t)cell_of[c];
        cdur[k] = cdur[k] + fabsf(times[c] - ctimes[k]);
    }
    for (size_t k = 0; k < nc; k++) if (ntimes[k] > 0.f) cdur[k] = 4.f / ntimes[k] * cdur[k];
}

// weights of the sub-faults and psm_to_tdsm_table_eikonal :640-712, from the down-sampled grid
bool prep_eikonal_table(const EikonalWork& w, const EikonalCoarse& cg, EikonalPrep* out) {
    EikonalPrep& o = *out;
    const float* p = w.p;
    const bool mt_variant = w.mt_variant;
    const float* rot_slip = w.rot_slip;
    const float maxdt = w.shortest_doi;
    const int nxc = cg.nxc, nyc = cg.nyc;
    const size_t nc = (size_t)nxc * nyc;
    const std::vector<float>&ntimes = cg.ntimes, &ctimes = cg.ctimes, &cpoints = cg.cpoints, &cdur = cg.cdur;
    int npf = 0;
    for (size_t k = 0; k < nc; k++) npf = npf + (int)ntimes[k];   // (the reference counts the fine points as it goes: an exact integer either way)
    std::vector<float> cweights(nc, 0.f);
    for (size_t k = 0; k < nc; k++) cweights[k] = ntimes[k] / (float)npf;
    // ---- psm_to_tdsm_table_eikonal :640-712 ----------------------------------------------------------------------
    const float origin_time = p[0];
    float centertime = 0.f;
    for (size_t k = 0; k < nc; k++)
        if (ctimes[k] >= 0.f) centertime = centertime + ctimes[k] * cweights[k];
    if (!mt_variant) {
        const float m_unrot[9] = {0, 0, -1, 0, 0, 0, -1, 0, 0};
        float trot[9], tmp[9], m_rot[9];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) trot[i * 3 + j] = rot_slip[j * 3 + i];
        auto matmul3 = [](const float* a, const float* b, float* c) {
            for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) {
                float v = 0.f;
                for (int j = 0; j < 3; j++) v = v + a[i * 3 + j] * b[j * 3 + k];
                c[i * 3 + k] = v;
            }
        };
        matmul3(m_unrot, trot, tmp);
        matmul3(rot_slip, tmp, m_rot);
        o.mhat[0] = m_rot[0]; o.mhat[1] = m_rot[4]; o.mhat[2] = m_rot[8]; o.mhat[3] = m_rot[1]; o.mhat[4] = m_rot[2]; o.mhat[5] = m_rot[5];
    } else {
        for (int i = 0; i < 6; i++) o.mhat[i] = p[13 + i];   // source_mt_eikonal.f90:697-702
    }
    std::vector<float> tw, toff;
    for (size_t k = 0; k < nc; k++) {   // iy outer, ix inner
        if (ctimes[k] < 0.f) continue;
        int nt;
        discretize_subfault_time(cdur[k], 0.f, maxdt, tw, toff, &nt);
        if (nt < 1 || nt > 32) { o.err = "too many time centroids in a sub-fault"; return false; }
        EikonalGroup g;
        g.north = cpoints[3 * k]; g.east = cpoints[3 * k + 1]; g.depth = cpoints[3 * k + 2];
        g.gw = cweights[k];
        g.tap_begin = (int)o.tap_time.size(); g.tap_count = nt;
        for (int it = 0; it < nt; it++) {
            o.tap_time.push_back(ctimes[k] + toff[it] + origin_time - centertime);   // :695
            o.tap_wt.push_back(tw[it]);
        }
        o.groups.push_back(g);
    }
    o.nx = nxc; o.ny = nyc;
    if (o.groups.empty()) { o.err = "Empty rupture area"; return false; }
    return true;
}

bool prep_eikonal_finish(EikonalWork* work, EikonalPrep* out) {
    EikonalCoarse cg;
    if (!prep_eikonal_coarse_dims(*work, &cg, &out->err)) return false;
    prep_eikonal_downsample_host(work, &cg);
    return prep_eikonal_table(*work, cg, out);
}

bool prep_eikonal(const float* p, bool mt_variant, float shortest_doi, double olat_rad, double olon_rad, const Crust2x2& crust,
                  const std::vector<Halfspace>& constraints, EikonalPrep* out) {
    EikonalWork w;
    if (!prep_eikonal_begin(p, mt_variant, shortest_doi, olat_rad, olon_rad, crust, constraints, &w, out)) return false;
    prep_eikonal_solve_host(&w);
    return prep_eikonal_finish(&w, out);
}

}  // namespace kh
