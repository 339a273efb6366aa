// Host-side Green's function database: container, flat binary file "KGF1", and the analytical
// full-space builder.  Replaces (for this engine) the storage half of gfdb.f90 and the tools
// gfdb_build / gfdb_build_ahfull of the reference; citations are file:line of /root/reference.
//
// Built with -ffp-contract=off: the builder's fp32 arithmetic follows elseis.f90 statement by
// statement so that the synthetic database is the one the reference's tool would write.
#include "kiwi_internal.hpp"
#include <cmath>
#include <cstring>
#include <cstdint>
#include <algorithm>
#include <thread>
#include <atomic>

static thread_local std::string g_errstr;

int kiwi_set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_errstr = buf;
    return 1;
}
extern "C" const char* kiwi_last_error(void) { return g_errstr.c_str(); }
extern "C" const char* kiwi_version(void) { return "kiwi_b200 0.1 (sm_100a)"; }

void kiwi_pack_window(const float* data, int n, int* first, int* last) {
    int i0 = -1, i1 = -1;
    for (int i = 0; i < n; i++) if (data[i] != 0.f) { if (i0 < 0) i0 = i; i1 = i; }
    if (i0 < 0) { *first = 0; *last = 0; return; }       // no data: one zero sample (sparse_trace.f90:493-512)
    if (i1 < n - 1) i1 = i1 + 1;                          // keep one of the trailing zeros (:539-547)
    *first = i0; *last = i1;
}

void kiwi_gfdb::flatten() {
    if (flat) return;
    size_t n = ntr();
    long long total = 0;
    for (size_t i = 0; i < n; i++) { offset[i] = total; total += len[i]; }
    data.resize((size_t)total);
    for (size_t i = 0; i < n; i++) {
        if (len[i] > 0) memcpy(&data[(size_t)offset[i]], pending[i].data(), sizeof(float) * len[i]);
        std::vector<float>().swap(pending[i]);
    }
    std::vector<std::vector<float>>().swap(pending);
    flat = true;
}

extern "C" {

kiwi_gfdb* kiwi_gfdb_create(int nx, int nz, int ng, float dt, float dx, float dz, float firstx, float firstz) {
    if (nx < 1 || nz < 1 || (ng != 8 && ng != 10) || !(dt > 0.f) || !(dx > 0.f) || !(dz > 0.f)) {
        kiwi_set_error("kiwi_gfdb_create: invalid grid (nx=%d nz=%d ng=%d dt=%g dx=%g dz=%g)", nx, nz, ng, dt, dx, dz);
        return nullptr;
    }
    kiwi_gfdb* db = new kiwi_gfdb();
    db->nx = nx; db->nz = nz; db->ng = ng; db->dt = dt; db->dx = dx; db->dz = dz; db->firstx = firstx; db->firstz = firstz;
    size_t n = db->ntr();
    db->span0.assign(n, 0); db->len.assign(n, 0); db->offset.assign(n, 0);
    db->pending.assign(n, {}); db->flat = false;
    return db;
}
void kiwi_gfdb_destroy(kiwi_gfdb* db) { delete db; }

int kiwi_gfdb_save_array(kiwi_gfdb* db, int ix, int iz, int ig, int span0, int n, const float* data) {
    if (!db) return kiwi_set_error("kiwi_gfdb_save_array: null database");
    if (ix < 1 || ix > db->nx || iz < 1 || iz > db->nz || ig < 1 || ig > db->ng)
        return kiwi_set_error("gfdb: invalid request: out of bounds: (%d,%d,%d)", ix, iz, ig);
    if (n < 1) return kiwi_set_error("kiwi_gfdb_save_array: empty trace");
    if (db->flat) {  // reopen for filling
        size_t nt = db->ntr();
        db->pending.assign(nt, {});
        for (size_t i = 0; i < nt; i++) if (db->len[i] > 0) db->pending[i].assign(&db->data[(size_t)db->offset[i]], &db->data[(size_t)db->offset[i]] + db->len[i]);
        std::vector<float>().swap(db->data);
        db->flat = false;
    }
    int f, l;
    kiwi_pack_window(data, n, &f, &l);
    size_t i = db->idx(ix, iz, ig);
    db->pending[i].assign(data + f, data + l + 1);
    db->span0[i] = span0 + f;
    db->len[i] = l - f + 1;
    return 0;
}

int kiwi_gfdb_meta(const kiwi_gfdb* db, int* nx, int* nz, int* ng, float* dt, float* dx, float* dz, float* firstx,
                   float* firstz, long long* ntraces, long long* nsamples) {
    if (!db) return kiwi_set_error("kiwi_gfdb_meta: null database");
    if (nx) *nx = db->nx; if (nz) *nz = db->nz; if (ng) *ng = db->ng;
    if (dt) *dt = db->dt; if (dx) *dx = db->dx; if (dz) *dz = db->dz; if (firstx) *firstx = db->firstx; if (firstz) *firstz = db->firstz;
    long long nt = 0, ns = 0;
    for (size_t i = 0; i < db->ntr(); i++) if (db->len[i] > 0) { nt++; ns += db->len[i]; }
    if (ntraces) *ntraces = nt; if (nsamples) *nsamples = ns;
    return 0;
}

int kiwi_gfdb_view(kiwi_gfdb* db, const int** span0, const int** len, const long long** offset, const float** data) {
    if (!db) return kiwi_set_error("kiwi_gfdb_view: null database");
    db->flatten();
    *span0 = db->span0.data(); *len = db->len.data(); *offset = db->offset.data(); *data = db->data.data();
    return 0;
}

// ---- file format KGF1: header | span0[ntr] | len[ntr] | offset[ntr] | data[nsamples] ---------
struct Kgf1Header { char magic[8]; int32_t nx, nz, ng, pad; float dt, dx, dz, firstx, firstz, padf; int64_t nsamples; };

int kiwi_gfdb_write(const kiwi_gfdb* cdb, const char* path) {
    kiwi_gfdb* db = const_cast<kiwi_gfdb*>(cdb);
    if (!db) return kiwi_set_error("kiwi_gfdb_write: null database");
    db->flatten();
    FILE* f = fopen(path, "wb");
    if (!f) return kiwi_set_error("can't open file %s", path);
    Kgf1Header h; memset(&h, 0, sizeof h);
    memcpy(h.magic, "KGF1\0\0\0\0", 8);
    h.nx = db->nx; h.nz = db->nz; h.ng = db->ng; h.dt = db->dt; h.dx = db->dx; h.dz = db->dz; h.firstx = db->firstx; h.firstz = db->firstz;
    h.nsamples = (int64_t)db->data.size();
    size_t n = db->ntr();
    bool ok = fwrite(&h, sizeof h, 1, f) == 1 && fwrite(db->span0.data(), sizeof(int), n, f) == n &&
              fwrite(db->len.data(), sizeof(int), n, f) == n && fwrite(db->offset.data(), sizeof(long long), n, f) == n &&
              fwrite(db->data.data(), sizeof(float), db->data.size(), f) == db->data.size();
    fclose(f);
    return ok ? 0 : kiwi_set_error("write error on %s", path);
}

kiwi_gfdb* kiwi_gfdb_read(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { kiwi_set_error("can't open file %s", path); return nullptr; }
    kiwi_gfdb* db = nullptr;
    auto fail = [&](const char* what) -> kiwi_gfdb* {
        fclose(f); delete db;
        kiwi_set_error("%s: %s", path, what);
        return nullptr;
    };
    try {
        Kgf1Header h;
        if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "KGF1", 4) != 0) return fail("not a KGF1 database");
        // the same rules as kiwi_gfdb_create, and sizes that agree with the length of the file
        if (h.nx < 1 || h.nz < 1 || (h.ng != 8 && h.ng != 10) || !(h.dt > 0.f) || !(h.dx > 0.f) || !(h.dz > 0.f) || h.nsamples < 0 ||
            !std::isfinite(h.firstx) || !std::isfinite(h.firstz))
            return fail("invalid header");
        const unsigned long long ntr = (unsigned long long)h.nx * (unsigned long long)h.nz * (unsigned long long)h.ng;
        if (ntr > (1ull << 31)) return fail("invalid header (too many traces)");
        if (fseek(f, 0, SEEK_END) != 0) return fail("read error");
        const long long fsize = ftell(f);
        const unsigned long long expect = sizeof h + ntr * (2 * sizeof(int) + sizeof(long long)) + (unsigned long long)h.nsamples * sizeof(float);
        if (fsize < 0 || (unsigned long long)fsize != expect) return fail("file size does not match the header (truncated or corrupt)");
        if (fseek(f, (long)sizeof h, SEEK_SET) != 0) return fail("read error");
        db = new kiwi_gfdb();
        db->nx = h.nx; db->nz = h.nz; db->ng = h.ng; db->dt = h.dt; db->dx = h.dx; db->dz = h.dz; db->firstx = h.firstx; db->firstz = h.firstz;
        const size_t n = (size_t)ntr;
        db->span0.resize(n); db->len.resize(n); db->offset.resize(n); db->data.resize((size_t)h.nsamples);
        const bool ok = fread(db->span0.data(), sizeof(int), n, f) == n && fread(db->len.data(), sizeof(int), n, f) == n &&
                        fread(db->offset.data(), sizeof(long long), n, f) == n &&
                        fread(db->data.data(), sizeof(float), db->data.size(), f) == db->data.size();
        if (!ok) return fail("read error");
        for (size_t i = 0; i < n; i++) {   // every trace inside the sample block, spans that cannot overflow the index arithmetic
            const long long len = db->len[i], off = db->offset[i];
            if (len < 0 || off < 0 || off > h.nsamples || len > h.nsamples - off || std::abs((long long)db->span0[i]) > (1ll << 28) || len > (1ll << 28))
                return fail("trace table out of range (corrupt file)");
        }
        fclose(f);
        db->flat = true;
        return db;
    } catch (const std::exception& e) {
        return fail(e.what());
    }
}

}  // extern "C"

// =================================================================================================
// Analytical homogeneous full-space builder (gfdb_build_ahfull.f90 + elseis.f90 + elseis_oo.f90 +
// differentiation.f90 + integration.f90).  fp32 throughout, statement order as in the reference.
// =================================================================================================
namespace {

inline int f_nint(float x) { return (int)lroundf(x); }

struct Elseis {
    float rho, alpha, beta, dt;
    std::vector<float> stf, dstf, istf, istftau;
    float matfac[5];
};

// integration.f90:27-59
void antiderivate(float dt, const std::vector<float>& f, std::vector<float>& ff) {
    size_t n = std::min(f.size(), ff.size());
    if (f.size() < 2) { std::fill(ff.begin(), ff.end(), 0.f); return; }
    ff[0] = 0.f;
    for (size_t i = 0; i + 1 < n; i++) ff[i + 1] = ff[i] + (f[i + 1] + f[i]) / 2.f * dt;
}
// differentiation.f90:27-70
void differentiate(float dt, const std::vector<float>& f, std::vector<float>& df) {
    size_t n = f.size();
    for (size_t i = 1; i + 1 < n; i++) df[i] = (f[i + 1] - f[i - 1]) / (dt * 2.f);
    df[0] = (f[1] - f[0]) / dt;
    df[n - 1] = (f[n - 1] - f[n - 2]) / dt;
}
// elseis.f90:434-452
void make_istfs(float dt, const std::vector<float>& stf, std::vector<float>& istf, std::vector<float>& istftau) {
    std::vector<float> stftau(stf.size());
    for (size_t i = 0; i < stf.size(); i++) stftau[i] = stf[i] * (float)i * dt;
    antiderivate(dt, stf, istf);
    antiderivate(dt, stftau, istftau);
}
// elseis.f90:382-396
void material_factors_mt(float rho, float alpha, float beta, float m[5]) {
    const float PI = 3.14159265358979f;
    m[0] = 1.0f / (4.0f * PI * rho);
    m[1] = 1.0f / (4.0f * PI * rho * (alpha * alpha));
    m[2] = 1.0f / (4.0f * PI * rho * (beta * beta));
    m[3] = 1.0f / (4.0f * PI * rho * (alpha * alpha * alpha));
    m[4] = 1.0f / (4.0f * PI * rho * (beta * beta * beta));
}
// elseis.f90:321-357; n,p,q 1-based
void radpat_mt(const float g[3], int n, int p, int q, float rpc[5]) {
    auto delta = [](int a, int b) { return a == b ? 1.f : 0.f; };
    float gn = g[n - 1], gp = g[p - 1], gq = g[q - 1];
    rpc[0] = (15.f * gn * gp * gq) - (3.f * gn * delta(p, q)) - (3.f * gp * delta(n, q)) - (3.f * gq * delta(n, p));
    rpc[1] = (6.f * gn * gp * gq) - (gn * delta(p, q)) - (gp * delta(n, q)) - (gq * delta(n, p));
    rpc[2] = -((6.f * gn * gp * gq) - (gn * delta(p, q)) - (gp * delta(n, q)) - (2.f * gq * delta(n, p)));
    rpc[3] = gn * gp * gq;
    rpc[4] = -(gn * gp - delta(n, p)) * gq;
}
// elseis.f90:293-305
void factors_mt(const float matfac[5], const float radpat[5], float r, float f[5]) {
    f[0] = matfac[0] * radpat[0] / ((r * r) * (r * r));  // gfortran expands r**4 by repeated squaring
    f[1] = matfac[1] * radpat[1] / (r * r);
    f[2] = matfac[2] * radpat[2] / (r * r);
    f[3] = matfac[3] * radpat[3] / r;
    f[4] = matfac[4] * radpat[4] / r;
}
// elseis.f90:133-209 with addweight present
void elseis_mt_add(const Elseis& es, const float factors[5], float r, float toffset, bool nfflag, bool ffflag, float* elseism,
                   int npt, float addweight) {
    const float dt = es.dt, alpha = es.alpha, beta = es.beta;
    const int lstf = (int)es.stf.size();
    const float *stf = es.stf.data(), *dstf = es.dstf.data(), *istf = es.istf.data(), *istftau = es.istftau.data();
    int ita_delta = f_nint(toffset / dt - r / alpha / dt);
    int itb_delta = f_nint(toffset / dt - r / beta / dt);
    for (int it = 1; it <= npt; it++) {
        float t = toffset + (float)(it - 1) * dt;
        float ta = t - r / alpha;
        float tb = t - r / beta;
        int ita = ita_delta + (it - 1);
        int itb = itb_delta + (it - 1);
        ita = std::min(std::max(ita, 0), lstf - 1);
        itb = std::min(std::max(itb, 0), lstf - 1);
        float ta_delta = 0.f, tb_delta = 0.f;
        if (nfflag) { ta_delta = ta - (float)ita * dt; tb_delta = tb - (float)itb * dt; }
        // ita, itb are now 0-based indices into stf (the Fortran adds 1 for its 1-based arrays)
        float term = 0.0f;
        if (nfflag) {
            float integral_term =
                t * (istf[ita] - istf[itb] + ta_delta * stf[ita] - tb_delta * stf[itb]) -
                (istftau[ita] + ta_delta * stf[ita] * (float)ita * dt + 0.5f * stf[ita] * (ta_delta * ta_delta) - istftau[itb] -
                 tb_delta * stf[itb] * (float)itb * dt - 0.5f * stf[itb] * (tb_delta * tb_delta));
            term = term + factors[0] * integral_term;
            term = term + factors[1] * stf[ita];
            term = term + factors[2] * stf[itb];
        }
        if (ffflag) {
            term = term + factors[3] * dstf[ita];
            term = term + factors[4] * dstf[itb];
        }
        elseism[it - 1] = elseism[it - 1] + term * addweight;
    }
}

// gfdb_build_ahfull.f90:34-37 (reshape is column-major: source(p,q) = list[(q-1)*3 + (p-1)])
const float source_a[9] = {1, 1, 0, 1, 0, 0, 0, 0, 0};
const float source_b[9] = {0, 0, 1, 0, 0, 1, 1, 1, 0};
const float source_c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 1};
const float source_d[9] = {0, 0, 0, 0, 1, 0, 0, 0, 0};
inline float src(const float* s, int p, int q) { return s[(q - 1) * 3 + (p - 1)]; }

// gfdb_build_ahfull.f90:70-191 addentry for node (x,z); traces go straight into the container
void addentry(kiwi_gfdb* db, const Elseis& es, int ix, int iz, bool nfflag, bool ffflag, std::vector<float>& seis) {
    const float dt = db->dt;
    float x = db->firstx + (float)(ix - 1) * db->dx;  // gfdb_get_position gfdb.f90:817-828
    float z = db->firstz + (float)(iz - 1) * db->dz;
    float s_loc[3] = {0.f, 0.f, z}, r_loc[3] = {x, 0.f, 0.f};
    float rel[3] = {r_loc[0] - s_loc[0], r_loc[1] - s_loc[1], r_loc[2] - s_loc[2]};
    float d = sqrtf((s_loc[0] - r_loc[0]) * (s_loc[0] - r_loc[0]) + (s_loc[1] - r_loc[1]) * (s_loc[1] - r_loc[1]) +
                    (s_loc[2] - r_loc[2]) * (s_loc[2] - r_loc[2]));
    float tstf = (float)((int)es.stf.size() - 1) * es.dt;
    auto snapdown = [](float t, float dt_) { return (float)((int)floorf(t / dt_)) * dt_; };
    auto snapup = [](float t, float dt_) { return (float)((int)ceilf(t / dt_)) * dt_; };
    float firstarrival_p = snapdown(d / es.alpha, dt);
    float lastarrival_p = snapup(d / es.alpha + tstf, dt);
    float firstarrival_s = snapdown(d / es.beta, dt);
    float lastarrival_s = snapup(d / es.beta + tstf, dt) + dt * 2.f;
    float tbegin_total = firstarrival_p, tend_total = lastarrival_s;
    int nwindows; float tbegin[2], tend[2];
    if (lastarrival_p >= firstarrival_s || nfflag) { nwindows = 1; tbegin[0] = firstarrival_p; tend[0] = lastarrival_s; }
    else { nwindows = 2; tbegin[0] = firstarrival_p; tend[0] = lastarrival_p; tbegin[1] = firstarrival_s; tend[1] = lastarrival_s; }
    int nsamples = f_nint((tend_total - tbegin_total) / es.dt + 1.f);
    seis.assign((size_t)12 * nsamples, 0.f);
    // elseis.f90:399-414 make_direction_cosine
    float r = sqrtf(rel[0] * rel[0] + rel[1] * rel[1] + rel[2] * rel[2]);
    float gamma[3] = {rel[0] / r, rel[1] / r, rel[2] / r};
    const float* sources[4] = {source_a, source_b, source_c, source_d};
    for (int n = 1; n <= 3; n++)
        for (int p = 1; p <= 3; p++)
            for (int q = 1; q <= 3; q++) {
                float radpat[5], factors[5];
                bool have = false;
                for (int iw = 0; iw < nwindows; iw++) {
                    int itbegin = f_nint((tbegin[iw] - tbegin_total) / es.dt) + 1;
                    int itend = f_nint((tend[iw] - tbegin_total) / es.dt) + 1;
                    for (int isrc = 0; isrc < 4; isrc++) {
                        float w = src(sources[isrc], p, q);
                        if (w == 0.f) continue;  // adds term*0 in the reference: no change
                        if (!have) { radpat_mt(gamma, n, p, q, radpat); factors_mt(es.matfac, radpat, r, factors); have = true; }
                        float* row = &seis[(size_t)(n - 1 + 3 * isrc) * nsamples];
                        elseis_mt_add(es, factors, r, tbegin[iw], nfflag, ffflag, row + (itbegin - 1), itend - itbegin + 1, w);
                    }
                }
            }
    // gfdb_build_ahfull.f90:166-175: GF component ig <- elementary seismogram row (1-based)
    static const int rowof[10] = {1, 4, 7, 2, 5, 3, 6, 9, 10, 12};
    int span0 = f_nint(tbegin_total / dt);  // gfdb_build_ahfull.f90:206
    for (int ig = 1; ig <= db->ng; ig++) {
        const float* row = &seis[(size_t)(rowof[ig - 1] - 1) * nsamples];
        int f, l;
        kiwi_pack_window(row, nsamples, &f, &l);
        size_t i = db->idx(ix, iz, ig);
        db->pending[i].assign(row + f, row + l + 1);
        db->span0[i] = span0 + f;
        db->len[i] = l - f + 1;
    }
}

}  // namespace

extern "C" int kiwi_gfdb_build_ahfull(kiwi_gfdb* db, float rho, float alpha, float beta, const float* stf, int nstf,
                                      int nfflag, int ffflag, int nthreads) {
    if (!db) return kiwi_set_error("kiwi_gfdb_build_ahfull: null database");
    if (nstf < 2) return kiwi_set_error("sizes of arrays for differentiation are too short.");
    if (db->ng != 10 && nfflag) return kiwi_set_error("near field terms need a 10 component database");
    if (db->flat) { db->pending.assign(db->ntr(), {}); std::vector<float>().swap(db->data); db->flat = false; }
    Elseis es;
    es.rho = rho; es.alpha = alpha; es.beta = beta; es.dt = db->dt;
    es.stf.assign(stf, stf + nstf);
    es.dstf.assign(nstf, 0.f); es.istf.assign(nstf, 0.f); es.istftau.assign(nstf, 0.f);
    make_istfs(es.dt, es.stf, es.istf, es.istftau);   // elseis_oo.f90:127-157 set_stf
    differentiate(es.dt, es.stf, es.dstf);
    material_factors_mt(rho, alpha, beta, es.matfac);
    if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::atomic<int> next(0);
    const int nx = db->nx, nz = db->nz;
    auto work = [&]() {
        std::vector<float> seis;
        for (;;) {
            int ix0 = next.fetch_add(1);
            if (ix0 >= nx) break;
            for (int iz = 1; iz <= nz; iz++) addentry(db, es, ix0 + 1, iz, nfflag != 0, ffflag != 0, seis);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    db->flatten();
    return 0;
}
