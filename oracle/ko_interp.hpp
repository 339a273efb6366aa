// TEST INFRASTRUCTURE ONLY (see ko_base.hpp).  Restates interpolation.f90 (Gulunay's generalised f-k interpolation) and
// the interpolating half of gfdb.f90 (gfdb_init nipx/nipz :163-264, gfdb_interpolate_block :1109-1232, interpolate3d
// :1234-1310).  FFTW3 (sfftw_plan_dft_r2c_2d/3d, c2r) is replaced by the radix-2 fp32 transform of ko_comparator.hpp with FFTW's
// conventions: unnormalised, the halved dimension is the first (time) one, the multi-dimensional c2r transforms the other
// dimensions first and ignores the imaginary parts of the time bins 0 and t/2.  PARITY UNPINNED: the reference has no test of the
// interpolation.  Two spots of the Fortran read memory it never wrote (the automatic arrays B and D: B outside (1::l) x (il::l) in
// gulunay3d, D beyond column s in gulunay2d and off the stride-l lattice in gulunay3d); they are taken as zero, the evident intent.
#pragma once
#include "ko_gfdb.hpp"
#include "ko_comparator.hpp"

namespace ko {

static const int nblockx_default = 128, nblockx_overlap_default = 32, nblockx_payload_default = 128 - 32;   // gfdb.f90:31-33
static const int nblockz_default = 32, nblockz_overlap_default = 8, nblockz_payload_default = 32 - 8;       // gfdb.f90:35-37

// complex array (n0 fastest, n1, n2), transform along one axis
static inline void fft_axis(std::vector<cfloat>& a, int n0, int n1, int n2, int axis, int sign) {
    const int n[3] = {n0, n1, n2};
    const size_t st[3] = {1, (size_t)n0, (size_t)n0 * n1};
    const int len = n[axis];
    if (len <= 1) return;
    const int o1 = (axis + 1) % 3, o2 = (axis + 2) % 3;
    std::vector<cfloat> line(len);
    for (int j = 0; j < n[o2]; j++)
        for (int i = 0; i < n[o1]; i++) {
            const size_t base = (size_t)i * st[o1] + (size_t)j * st[o2];
            for (int k = 0; k < len; k++) line[k] = a[base + (size_t)k * st[axis]];
            fft_c(line, sign);
            for (int k = 0; k < len; k++) a[base + (size_t)k * st[axis]] = line[k];
        }
}
// sfftw_plan_dft_r2c_{2,3}d of a real array (nt, n1, n2): rows 0..nkeep-1 of the halved time dimension
static inline void r2c_keep(const std::vector<float>& r, int nt, int n1, int n2, int nkeep, std::vector<cfloat>& out) {
    out.assign((size_t)nkeep * n1 * n2, cfloat(0.f, 0.f));
    std::vector<cfloat> line(nt);
    for (size_t l = 0; l < (size_t)n1 * n2; l++) {
        for (int i = 0; i < nt; i++) line[i] = cfloat(r[l * nt + i], 0.f);
        fft_c(line, -1);
        for (int f = 0; f < nkeep; f++) out[l * nkeep + f] = line[f];
    }
    fft_axis(out, nkeep, n1, n2, 1, -1);
    fft_axis(out, nkeep, n1, n2, 2, -1);
}
// abs() of a default complex: sqrt(re^2 + im^2) without spurious overflow; written through a double square root so that it is
// reproducible anywhere (glibc's cabsf agrees except for rare last-bit cases)
static inline float cabs_f(cfloat z) { return (float)sqrt((double)z.real() * (double)z.real() + (double)z.imag() * (double)z.imag()); }
// complex division as gfortran emits it (-fcx-fortran-rules: Smith's method)
static inline cfloat cdiv_fortran(cfloat a, cfloat b) {
    if (std::fabs(b.real()) >= std::fabs(b.imag())) {
        const float r = b.imag() / b.real(), den = b.real() + b.imag() * r;
        return cfloat((a.real() + a.imag() * r) / den, (a.imag() - a.real() * r) / den);
    }
    const float r = b.real() / b.imag(), den = b.real() * r + b.imag();
    return cfloat((a.real() * r + a.imag()) / den, (a.imag() * r - a.real()) / den);
}
static inline float taper_w(int k, float width) { return (1.f - cosf(2.f * pi * ((float)k / width))) / 2.f; }   // interpolation.f90:69-83

// gulunay2d (interpolation.f90:29-159) and gulunay3d (:161-311) in one: A (t, s1, s2) -> Inter (t, s1*l1, s2*l2) with l1, l2 in {1, l}.
// gulunay2d is (l1, l2) = (l, 1) with s2 = 1; margin2 is unused then.
static inline void gulunay(std::vector<float>& A, int t, int s1, int s2, int l1, int l2, std::vector<float>& Inter, int ntmargin, int margin1,
                           int margin2) {
    const int l = std::max(l1, l2), kk1 = s1 * l1, kk2 = s2 * l2, ff = l * t, fny = t / 2 + 1;
    auto a = [&](int it, int i1, int i2) -> float& { return A[((size_t)i2 * s1 + i1) * t + it]; };
    // --- taper (in place, the caller's array is changed): last dimension, then the middle one, then time
    if (l2 > 1) {
        const int m = margin2 / l; const float w = 2.f * (float)margin2 / (float)l;
        for (int x = 1; x <= m; x++) for (int i1 = 0; i1 < s1; i1++) for (int it = 0; it < t; it++) a(it, i1, x - 1) = a(it, i1, x - 1) * taper_w(x - 1, w);
        for (int x = s2 - m + 1; x <= s2; x++) for (int i1 = 0; i1 < s1; i1++) for (int it = 0; it < t; it++) a(it, i1, x - 1) = a(it, i1, x - 1) * taper_w(s2 - x, w);
    }
    if (l1 > 1) {
        const int m = margin1 / l; const float w = 2.f * (float)margin1 / (float)l;
        for (int x = 1; x <= m; x++) for (int i2 = 0; i2 < s2; i2++) for (int it = 0; it < t; it++) a(it, x - 1, i2) = a(it, x - 1, i2) * taper_w(x - 1, w);
        for (int x = s1 - m + 1; x <= s1; x++) for (int i2 = 0; i2 < s2; i2++) for (int it = 0; it < t; it++) a(it, x - 1, i2) = a(it, x - 1, i2) * taper_w(s1 - x, w);
    }
    {
        const int m = ntmargin / l; const float w = 2.f * (float)ntmargin / (float)l;
        for (int x = 1; x <= m; x++) for (size_t c = 0; c < (size_t)s1 * s2; c++) A[c * t + x - 1] = A[c * t + x - 1] * taper_w(x - 1, w);
        for (int x = t - m + 1; x <= t; x++) for (size_t c = 0; c < (size_t)s1 * s2; c++) A[c * t + x - 1] = A[c * t + x - 1] * taper_w(t - x, w);
    }
    // --- B: zero traces inserted; C: zero padded in all dimensions; D: C with only every l-th trace kept
    std::vector<float> B((size_t)t * kk1 * kk2, 0.f), C((size_t)ff * kk1 * kk2, 0.f), D((size_t)ff * kk1 * kk2, 0.f);
    for (int i2 = 0; i2 < s2; i2++)
        for (int i1 = 0; i1 < s1; i1++)
            for (int it = 0; it < t; it++) {
                const float v = a(it, i1, i2);
                B[((size_t)(i2 * l2) * kk1 + i1 * l1) * t + it] = v;
                C[((size_t)i2 * kk1 + i1) * ff + it] = v;
                if (i1 % l1 == 0 && i2 % l2 == 0) D[((size_t)i2 * kk1 + i1) * ff + it] = v;
            }
    std::vector<cfloat> fB, fC, fD;
    r2c_keep(B, t, kk1, kk2, fny, fB);
    r2c_keep(C, ff, kk1, kk2, fny, fC);      // only fC(1:fny,...) and fD(1:fny,...) are used
    r2c_keep(D, ff, kk1, kk2, fny, fD);
    // --- white noise (:119-127, :268-276)
    float mx = 0.f;
    for (size_t c = 0; c < (size_t)kk1 * kk2; c++) mx = std::max(mx, cabs_f(fD[c * fny + (fny - 1)]));
    const float m = 0.01f * mx;
    for (auto& v : fD) if (cabs_f(v) < m / 1000.f) v = cfloat(m, v.imag());
    for (auto& v : fD) { const float av = cabs_f(v); if (av < m) { const float sc = m / av; v = cfloat(sc * v.real(), sc * v.imag()); } }
    // --- operator, clipped (:129-144, :278-296)
    const float ls = (float)(l1 * l2), lowcut = (l1 > 1 && l2 > 1) ? 0.5f * (float)(l * l) : (float)l * 0.5f;
    const float norm = (float)(t * kk1 * kk2);   // t*kk, t*kkx*kkz: default integer products
    std::vector<cfloat> fI(fB.size());
    for (size_t i = 0; i < fB.size(); i++) {
        cfloat op = cdiv_fortran(fC[i], fD[i]);
        { const float ao = cabs_f(op); if (ao > ls) { const float sc = ls / ao; op = cfloat(sc * op.real(), sc * op.imag()); } }
        if (cabs_f(op) < lowcut) op = cfloat(0.f, 0.f);
        const cfloat pr(fB[i].real() * op.real() - fB[i].imag() * op.imag(), fB[i].real() * op.imag() + fB[i].imag() * op.real());
        fI[i] = cfloat(pr.real() / norm, pr.imag() / norm);
    }
    // --- back to the time domain: c2r transforms the trace dimensions first, the halved one last
    fft_axis(fI, fny, kk1, kk2, 2, +1);
    fft_axis(fI, fny, kk1, kk2, 1, +1);
    Inter.assign((size_t)t * kk1 * kk2, 0.f);
    std::vector<cfloat> half(fny);
    std::vector<sreal> line(t);
    for (size_t c = 0; c < (size_t)kk1 * kk2; c++) {
        for (int f = 0; f < fny; f++) half[f] = fI[c * fny + f];
        fft_c2r(half, t, line.data());
        for (int it = 0; it < t; it++) Inter[c * t + it] = (float)line[it];
    }
}

// interpolate3d (gfdb.f90:1234-1310): fin (nt, nz_in, nx_in) -> fout (nt, nz_out, nx_out)
static inline void interpolate3d(std::vector<float>& fin, int nt, int nz_in, int nx_in, std::vector<float>& fout, int nz_out, int nx_out, int ntmargin,
                                 int nxmargin, int nzmargin) {
    const int nipx = nx_out / nx_in, nipz = nz_out / nz_in;
    if (nipz == 1) { gulunay(fin, nt, nx_in, 1, nipx, 1, fout, ntmargin, nxmargin, 0); return; }     // fin(:,1,:)
    if (nipx == 1) { gulunay(fin, nt, nz_in, 1, nipz, 1, fout, ntmargin, nzmargin, 0); return; }     // fin(:,:,1)
    if (nipx == 4 && nipz == 4) {
        std::vector<float> finter;
        gulunay(fin, nt, nz_in, nx_in, 2, 2, finter, ntmargin, nzmargin / 2, nxmargin / 2);
        gulunay(finter, nt, nz_out / 2, nx_out / 2, 2, 2, fout, ntmargin, nzmargin, nxmargin);
        return;
    }
    if (nipx == nipz) { gulunay(fin, nt, nz_in, nx_in, nipz, nipx, fout, ntmargin, nzmargin, nxmargin); return; }
    // pseudo 3-D: horizontal, then vertical 2-D passes, statement by statement (including the ix_in test and the x margin for z)
    fout.assign((size_t)nt * nz_out * nx_out, 0.f);
    std::vector<float> in, out;
    for (int iz_in = 1; iz_in <= nz_in; iz_in++) {
        const int iz_out = (iz_in - 1) * nipz + 1;
        in.assign((size_t)nt * nx_in, 0.f);
        for (int ix = 0; ix < nx_in; ix++) for (int it = 0; it < nt; it++) in[(size_t)ix * nt + it] = fin[((size_t)ix * nz_in + iz_in - 1) * nt + it];
        gulunay(in, nt, nx_in, 1, nipx, 1, out, ntmargin, nxmargin, 0);
        for (int ix = 0; ix < nx_out; ix++) for (int it = 0; it < nt; it++) fout[((size_t)ix * nz_out + iz_out - 1) * nt + it] = out[(size_t)ix * nt + it];
    }
    for (int ix_out = 1; ix_out <= nx_out; ix_out++) {
        const int ix_in = (ix_out - 1) / nipx + 1;
        in.assign((size_t)nt * nz_in, 0.f);
        if ((ix_in - 1) % nipx == 0) {
            for (int iz = 0; iz < nz_in; iz++) for (int it = 0; it < nt; it++) in[(size_t)iz * nt + it] = fin[((size_t)(ix_in - 1) * nz_in + iz) * nt + it];
        } else {
            for (int iz = 0; iz < nz_in; iz++) for (int it = 0; it < nt; it++) in[(size_t)iz * nt + it] = fout[((size_t)(ix_out - 1) * nz_out + iz * nipz) * nt + it];
        }
        gulunay(in, nt, nz_in, 1, nipz, 1, out, ntmargin, nxmargin, 0);
        for (int iz = 0; iz < nz_out; iz++) for (int it = 0; it < nt; it++) fout[((size_t)(ix_out - 1) * nz_out + iz) * nt + it] = out[(size_t)iz * nt + it];
    }
}

struct InterpGfdb {   // the t_gfdb fields of an interpolating database (gfdb.f90:116-131)
    Gfdb db;          // pretends nx*nipx x nz*nipz traces
    int nipx = 1, nipz = 1;
    int nblockx = 1, nblockx_overlap = 0, nblockx_payload = 1, nblockz = 1, nblockz_overlap = 0, nblockz_payload = 1;
};
// gfdb_init with nipx / nipz (gfdb.f90:205-246): real traces sit at ((ix-1)*nipx+1, (iz-1)*nipz+1)
static inline void interp_gfdb_init(InterpGfdb& g, const Gfdb& src, int nipx, int nipz) {
    g = InterpGfdb();
    g.nipx = nipx; g.nipz = nipz;
    g.db.dt = src.dt; g.db.firstx = src.firstx; g.db.firstz = src.firstz; g.db.ng = src.ng;
    g.db.nx = src.nx * nipx; g.db.dx = src.dx / (float)nipx;
    g.db.nz = src.nz * nipz; g.db.dz = src.dz / (float)nipz;
    if (nipx != 1) { g.nblockx = nblockx_default; g.nblockx_payload = nblockx_payload_default; g.nblockx_overlap = nblockx_overlap_default; }
    if (nipz != 1) { g.nblockz = nblockz_default; g.nblockz_payload = nblockz_payload_default; g.nblockz_overlap = nblockz_overlap_default; }
    g.db.traces.assign((size_t)g.db.nx * g.db.nz * g.db.ng, Trace());
    for (int ix = 1; ix <= src.nx; ix++)
        for (int iz = 1; iz <= src.nz; iz++)
            for (int ig = 1; ig <= src.ng; ig++)
                g.db.tr((ix - 1) * nipx + 1, (iz - 1) * nipz + 1, ig) = src.traces[((size_t)(ix - 1) * src.nz + (iz - 1)) * src.ng + (ig - 1)];
}
static inline void gfdb_allowed_span(const int span[2], int minlength, int out[2]) {   // gfdb.f90:1313-1330
    int length = span[1] - span[0] + 1;
    if (length < minlength) length = minlength;
    const int lengthp = 1 << f_ceiling(logf((float)length) / logf(2.f));
    out[0] = span[0] - f_floor((float)(lengthp - length) / 2.f);
    out[1] = out[0] + lengthp - 1;
}
// gfdb_interpolate_block (gfdb.f90:1109-1232): fills the payload of the block that contains (ix_in, iz_in)
static inline void gfdb_interpolate_block(InterpGfdb& g, int ix_in, int iz_in) {
    Gfdb& db = g.db;
    const int nbx = g.nblockx, nbz = g.nblockz, nipx = g.nipx, nipz = g.nipz;
    std::vector<int> spans((size_t)2 * nbz * nbx, 0);
    auto sp = [&](int k, int bz, int bx) -> int& { return spans[((size_t)(bx - 1) * nbz + (bz - 1)) * 2 + k]; };
    const int ibx = (ix_in - 1) / g.nblockx_payload + 1, ibz = (iz_in - 1) / g.nblockz_payload + 1;
    const int ixfirst = (ibx - 1) * g.nblockx_payload + 1 - g.nblockx_overlap / 2, izfirst = (ibz - 1) * g.nblockz_payload + 1 - g.nblockz_overlap / 2;
    const int ixlast = ixfirst + nbx - 1, izlast = izfirst + nbz - 1;
    auto real_ix = [&](int ix) { return (std::min(std::max(ix, 1), db.nx) - 1) / nipx * nipx + 1; };   // repeat end points
    auto real_iz = [&](int iz) { return (std::min(std::max(iz, 1), db.nz) - 1) / nipz * nipz + 1; };
    int span[2] = {std::numeric_limits<int>::max(), -std::numeric_limits<int>::max()};
    for (int ix = ixfirst; ix <= ixlast; ix += nipx)
        for (int iz = izfirst; iz <= izlast; iz += nipz)
            for (int ig = 1; ig <= db.ng; ig++) {
                Trace* tp = gfdb_get_trace(db, real_ix(ix), real_iz(iz), ig);
                if (!tp) { sp(0, iz - izfirst + 1, ix - ixfirst + 1) = 0; sp(1, iz - izfirst + 1, ix - ixfirst + 1) = 0; continue; }
                span[0] = std::min(tp->span[0], span[0]); span[1] = std::max(tp->span[1], span[1]);
                sp(0, iz - izfirst + 1, ix - ixfirst + 1) = tp->span[0]; sp(1, iz - izfirst + 1, ix - ixfirst + 1) = tp->span[1];
            }
    { int t[2]; gfdb_allowed_span(span, std::min(64, (int)((float)(span[1] - span[0]) * 1.2f)), t); span[0] = t[0]; span[1] = t[1]; }
    const int nblockt = span[1] - span[0] + 1;
    if (nblockt <= 1) return;
    const int nzo = nbz / nipz, nxo = nbx / nipx;
    std::vector<float> field_orig, field_interpol;
    std::vector<sreal> line(nblockt);
    const int ovx = g.nblockx_overlap / 2, ovz = g.nblockz_overlap / 2;
    for (int ig = 1; ig <= db.ng; ig++) {
        field_orig.assign((size_t)nblockt * nzo * nxo, 0.f);
        for (int iz = izfirst; iz <= izlast; iz += nipz) {
            const int bz = iz - izfirst + 1;
            for (int ix = ixfirst; ix <= ixlast; ix += nipx) {
                const int bx = ix - ixfirst + 1;
                Trace* tp = gfdb_get_trace(db, real_ix(ix), real_iz(iz), ig);
                if (!tp) continue;
                float* dst = &field_orig[((size_t)((bx - 1) / nipx) * nzo + (bz - 1) / nipz) * nblockt];
                for (int i = 0; i < nblockt; i++) line[i] = dst[i];
                trace_multiply_add_nogrow(*tp, line.data(), span[0], span[1]);
                for (int i = 0; i < nblockt; i++) dst[i] = (float)line[i];
            }
        }
        interpolate3d(field_orig, nblockt, nzo, nxo, field_interpol, nbz, nbx, (int)(0.1f * (float)(span[1] - span[0])), ovx, ovz);
        for (int iz = izfirst + ovz; iz <= izlast - ovz; iz++) {
            const int bz = iz - izfirst + 1;
            for (int ix = ixfirst + ovx; ix <= ixlast - ovx; ix++) {
                const int bx = ix - ixfirst + 1;
                if ((ix - 1) % nipx == 0 && (iz - 1) % nipz == 0) continue;
                if (ix < 1 || db.nx < ix || iz < 1 || db.nz < iz) continue;
                const int lrx = ((bx - 1) / nipx) * nipx + 1, lrz = ((bz - 1) / nipz) * nipz + 1, nrx = lrx + nipx, nrz = lrz + nipz;
                int ds[2] = {sp(0, lrz, lrx), sp(1, lrz, lrx)};
                if (nrx <= nbx) { ds[0] = std::min(sp(0, lrz, nrx), ds[0]); ds[1] = std::max(sp(1, lrz, nrx), ds[1]); }
                if (nrz <= nbz) { ds[0] = std::min(sp(0, nrz, lrx), ds[0]); ds[1] = std::max(sp(1, nrz, lrx), ds[1]); }
                if (nrx <= nbx && nrz <= nbz) { ds[0] = std::min(sp(0, nrz, nrx), ds[0]); ds[1] = std::max(sp(1, nrz, nrx), ds[1]); }
                Trace& dst = db.tr(ix, iz, ig);
                if (dst.alloc) continue;   // "trace already exists"
                trace_create_simple(dst, &field_interpol[((size_t)(bx - 1) * nbz + (bz - 1)) * nblockt + (ds[0] - span[0])], ds[0], ds[1]);
            }
        }
    }
}
// every block, eagerly (the reference interpolates a block the first time one of its traces is asked for, gfdb.f90:996-1002)
static inline void gfdb_interpolate_all(InterpGfdb& g) {
    for (int ix = 1; ix <= g.db.nx; ix += g.nblockx_payload)
        for (int iz = 1; iz <= g.db.nz; iz += g.nblockz_payload)
            gfdb_interpolate_block(g, ix, iz);
}

}  // namespace ko
