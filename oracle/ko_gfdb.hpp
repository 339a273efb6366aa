// TEST INFRASTRUCTURE ONLY (see ko_base.hpp).  Restates the arithmetic part of gfdb.f90
// (grid metadata, index computation, trace lookup, bilinear trace fetch).  The HDF5 chunk cache
// (gfdb.f90:163-780, 952-1107; gfdb_io_hdf.f90) is storage, not arithmetic: it is replaced by an
// in-memory table of traces that are filled from the flat KGF1 array form (see include/kiwi_b200.h)
// by running every dense trace through trace_pack, exactly as gfdb_build_ahfull.f90:193-216 does.
#pragma once
#include "ko_trace.hpp"

namespace ko {

struct Gfdb {  // gfdb.f90:93-146 (fields used by the hot path)
    float dt = 0.f, dx = 0.f, dz = 0.f, firstx = 0.f, firstz = 0.f;
    int nx = 0, nz = 0, ng = 0;
    std::vector<Trace> traces;  // (ig, iz, ix) with ig fastest, 0-based storage
    std::vector<Trace> interpolated_traces;  // per-thread scratch, gfdb.f90:913-931
    long nwarn_oob = 0;
    Trace& tr(int ix, int iz, int ig) { return traces[((size_t)(ix - 1) * nz + (iz - 1)) * ng + (ig - 1)]; }
};

// gfdb.f90:781-792
static inline void gfdb_get_indices(const Gfdb& c, float x, float z, int& ix, int& iz) {
    ix = f_nint((x - c.firstx) / c.dx) + 1;
    iz = f_nint((z - c.firstz) / c.dz) + 1;
}
// gfdb.f90:794-815
static inline void gfdb_get_indices_bilin(const Gfdb& c, float x, float z, int nxu, int nzu, int ix[2], int iz[2],
                                          float& dix, float& diz) {
    ix[0] = (int)f_floor((x - c.firstx) / (c.dx * (float)nxu)) * nxu + 1;
    iz[0] = (int)f_floor((z - c.firstz) / (c.dz * (float)nzu)) * nzu + 1;
    ix[1] = ix[0] + nxu;
    iz[1] = iz[0] + nzu;
    dix = (x - c.firstx - (float)(ix[0] - 1) * c.dx) / (c.dx * (float)nxu);
    diz = (z - c.firstz - (float)(iz[0] - 1) * c.dz) / (c.dz * (float)nzu);
}
// gfdb.f90:830-863 (+ chunk_get_trace :952-1031: missing trace -> null)
static inline Trace* gfdb_get_trace(Gfdb& db, int ix, int iz, int ig) {
    if (db.traces.empty() || ix > db.nx || ix < 1 || iz > db.nz || iz < 1 || ig > db.ng || ig < 1) {
#pragma omp atomic
        db.nwarn_oob++;
        return nullptr;
    }
    Trace* t = &db.tr(ix, iz, ig);
    if (!t->alloc) return nullptr;  // "no trace available for index"
    return t;
}
// gfdb.f90:865-950; `scratch` is this thread's db%interpolated_traces(iipt)%p
static inline Trace* gfdb_get_trace_bilin(Gfdb& db, const int ix[2], const int iz[2], int ig, float dix, float diz,
                                          Trace& scratch) {
    if (dix == 0.f && diz == 0.f) return gfdb_get_trace(db, ix[0], iz[0], ig);
    Trace* t00 = gfdb_get_trace(db, ix[0], iz[0], ig);
    Trace* t01 = gfdb_get_trace(db, ix[0], iz[1], ig);
    Trace* t10 = gfdb_get_trace(db, ix[1], iz[0], ig);
    Trace* t11 = gfdb_get_trace(db, ix[1], iz[1], ig);
    if (!(t00 && t01 && t10 && t11)) return nullptr;
    int span[2];
    span[0] = std::min(std::min(t00->span[0], t01->span[0]), std::min(t10->span[0], t11->span[0]));
    span[1] = std::max(std::max(t00->span[1], t01->span[1]), std::max(t10->span[1], t11->span[1]));
    Trace* tp = &scratch;
    if (trace_is_empty(*tp)) {
        trace_create_simple_nodata(*tp, span[0], span[1]);
    } else {
        resize(tp->strips[0], span[0], span[1] - span[0] + 1);
        tp->span[0] = span[0]; tp->span[1] = span[1];
    }
    std::fill(tp->strips[0].d.begin(), tp->strips[0].d.end(), 0.f);
    sreal* arr = tp->strips[0].d.data();
    trace_multiply_add_nogrow(*t00, arr, span[0], span[1], (1.f - dix) * (1.f - diz));
    trace_multiply_add_nogrow(*t01, arr, span[0], span[1], (1.f - dix) * diz);
    trace_multiply_add_nogrow(*t10, arr, span[0], span[1], dix * (1.f - diz));
    trace_multiply_add_nogrow(*t11, arr, span[0], span[1], dix * diz);
    return tp;
}

// Fill from the flat array form: for every (ix,iz,ig) a dense sample array that starts at sample
// index span0 (gfdb_build_ahfull.f90:206 convention), len==0 meaning "no trace stored".
static inline void gfdb_from_arrays(Gfdb& db, int nx, int nz, int ng, float dt, float dx, float dz, float firstx,
                                    float firstz, const int* span0, const int* len, const long long* offset,
                                    const float* data) {
    db.nx = nx; db.nz = nz; db.ng = ng; db.dt = dt; db.dx = dx; db.dz = dz; db.firstx = firstx; db.firstz = firstz;
    size_t n = (size_t)nx * nz * ng;
    db.traces.assign(n, Trace());
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; i++) {
        if (len[i] <= 0) continue;
        Strip conti;
        strip_init(span0[i], span0[i] + len[i] - 1, data + offset[i], len[i], conti);
        trace_pack(conti, db.traces[i]);
    }
}

}  // namespace ko
