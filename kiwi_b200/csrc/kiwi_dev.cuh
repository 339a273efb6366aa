// Device-side data layout of the engine (HBM-resident structures shared by the kernels).
// Reference citations are file:line of /root/reference.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define KIWI_INTERNAL_XCORR 9     // k_misfit_general: cross-correlation mode (autoshift_ref_seismogram), not a norm id of the ABI
#define KIWI_MAX_COMP 5          // receiver.f90:35-48: at most a/c r/l d/u n/s e/w
#define KIWI_NG_MAX 10
#define KIWI_MAX_PLF 8           // points of a taper / filter piecewise linear function kept on the device

// ---- Green's function database in HBM -----------------------------------------------------------
// One slab per grid node (ix,iz): ng rows of `wn` fp32 samples, all rows on the node's common
// sample window [w0, w0+wn).  w0 and wn are multiples of 4 and every slab starts on a 16-byte
// boundary, so a lane that owns the absolute sample quad 4q..4q+3 reads row + (4q - w0) with one
// 128-bit load.  Inside the window a row holds the trace with the reference's implicit
// continuation made explicit (sparse_trace.f90:29-50): zeros left of the trace span and in
// inter-strip gaps, the last sample repeated to the right.
struct __align__(16) NodeInfo {
    unsigned long long off;  // first float of the slab, multiple of 4; ~0ull = node has no traces
    int w0;                  // first sample index of the window (multiple of 4)
    int wn;                  // window length in samples (multiple of 4)
};
struct GfdbDev {
    float dt, dx, dz, firstx, firstz;
    int nx, nz, ng;
    const float* slabs;       // all slabs
    const NodeInfo* nodes;    // [nx*nz], inode = (ix-1)*nz + (iz-1)
    const int2* tspan;        // [nx*nz*ng] first/last sample index of every trace (trace%span)
    const int4* nspan;        // [nx*nz][2] span unions of a node's component sets {lo1, hi1, lo2, hi2} {lo3, hi3, -, -}: set 1 = g1 g2 g3 (g9)
                              // -> radial, set 2 = g4 g5 -> transverse, set 3 = g6 g7 g8 (g10) -> vertical (seismogram.f90:167-250)
    const float* lastval;     // [nx*nz*ng] last stored sample of every trace
};

// ---- receivers ------------------------------------------------------------------------------------
struct ReceiverDev {
    double azi0, bazi0, dist0;   // azibazi + distance_accurate50m to the source origin (seismogram.f90:99-100)
    float depth;
    float cl0, sl0;              // cos/sin(bazi0 + pi), seismogram.f90:270-271
    int enabled;
    int ncomp;
    int comp[KIWI_MAX_COMP];     // component ids (+-1..5) in the receiver's order
    int misfit_base;             // index of this receiver's first misfit pair among enabled receivers
    int ja, jr, jd, jn, je;      // 1-based component index or 0 (seismogram.f90:82-86)
    float sa, sr, sd, sn, se;    // component signs (:88-92)
    // reference traces / probes (comparator.f90:55-80), one per component
    int ref_ds0[KIWI_MAX_COMP], ref_ds1[KIWI_MAX_COMP];   // ref probe dataspan
    int ref_sp0[KIWI_MAX_COMP], ref_sp1[KIWI_MAX_COMP];   // ref probe span after set_ref_seismograms
    long long ref_off[KIWI_MAX_COMP];                     // offset of the ref samples in d_refdata
    float ref_rs[KIWI_MAX_COMP];                          // power of two that brings the rms of the reference trace to order one: the
                                                          // tensor-core misfit kernels square fp32 residuals scaled by it (k_mt_fused)
    double ref_ss[KIWI_MAX_COMP], ref_sa[KIWI_MAX_COMP];  // sum of squares / of magnitudes of the reference's data span in double (the
                                                          // untapered reference-only norm, comparator.f90:639-659: the same for every candidate)
    // taper (piecewise_linear_function.f90:195-237), tabulated on [tp0, tp1] by the host
    int has_taper, tp0, tp1;     // tp0 = floor(x1/dt)+1, tp1 = floor(xn/dt): support of the taper
    int dps0, dps1;              // discrete_plf_span (comparator.f90:1145-1157)
    long long taper_off;         // offset into d_taper
    int has_filter;
    int fs0, fs1;                // floating shift range in samples
    // the piecewise linear functions themselves (piecewise_linear_function.f90:27-35): the filter is
    // evaluated at k*df on the device because df depends on the padded span of each candidate
    int ntp, nfp;
    float tpx[KIWI_MAX_PLF], tpy[KIWI_MAX_PLF];
    float fpx[KIWI_MAX_PLF], fpy[KIWI_MAX_PLF];
};

// ---- discretised sources (device SoA; discrete_source.f90:27-45 made column-wise) -----------------
// A "group" is a set of centroids that share position and moment-tensor shape and differ only in
// time and scalar weight (the nt time-centroids of one bilateral sub-fault, source_bilat.f90:440-457;
// all nt centroids of a point moment tensor, source_moment_tensor.f90:256-263).  The reference's
// centroid table is recovered as: for each group, for each of its taps:
//   (north, east, depth, tbase (+) toff, mhat * wt).
struct CandDev {
    int group_begin, ngroups;   // into the group arrays
    int tap_begin, ntaps_total; // into the tap arrays
    float moment, risetime;     // psm%moment / psm%risetime applied after synthesis (receiver.f90:853-904)
    int nx, ny, nt;             // grid_size (source_bilat.f90:266-268)
    int status;
    int walk_ny;                // > 0: the groups are an nx x walk_ny lattice listed with the second (down-dip) index fastest
                                // (source_bilat.f90:349-371); 0: listed row by row or without order (depth bands of k_synth = slices)
    int nbands;                 // depth bands k_synth works through this candidate in (>= 1): a function of the candidate, the
                                // database and the receivers only, so that a result does not depend on the rest of the batch
};
struct GroupSoA {
    float *north, *east, *depth, *tbase;
    float* mhat;                // [6][ngroups_total] : mxx myy mzz mxy mxz myz
    float* gw;                  // scalar weight of the group (sub-fault weight of an eikonal source, source_eikonal.f90:697; 1 otherwise)
    float* lam;                 // atan2f(east, north) of the group from the host library (orthodrome.f90:121), or null = the device's own
    int *tap_begin, *tap_count; // taps of this group
    int *its_min, *its_max;     // min/max of floor((tbase (+) toff)/dt) over the taps
    // receiver-independent shift table of the group (k_tap_table): what trace_multiply_add derives from
    // (time, weight) of every centroid, sparse_trace.f90:639-646
    int* tt_begin;              // first entry of the group in taprec (room for one entry per tap)
    int* nstep;                 // entries of the group: distinct quad shifts (sample shift div 4) of its taps
    float4* taprec;             // two per entry: {quad shift (int bits), h0, h1, h2} {h3, h4, W, -}, see k_tap_table
};
struct TapSoA {
    float *toff, *wt;
};

// ---- per (candidate, receiver, group) geometry record written by the pre-pass ---------------------
struct __align__(16) GeoRec {   // 128 bytes: everything k_synth needs to start streaming a group
    int ix1, iz1;       // gfdb_get_indices[_bilin] gfdb.f90:781-815
    float dix, diz;
    float f[6];         // make_weights(real(azi), mhat) seismogram.f90:316-336 (tap weight applied later)
    float cl, sl;       // real(cos/sin(bazi - bazi0)), seismogram.f90:163-164
    int flags;
    int tt_begin, nstep;    // copy of the group's shift-table header (GroupSoA), so that it arrives with the record
    int pad[1];
    NodeInfo node[4];   // slabs of the corners (ix1,iz1) (ix1,iz2) (ix2,iz1) (ix2,iz2); all = corner 0 if GEO_SINGLE
};
#define GEO_SKIP 1      // a needed node is outside the database: centroid skipped (seismogram.f90:172)
#define GEO_ROT 2       // lambda /= 0: per-centroid rotation branch (seismogram.f90:160)
#define GEO_NEAR 4      // scaled coordinate within 4 ulps of an integer
#define GEO_SINGLE 8    // dix == 0 and diz == 0: node trace used directly (gfdb.f90:893-896)

// one grid location of a point moment-tensor grid search: its candidates in the location-sorted tensor list
struct MtLoc {
    int mt_begin, mt_count;
};

// misfit-stage view of a candidate whose synthesis is shared with others (only the moment differs,
// minimizer_engine.f90:511-521 `only_moment_changed`): where its results go, whose synthetics it uses, its moment
struct CandMap {
    int out;        // candidate index for the misfit block / status (relative to the pointers passed)
    int syn;        // synthesised candidate whose rows are read (relative to cands / seis / shdrs passed)
    float moment;   // psm%moment of this candidate (receiver.f90:853-904)
    int pad;
};

// per (candidate, receiver) header
struct PairHdr {
    int s1lo, s1hi;     // span of displacement_ar(1)
    int s2lo, s2hi;     // span of displacement_ar(2)
    int s3lo, s3hi;     // span of the vertical strip
    int out0, T;        // union window: first sample, length (0 = nothing synthesised)
};

// per (candidate, receiver, component) output descriptor
struct SeisHdr {
    int lo, hi;         // strip span of receiver%displacement(icomp) (fresh state)
    int base;           // sample index of element 0 of the stored row (multiple of 4)
    int pad;
};
