// Hand-written sm_100a kernels of the Kiwi source-inversion hot path.
//   K1 k_bilat_groups / k_group_tap_range : sub-source groups -> device SoA   (source_bilat.f90:349-377)
//   K2 k_geometry                         : per (candidate, receiver, group) azimuth/distance, GF
//                                           indices, rotation, output spans  (orthodrome.f90:77-156,
//                                           seismogram.f90:139-165, gfdb.f90:781-815)
//   K3 k_synth                            : GF gather-FMA synthesis           (gfdb.f90:865-950,
//                                           sparse_trace.f90:597-707, seismogram.f90:131-289)
//   K5 k_misfit_td                        : scaling + time-domain misfits     (receiver.f90:853-904,
//                                           comparator.f90:222-271, 464-486, 627-697, 770-859)
// Reference citations are file:line of /root/reference.  All kernels are memory- or latency-
// bound integer/fp32 work; none is GEMM-shaped, so no tensor-core path is used here.
#include "kiwi_dev.cuh"
#include "kernels.cuh"
#include <cfloat>
#include <climits>
#include <cstdlib>
#include <algorithm>

// ---- exactly-rounded fp32 helpers: never contracted into FMA by ptxas -------------------------
__device__ __forceinline__ float A_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float S_(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float M_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float D_(float a, float b) { return __fdiv_rn(a, b); }

__device__ __forceinline__ NodeInfo ld_node(const NodeInfo* p) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    NodeInfo n; n.off = ((unsigned long long)u.y << 32) | u.x; n.w0 = (int)u.z; n.wn = (int)u.w;
    return n;
}
__device__ __forceinline__ int warp_min_i(int v) { for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
__device__ __forceinline__ int warp_max_i(int v) { for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
__device__ __forceinline__ int floordiv4(int x) { return x >> 2; }            // arithmetic shift = floor
__device__ __forceinline__ int floor4(int x) { return x & ~3; }

// =================================================================================================
// K1: bilateral sub-fault grid (source_bilat.f90:349-377); one CTA per candidate.
// Transcendentals (euler matrices) and the STF taps were evaluated on the host, everything here is
// + - * / abs max floor in IEEE fp32 with the reference's operation order.
// =================================================================================================
__global__ void k_bilat_groups(const BilatCand* __restrict__ cands, GroupSoA g, TapSoA taps, float dt, int ngroups_total) {
    const BilatCand c = cands[blockIdx.x];
    const int np = c.nx * c.ny;
    const float length = A_(c.length_a, c.length_b);
    for (int ip = threadIdx.x; ip < np; ip += blockDim.x) {
        const int ix = ip / c.ny + 1, iy = ip % c.ny + 1;   // do ix / do iy, ip = ip+1 (source_bilat.f90:349-371)
        // grid(1,ip) = (2.*(ix-1.)-nx+1.)/(2.*nx) * length
        float g0 = M_(D_(A_(S_(M_(2.f, S_((float)ix, 1.f)), (float)c.nx), 1.f), M_(2.f, (float)c.nx)), length);
        float g1 = M_(D_(A_(S_(M_(2.f, S_((float)iy, 1.f)), (float)c.ny), 1.f), M_(2.f, (float)c.ny)), c.width);
        float g2 = 0.f;
        // tshift = abs(length/2. - length_b + grid(1,ip))/rupvel + params(1) - max(la,lb)/2./rupvel
        float tshift = S_(A_(D_(fabsf(A_(S_(D_(length, 2.f), c.length_b), g0)), c.rupvel), c.time),
                          D_(D_(fmaxf(c.length_a, c.length_b), 2.f), c.rupvel));
        // p = matmul(rotmat_rup, grid(:,ip)); row-major rot[i*3+j] = rotmat(i+1,j+1)
        float p[3];
#pragma unroll
        for (int i = 0; i < 3; i++)
            p[i] = A_(A_(A_(0.f, M_(c.rot_rup[i * 3 + 0], g0)), M_(c.rot_rup[i * 3 + 1], g1)), M_(c.rot_rup[i * 3 + 2], g2));
        const int gi = c.group_begin + ip;
        g.north[gi] = A_(p[0], c.north);
        g.east[gi] = A_(p[1], c.east);
        g.depth[gi] = A_(p[2], c.depth);
        g.tbase[gi] = tshift;
#pragma unroll
        for (int k = 0; k < 6; k++) g.mhat[(size_t)k * ngroups_total + gi] = c.mhat[k];
        g.gw[gi] = 1.f;
        g.tap_begin[gi] = c.tap_begin;
        g.tap_count[gi] = c.nt;
        g.tt_begin[gi] = c.tt_begin + ip * c.nt;
        int lo = INT_MAX, hi = INT_MIN;
        for (int k = 0; k < c.nt; k++) {
            // time = tshift(ip) + toff(it) (source_bilat.f90:446); rshift = time/dt (seismogram.f90:139)
            int its = (int)floorf(D_(A_(tshift, taps.toff[c.tap_begin + k]), dt));
            lo = min(lo, its); hi = max(hi, its);
        }
        g.its_min[gi] = lo; g.its_max[gi] = hi;
    }
}

// generic: sample-shift range of groups whose position/taps were filled by the host
__global__ void k_group_tap_range(GroupSoA g, TapSoA taps, float dt, int gbegin, int gend) {
    int gi = gbegin + blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= gend) return;
    int lo = INT_MAX, hi = INT_MIN;
    const int tb = g.tap_begin[gi], tn = g.tap_count[gi];
    const float tbase = g.tbase[gi];
    for (int k = 0; k < tn; k++) {
        int its = (int)floorf(D_(A_(taps.toff[tb + k], tbase), dt));
        lo = min(lo, its); hi = max(hi, its);
    }
    g.its_min[gi] = lo; g.its_max[gi] = hi;
}

// Shift table: one warp per group, lane k = tap k.  What trace_multiply_add derives from the centroid time
// (sparse_trace.f90:639-646: rshift = time/dt, its = floor(rshift), wr = rshift-its, wl = 1-wr, both times the
// weight) does not depend on the receiver, so it is tabulated once per candidate -- per distinct quad shift
// m = its div 4 of the group, because the taps of one m update the same output quad of a lane:
//     out(4(q+m)+j) += sum_t h[t] * A(4q+j-t),  t = 0..4,   h[s] += wl, h[s+1] += wr for a tap with its mod 4 = s
// (summed in tap order), so k_synth reads and writes that quad once instead of once per tap.  W = sum of (wl+wr):
// the step every sample right of the group's window receives (:696-703).
__global__ void __launch_bounds__(256) k_tap_table(GroupSoA g, TapSoA taps, float dt, int ngroups) {
    const int gi = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (gi >= ngroups) return;
    const int tb = g.tap_begin[gi], tn_all = g.tap_count[gi], tt = g.tt_begin[gi];
    const unsigned lt = (1u << lane) - 1u;
    int nentries = 0, lo = INT_MAX, hi = INT_MIN;
    // rounds of 32 taps (source_bilat.f90:274-315 puts no bound on nt)
    for (int t0 = 0; t0 < tn_all; t0 += 32) {
        const int tn = min(tn_all - t0, 32);
        const bool on = lane < tn;
        int its = 0; float wl = 0.f, wr = 0.f;
        if (on) {
            const float time = A_(g.tbase[gi], taps.toff[tb + t0 + lane]);
            const float rshift = D_(time, dt);
            its = (int)floorf(rshift);
            const float wr0 = S_(rshift, (float)its);
            const float wl0 = S_(1.f, wr0);
            const float wt = taps.wt[tb + t0 + lane];
            wr = M_(wr0, wt); wl = M_(wl0, wt);
        }
        const int cls = its & 3, qoff = its >> 2;
        // leader = first tap of every distinct quad shift; it gathers the taps of its quad shift in tap order
        float h[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, W = 0.f;
        bool leader = on;
        for (int j = 0; j < tn; j++) {
            const float wlj = __shfl_sync(0xffffffffu, wl, j), wrj = __shfl_sync(0xffffffffu, wr, j);
            const int qj = __shfl_sync(0xffffffffu, qoff, j), sj = __shfl_sync(0xffffffffu, cls, j);
            if (qj == qoff) {
                if (j < lane) leader = false;
#pragma unroll
                for (int t = 0; t < 5; t++) { if (sj == t) h[t] += wlj; if (sj + 1 == t) h[t] += wrj; }
                W += wlj + wrj;
            }
        }
        // a quad shift that already has an entry from an earlier round is added to it (entries stay distinct: k_synth lets one lane
        // per entry update the end-value step, and two entries of one quad shift would meet in the same word)
        int found = -1;
        if (leader && t0 > 0)
            for (int i = 0; i < nentries && found < 0; i++)
                if (__float_as_int(g.taprec[2 * ((size_t)tt + i)].x) == qoff) found = i;
        if (leader && found >= 0) {
            float4* e = g.taprec + 2 * ((size_t)tt + found);
            const float4 a = e[0], b = e[1];
            e[0] = make_float4(a.x, a.y + h[0], a.z + h[1], a.w + h[2]);
            e[1] = make_float4(b.x + h[3], b.y + h[4], b.z + W, 0.f);
        }
        const bool fresh = leader && found < 0;
        const unsigned lm = __ballot_sync(0xffffffffu, fresh);
        if (fresh) {
            float4* e = g.taprec + 2 * ((size_t)tt + nentries + __popc(lm & lt));
            e[0] = make_float4(__int_as_float(qoff), h[0], h[1], h[2]);
            e[1] = make_float4(h[3], h[4], W, 0.f);
        }
        nentries += __popc(lm);
        __syncwarp();   // (the entries written in this round are read by the next one)
        lo = min(lo, warp_min_i(on ? its : INT_MAX)); hi = max(hi, warp_max_i(on ? its : INT_MIN));
    }
    if (lane == 0) { g.nstep[gi] = nentries; g.its_min[gi] = lo; g.its_max[gi] = hi; }
}

// expand the SoA of one candidate back into the reference's centroid table (test/inspection only)
__global__ void k_expand_centroids(CandDev cand, GroupSoA g, TapSoA taps, int ngroups_total, float* __restrict__ table, int cap) {
    for (int ig = blockIdx.x * blockDim.x + threadIdx.x; ig < cand.ngroups; ig += gridDim.x * blockDim.x) {
        const int gi = cand.group_begin + ig;
        const int tb = g.tap_begin[gi], tn = g.tap_count[gi];
        // centroid index = sum of tap counts of earlier groups; groups of one candidate all have
        // the same count for the source types built so far, eikonal stores explicit offsets later
        int first = 0;
        for (int j = 0; j < ig; j++) first += g.tap_count[cand.group_begin + j];
        for (int k = 0; k < tn; k++) {
            int id = first + k;
            if (id >= cap) break;
            float* t = table + (size_t)id * 10;
            t[0] = g.north[gi]; t[1] = g.east[gi]; t[2] = g.depth[gi];
            t[3] = A_(g.tbase[gi], taps.toff[tb + k]);
            for (int m = 0; m < 6; m++) t[4 + m] = M_(M_(g.mhat[(size_t)m * ngroups_total + gi], taps.wt[tb + k]), g.gw[gi]);
        }
    }
}

// =================================================================================================
// K2: geometry + index pre-pass.  One CTA per (candidate, receiver); threads stride over groups.
// =================================================================================================
__device__ __forceinline__ double clipd(double x, double mi, double ma) { return fmin(fmax(mi, x), ma); }
__device__ __forceinline__ double wrapd(double x, double mi, double ma) { return x - floor((x - mi) / (ma - mi)) * (ma - mi); }

// orthodrome.f90:77-156 (flat approximation and const-azimuth approximation are switched off by
// constants, :67,72; r == 0 still takes the const-azimuth branch because dist/0 = +Inf > huge)
__device__ void approx_differential_azidist(float delta_x, float delta_y, double azimuth, double backazimuth, double dist,
                                            double& new_azimuth, double& new_backazimuth, double& new_dist, const float* host_lambda = nullptr) {
    const double pi_ = (double)3.14159265358979f;       // constants.f90:22: default-real literal
    const double earthradius = (double)(6371.f * 1000.f);
    double r = (double)__fsqrt_rn(A_(M_(delta_x, delta_x), M_(delta_y, delta_y)));
    if (dist / r > DBL_MAX) {
        new_azimuth = azimuth;
        new_backazimuth = backazimuth;
        new_dist = dist - ((double)delta_x * cos(azimuth) + (double)delta_y * sin(azimuth));
    } else {
        double a = r / earthradius;
        double b = dist / earthradius;
        // (the host library's atan2f of the sub-source position, handed in: the device's own is an ulp off it often enough to move
        //  the epicentral distance of ~2 % of the sub-sources by one fp32 ulp -- at 50 km an ulp of the angle is a few millimetres)
        double lambda = host_lambda ? (double)*host_lambda : (double)atan2f(delta_y, delta_x);
        double gamma = azimuth - lambda;
        double sa, ca, sb, cb, sg, cg;
        sincos(a, &sa, &ca); sincos(b, &sb, &cb); sincos(gamma, &sg, &cg);
        double c = acos(clipd(__dadd_rn(__dmul_rn(ca, cb), __dmul_rn(__dmul_rn(sa, sb), cg)), -1., 1.));
        double sc = sin(c), cc = cos(c);
        double alpha = asin(clipd(__dmul_rn(sa, sg) / sc, -1., 1.));
        double beta = asin(clipd(__dmul_rn(sb, sg) / sc, -1., 1.));
        if (__dsub_rn(ca, __dmul_rn(cb, cc)) < 0) alpha = (alpha > 0) ? pi_ - alpha : -pi_ - alpha;
        if (__dsub_rn(cb, __dmul_rn(ca, cc)) < 0) beta = (beta > 0) ? pi_ - beta : -pi_ - beta;
        new_dist = c * earthradius;
        new_backazimuth = wrapd(backazimuth + alpha, -pi_, pi_);
        new_azimuth = wrapd(lambda - pi_ - beta, -pi_, pi_);
    }
}

struct SpanAcc {
    int lo, hi;
    __device__ void init() { lo = INT_MAX; hi = INT_MIN; }
    __device__ void add(int l, int h) { lo = min(lo, l); hi = max(hi, h); }
    __device__ void merge(const SpanAcc& o) { lo = min(lo, o.lo); hi = max(hi, o.hi); }
};
__device__ __forceinline__ int warp_min(int v) { for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
__device__ __forceinline__ int warp_max(int v) { for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }

// trace-span unions of the three component sets over the corners of one group (tabulated per node by kiwi_set_database)
struct SetSpans { int lo1, hi1, lo2, hi2, lo3, hi3; };
__device__ __forceinline__ SetSpans corner_spans(const GfdbDev& db, const int inode[4], int ncorner) {
    SetSpans u = {INT_MAX, INT_MIN, INT_MAX, INT_MIN, INT_MAX, INT_MIN};
    for (int c = 0; c < ncorner; c++) {
        const int4 a = __ldg(&db.nspan[2 * (size_t)inode[c]]), b = __ldg(&db.nspan[2 * (size_t)inode[c] + 1]);
        u.lo1 = min(u.lo1, a.x); u.hi1 = max(u.hi1, a.y); u.lo2 = min(u.lo2, a.z); u.hi2 = max(u.hi2, a.w);
        u.lo3 = min(u.lo3, b.x); u.hi3 = max(u.hi3, b.y);
    }
    return u;
}

// SINGLE: every candidate has one group (point sources: the 6 x nloc x nrcv basis syntheses of a moment-tensor grid search are
// 6e5 pairs of one group each): one THREAD per (candidate, receiver) instead of one CTA, no block reductions.
template <bool SINGLE>
__global__ void __launch_bounds__(256, 3) k_geometry(GfdbDev db, const ReceiverDev* __restrict__ rcv, int nrcv,
                                                   const CandDev* __restrict__ cands, GroupSoA g, int ngroups_total, int interpolate,
                                                   int xunder, int zunder, GeoRec* __restrict__ recs, size_t rec_stride,
                                                   PairHdr* __restrict__ hdrs, int* __restrict__ tmax, int npairs, int trig_only,
                                                   float* __restrict__ azf_out) {
    const int pair_ = SINGLE ? (int)(blockIdx.x * blockDim.x + threadIdx.x) : (int)blockIdx.x;
    const bool valid = !SINGLE || pair_ < npairs;      // (threads past the end stay for the warp shuffles at the bottom)
    const int pair = valid ? pair_ : 0;
    const int b = pair / nrcv, ir = pair % nrcv;
    const ReceiverDev& R = rcv[ir];
    const CandDev cand = cands[b];
    GeoRec* myrecs = recs + (size_t)pair * rec_stride;
    __shared__ int red[8][10];
    __shared__ int s_lastrot;

    const bool need_h = (R.ja | R.jr | R.jn | R.je) != 0;
    const bool need_v = R.jd != 0;

    SpanAcc urot, n1all, n2all, s3; urot.init(); n1all.init(); n2all.init(); s3.init();
    int last_rot = -1, any_nonrot = 0;

    if (valid && R.enabled && cand.status == 0) {
        for (int ip = SINGLE ? 0 : (int)threadIdx.x; ip < (SINGLE ? min(cand.ngroups, 1) : cand.ngroups); ip += SINGLE ? 1 : (int)blockDim.x) {
            const int gi = cand.group_begin + ip;
            const float dnorth = g.north[gi], deast = g.east[gi], depth = g.depth[gi];
            double azi, bazi, dist;
            approx_differential_azidist(dnorth, deast, R.azi0, R.bazi0, R.dist0, azi, bazi, dist, g.lam ? g.lam + gi : nullptr);
            GeoRec rec;
            {   // make_weights seismogram.f90:316-336 on the group's moment-tensor shape; the scalar tap
                // weight wt multiplies the result later (the reference applies it to m first)
                float m[6];
#pragma unroll
                const float gwt = g.gw[gi];
                for (int k = 0; k < 6; k++) m[k] = g.mhat[(size_t)k * ngroups_total + gi] * gwt;
                const float azf = (float)azi;
                float sa, ca, s2a, c2a;
                sincosf(azf, &sa, &ca);
                sincosf(2.f * azf, &s2a, &c2a);
                rec.f[0] = m[0] * (ca * ca) + m[1] * (sa * sa) + m[3] * s2a;
                rec.f[1] = m[4] * ca + m[5] * sa;
                rec.f[2] = m[2];
                rec.f[3] = 0.5f * (m[1] - m[0]) * s2a + m[3] * c2a;
                rec.f[4] = m[5] * ca - m[4] * sa;
                rec.f[5] = m[0] * (sa * sa) + m[1] * (ca * ca) - m[3] * s2a;
                // reference-order synthesis (synth_exact.cu) forms the weights per centroid itself: it gets the azimuth functions
                // (rounded from double: the device's sincosf may be an ulp off the host library's, which is correctly rounded nearly always;
                //  with azf_out the caller replaces them by the host library's own values, computed from the azimuth handed back)
                if (azf_out) azf_out[(size_t)pair * rec_stride + ip] = azf;
                if (trig_only) {
                    const double ad = (double)azf, a2 = (double)(2.f * azf);
                    rec.f[0] = (float)cos(ad); rec.f[1] = (float)sin(ad); rec.f[2] = (float)sin(a2); rec.f[3] = (float)cos(a2); rec.f[4] = 0.f; rec.f[5] = 0.f;
                }
                rec.tt_begin = g.tt_begin[gi]; rec.nstep = g.nstep[gi]; rec.pad[0] = 0;
            }
            const float x = (float)dist;
            const float z = S_(depth, R.depth);
            int ix1, iz1, ix2, iz2; float dix, diz; int flags = 0;
            if (interpolate) {  // gfdb_get_indices_bilin gfdb.f90:794-815
                const float denx = M_(db.dx, (float)xunder), denz = M_(db.dz, (float)zunder);
                const float ax = D_(S_(x, db.firstx), denx), az = D_(S_(z, db.firstz), denz);
                ix1 = (int)floorf(ax) * xunder + 1;
                iz1 = (int)floorf(az) * zunder + 1;
                ix2 = ix1 + xunder; iz2 = iz1 + zunder;
                dix = D_(S_(S_(x, db.firstx), M_((float)(ix1 - 1), db.dx)), denx);
                diz = D_(S_(S_(z, db.firstz), M_((float)(iz1 - 1), db.dz)), denz);
                // (only the distance goes through fp64 libm, where the device may differ from glibc in the last bits; the depth
                //  coordinate is exactly rounded fp32 arithmetic on both sides)
                const float fx = ax - floorf(ax);
                const float tx = 4.f * 1.1920929e-7f * fmaxf(fabsf(ax), 1.f);
                if (fx < tx || 1.f - fx < tx) flags |= GEO_NEAR;
            } else {            // gfdb_get_indices gfdb.f90:781-792 (nint: half away from zero)
                const float ax = D_(S_(x, db.firstx), db.dx), az = D_(S_(z, db.firstz), db.dz);
                ix1 = (int)roundf(ax) + 1; iz1 = (int)roundf(az) + 1;
                ix2 = ix1 + 1; iz2 = iz1 + 1; dix = 0.f; diz = 0.f;
                const float fx = fabsf(fabsf(ax - floorf(ax)) - 0.5f);
                const float tx = 4.f * 1.1920929e-7f * fmaxf(fabsf(ax), 1.f);
                if (fx < tx) flags |= GEO_NEAR;
            }
            const bool single = (dix == 0.f && diz == 0.f);
            if (single) flags |= GEO_SINGLE;
            // horizontal rotation (seismogram.f90:158-165)
            const double lambda = bazi - R.bazi0;
            rec.cl = 1.f; rec.sl = 0.f;
            if (lambda != 0.) { flags |= GEO_ROT; rec.cl = (float)cos(lambda); rec.sl = (float)sin(lambda); }
            // node availability (gfdb_get_trace gfdb.f90:843-855, chunk_get_trace :1005-1010)
            int inode[4]; int ncorner = single ? 1 : 4;
            const int cx[4] = {ix1, ix1, ix2, ix2}, cz[4] = {iz1, iz2, iz1, iz2};
            bool ok = true;
            for (int c = 0; c < 4; c++) {
                if (c >= ncorner) { rec.node[c] = rec.node[0]; continue; }
                if (cx[c] < 1 || cx[c] > db.nx || cz[c] < 1 || cz[c] > db.nz) { ok = false; inode[c] = 0; rec.node[c].off = ~0ull; rec.node[c].w0 = 0; rec.node[c].wn = 4; continue; }
                inode[c] = (cx[c] - 1) * db.nz + (cz[c] - 1);
                rec.node[c] = ld_node(&db.nodes[inode[c]]);
                if (rec.node[c].off == ~0ull) ok = false;
            }
            if (!ok) flags |= GEO_SKIP;
            rec.ix1 = ix1; rec.iz1 = iz1; rec.dix = dix; rec.diz = diz; rec.flags = flags;
            {   // 128-byte record, eight 128-bit stores
                const uint4* sp = reinterpret_cast<const uint4*>(&rec);
                uint4* dp = reinterpret_cast<uint4*>(myrecs + ip);
#pragma unroll
                for (int k = 0; k < 8; k++) dp[k] = sp[k];
            }
            if (ok && (need_h || need_v)) {
                const int smin = g.its_min[gi], smax = g.its_max[gi];
                const SetSpans u = corner_spans(db, inode, ncorner);
                if (need_h) {
                    const int lo1 = u.lo1 + smin, hi1 = u.hi1 + smax + 1, lo2 = u.lo2 + smin, hi2 = u.hi2 + smax + 1;   // sparse_trace.f90:649-653
                    if (flags & GEO_ROT) { urot.add(lo1, hi1); urot.add(lo2, hi2); last_rot = max(last_rot, ip); }
                    else { n1all.add(lo1, hi1); n2all.add(lo2, hi2); any_nonrot = 1; }
                }
                if (need_v) s3.add(u.lo3 + smin, u.hi3 + smax + 1);
            }
        }
    }
    if (SINGLE) {   // the pair's header straight from this thread's spans (S1 = Urot u N1, S2 = Urot u N2, see below)
        PairHdr h;
        h.s1lo = min(urot.lo, n1all.lo); h.s1hi = max(urot.hi, n1all.hi);
        h.s2lo = min(urot.lo, n2all.lo); h.s2hi = max(urot.hi, n2all.hi);
        h.s3lo = s3.lo; h.s3hi = s3.hi;
        const int lo = min(min(h.s1lo, h.s2lo), h.s3lo), hi = max(max(h.s1hi, h.s2hi), h.s3hi);
        if (hi < lo || !valid) { h.out0 = 0; h.T = 0; } else { h.out0 = lo; h.T = hi - lo + 1; }
        if (valid) hdrs[pair] = h;
        const int wt = warp_max(h.T), wlo = warp_min(h.T > 0 ? lo : INT_MAX), whi = warp_max(h.T > 0 ? hi : INT_MIN);
        if ((threadIdx.x & 31) == 0 && wt > 0) { atomicMax(tmax, wt); atomicMin(tmax + 1, wlo); atomicMax(tmax + 2, whi); }
        return;
    }
    // block reduction
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int v[10] = {urot.lo, n1all.lo, n2all.lo, s3.lo, INT_MAX, urot.hi, n1all.hi, n2all.hi, s3.hi, last_rot};
#pragma unroll
    for (int i = 0; i < 5; i++) v[i] = warp_min(v[i]);
#pragma unroll
    for (int i = 5; i < 10; i++) v[i] = warp_max(v[i]);
    any_nonrot = __syncthreads_or(any_nonrot);
    if (lane == 0) for (int i = 0; i < 10; i++) red[wid][i] = v[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) {
            for (int i = 0; i < 5; i++) red[0][i] = min(red[0][i], red[w][i]);
            for (int i = 5; i < 10; i++) red[0][i] = max(red[0][i], red[w][i]);
        }
        s_lastrot = red[0][9];
    }
    __syncthreads();
    // non-rotating groups that precede the last rotating one also widen the *other* strip, because
    // strip_extend_to_same_span_4 (seismogram.f90:196) unifies all four strips at every rotating centroid
    SpanAcc n1before, n2before; n1before.init(); n2before.init();
    if (any_nonrot && s_lastrot >= 0 && need_h) {
        for (int ip = threadIdx.x; ip < s_lastrot; ip += blockDim.x) {
            const GeoRec rec = myrecs[ip];
            if ((rec.flags & (GEO_ROT | GEO_SKIP)) != 0) continue;
            const int gi = cand.group_begin + ip;
            const bool single = rec.flags & GEO_SINGLE;
            const int ix2 = rec.ix1 + (interpolate ? xunder : 1), iz2 = rec.iz1 + (interpolate ? zunder : 1);
            const int cx[4] = {rec.ix1, rec.ix1, ix2, ix2}, cz[4] = {rec.iz1, iz2, rec.iz1, iz2};
            int inode[4]; const int ncorner = single ? 1 : 4;
            for (int c = 0; c < ncorner; c++) inode[c] = (cx[c] - 1) * db.nz + (cz[c] - 1);
            const int smin = g.its_min[gi], smax = g.its_max[gi];
            const SetSpans u = corner_spans(db, inode, ncorner);
            n1before.add(u.lo1 + smin, u.hi1 + smax + 1); n2before.add(u.lo2 + smin, u.hi2 + smax + 1);
        }
    }
    int w4[4] = {n1before.lo, n2before.lo, n1before.hi, n2before.hi};
    w4[0] = warp_min(w4[0]); w4[1] = warp_min(w4[1]); w4[2] = warp_max(w4[2]); w4[3] = warp_max(w4[3]);
    __shared__ int red2[8][4];
    if (lane == 0) for (int i = 0; i < 4; i++) red2[wid][i] = w4[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) {
            red2[0][0] = min(red2[0][0], red2[w][0]); red2[0][1] = min(red2[0][1], red2[w][1]);
            red2[0][2] = max(red2[0][2], red2[w][2]); red2[0][3] = max(red2[0][3], red2[w][3]);
        }
        PairHdr h;
        // S1 = Urot u N1(all) u N2(before last rot), S2 = Urot u N2(all) u N1(before last rot)
        h.s1lo = min(min(red[0][0], red[0][1]), red2[0][1]); h.s1hi = max(max(red[0][5], red[0][6]), red2[0][3]);
        h.s2lo = min(min(red[0][0], red[0][2]), red2[0][0]); h.s2hi = max(max(red[0][5], red[0][7]), red2[0][2]);
        h.s3lo = red[0][3]; h.s3hi = red[0][8];
        int lo = min(min(h.s1lo, h.s2lo), h.s3lo), hi = max(max(h.s1hi, h.s2hi), h.s3hi);
        if (hi < lo) { h.out0 = 0; h.T = 0; } else { h.out0 = lo; h.T = hi - lo + 1; }
        hdrs[pair] = h;
        if (h.T > 0) { atomicMax(tmax, h.T); atomicMin(tmax + 1, lo); atomicMax(tmax + 2, hi); }
    }
}

// =================================================================================================
// K3: synthesis.  One CTA per (candidate, receiver); every warp owns a private set of three
// accumulator strips (displacement_ar(1), displacement_ar(2), vertical) in shared memory and works
// through its share of the groups.  Per group and 128-sample chunk a lane owns one aligned sample
// quad: 4 corners x ng rows are staged HBM -> shared memory with 128-bit cp.async copies (a ring of
// SYN_STAGES components per warp that keeps running across chunk and group boundaries), combined
// bilinearly (gfdb.f90:943-948), weighted with the moment-tensor/azimuth factors (make_weights
// seismogram.f90:316-336, precomputed by k_geometry), rotated by the centroid's back-azimuth
// difference (:196-203), and then added nt times with the sample shift and linear sub-sample
// interpolation of trace_multiply_add (sparse_trace.f90:639-705; shifts and weights tabulated by
// k_tap_table).  The "last sample repeats for ever" rule (:696-703) becomes a step per distinct quad
// shift of the group that is prefix-summed once at the end.
// All fp32 arithmetic of the inner loops is issued as packed pairs (FFMA2, fma.rn.f32x2): same
// roundings as the scalar fma, half the issue slots.
// =================================================================================================
#define SYN_STAGES 3      // ring depth per warp: items (one GF component x four corners) in flight
#define SYN_ITEM_BYTES (4 * 32 * 16)

typedef unsigned long long u64;
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
struct Q2 { u64 lo, hi; };   // one sample quad as two packed fp32 pairs (x,y) (z,w)

__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpk2(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
// acc (+)= (s,s) * v, both halves rounded like fmaf
__device__ __forceinline__ void ffma2(u64& acc, float s, u64 v) {
    asm("{\n\t.reg .b64 ss;\n\tmov.b64 ss, {%1, %1};\n\tfma.rn.f32x2 %0, ss, %2, %0;\n\t}" : "+l"(acc) : "f"(s), "l"(v));
}
__device__ __forceinline__ void q2_fma(Q2& a, float s, const Q2& v) { ffma2(a.lo, s, v.lo); ffma2(a.hi, s, v.hi); }
__device__ __forceinline__ Q2 q2_zero() { Q2 r; r.lo = 0ull; r.hi = 0ull; return r; }
template <int OFF>
__device__ __forceinline__ Q2 lds_q2(unsigned addr) {
    Q2 r;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2+%3];" : "=l"(r.lo), "=l"(r.hi) : "r"(addr), "n"(OFF) : "memory");
    return r;
}
// the four corner quads of one ring slot; lanes with on == 0 do not read (and keep what t0..t3 held)
__device__ __forceinline__ void lds_ring(unsigned addr, int on, Q2& t0, Q2& t1, Q2& t2, Q2& t3) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %9, 0;\n\t"
        "@p ld.shared.v2.b64 {%0, %1}, [%8];\n\t@p ld.shared.v2.b64 {%2, %3}, [%8+512];\n\t"
        "@p ld.shared.v2.b64 {%4, %5}, [%8+1024];\n\t@p ld.shared.v2.b64 {%6, %7}, [%8+1536];\n\t}"
        : "+l"(t0.lo), "+l"(t0.hi), "+l"(t1.lo), "+l"(t1.hi), "+l"(t2.lo), "+l"(t2.hi), "+l"(t3.lo), "+l"(t3.hi)
        : "r"(addr), "r"(on)
        : "memory");
}
__device__ __forceinline__ float4 lds_f4(unsigned addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr) : "memory");
    return r;
}
__device__ __forceinline__ void sts_q2(unsigned addr, const Q2& v) {
    asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(addr), "l"(v.lo), "l"(v.hi) : "memory");
}
__device__ __forceinline__ u64 shfl_up_u64(u64 v, int d) {
    float a, b; unpk2(v, a, b);
    return pk2(__shfl_up_sync(0xffffffffu, a, d), __shfl_up_sync(0xffffffffu, b, d));
}
__device__ __forceinline__ u64 shfl_u64(u64 v, int src) {
    float a, b; unpk2(v, a, b);
    return pk2(__shfl_sync(0xffffffffu, a, src), __shfl_sync(0xffffffffu, b, src));
}

// ---- asynchronous staging (cp.async): group records and GF quads travel HBM -> shared memory without
// passing through registers, so a warp keeps SYN_STAGES x 4 128-bit loads in flight also while it is
// busy with the tap phase of the previous chunk -------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16s(unsigned smem_addr, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// The groups a CTA walks in one launch.  k_synth can be launched in `nbands` depth bands: every launch then covers the sub-faults of
// a few depth rows of the database only, so that the slice all its CTAs gather from stays resident in L2.
// ny > 0 (the groups form an nx x ny lattice listed with the down-dip index fastest, source_bilat.f90:349-371): the lattice rows
// [row0, row0 + rows); ny == 0 (listed row by row, or in no particular order): the slice [first, first + n) of the list.
struct GroupWalk {
    int n;              // groups walked
    int ny, row0, rows, first;
    __device__ __forceinline__ int rec(int idx) const { return ny > 0 ? (idx / rows) * ny + row0 + idx % rows : first + idx; }
};
// copy of one 128-byte group record (lanes 0..7, 16 bytes each); not committed here
__device__ __forceinline__ void rec_copy_async(const GeoRec* __restrict__ recs, int idx, const GroupWalk& gw, GeoRec* dst_slot, int lane) {
    if (idx < gw.n && lane < 8) cp_async16(reinterpret_cast<uint4*>(dst_slot) + lane, reinterpret_cast<const uint4*>(recs + gw.rec(idx)) + lane);
}

// GF components a receiver with horizontal (H) / vertical (V) components needs, in the reference's order of
// accumulation (seismogram.f90:167-250): g1 g2 g3 (g9) -> radial, g4 g5 -> transverse, g6 g7 g8 (g10) -> vertical
template <bool H, bool V, bool NG10>
struct CompSeq {
    static constexpr int N = (H && V) ? (NG10 ? 10 : 8) : (H ? (NG10 ? 6 : 5) : (NG10 ? 4 : 3));
    __host__ __device__ static constexpr int comp(int j) { return (H && V) ? j : (H ? (j < 5 ? j : 8) : (j < 3 ? 5 + j : 9)); }
    __host__ __device__ static constexpr int step(int j) { return j + 1 < N ? comp(j + 1) - comp(j) : 0; }   // rows from item j to item j+1
};

// where a lane reads a chunk from: for the four corners the address of the lane's quad in the row of the
// component that is issued next (clamped into each window: quad 0 holds zeros, the last quad the continuation)
// and the row strides in quads
struct ChunkSrc {
    const char *p0, *p1, *p2, *p3;
    unsigned s0, s1, s2, s3;
    bool active;
};
__device__ __forceinline__ const char* corner_ptr(const float* slabs, const NodeInfo& n, int q, int comp0, unsigned& stride_quads) {
    const int nq = n.wn >> 2;
    stride_quads = (unsigned)nq;
    const int qi = min(max(q - (n.w0 >> 2), 0), nq - 1) + comp0 * nq;
    return reinterpret_cast<const char*>(slabs + n.off) + ((size_t)(unsigned)qi << 4);
}
__device__ __forceinline__ void window_quads(const NodeInfo& n0, const NodeInfo& n1, const NodeInfo& n2, const NodeInfo& n3, int& q_first,
                                             int& q_last) {
    // every window starts with one quad of zeros (left continuation): the earliest window's zero quad need not be read
    q_first = (min(min(n0.w0, n1.w0), min(n2.w0, n3.w0)) >> 2) + 1;
    // last quad of the longest window: continuation only, for every corner
    q_last = (max(max(n0.w0 + n0.wn, n1.w0 + n1.wn), max(n2.w0 + n2.wn, n3.w0 + n3.wn)) >> 2) - 1;
}
// source of the chunk starting at quad q0 (or, q0 < 0, of the first chunk) of the group whose record is *rec;
// also returns the group's quad range.  Lane 0 re-reads the quad left of the chunk (for the first chunk: the zeros
// left of the windows), lanes 1..31 own the chunk's 31 quads.
__device__ __forceinline__ void chunk_src(ChunkSrc& c, const float* slabs, const GeoRec* rec, int q0, int lane, int comp0, int& q_first, int& q_last) {
    const NodeInfo n0 = rec->node[0], n1 = rec->node[1], n2 = rec->node[2], n3 = rec->node[3];
    window_quads(n0, n1, n2, n3, q_first, q_last);
    const int q = (q0 < 0 ? q_first : q0) + lane - 1;
    c.p0 = corner_ptr(slabs, n0, q, comp0, c.s0);
    c.p1 = corner_ptr(slabs, n1, q, comp0, c.s1);
    c.p2 = corner_ptr(slabs, n2, q, comp0, c.s2);
    c.p3 = corner_ptr(slabs, n3, q, comp0, c.s3);
    c.active = q <= q_last;
}
// (multiply-add on the whole 64-bit pointer: the result is a register pair the copy can use as it is)
__device__ __forceinline__ void advance_rows(const char*& p, unsigned stride_quads, int rows) {
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(p) : "r"(stride_quads), "r"(16u * (unsigned)rows));
}
// one item = the next GF component of the chunk -> ring slot (4 corners x 32 lanes x 16 bytes); the source then moves
// on by `dcomp` rows
__device__ __forceinline__ void issue_item(ChunkSrc& c, unsigned slot, int dcomp /* constant after unrolling */) {
    if (c.active) {
        cp_async16s(slot, c.p0);
        cp_async16s(slot + 32 * 16, c.p1);
        cp_async16s(slot + 64 * 16, c.p2);
        cp_async16s(slot + 96 * 16, c.p3);
    }
    if (dcomp != 0) {   // running pointers (opaque to the compiler, which would otherwise re-derive every address from the chunk base)
        advance_rows(c.p0, c.s0, dcomp); advance_rows(c.p1, c.s1, dcomp); advance_rows(c.p2, c.s2, dcomp); advance_rows(c.p3, c.s3, dcomp);
    }
}

// out quad(q + m) (+)= sum_t h[t] * A(4q + j - t): E[i] = samples 2i, 2i+1 of the eight samples (previous quad, own quad),
// G[i] = samples 2i+1, 2i+2.  Lanes without a quad of their own (right of the windows, lane 0, outside the strips; ok == 0)
// are handed a harmless address to read and do not store.
__device__ __forceinline__ void rmw_quad(unsigned addr, int ok, const u64* E, const u64* G, float h0, float h1, float h2, float h3, float h4) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 a0, a1, w;\n\t"
        "setp.ne.s32 p, %1, 0;\n\t"
        "ld.shared.v2.b64 {a0, a1}, [%0];\n\t"
        "mov.b64 w, {%2, %2};\n\tfma.rn.f32x2 a0, w, %9, a0;\n\tfma.rn.f32x2 a1, w, %10, a1;\n\t"
        "mov.b64 w, {%3, %3};\n\tfma.rn.f32x2 a0, w, %12, a0;\n\tfma.rn.f32x2 a1, w, %13, a1;\n\t"
        "mov.b64 w, {%4, %4};\n\tfma.rn.f32x2 a0, w, %8, a0;\n\tfma.rn.f32x2 a1, w, %9, a1;\n\t"
        "mov.b64 w, {%5, %5};\n\tfma.rn.f32x2 a0, w, %11, a0;\n\tfma.rn.f32x2 a1, w, %12, a1;\n\t"
        "mov.b64 w, {%6, %6};\n\tfma.rn.f32x2 a0, w, %7, a0;\n\tfma.rn.f32x2 a1, w, %8, a1;\n\t"
        "@p st.shared.v2.b64 [%0], {a0, a1};\n\t}"
        ::"r"(addr), "r"(ok), "f"(h0), "f"(h1), "f"(h2), "f"(h3), "f"(h4), "l"(E[0]), "l"(E[1]), "l"(E[2]), "l"(E[3]), "l"(G[0]), "l"(G[1]), "l"(G[2])
        : "memory");
}
// nz = -0.0f handed in as a kernel argument: x + nz == x bit for bit, but the assembler cannot see that and so gives the
// odd pairs registers of their own once per chunk, instead of re-assembling them from the halves of E at every use
__device__ __forceinline__ void make_pairs(const Q2& P, const Q2& A, u64* E, u64* G, float nz) {
    float p0, p1, p2, p3, a0, a1, a2, a3;
    unpk2(P.lo, p0, p1); unpk2(P.hi, p2, p3); unpk2(A.lo, a0, a1); unpk2(A.hi, a2, a3);
    E[0] = P.lo; E[1] = P.hi; E[2] = A.lo; E[3] = A.hi;
    G[0] = pk2(__fadd_rn(p1, nz), __fadd_rn(p2, nz)); G[1] = pk2(__fadd_rn(p3, nz), __fadd_rn(a0, nz)); G[2] = pk2(__fadd_rn(a1, nz), __fadd_rn(a2, nz));
    (void)p0; (void)a3;
}

// One warp works through its share of the groups of one (candidate, receiver) pair.
template <bool H, bool V, bool NG10>
__device__ __forceinline__ void synth_warp(const GfdbDev& db, const GeoRec* __restrict__ myrecs, const GroupWalk walk, const float4* __restrict__ taprec,
                                           float sd, unsigned acc_s /* shared address of the warp's strips */, unsigned strip_bytes,
                                           float* __restrict__ step, int nq, int baseq, GeoRec* slot /* [3] */,
                                           unsigned ring_s /* shared address of the warp's ring */, int warp, int nwarps, int lane, float nz) {
    typedef CompSeq<H, V, NG10> Seq;
    constexpr int N = Seq::N, S = SYN_STAGES;
    static_assert(N >= S && S == 3, "ring depth");
    // records of the first two groups synchronously; from then on two groups ahead
    const int ngroups = walk.n;
    rec_copy_async(myrecs, warp, walk, slot, lane);
    rec_copy_async(myrecs, warp + nwarps, walk, slot + 1, lane);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    // ring slots of this lane in the order the next items use them (rotated after every chunk)
    unsigned r0 = ring_s + lane * 16, r1 = r0 + SYN_ITEM_BYTES, r2 = r1 + SYN_ITEM_BYTES;
    bool primed = false;    // the first S items of the chunk about to be processed are already in flight
    ChunkSrc src;           // where the items still to be issued come from
    src.active = false;
    int qf_n = 0, ql_n = 0; // quad range of the group the look-ahead has started
    int sl = 0;             // slot of the current group's record
    for (int ip = warp; ip < ngroups; ip += nwarps, sl = (sl == 2 ? 0 : sl + 1)) {
        // record of the group after next; rides in the next commit group
        rec_copy_async(myrecs, ip + 2 * nwarps, walk, slot + (sl == 0 ? 2 : sl - 1), lane);
        const GeoRec* rec = slot + sl;
        const int flags = rec->flags;
        if (flags & GEO_SKIP) {   // (never primed: the look-ahead does not cross skipped groups)
            cp_async_commit(); cp_async_wait<0>(); __syncwarp();   // rare path: make sure the records fetched ahead have landed
            continue;
        }
        // ---- corners (gfdb.f90:943-948 weights in the reference's association) -------------------------
        // next group of this warp (for the look-ahead at the end of this group's last chunk)
        const GeoRec* nrec = slot + (sl == 2 ? 0 : sl + 1);
        // (whether the look-ahead may run into the next group is decided where it is needed, inside the last chunk: the record
        //  after next travels with item 0 of this group's first chunk and is only complete after that item's wait)
        const int tt = rec->tt_begin, nstep = rec->nstep;
        const unsigned rec_s = (unsigned)__cvta_generic_to_shared(rec);
        int q_first = qf_n, q_last = ql_n;
        if (!primed) {   // pipeline (re)start: first S items of this group
            chunk_src(src, db.slabs, rec, -1, lane, Seq::comp(0), q_first, q_last);
            issue_item(src, r0, Seq::step(0)); cp_async_commit();
            issue_item(src, r1, Seq::step(1)); cp_async_commit();
            issue_item(src, r2, Seq::step(2)); cp_async_commit();
        }

        // lane m keeps the m-th distinct quad shift of the group (tabulated by k_tap_table); broadcast by shuffles below
        int my_q = 0; float my_h0 = 0.f, my_h1 = 0.f, my_h2 = 0.f, my_h3 = 0.f, my_h4 = 0.f, my_W = 0.f;
        if (lane < nstep) {
            const float4 ta = __ldg(taprec + 2 * ((size_t)tt + lane)), tb = __ldg(taprec + 2 * ((size_t)tt + lane) + 1);
            my_q = __float_as_int(ta.x); my_h0 = ta.y; my_h1 = ta.z; my_h2 = ta.w; my_h3 = tb.x; my_h4 = tb.y; my_W = tb.z;
        }

        for (int q0 = q_first; q0 <= q_last; q0 += 31) {
            const int q = q0 + lane - 1;   // lane 0: the quad left of the chunk, only read
            const bool active = lane > 0 && q <= q_last;
            const bool more = q0 + 31 <= q_last;
            bool have_next = more;
            const int lane_on = q <= q_last;   // lanes that hold a quad of the windows (lane 0: the quad left of the chunk)
            // the group's weights are re-read from its record for every chunk rather than kept in registers across the tap loops
            float wc0, wc1, wc2, wc3, f1, f2, f3, f4, f5, f6, cl, sl_;
            {
                const float4 ra = lds_f4(rec_s), rb = lds_f4(rec_s + 16), rc = lds_f4(rec_s + 32);
                // corners (gfdb.f90:943-948 weights in the reference's association)
                const bool single = flags & GEO_SINGLE;
                const float dix = ra.z, diz = ra.w;
                wc0 = single ? 1.f : (1.f - dix) * (1.f - diz); wc1 = single ? 0.f : (1.f - dix) * diz;
                wc2 = single ? 0.f : dix * (1.f - diz); wc3 = single ? 0.f : dix * diz;
                f1 = rb.x; f2 = rb.y; f3 = rb.z; f4 = rb.w; f5 = rc.x; f6 = rc.y; cl = rc.z; sl_ = rc.w;
            }
            Q2 A1 = q2_zero(), A2 = q2_zero(), A3 = q2_zero(), Rr = q2_zero(), Tt = q2_zero();
            Q2 t0 = q2_zero(), t1 = q2_zero(), t2 = q2_zero(), t3 = q2_zero();   // (defined here so that the predicated reads below do not keep them alive across chunks)
#pragma unroll
            for (int j = 0; j < N; j++) {
                const unsigned rs = (j % 3 == 0) ? r0 : (j % 3 == 1 ? r1 : r2);
                cp_async_wait<S - 1>();    // item j has landed (this lane's own copies; no other lane reads them)
                {   // (lanes right of the windows skip the reads and combine whatever their registers hold; nothing of it is stored)
                    lds_ring(rs, lane_on, t0, t1, t2, t3);
                    Q2 r = q2_zero();
                    q2_fma(r, wc0, t0); q2_fma(r, wc1, t1); q2_fma(r, wc2, t2); q2_fma(r, wc3, t3);   // a read clamped to quad 0 of a row returns the zeros left of the trace
                    const int k = Seq::comp(j);   // constant after unrolling
                    if (k == 0) q2_fma(Rr, f1, r);
                    else if (k == 1) q2_fma(Rr, f2, r);
                    else if (k == 2) q2_fma(Rr, f3, r);
                    else if (k == 3) q2_fma(Tt, f4, r);
                    else if (k == 4) q2_fma(Tt, f5, r);
                    else if (k == 5) q2_fma(A3, f1 * sd, r);
                    else if (k == 6) q2_fma(A3, f2 * sd, r);
                    else if (k == 7) q2_fma(A3, f3 * sd, r);
                    else if (k == 8) q2_fma(Rr, f6, r);
                    else q2_fma(A3, f6 * sd, r);
                }
                // refill the slot just consumed: a later item of this chunk, or one of the first S items of the next chunk
                if (j + S < N) {
                    issue_item(src, rs, Seq::step(j + S));
                } else {
                    const int jj = j + S - N;     // constant after unrolling, < S
                    if (jj == 0) {                // all items of this chunk are on their way: move the source on
                        if (!more) {
                            __syncwarp();         // lanes 0..7 have waited for their copies of the next record: visible to all lanes now
                            have_next = (ip + nwarps < ngroups) && !(nrec->flags & GEO_SKIP);
                        }
                        if (have_next) {
                            int qf, ql;
                            chunk_src(src, db.slabs, more ? rec : nrec, more ? q0 + 31 : -1, lane, Seq::comp(0), qf, ql);
                            if (!more) { qf_n = qf; ql_n = ql; }
                        }
                    }
                    if (have_next) issue_item(src, rs, Seq::step(jj));
                }
                cp_async_commit();
            }
            {   // the next chunk's item 0 uses the slot after the one item N-1 used
                const unsigned a = r0, b = r1, c = r2;
                if (N % 3 == 1) { r0 = b; r1 = c; r2 = a; }
                else if (N % 3 == 2) { r0 = c; r1 = a; r2 = b; }
            }
            primed = have_next;
            if (H) {   // seismogram.f90:200-203: ar1 += cl*temp1 - sl*temp2; ar2 += cl*temp2 + sl*temp1
                q2_fma(A1, cl, Rr); q2_fma(A1, -sl_, Tt);
                q2_fma(A2, cl, Tt); q2_fma(A2, sl_, Rr);
            }
            // previous quad: lane-1
            u64 E1[4], G1[3], E2[4], G2[3], E3[4], G3[3];
            if (H) {
                Q2 P1, P2;
                P1.lo = shfl_up_u64(A1.lo, 1); P1.hi = shfl_up_u64(A1.hi, 1); P2.lo = shfl_up_u64(A2.lo, 1); P2.hi = shfl_up_u64(A2.hi, 1);
                make_pairs(P1, A1, E1, G1, nz); make_pairs(P2, A2, E2, G2, nz);
            }
            if (V) {
                Q2 P3;
                P3.lo = shfl_up_u64(A3.lo, 1); P3.hi = shfl_up_u64(A3.hi, 1);
                make_pairs(P3, A3, E3, G3, nz);
            }
            // ---- taps (sparse_trace.f90:647-695): one read-modify-write of the strips per distinct quad shift ----------
            const int qb = q - baseq;
            const int srcl = q_last - q0 + 1;   // lane that holds the last quad of the windows (last chunk only)
            float e1 = 0.f, e2 = 0.f, e3 = 0.f;
            if (!more) {
                float d0;
                if (H) { unpk2(A1.hi, d0, e1); unpk2(A2.hi, d0, e2); e1 = __shfl_sync(0xffffffffu, e1, srcl); e2 = __shfl_sync(0xffffffffu, e2, srcl); }
                if (V) { unpk2(A3.hi, d0, e3); e3 = __shfl_sync(0xffffffffu, e3, srcl); }
                (void)d0;
            }
            // the entries [r0, r0 + nr) of the group's shift table, entry m held by lane m
            auto apply_steps = [&](int nr, int t_q, float t_h0, float t_h1, float t_h2, float t_h3, float t_h4, float t_W) {
                for (int m = 0; m < nr; m++) {
                    const int qrel = qb + __shfl_sync(0xffffffffu, t_q, m);
                    const float h0 = __shfl_sync(0xffffffffu, t_h0, m), h1 = __shfl_sync(0xffffffffu, t_h1, m), h2 = __shfl_sync(0xffffffffu, t_h2, m),
                                h3 = __shfl_sync(0xffffffffu, t_h3, m), h4 = __shfl_sync(0xffffffffu, t_h4, m);
                    const int ok = active && (unsigned)qrel < (unsigned)nq;
                    // lanes without a quad read the group's record instead (nobody writes it now) and do not store
                    const unsigned a = acc_s + ((unsigned)qrel << 4);
                    if (H) { rmw_quad(ok ? a : rec_s, ok, E1, G1, h0, h1, h2, h3, h4); rmw_quad(ok ? a + strip_bytes : rec_s, ok, E2, G2, h0, h1, h2, h3, h4); }
                    if (V) rmw_quad(ok ? a + 2 * strip_bytes : rec_s, ok, E3, G3, h0, h1, h2, h3, h4);
                    __syncwarp();
                }
                if (!more) {
                    // ---- end-value repetition (sparse_trace.f90:696-703): every sample right of the last processed quad gets
                    // (wl+wr)*A_end; recorded as a step at quad q_last+1+shift, prefix-summed at the end.  Lane j owns the j-th
                    // distinct quad shift of the round: no two lanes share a word.
                    const int qs = q_last + 1 + t_q - baseq;
                    if (lane < nr && (unsigned)qs < (unsigned)nq) {
                        if (H) { step[qs] += t_W * e1; step[nq + qs] += t_W * e2; }
                        if (V) step[2 * nq + qs] += t_W * e3;
                    }
                    __syncwarp();
                }
            };
            apply_steps(min(nstep, 32), my_q, my_h0, my_h1, my_h2, my_h3, my_h4, my_W);
            for (int r0 = 32; r0 < nstep; r0 += 32) {   // more than 32 distinct quad shifts (long rise times): further rounds
                int t_q = 0; float t_h0 = 0.f, t_h1 = 0.f, t_h2 = 0.f, t_h3 = 0.f, t_h4 = 0.f, t_W = 0.f;
                if (r0 + lane < nstep) {
                    const float4 ta = __ldg(taprec + 2 * ((size_t)tt + r0 + lane)), tb = __ldg(taprec + 2 * ((size_t)tt + r0 + lane) + 1);
                    t_q = __float_as_int(ta.x); t_h0 = ta.y; t_h1 = ta.z; t_h2 = ta.w; t_h3 = tb.x; t_h4 = tb.y; t_W = tb.z;
                }
                apply_steps(min(nstep - r0, 32), t_q, t_h0, t_h1, t_h2, t_h3, t_h4, t_W);
            }
        }
        __syncwarp();
    }
    cp_async_wait<0>();
}

__global__ void __launch_bounds__(256, 2) k_synth(GfdbDev db, const ReceiverDev* __restrict__ rcv, int nrcv,
                                                   const CandDev* __restrict__ cands, GroupSoA g, const GeoRec* __restrict__ recs,
                                                   size_t rec_stride, const PairHdr* __restrict__ hdrs, int nq_alloc, int margin_q,
                                                   float* __restrict__ seis, size_t seis_stride /* floats per component row */,
                                                   SeisHdr* __restrict__ shdrs, float neg_zero /* -0.0f, see make_pairs */,
                                                   int band, float4* __restrict__ partial /* [pair][3*nq float4 + 3*nq float] */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // CTAs that run at the same time work on the same receiver for neighbouring candidates: candidates of a grid search
    // that share (part of) their sub-fault geometry then find each other's Green's function rows in L2
    const int ncand = gridDim.x / nrcv;
    const int ir = blockIdx.x / ncand, b = blockIdx.x % ncand;
    const int pair = b * nrcv + ir;
    const ReceiverDev& R = rcv[ir];
    const CandDev cand = cands[b];
    const PairHdr H = hdrs[pair];
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    SeisHdr* myshdr = shdrs + (size_t)pair * KIWI_MAX_COMP;
    if (H.T <= 0 || !R.enabled) {
        if (threadIdx.x < KIWI_MAX_COMP) { SeisHdr e; e.lo = 0; e.hi = -1; e.base = 0; e.pad = 0; myshdr[threadIdx.x] = e; }
        return;
    }
    // depth bands: launch `band` adds band `band` of every candidate that has that many to the pair's partial strips; the launch of
    // a candidate's last band finishes its seismograms
    const int nbands = cand.nbands;
    if (band >= nbands) return;
    GroupWalk walk;
    walk.n = cand.ngroups; walk.ny = 0; walk.row0 = 0; walk.rows = 1; walk.first = 0;
    if (nbands > 1) {
        if (cand.walk_ny > 0) {
            const int nx = cand.ngroups / cand.walk_ny;
            walk.ny = cand.walk_ny; walk.row0 = (int)(((long long)walk.ny * band) / nbands);
            walk.rows = (int)(((long long)walk.ny * (band + 1)) / nbands) - walk.row0;
            walk.n = nx * walk.rows;
            if (walk.rows <= 0) { walk.rows = 1; walk.n = 0; }
        } else {
            walk.first = (int)(((long long)cand.ngroups * band) / nbands);
            walk.n = (int)(((long long)cand.ngroups * (band + 1)) / nbands) - walk.first;
        }
    }
    const int base = floor4(H.out0) - 4 * margin_q;   // room on the left for the rise-time fold (k_fold)
    const int baseq = base >> 2;
    const int nq = nq_alloc;    // quads per accumulator strip
    const int nqs = nq;
    // shared memory: per warp 3 strips of nqs float4 + 3 step rows of nq floats, 3 group records, the cp.async ring
    float4* acc_all = reinterpret_cast<float4*>(smem_raw);
    float* step_all = reinterpret_cast<float*>(acc_all + (size_t)nwarps * 3 * nqs);
    float4* acc = acc_all + (size_t)warp * 3 * nqs;
    float* step = step_all + (size_t)warp * 3 * nq;
    for (int i = lane; i < 3 * nqs; i += 32) acc[i] = f4zero();
    for (int i = lane; i < 3 * nq; i += 32) step[i] = 0.f;
    __syncwarp();

    const bool need_h = (R.ja | R.jr | R.jn | R.je) != 0;
    const bool need_v = R.jd != 0;
    const bool ng10 = db.ng == 10;
    const GeoRec* myrecs = recs + (size_t)pair * rec_stride;
    // 16-byte aligned carve-up behind the step rows (3*nq floats per warp may end on an 8-byte boundary)
    unsigned char* tail = reinterpret_cast<unsigned char*>(step_all + (size_t)nwarps * 3 * nq);
    tail += (16 - (reinterpret_cast<size_t>(tail) & 15)) & 15;
    GeoRec* slot = reinterpret_cast<GeoRec*>(tail) + 3 * warp;
    float4* ring = reinterpret_cast<float4*>(reinterpret_cast<GeoRec*>(tail) + 3 * nwarps) + (size_t)warp * SYN_STAGES * 4 * 32;
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);

#define KIWI_SYNTH(HH, VV, NG) \
    synth_warp<HH, VV, NG>(db, myrecs, walk, g.taprec, R.sd, (unsigned)__cvta_generic_to_shared(acc), (unsigned)nq * 16u, step, nq, baseq, slot, ring_s, warp, nwarps, lane, neg_zero)
    if (need_h && need_v) { if (ng10) KIWI_SYNTH(true, true, true); else KIWI_SYNTH(true, true, false); }
    else if (need_h) { if (ng10) KIWI_SYNTH(true, false, true); else KIWI_SYNTH(true, false, false); }
    else if (need_v) { if (ng10) KIWI_SYNTH(false, true, true); else KIWI_SYNTH(false, true, false); }
#undef KIWI_SYNTH
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // ---- reduce the warps' strips (fixed order: deterministic) -------------------------------------
    for (int i = threadIdx.x; i < 3 * nqs; i += blockDim.x) {
        float4 s = acc_all[i];
        for (int w = 1; w < nwarps; w++) {
            const float4 v = acc_all[(size_t)w * 3 * nqs + i];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        acc_all[i] = s;
    }
    for (int i = threadIdx.x; i < 3 * nq; i += blockDim.x) {
        float st = step_all[i];
        for (int w = 1; w < nwarps; w++) st += step_all[(size_t)w * 3 * nq + i];
        step_all[i] = st;
    }
    if (nbands > 1) {   // running sums of the bands in global memory (one CTA per pair and launch, launches in stream order)
        float4* pacc = partial + (size_t)pair * (3 * (size_t)nqs + (3 * (size_t)nq + 3) / 4);
        float* pstep = reinterpret_cast<float*>(pacc + 3 * (size_t)nqs);
        __syncthreads();
        for (int i = threadIdx.x; i < 3 * nqs; i += blockDim.x) {
            float4 s = acc_all[i];
            if (band > 0) { const float4 v = pacc[i]; s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
            if (band + 1 < nbands) pacc[i] = s; else acc_all[i] = s;
        }
        for (int i = threadIdx.x; i < 3 * nq; i += blockDim.x) {
            float st = step_all[i];
            if (band > 0) st += pstep[i];
            if (band + 1 < nbands) pstep[i] = st; else step_all[i] = st;
        }
        if (band + 1 < nbands) return;
    }
    __syncthreads();
    // inclusive prefix sum of the steps over quads, one warp per strip
    for (int strip = warp; strip < 3; strip += nwarps) {
        float* st = step_all + (size_t)strip * nq;
        float run = 0.f;
        for (int q0 = 0; q0 < nq; q0 += 32) {
            const int q = q0 + lane;
            float v = q < nq ? st[q] : 0.f;
            for (int o = 1; o < 32; o <<= 1) { float t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
            v += run;
            if (q < nq) st[q] = v;
            run = __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    // ---- components: signs and the (away,right) -> (north,east) rotation, seismogram.f90:256-289 ----
    const float* a1 = reinterpret_cast<const float*>(acc_all);
    const float* a2 = a1 + (size_t)4 * nqs;
    const float* a3 = a2 + (size_t)4 * nqs;
    const float* st1 = step_all; const float* st2 = step_all + nq; const float* st3 = step_all + 2 * nq;
    const int nsamp = 4 * nq;
    const int s12lo = min(H.s1lo, H.s2lo), s12hi = max(H.s1hi, H.s2hi);
    for (int ic = 0; ic < R.ncomp; ic++) {
        const int id = R.comp[ic];
        const int aid = id < 0 ? -id : id;
        const float sg = id < 0 ? -1.f : 1.f;
        float* row = seis + ((size_t)pair * KIWI_MAX_COMP + ic) * seis_stride;
        int lo, hi;
        if (aid == 1) { lo = H.s1lo; hi = H.s1hi; }
        else if (aid == 2) { lo = H.s2lo; hi = H.s2hi; }
        else if (aid == 3) { lo = H.s3lo; hi = H.s3hi; }
        else { lo = s12lo; hi = s12hi; }
        for (int i = threadIdx.x; i < nsamp && i < (int)seis_stride; i += blockDim.x) {
            const int qi = i >> 2;
            float v;
            if (aid == 3) v = a3[i] + st3[qi];
            else {
                const float u1 = a1[i] + st1[qi], u2 = a2[i] + st2[qi];
                if (aid == 1) v = u1 * sg;
                else if (aid == 2) v = u2 * sg;
                else if (aid == 4) v = (R.cl0 * u1 - R.sl0 * u2) * sg;
                else v = (R.cl0 * u2 + R.sl0 * u1) * sg;
            }
            row[i] = v;
        }
        if (threadIdx.x == 0) { SeisHdr e; e.lo = lo; e.hi = hi; e.base = base; e.pad = 0; myshdr[ic] = e; }
    }
}

// =================================================================================================
// K4: rise-time fold.  receiver_scaled_seismograms_to_probes (receiver.f90:853-904) convolves the
// synthetic of a source with psm%risetime > 0 (eikonal sources) with a boxcar, written as a sum of
// shifted copies of the data-span part of the strip (strip_fold sparse_trace.f90:379-402, strip_dataspan
// :347-377).  One CTA per (candidate, receiver, component); the row is rewritten in place.
// =================================================================================================
#define FOLD_TILE 1024     // shifts whose weights are tabulated in shared memory at a time
__global__ void __launch_bounds__(256) k_fold(const ReceiverDev* __restrict__ rcv, int nrcv, const CandDev* __restrict__ cands,
                                               float* __restrict__ seis, size_t seis_stride, SeisHdr* __restrict__ shdrs, float dt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* d = reinterpret_cast<float*>(smem_raw);   // copy of the strip
    float* vacc = d + seis_stride;                   // the folded row while it is being summed
    __shared__ float s_wl[FOLD_TILE], s_wr[FOLD_TILE], s_w[FOLD_TILE];
    __shared__ int s_its[FOLD_TILE];
    __shared__ int s_ds0, s_ds1;
    __shared__ float s_sum;
    const int item = blockIdx.x;
    const int ic = item % KIWI_MAX_COMP, pair = item / KIWI_MAX_COMP;
    const int b = pair / nrcv, ir = pair % nrcv;
    const ReceiverDev& R = rcv[ir];
    if (!R.enabled || ic >= R.ncomp) return;
    const float risetime = cands[b].risetime;
    if (!(risetime > 0.f) || cands[b].status != 0) return;
    SeisHdr sh = shdrs[item];
    if (sh.hi < sh.lo) return;
    float* row = seis + (size_t)item * seis_stride;
    const int n = sh.hi - sh.lo + 1;
    if (threadIdx.x == 0) { s_ds0 = n - 1; s_ds1 = 0; }
    for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = row[sh.lo - sh.base + i];
    __syncthreads();
    {   // strip_dataspan: first sample that is not zero .. first sample of the trailing constant run
        const float lastvalue = d[n - 1];
        int first = n - 1, lastdiff = -1;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            if (d[i] != 0.f) first = min(first, i);
            if (d[i] != lastvalue) lastdiff = max(lastdiff, i);
        }
        atomicMin(&s_ds0, first);
        atomicMax(&s_ds1, lastdiff + 1);
    }
    // receiver.f90:868-885: boxcar weights of the shifted copies
    const float r0 = -risetime / 2.f, r1 = risetime / 2.f;
    const int nshifts = 1 + 2 * (int)roundf(0.5f * risetime / dt);
    auto shift_time = [&](int is /* 1-based */) { return ((float)(is - 1) - 0.5f * (float)(nshifts - 1)) * dt; };
    auto raw_weight = [&](float ts) { const float a0 = ts - dt / 2.f, a1 = ts + dt / 2.f; return fmaxf(0.f, fminf(r1, a1) - fmaxf(r0, a0)); };
    if (threadIdx.x == 0) {   // (summed in shift order, as the reference does)
        float sum = 0.f;
        for (int is = 1; is <= nshifts; is++) sum = A_(sum, raw_weight(shift_time(is)));
        s_sum = sum;
    }
    __syncthreads();
    const int ds0 = s_ds0, ds1 = s_ds1;
    if (ds1 < ds0) return;
    const float sum = s_sum;
    // sample shifts grow with the shift number: the extreme ones are the first and the last
    const int itsmin = (int)floorf(D_(shift_time(1), dt)), itsmax = (int)floorf(D_(shift_time(nshifts), dt));
    // new strip span (growth of trace_multiply_add, sparse_trace.f90:649-668), clamped to the row
    const int nlo = max(min(sh.lo, sh.lo + ds0 + itsmin), sh.base);
    const int nhi = min(max(sh.hi, sh.lo + ds1 + itsmax + 1), sh.base + (int)seis_stride - 1);
    const float lastval = d[ds1];
    for (int x = nlo + (int)threadIdx.x; x <= nhi; x += blockDim.x) vacc[x - nlo] = 0.f;
    for (int t0 = 0; t0 < nshifts; t0 += FOLD_TILE) {
        const int nt = min(FOLD_TILE, nshifts - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < nt; i += blockDim.x) {
            const float ts = shift_time(t0 + i + 1);
            const float w = D_(raw_weight(ts), sum);
            const float rshift = D_(ts, dt);
            const int its = (int)floorf(rshift);
            const float wr0 = S_(rshift, (float)its), wl0 = S_(1.f, wr0);   // sparse_trace.f90:639-646
            s_its[i] = its; s_w[i] = w; s_wr[i] = M_(wr0, w); s_wl[i] = M_(wl0, w);
        }
        __syncthreads();
        for (int x = nlo + (int)threadIdx.x; x <= nhi; x += blockDim.x) {
            const int xr = x - sh.lo;   // index relative to the old strip start
            float v = vacc[x - nlo];
            for (int i = 0; i < nt; i++) {
                const int y = xr - s_its[i];          // sample of the data-span trace under the left weight
                if (y > ds1) { if (lastval != 0.f) v = A_(v, M_(s_w[i], lastval)); }                   // :696-703
                else if (y >= ds0) {
                    v = A_(v, M_(s_wl[i], d[y]));
                    if (y - 1 >= ds0) v = A_(v, M_(s_wr[i], d[y - 1]));
                }
            }
            vacc[x - nlo] = v;
        }
    }
    for (int x = nlo + (int)threadIdx.x; x <= nhi; x += blockDim.x) row[x - sh.base] = vacc[x - nlo];
    if (threadIdx.x == 0) { sh.lo = nlo; sh.hi = nhi; shdrs[item] = sh; }
}

// =================================================================================================
// K5: scaling + time-domain misfit.  One warp per (candidate, receiver, component).
// =================================================================================================
__device__ __forceinline__ int next_pow2(int n) { return n <= 1 ? 1 : 1 << (32 - __clz(n - 1)); }   // comparator.f90:1111-1118
// comparator.f90:1092-1109
__device__ __forceinline__ void allowed_span(int s0, int s1, int minlength, int& n0, int& n1) {
    int slen = s1 - s0 + 1;
    int length = max(slen, minlength);
    int lengthp = next_pow2(length);
    n0 = s0 - (int)floorf((float)(lengthp - slen) / 2.f);
    n1 = n0 + lengthp - 1;
}
__device__ __forceinline__ int ceil_len2(int len) { return (int)ceilf((float)len * 2.f); }   // ceiling(datalength*paddingfactor)

// final common probe span after probe_set_array(syn) + probes_adjust_spans(ref, syn), fresh state
// (comparator.f90:222-271, 464-486, 291-330)
__device__ void probe_spans(int rds0, int rds1, int rsp0, int rsp1, int sds0, int sds1, int& F0, int& F1) {
    int bs0, bs1;
    allowed_span(sds0, sds1, ceil_len2(sds1 - sds0 + 1), bs0, bs1);             // syn probe_set_array
    const int u0 = min(rds0, sds0), u1 = max(rds1, sds1);
    const int minlength = max(ceil_len2(rds1 - rds0 + 1), ceil_len2(sds1 - sds0 + 1));
    int n0, n1;
    allowed_span(u0, u1, minlength, n0, n1);
    const bool same = (rsp0 == bs0 && rsp1 == bs1) && ((rsp1 - rsp0) == (n1 - n0)) &&
                      (rsp0 <= sds0 && sds1 <= rsp1) && (bs0 <= rds0 && rds1 <= bs1);
    if (same) { F0 = rsp0; F1 = rsp1; } else { F0 = n0; F1 = n1; }
}

__device__ __forceinline__ double warp_sum_d(double v) { for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }
__device__ __forceinline__ double warp_max_d(double v) { for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
__device__ __forceinline__ float warp_max_f(float v) { for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }

__global__ void __launch_bounds__(128) k_misfit_td(const ReceiverDev* __restrict__ rcv, int nrcv, const CandDev* __restrict__ cands,
                                                    int ncand, const float* __restrict__ seis, size_t seis_stride,
                                                    const SeisHdr* __restrict__ shdrs, const float* __restrict__ refdata,
                                                    const float* __restrict__ taperdata, int method, float dt, float syn_factor,
                                                    int nmisfits, float* __restrict__ out /* [ncand][nmisfits][2] */,
                                                    int* __restrict__ status, const CandMap* __restrict__ map) {
    const int lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nitems = (long long)ncand * nrcv * KIWI_MAX_COMP;
    if (item >= nitems) return;
    const int ic = (int)(item % KIWI_MAX_COMP);
    const int slot = (int)(item / KIWI_MAX_COMP) / nrcv, ir = (int)(item / KIWI_MAX_COMP) % nrcv;
    const ReceiverDev& R = rcv[ir];
    if (!R.enabled || ic >= R.ncomp) return;
    const int b = map ? map[slot].out : slot;         // where the result goes
    const int bs = map ? map[slot].syn : slot;        // whose synthetics are used
    const int pair = bs * nrcv + ir;
    const CandDev cand = cands[bs];
    float* o = out + ((size_t)b * nmisfits + R.misfit_base + ic) * 2;
    const SeisHdr sh = shdrs[(size_t)pair * KIWI_MAX_COMP + ic];
    if (cand.status != 0 || sh.hi < sh.lo) {
        if (lane == 0) { o[0] = nanf(""); o[1] = nanf(""); if (cand.status == 0) atomicMax(&status[b], 1); }
        return;
    }
    const float* srow = seis + ((size_t)pair * KIWI_MAX_COMP + ic) * seis_stride;
    const float moment = map ? map[slot].moment : cand.moment;
    const int sds0 = sh.lo, sds1 = sh.hi;
    const int rds0 = R.ref_ds0[ic], rds1 = R.ref_ds1[ic];
    const float* rdat = refdata + R.ref_off[ic];
    int F0, F1;
    probe_spans(rds0, rds1, R.ref_sp0[ic], R.ref_sp1[ic], sds0, sds1, F0, F1);
    // summation span (comparator.f90:784-801) and the span of the reference-only norm (:838-846)
    int p0, p1, q0, q1;
    if (R.has_taper) {
        p0 = max(R.dps0, F0); p1 = min(R.dps1, F1);
        q0 = p0; q1 = p1;   // ref probe span == F after probes_adjust_spans
    } else {
        p0 = min(rds0, sds0); p1 = max(rds1, sds1);
        q0 = rds0; q1 = rds1;
    }
    const float* tp = taperdata + R.taper_off;
    const float fa = 1.f, fb = syn_factor;
    const bool unit = (fa == 1.f && fb == 1.f);
    // array values with the probe continuation rule (comparator.f90:264-267): 0 left of the data
    // span, last data value repeated to the right; synthetic scaled by the moment (:265)
    auto refval = [&](int x) -> float {
        if (x < rds0) return 0.f;
        float v = rdat[min(x, rds1) - rds0];
        if (R.has_taper) v = (x >= R.tp0 && x <= R.tp1) ? v * tp[x - R.tp0] : 0.f;
        return v;
    };
    auto synval = [&](int x) -> float {
        if (x < sds0) return 0.f;
        float v = srow[min(x, sds1) - sh.base] * moment;
        if (R.has_taper) v = (x >= R.tp0 && x <= R.tp1) ? v * tp[x - R.tp0] : 0.f;
        return v;
    };
    double acc = 0., accn = 0.;
    if (method == 6) { acc = -DBL_MAX; }
    float accn_peak = -FLT_MAX;
    for (int x = p0 + lane; x <= p1; x += 32) {
        const float a = refval(x), bb = synval(x);
        if (method == 1) { const double d = (double)(unit ? a - bb : fa * a - fb * bb); acc += d * d; }
        else if (method == 2) { acc += (double)fabsf(unit ? a - bb : fa * a - fb * bb); }
        else if (method == 5) { acc += (double)(unit ? a * bb : a * fa * bb * fb); }
        else if (method == 6) { const double xx = (double)(fa * a), yy = (double)(fb * bb); acc = fmax(acc, sqrt(xx * xx + yy * yy)); }
    }
    for (int x = q0 + lane; x <= q1; x += 32) {
        const float a = refval(x);
        if (method == 1) { const double d = (double)a; accn += d * d; }
        else if (method == 2) accn += (double)fabsf(a);
        else if (method == 5) accn += (double)(a * a);
        else if (method == 6) accn_peak = fmaxf(accn_peak, fabsf(a));
    }
    float mis, nf;
    if (method == 6) { acc = warp_max_d(acc); accn_peak = warp_max_f(accn_peak); mis = (float)acc; nf = fa * accn_peak; }
    else {
        acc = warp_sum_d(acc); accn = warp_sum_d(accn);
        if (method == 1) { mis = (float)sqrt((double)dt * acc); nf = fa * (float)sqrt((double)dt * accn); }
        else if (method == 2) { mis = (float)((double)dt * acc); nf = fa * (float)((double)dt * accn); }
        else { mis = (float)acc; nf = fa * fa * (float)accn; }
    }
    if (p1 < p0) mis = 0.f;   // "applying timedomain norm to empty region" (comparator.f90:803-807)
    if (lane == 0) {
        o[0] = mis; o[1] = nf;
        if (!isfinite(mis) || !isfinite(nf)) atomicMax(&status[b], 2);
    }
}

// =================================================================================================
// K6/K7: general misfit kernel -- everything k_misfit_td does not cover: amplitude-spectrum norms
// (comparator.f90:861-909), norms of band-pass filtered traces (make_spectrum_filtered /
// make_array_filtered :1217-1263) and floating-shift norms (receiver.f90:439-510).  One CTA per
// (candidate, receiver); reference and synthetic of a component are transformed together as
// z = ref + i*syn with one complex FFT in shared memory (decimation in frequency forward, natural ->
// bit-reversed; decimation in time inverse, bit-reversed -> natural), so the real-filter product and
// the inverse transform act on both traces at once.
// =================================================================================================
__device__ __forceinline__ float ip_cos_dev(float x0, float y0, float x1, float y1, float xi) {   // piecewise_linear_function.f90:308-316
    if (y1 != y0) return y0 + (y1 - y0) * (0.5f - 0.5f * cosf((xi - x0) / (x1 - x0) * 3.14159265358979f));
    return y0;
}
// multiplier that plf_taper_array_{r,c} (piecewise_linear_function.f90:195-282) applies to element j
// of an array sampled at dx: 0 for j <= floor(x1/dx) and j >= floor(xn/dx)+1, the interpolated
// flank value in between (ip_cos) or the 0/1 mask of ip_zero_one (:318-327)
__device__ float plf_factor(const float* px, const float* py, int np, float dx, int j, bool zero_one) {
    if (j <= (int)floorf(px[0] / dx)) return 0.f;
    if (j >= (int)floorf(px[np - 1] / dx) + 1) return 0.f;
    for (int i = 0; i < np - 1; i++) {
        const int ibeg = (int)floorf(px[i] / dx) + 1, iend = (int)floorf(px[i + 1] / dx);
        if (j >= ibeg && j <= iend) {
            if (zero_one) return (py[i] == 0.f && py[i + 1] == 0.f) ? 0.f : 1.f;
            return ip_cos_dev(px[i], py[i], px[i + 1], py[i + 1], (float)j * dx);
        }
    }
    return 1.f;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// forward transform exp(-i...), natural order in, bit-reversed order out
__device__ void fft_dif_forward(float2* z, int n, const float2* __restrict__ tw, int tw_n) {
    for (int half = n >> 1; half >= 1; half >>= 1) {
        const int tstride = tw_n / (2 * half);
        for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
            const int j = t & (half - 1);
            const int i0 = ((t - j) << 1) + j, i1 = i0 + half;
            const float2 a = z[i0], b = z[i1];
            z[i0] = make_float2(a.x + b.x, a.y + b.y);
            z[i1] = cmul(make_float2(a.x - b.x, a.y - b.y), __ldg(&tw[j * tstride]));
        }
        __syncthreads();
    }
}
// inverse transform exp(+i...), bit-reversed order in, natural order out, unnormalised
__device__ void fft_dit_inverse(float2* z, int n, const float2* __restrict__ tw, int tw_n) {
    for (int half = 1; half < n; half <<= 1) {
        const int tstride = tw_n / (2 * half);
        for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
            const int j = t & (half - 1);
            const int i0 = ((t - j) << 1) + j, i1 = i0 + half;
            float2 w = __ldg(&tw[j * tstride]); w.y = -w.y;
            const float2 a = z[i0], b = cmul(z[i1], w);
            z[i0] = make_float2(a.x + b.x, a.y + b.y);
            z[i1] = make_float2(a.x - b.x, a.y - b.y);
        }
        __syncthreads();
    }
}

// block-wide reductions of two doubles at once (sum or max); result valid in all threads
__device__ void block_reduce2(double& a, double& b, bool is_max, double* scratch /* [2*32] */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int o = 16; o; o >>= 1) {
        const double ta = __shfl_xor_sync(0xffffffffu, a, o), tb = __shfl_xor_sync(0xffffffffu, b, o);
        if (is_max) { a = fmax(a, ta); b = fmax(b, tb); } else { a += ta; b += tb; }
    }
    __syncthreads();
    if (lane == 0) { scratch[wid] = a; scratch[32 + wid] = b; }
    __syncthreads();
    a = scratch[0]; b = scratch[32];
    for (int w = 1; w < nw; w++) {
        if (is_max) { a = fmax(a, scratch[w]); b = fmax(b, scratch[32 + w]); } else { a += scratch[w]; b += scratch[32 + w]; }
    }
    __syncthreads();
}

// accumulate one element of the two-trace norm and of the reference-only norm (comparator.f90:627-697)
__device__ __forceinline__ void norm_accum2(int bm, float a, float b, float fa, float fb, bool unit, double& acc) {
    if (bm == 1) { const double d = (double)(unit ? a - b : fa * a - fb * b); acc += d * d; }
    else if (bm == 2) acc += (double)fabsf(unit ? a - b : fa * a - fb * b);
    else if (bm == 5) acc += (double)(unit ? a * b : a * fa * b * fb);
    else { const double xx = (double)(fa * a), yy = (double)(fb * b); acc = fmax(acc, sqrt(xx * xx + yy * yy)); }
}
__device__ __forceinline__ void norm_accum1(int bm, float a, double& acc) {
    if (bm == 1) { const double d = (double)a; acc += d * d; }
    else if (bm == 2) acc += (double)fabsf(a);
    else if (bm == 5) acc += (double)(a * a);
    else acc = fmax(acc, (double)fabsf(a));
}
__device__ __forceinline__ void norm_finish(int bm, double acc, double accn, float dx, float fa, float& mis, float& nf) {
    if (bm == 1) { mis = (float)sqrt((double)dx * acc); nf = fa * (float)sqrt((double)dx * accn); }
    else if (bm == 2) { mis = (float)((double)dx * acc); nf = fa * (float)((double)dx * accn); }
    else if (bm == 5) { mis = (float)acc; nf = fa * fa * (float)accn; }
    else { mis = (float)acc; nf = fa * (float)accn; }
}

__global__ void __launch_bounds__(256) k_misfit_general(const ReceiverDev* __restrict__ rcv, int nrcv, const CandDev* __restrict__ cands,
                                                         const float* __restrict__ seis, size_t seis_stride,
                                                         const SeisHdr* __restrict__ shdrs, const float* __restrict__ refdata,
                                                         const float* __restrict__ taperdata, const float2* __restrict__ tw, int tw_n,
                                                         int method, float dt, float syn_factor, int nmisfits, float* __restrict__ out,
                                                         int* __restrict__ status, int* __restrict__ fshift, int n_alloc, int nshift_alloc,
                                                         const CandMap* __restrict__ map, int xs0, int xs1, int premethod,
                                                         float2* __restrict__ zscratch /* transforms too long for shared memory: [CTA][n_alloc] */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* z = zscratch ? zscratch + (size_t)blockIdx.x * n_alloc : reinterpret_cast<float2*>(smem_raw);
    float* sm_m = reinterpret_cast<float*>(smem_raw) + (zscratch ? 0 : 2 * (size_t)n_alloc);
    float* sm_n = sm_m + (size_t)nshift_alloc * KIWI_MAX_COMP;
    __shared__ double scratch[64];
    const int slot = blockIdx.x / nrcv, ir = blockIdx.x % nrcv;
    const ReceiverDev& R = rcv[ir];
    if (!R.enabled || R.ncomp == 0) return;
    const int b = map ? map[slot].out : slot;
    const int bs = map ? map[slot].syn : slot;
    const int pair = bs * nrcv + ir;                  // rows of the synthetics
    const int opair = b * nrcv + ir;                  // floating shift of this candidate
    const CandDev cand = cands[bs];
    float* o = out + ((size_t)b * nmisfits + R.misfit_base) * 2;
    const bool floating = method >= 7;
    // method 9 (internal): windowed cross-correlation of the synthetics with the references pulled through the shifts xs0..xs1
    // (probes_windowed_cross_corr comparator.f90:1061-1090), best shift by receiver_autoshift_ref_seismogram's rule into fshift
    const bool xcorr = method == KIWI_INTERNAL_XCORR;
    const int bm = xcorr ? 5 : (method == 7 ? 1 : (method == 8 ? 2 : (method == 3 ? 1 : (method == 4 ? 2 : method))));   // norm applied
    const bool freq = method == 3 || method == 4;
    const int fs0 = xcorr ? xs0 : R.fs0, fs1 = xcorr ? xs1 : R.fs1;
    const int nshift = floating ? (fs1 - fs0 + 1) : 1;
    bool fail = cand.status != 0 || nshift < 1 || nshift > nshift_alloc;
    for (int ic = 0; ic < R.ncomp && !fail; ic++) if (shdrs[(size_t)pair * KIWI_MAX_COMP + ic].hi < shdrs[(size_t)pair * KIWI_MAX_COMP + ic].lo) fail = true;
    if (fail && xcorr) {
        if (threadIdx.x == 0) fshift[opair] = 0;
        for (int j = threadIdx.x; j < R.ncomp * max(nshift, 0); j += blockDim.x) out[(size_t)opair * KIWI_MAX_COMP * nshift + j] = nanf("");
        return;
    }
    if (fail) {
        if (threadIdx.x < R.ncomp) { o[2 * threadIdx.x] = nanf(""); o[2 * threadIdx.x + 1] = nanf(""); }
        if (threadIdx.x == 0) { if (cand.status == 0) atomicMax(&status[b], 1); if (fshift) fshift[opair] = 0; }
        return;
    }
    const float* tp = taperdata + R.taper_off;
    const float fa = 1.f, fb = syn_factor;
    const bool unit = (fa == 1.f && fb == 1.f);
    const float moment = map ? map[slot].moment : cand.moment;
    const bool tapered = R.has_taper != 0, filtered = R.has_filter != 0;

    for (int ic = 0; ic < R.ncomp; ic++) {
        const SeisHdr sh = shdrs[(size_t)pair * KIWI_MAX_COMP + ic];
        const float* srow = seis + ((size_t)pair * KIWI_MAX_COMP + ic) * seis_stride;
        const float* rdat = refdata + R.ref_off[ic];
        const int sds0 = sh.lo, sds1 = sh.hi;
        int rds0 = R.ref_ds0[ic], rds1 = R.ref_ds1[ic];
        int rsp0 = R.ref_sp0[ic], rsp1 = R.ref_sp1[ic];
        int ssp0, ssp1;
        allowed_span(sds0, sds1, ceil_len2(sds1 - sds0 + 1), ssp0, ssp1);   // probe_set_array(syn) on a fresh probe
        const int rlen = rds1 - rds0 + 1;
        auto shift_ref = [&](int ishift) {   // probe_shift (comparator.f90:273-288): the span only grows (:245-249)
            rds0 += ishift; rds1 += ishift;
            allowed_span(min(rds0, rsp0), max(rds1, rsp1), ceil_len2(rlen), rsp0, rsp1);
        };
        auto adjust_spans = [&]() {   // probes_adjust_spans (comparator.f90:464-486)
            const int u0 = min(rds0, sds0), u1 = max(rds1, sds1);
            const int minlength = max(ceil_len2(rlen), ceil_len2(sds1 - sds0 + 1));
            int n0, n1;
            allowed_span(u0, u1, minlength, n0, n1);
            const bool same = (rsp0 == ssp0 && rsp1 == ssp1) && ((rsp1 - rsp0) == (n1 - n0)) && (rsp0 <= sds0 && sds1 <= rsp1) &&
                              (ssp0 <= rds0 && rds1 <= ssp1);
            if (!same) { rsp0 = ssp0 = n0; rsp1 = ssp1 = n1; }
        };
        if (xcorr) {   // the probe spans as update_misfits (minimizer_engine.f90:390) leaves them in a fresh process
            if (premethod < 0) {
            } else if (premethod >= 7) {
                for (int i = 0; i <= R.fs1 - R.fs0; i++) { shift_ref(i == 0 ? R.fs0 : 1); adjust_spans(); }
                shift_ref(-R.fs1);
            } else adjust_spans();
        }
        for (int i = 0; i < nshift; i++) {
            if (floating) shift_ref((i == 0) ? fs0 : 1);
            adjust_spans();
            const int F0 = rsp0, F1 = rsp1, n = F1 - F0 + 1;
            // element x of the probe arrays with the continuation rule (comparator.f90:264-267) and the taper
            auto refval = [&](int x) -> float {
                if (x < rds0) return 0.f;
                float v = rdat[min(x, rds1) - rds0];
                if (tapered) v = (x >= R.tp0 && x <= R.tp1) ? v * tp[x - R.tp0] : 0.f;
                return v;
            };
            auto synval = [&](int x) -> float {
                if (x < sds0) return 0.f;
                float v = srow[min(x, sds1) - sh.base] * moment;
                if (tapered) v = (x >= R.tp0 && x <= R.tp1) ? v * tp[x - R.tp0] : 0.f;
                return v;
            };
            // summation spans of the time-domain norms (comparator.f90:784-801, 838-846)
            int p0, p1, q0, q1;
            if (tapered) { p0 = max(R.dps0, F0); p1 = min(R.dps1, F1); q0 = p0; q1 = p1; }
            else { p0 = min(rds0, sds0); p1 = max(rds1, sds1); q0 = rds0; q1 = rds1; }
            double acc = (bm == 6) ? -DBL_MAX : 0., accn = (bm == 6) ? -DBL_MAX : 0.;
            float dx = dt;
            bool bad = false;
            if (!freq && !filtered) {
                if (xcorr) for (int x = p0 + threadIdx.x; x <= p1; x += blockDim.x) norm_accum2(bm, synval(x), refval(x), fb, fa, unit, acc);   // (syn, ref) order
                else for (int x = p0 + threadIdx.x; x <= p1; x += blockDim.x) norm_accum2(bm, refval(x), synval(x), fa, fb, unit, acc);
                for (int x = q0 + threadIdx.x; x <= q1; x += blockDim.x) norm_accum1(bm, refval(x), accn);
            } else if (n > n_alloc || n < 2 || (n & (n - 1)) != 0) {
                bad = true;
            } else {
                for (int j = threadIdx.x; j < n; j += blockDim.x) z[j] = make_float2(refval(F0 + j), synval(F0 + j));
                __syncthreads();
                fft_dif_forward(z, n, tw, tw_n);
                int log2n = 0; while ((1 << log2n) < n) log2n++;
                const float df = 1.f / ((float)n * dt);   // comparator.f90:1213
                if (freq) {
                    dx = df;
                    for (int k = threadIdx.x; k <= (n >> 1); k += blockDim.x) {
                        const int pk = (int)(__brev((unsigned)k) >> (32 - log2n));
                        const int pm = (int)(__brev((unsigned)((n - k) & (n - 1))) >> (32 - log2n));
                        const float2 zk = z[pk], zm = z[pm];
                        // Ref_k = (Z_k + conj Z_{n-k})/2, Syn_k = (Z_k - conj Z_{n-k})/(2i)
                        const float rr = 0.5f * (zk.x + zm.x), ri = 0.5f * (zk.y - zm.y);
                        const float sr = 0.5f * (zk.y + zm.y), si = -0.5f * (zk.x - zm.x);
                        float A = sqrtf(rr * rr + ri * ri), B = sqrtf(sr * sr + si * si);
                        if (filtered) { const float hfac = plf_factor(R.fpx, R.fpy, R.nfp, df, k, false); A *= hfac; B *= hfac; }
                        norm_accum2(bm, A, B, fa, fb, unit, acc);
                        norm_accum1(bm, A, accn);
                    }
                } else {
                    // spectrum_filtered = spectrum * filter(k df) (comparator.f90:1217-1231), applied to bins k and n-k alike
                    for (int pz = threadIdx.x; pz < n; pz += blockDim.x) {
                        const int k = (int)(__brev((unsigned)pz) >> (32 - log2n));
                        const int kk = k <= (n >> 1) ? k : n - k;
                        const float hfac = plf_factor(R.fpx, R.fpy, R.nfp, df, kk, false);
                        z[pz].x *= hfac; z[pz].y *= hfac;
                    }
                    __syncthreads();
                    fft_dit_inverse(z, n, tw, tw_n);
                    const float fn = (float)n;
                    for (int j = threadIdx.x; j < n; j += blockDim.x) {   // comparator.f90:1250-1261
                        float2 v = z[j]; v.x = v.x / fn; v.y = v.y / fn;
                        if (tapered) { const float m01 = plf_factor(R.tpx, R.tpy, R.ntp, dt, F0 + j, true); v.x *= m01; v.y *= m01; }
                        z[j] = v;
                    }
                    __syncthreads();
                    if (xcorr) for (int x = p0 + threadIdx.x; x <= p1; x += blockDim.x) { const float2 v = z[x - F0]; norm_accum2(bm, v.y, v.x, fb, fa, unit, acc); }
                    else for (int x = p0 + threadIdx.x; x <= p1; x += blockDim.x) { const float2 v = z[x - F0]; norm_accum2(bm, v.x, v.y, fa, fb, unit, acc); }
                    for (int x = q0 + threadIdx.x; x <= q1; x += blockDim.x) norm_accum1(bm, z[x - F0].x, accn);
                }
            }
            block_reduce2(acc, accn, bm == 6, scratch);
            if (threadIdx.x == 0) {
                float mis, nf;
                norm_finish(bm, acc, accn, dx, fa, mis, nf);
                if (!freq && p1 < p0) mis = 0.f;          // "applying timedomain norm to empty region" (comparator.f90:803-807)
                if (!freq && q1 < q0) nf = (bm == 6) ? fa * -FLT_MAX : 0.f;
                if (bad) { mis = nanf(""); nf = nanf(""); atomicMax(&status[b], 3); }
                sm_m[i * KIWI_MAX_COMP + ic] = mis; sm_n[i * KIWI_MAX_COMP + ic] = nf;
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0 && xcorr) {   // receiver.f90:827: maxloc(sum(max(cc/max(1.,maxval(cc)),0.)**2,2),1), first maximum
        float mx = -FLT_MAX;
        for (int i = 0; i < nshift; i++) for (int ic = 0; ic < R.ncomp; ic++) mx = fmaxf(mx, sm_m[i * KIWI_MAX_COMP + ic]);
        const float den = fmaxf(1.f, mx);
        int iloc = 0; float best = 0.f;
        for (int i = 0; i < nshift; i++) {
            float s = 0.f;
            for (int ic = 0; ic < R.ncomp; ic++) { const float v = fmaxf(__fdiv_rn(sm_m[i * KIWI_MAX_COMP + ic], den), 0.f); s = __fadd_rn(s, __fmul_rn(v, v)); }
            if (i == 0 || s > best) { best = s; iloc = i; }
        }
        fshift[opair] = fs0 + iloc;
        for (int ic = 0; ic < R.ncomp; ic++)   // the correlations themselves: out[pair][component][shift] (output_cross_correlations)
            for (int i = 0; i < nshift; i++) out[((size_t)opair * KIWI_MAX_COMP + ic) * nshift + i] = sm_m[i * KIWI_MAX_COMP + ic];
        return;
    }
    if (threadIdx.x == 0) {
        int iloc = 0;
        if (floating) {   // minloc(sum(misfits[**2],1),1): first minimum (receiver.f90:486-494)
            float best = 0.f;
            for (int i = 0; i < nshift; i++) {
                float s = 0.f;
                for (int ic = 0; ic < R.ncomp; ic++) { const float m = sm_m[i * KIWI_MAX_COMP + ic]; s = s + (bm == 2 ? m : m * m); }
                if (i == 0 || s < best) { best = s; iloc = i; }
            }
            if (fshift) fshift[opair] = fs0 + iloc;
        } else if (fshift) fshift[opair] = 0;
        for (int ic = 0; ic < R.ncomp; ic++) {
            const float mis = sm_m[iloc * KIWI_MAX_COMP + ic];
            float nf;
            if (floating) { float s = 0.f; for (int i = 0; i < nshift; i++) s = s + sm_n[i * KIWI_MAX_COMP + ic]; nf = s / (float)nshift; }
            else nf = sm_n[ic];
            o[2 * ic] = mis; o[2 * ic + 1] = nf;
            if (!isfinite(mis) || !isfinite(nf)) atomicMax(&status[b], 2);
        }
    }
}

// probe_get / probe_get_amp_spectrum (comparator.f90:332-433) of ONE probe: the synthetic (which_probe 0) or the reference (1) of
// component ic of receiver ir, with the probe's own span as in a fresh process (no probes_adjust_spans has run).  processing 0 plain,
// 1 tapered, 2 filtered; spectrum != 0: amplitude spectrum instead of the trace.  hdr: [0] first index, [1] length, [2] df (bits),
// [3] status (0 ok, 1 no trace, 2 span does not fit).  One CTA.
__global__ void __launch_bounds__(256) k_probe_export(const ReceiverDev* __restrict__ rcv, int ir, int ic, const CandDev* __restrict__ cands,
                                                       const float* __restrict__ seis, size_t seis_stride, const SeisHdr* __restrict__ shdrs, int nrcv,
                                                       const float* __restrict__ refdata, const float* __restrict__ taperdata,
                                                       const float2* __restrict__ tw, int tw_n, int which_probe, int processing, int spectrum, float dt,
                                                       int n_alloc, int* __restrict__ hdr, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* z = reinterpret_cast<float2*>(smem_raw);
    const ReceiverDev& R = rcv[ir];
    const CandDev cand = cands[0];
    const SeisHdr sh = shdrs[(size_t)ir * KIWI_MAX_COMP + ic];
    const float* srow = seis + ((size_t)ir * KIWI_MAX_COMP + ic) * seis_stride;
    const float* rdat = refdata + R.ref_off[ic];
    const float* tp = taperdata + R.taper_off;
    const bool tapered = R.has_taper != 0, filtered = R.has_filter != 0;
    int ds0, ds1, sp0, sp1;
    if (which_probe == 0) {
        ds0 = sh.lo; ds1 = sh.hi;
        if (ds1 >= ds0) allowed_span(ds0, ds1, ceil_len2(ds1 - ds0 + 1), sp0, sp1);
    } else {
        ds0 = R.ref_ds0[ic]; ds1 = R.ref_ds1[ic]; sp0 = R.ref_sp0[ic]; sp1 = R.ref_sp1[ic];
    }
    if (ds1 < ds0 || cand.status != 0) { if (threadIdx.x == 0) { hdr[0] = 1; hdr[1] = 0; hdr[2] = 0; hdr[3] = 1; } return; }
    auto plain = [&](int x) -> float {   // element x of probe%array (comparator.f90:264-267)
        if (x < ds0) return 0.f;
        return which_probe == 0 ? srow[min(x, ds1) - sh.base] * cand.moment : rdat[min(x, ds1) - ds0];
    };
    auto tap = [&](int x) -> float {     // element x of probe%array_tapered
        const float v = plain(x);
        if (!tapered) return v;
        return (x >= R.tp0 && x <= R.tp1) ? v * tp[x - R.tp0] : 0.f;
    };
    const int n = sp1 - sp0 + 1;
    const bool need_fft = spectrum || (processing == 2 && filtered);
    if (need_fft && (n > n_alloc || n < 2 || (n & (n - 1)) != 0)) { if (threadIdx.x == 0) { hdr[0] = 1; hdr[1] = 0; hdr[2] = 0; hdr[3] = 2; } return; }
    if (!need_fft) {
        int e0 = ds0, e1 = ds1;
        if (processing >= 1 && tapered) { e0 = max(R.dps0, ds0); e1 = min(R.dps1, ds1); if (e0 > e1) { e0 = ds0; e1 = ds1; } }   // :385-386
        for (int x = e0 + threadIdx.x; x <= e1; x += blockDim.x) out[x - e0] = (processing >= 1) ? tap(x) : plain(x);
        if (threadIdx.x == 0) { hdr[0] = e0; hdr[1] = e1 - e0 + 1; hdr[2] = 0; hdr[3] = 0; }
        return;
    }
    for (int j = threadIdx.x; j < n; j += blockDim.x) z[j] = make_float2(tap(sp0 + j), 0.f);   // make_spectrum transforms the tapered array (:1206-1210)
    __syncthreads();
    fft_dif_forward(z, n, tw, tw_n);
    int log2n = 0; while ((1 << log2n) < n) log2n++;
    const float df = 1.f / ((float)n * dt);
    if (spectrum) {
        for (int k = threadIdx.x; k <= (n >> 1); k += blockDim.x) {
            const float2 v = z[(int)(__brev((unsigned)k) >> (32 - log2n))];
            float a = sqrtf(v.x * v.x + v.y * v.y);
            if (filtered && processing == 2) a *= plf_factor(R.fpx, R.fpy, R.nfp, df, k, false);
            out[k] = a;
        }
        if (threadIdx.x == 0) { hdr[0] = 1; hdr[1] = (n >> 1) + 1; hdr[2] = __float_as_int(df); hdr[3] = 0; }
        return;
    }
    for (int pz = threadIdx.x; pz < n; pz += blockDim.x) {   // spectrum_filtered, as in k_misfit_general
        const int k = (int)(__brev((unsigned)pz) >> (32 - log2n));
        const int kk = k <= (n >> 1) ? k : n - k;
        const float hfac = plf_factor(R.fpx, R.fpy, R.nfp, df, kk, false);
        z[pz].x *= hfac; z[pz].y *= hfac;
    }
    __syncthreads();
    fft_dit_inverse(z, n, tw, tw_n);
    int e0 = ds0, e1 = ds1;
    if (tapered) { e0 = max(R.dps0, sp0); e1 = min(R.dps1, sp1); if (e0 > e1) { e0 = ds0; e1 = ds1; } }   // :408-413
    const float fn = (float)n;
    for (int x = e0 + threadIdx.x; x <= e1; x += blockDim.x) {
        float v = z[x - sp0].x / fn;
        if (tapered) v *= plf_factor(R.tpx, R.tpy, R.ntp, dt, x, true);
        out[x - e0] = v;
    }
    if (threadIdx.x == 0) { hdr[0] = e0; hdr[1] = e1 - e0 + 1; hdr[2] = 0; hdr[3] = 0; }
}

// =================================================================================================
// K8: point moment-tensor grid search (config C2).  For a point source the synthetic is linear in the
// six moment-tensor components (make_weights seismogram.f90:316-336 is linear in m, everything after
// it too), so for every grid location the six unit-tensor "basis" seismograms B_i are synthesised
// once by k_synth and every candidate tensor m_j of that location costs only
//     syn_j(t) = sum_i m_j,i B_i(t)           -- a dense [nmt x 6] . [6 x T] contraction
// followed by the misfit reduction over t.  The contraction runs on the 5th-generation tensor cores:
// tcgen05.mma kind::tf32, A (tensors) and B (basis samples) in shared memory in the canonical K-major
// no-swizzle layout, fp32 accumulator tile 128 x 128 in tensor memory.  TF32 keeps 10 mantissa bits, so
// both operands are split hi + lo and K is laid out as [hi | hi | lo] x [hi | lo | hi] (3xTF32,
// K = 18 padded to 24): the dropped lo*lo term is 2^-22 relative.  Each of the 128 threads then owns one
// candidate (one TMEM lane) and streams its row of the accumulator through the misfit norm
// (comparator.f90:627-697, 770-859).  One CTA per (location, receiver).
// =================================================================================================
#define MTC_M 128          // candidates (TMEM lanes) per tile
#define MTC_N 128          // samples (TMEM columns) per chunk
#define MTC_K 24           // 4 x 6 split products (hi*hi, hi*lo, lo*hi, lo*lo) = 3 MMA k-steps of 8

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE: 8x16-byte core matrices; SBO = distance
// between core matrices along M/N, LBO = distance between the two core matrices along K (16-byte units)
__device__ __forceinline__ unsigned long long umma_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr >> 4) & 0x3fff);
    d |= (unsigned long long)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (unsigned long long)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= 1ull << 46;                                      // descriptor version (Blackwell)
    return d;                                             // base offset 0, layout type 0 = no swizzle
}
// byte offset of element (row, k) of an [rows x MTC_K] K-major operand tile stored as
// [k-core (4 elements)][row-core (8 rows)][8 rows][4 elements]
__device__ __forceinline__ unsigned operand_off(int row, int k, int rows) {
    return (unsigned)(((k >> 2) * (rows >> 3) + (row >> 3)) * 128 + (row & 7) * 16 + (k & 3) * 4);
}

__global__ void __launch_bounds__(128) k_mt_contract(const ReceiverDev* __restrict__ rcv, int nrcv, const MtLoc* __restrict__ locs,
                                                      const float* __restrict__ mts /* [n][6] sorted by location */,
                                                      const int* __restrict__ cand_of /* [n] original candidate index */,
                                                      const float* __restrict__ seis, size_t seis_stride,
                                                      const SeisHdr* __restrict__ shdrs, const float* __restrict__ refdata,
                                                      const float* __restrict__ taperdata, int method, float dt, float syn_factor,
                                                      int nmisfits, float* __restrict__ out, int rcv_per_cta) {
    __shared__ __align__(128) float sA[MTC_M * MTC_K];       // 12 KiB
    __shared__ __align__(128) float sB[MTC_N * MTC_K];       // 12 KiB
    __shared__ __align__(16) float s_ref[MTC_N];             // fa * (tapered) reference of the chunk's columns, 0 beyond the chunk
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ unsigned s_tmem;
    __shared__ double s_red[2][4];

    // one CTA per (grid location, block of receivers): tensor memory, barrier and the candidates' operand tile are set up
    // once and reused for every receiver of the block
    const int nrblk = (nrcv + rcv_per_cta - 1) / rcv_per_cta;
    const int loc = blockIdx.x / nrblk, ir_begin = (blockIdx.x % nrblk) * rcv_per_cta, ir_end = min(nrcv, ir_begin + rcv_per_cta);
    const MtLoc L = locs[loc];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool l1 = method == 2;

    // ---- one-time setup: mbarrier, tensor memory (128 columns = one 128 x 128 fp32 accumulator) ----
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(MTC_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = s_tmem;
    unsigned phase = 0;
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
    const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(MTC_N >> 3) << 17) | ((unsigned)(MTC_M >> 4) << 24);
    const float fa = 1.f, fb = syn_factor;
    unsigned nred = 0;

    for (int m0 = 0; m0 < L.mt_count; m0 += MTC_M) {
        // ---- A tile: this thread's candidate tensor, split hi/lo, zero rows beyond the list ----------
        const int j = m0 + tid;
        const bool have = j < L.mt_count;
        {
            float m[6];
#pragma unroll
            for (int k = 0; k < 6; k++) m[k] = have ? __ldg(&mts[(size_t)(L.mt_begin + j) * 6 + k]) : 0.f;
            // row `tid` of the K-major tile: k-core c (4 values = one 16-byte store) at (c*(rows/8) + row/8)*128 + (row%8)*16
            float hi[6], lo[6];
#pragma unroll
            for (int k = 0; k < 6; k++) { hi[k] = tf32_hi(m[k]); lo[k] = tf32_hi(m[k] - hi[k]); }
            char* a = reinterpret_cast<char*>(sA);
            const float v24[24] = {hi[0], hi[1], hi[2], hi[3], hi[4], hi[5], hi[0], hi[1], hi[2], hi[3], hi[4], hi[5],
                                   lo[0], lo[1], lo[2], lo[3], lo[4], lo[5], lo[0], lo[1], lo[2], lo[3], lo[4], lo[5]};
#pragma unroll
            for (int c4 = 0; c4 < 6; c4++)
                *reinterpret_cast<float4*>(a + operand_off(tid, 4 * c4, MTC_M)) = make_float4(v24[4 * c4], v24[4 * c4 + 1], v24[4 * c4 + 2], v24[4 * c4 + 3]);
        }
        for (int ir = ir_begin; ir < ir_end; ir++) {
        const ReceiverDev& R = rcv[ir];
        if (!R.enabled || R.ncomp == 0) continue;
        const float* tp = taperdata + R.taper_off;
        const bool tapered = R.has_taper != 0;
        for (int ic = 0; ic < R.ncomp; ic++) {
            const size_t item0 = ((size_t)(loc * 6) * nrcv + ir) * KIWI_MAX_COMP + ic;     // basis tensor 0 of this location
            const size_t item_stride = (size_t)nrcv * KIWI_MAX_COMP;                        // next basis tensor
            const SeisHdr sh = shdrs[item0];
            float* o = have ? out + ((size_t)cand_of[L.mt_begin + j] * nmisfits + R.misfit_base + ic) * 2 : nullptr;
            if (sh.hi < sh.lo) { if (o) { o[0] = nanf(""); o[1] = nanf(""); } continue; }
            const int sds0 = sh.lo, sds1 = sh.hi;
            const int rds0 = R.ref_ds0[ic], rds1 = R.ref_ds1[ic];
            const float* rdat = refdata + R.ref_off[ic];
            int F0, F1;
            probe_spans(rds0, rds1, R.ref_sp0[ic], R.ref_sp1[ic], sds0, sds1, F0, F1);
            int p0, p1, q0, q1;
            if (tapered) { p0 = max(R.dps0, F0); p1 = min(R.dps1, F1); q0 = p0; q1 = p1; }
            else { p0 = min(rds0, sds0); p1 = max(rds1, sds1); q0 = rds0; q1 = rds1; }
            auto refval = [&](int x) -> float {
                if (x < rds0) return 0.f;
                float v = rdat[min(x, rds1) - rds0];
                if (tapered) v = (x >= R.tp0 && x <= R.tp1) ? v * tp[x - R.tp0] : 0.f;
                return v;
            };
            auto tapval = [&](int x) -> float { return tapered ? ((x >= R.tp0 && x <= R.tp1) ? tp[x - R.tp0] : 0.f) : 1.f; };
            // The residuals of a chunk are squared and summed 32 at a time in fp32 before they enter the fp64 sum (comparator.f90:639-659 sums in
            // double): with moments of 1e18 and Green's functions of order one a square would leave the fp32 range.  Both operand rows are
            // therefore scaled by a power of two (exact) that brings the reference trace to order one; the sum is scaled back in double.
            const float rs = R.ref_rs[ic];
            double acc = 0.;                   // sum of the scaled residuals (or of their squares)
            // (i) left of the synthetic's data span the synthetic is zero: reference only
            for (int x = p0; x <= min(p1, sds0 - 1); x++) { const float a = (fa * refval(x)) * rs; acc += l1 ? (double)fabsf(a) : (double)a * (double)a; }
            // (ii) columns x in [xs, xe] come out of the tensor-core contraction, 128 at a time; xe = sds1 is
            // always included when the span reaches past it, its value is the continuation (comparator.f90:264-267)
            const int xs = max(p0, sds0), xe = min(p1, sds1) < xs ? -1 : ((p1 > sds1) ? sds1 : min(p1, sds1));
            // last sample of this thread's candidate (continued to the right, comparator.f90:264-267): six fp32 fmas
            float e_last = 0.f;
            if (p1 > sds1 && sds1 >= sds0 && have) {
#pragma unroll
                for (int k = 0; k < 6; k++)
                    e_last = fmaf(__ldg(&mts[(size_t)(L.mt_begin + j) * 6 + k]), __ldg(seis + (item0 + (size_t)k * item_stride) * seis_stride + (sds1 - sh.base)), e_last);
            }
            // column x = c0 + tid of the chunk: six basis samples, already multiplied by the taper and the synthetics factor
            // (moment = 1 for a moment-tensor source, source_moment_tensor.f90:199), and fa * reference: the epilogue is
            // r = fa*ref - acc and nothing else.  The next chunk's column is fetched while this chunk is in the tensor core.
            float colv[6], colref;
            auto fetch_col = [&](int c0) {
                const int x = c0 + tid;
                const bool in = x <= xe;
                const float scale = in ? (fb * tapval(x)) * rs : 0.f;
#pragma unroll
                for (int k = 0; k < 6; k++) colv[k] = in ? __ldg(seis + (item0 + (size_t)k * item_stride) * seis_stride + (x - sh.base)) * scale : 0.f;
                colref = in ? (fa * refval(x)) * rs : 0.f;
            };
            if (xe >= xs) fetch_col(xs);
            for (int c0 = xs; xe >= xs && c0 <= xe; c0 += MTC_N) {
                // ---- B chunk: six basis rows, split hi/lo -------------------------------------------------
                {
                    char* bsm = reinterpret_cast<char*>(sB);
                    float hi[6], lo[6];
#pragma unroll
                    for (int k = 0; k < 6; k++) { hi[k] = tf32_hi(colv[k]); lo[k] = tf32_hi(colv[k] - hi[k]); }
                    const float v24[24] = {hi[0], hi[1], hi[2], hi[3], hi[4], hi[5], lo[0], lo[1], lo[2], lo[3], lo[4], lo[5],
                                           hi[0], hi[1], hi[2], hi[3], hi[4], hi[5], lo[0], lo[1], lo[2], lo[3], lo[4], lo[5]};
#pragma unroll
                    for (int c4 = 0; c4 < 6; c4++)
                        *reinterpret_cast<float4*>(bsm + operand_off(tid, 4 * c4, MTC_N)) = make_float4(v24[4 * c4], v24[4 * c4 + 1], v24[4 * c4 + 2], v24[4 * c4 + 3]);
                    s_ref[tid] = colref;
                }
                if (c0 + MTC_N <= xe) fetch_col(c0 + MTC_N);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> async proxy (tensor core)
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncthreads();
                if (tid == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll
                    for (int ks = 0; ks < MTC_K / 8; ks++) {
                        // k-step ks covers k-cores 2ks, 2ks+1; a k-core block is (rows/8)*128 bytes
                        const unsigned long long da = umma_desc(a0 + ks * 2 * (MTC_M / 8) * 128, (MTC_M / 8) * 128, 128);
                        const unsigned long long dbb = umma_desc(b0 + ks * 2 * (MTC_N / 8) * 128, (MTC_N / 8) * 128, 128);
                        const unsigned accum = ks > 0 ? 1u : 0u;
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
                            "l"(da), "l"(dbb), "r"(idesc), "r"(accum)
                            : "memory");
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
                }
                // ---- wait for the accumulator, then every thread folds its row into the norm -----------------
                {
                    unsigned done = 0, spins = 0;
                    while (!done) {
                        if (++spins > (1u << 26)) { asm volatile("trap;"); }   // never spin for ever on a lost completion
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                            : "=r"(done)
                            : "r"(smem_u32(&s_bar)), "r"(phase)
                            : "memory");
                    }
                    phase ^= 1;
                }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int ncol = min(MTC_N, xe - c0 + 1);
                for (int cc = 0; cc < ncol; cc += 32) {   // (columns beyond ncol: zero basis rows and zero reference, r = 0)
                    unsigned v[32];
                    const unsigned taddr = tmem + ((unsigned)(warp * 32) << 16) + (unsigned)cc;
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(taddr)
                        : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    // 32 terms in fp32 (two interleaved partial sums, packed arithmetic), then into the fp64 sum
                    // (comparator.f90:639-659 sums in double)
                    u64 part2 = 0ull;
                    const float4* rp = reinterpret_cast<const float4*>(s_ref + cc);
#pragma unroll
                    for (int u4 = 0; u4 < 8; u4++) {
                        const float4 a4 = rp[u4];
                        u64 r01 = pk2(a4.x, a4.y), r23 = pk2(a4.z, a4.w);
                        ffma2(r01, -1.f, pk2(__uint_as_float(v[4 * u4]), __uint_as_float(v[4 * u4 + 1])));
                        ffma2(r23, -1.f, pk2(__uint_as_float(v[4 * u4 + 2]), __uint_as_float(v[4 * u4 + 3])));
                        if (l1) {
                            r01 &= 0x7fffffff7fffffffull; r23 &= 0x7fffffff7fffffffull;
                            ffma2(part2, 1.f, r01); ffma2(part2, 1.f, r23);
                        } else {
                            asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(part2) : "l"(r01));
                            asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(part2) : "l"(r23));
                        }
                    }
                    float pa, pb;
                    unpk2(part2, pa, pb);
                    acc += (double)(pa + pb);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncthreads();      // accumulator and operand tiles are free again
            }
            // (iii) right of the data span the synthetic repeats its last sample
            for (int x = max(max(p0, sds0), sds1 + 1); x <= p1; x++) {
                const float r = (fa * refval(x) - fb * (e_last * tapval(x))) * rs;
                acc += l1 ? (double)fabsf(r) : (double)r * (double)r;
            }
            acc = l1 ? acc / (double)rs : acc / ((double)rs * (double)rs);
            // reference-only norm (the same for all candidates): block reduction
            double accn = 0.;
            for (int x = q0 + tid; x <= q1; x += MTC_M) { const float a = refval(x); accn += l1 ? (double)fabsf(a) : (double)a * (double)a; }
            for (int ofs = 16; ofs; ofs >>= 1) accn += __shfl_xor_sync(0xffffffffu, accn, ofs);
            const unsigned rb = nred++ & 1u;   // alternating buffers: one barrier between the writes and the reads is enough
            if (lane == 0) s_red[rb][warp] = accn;
            __syncthreads();
            accn = s_red[rb][0] + s_red[rb][1] + s_red[rb][2] + s_red[rb][3];
            if (o) {
                float mis, nf;
                if (l1) { mis = (float)((double)dt * acc); nf = fa * (float)((double)dt * accn); }
                else { mis = (float)sqrt((double)dt * acc); nf = fa * (float)sqrt((double)dt * accn); }
                if (p1 < p0) mis = 0.f;
                o[0] = mis; o[1] = nf;
            }
        }
        }
        __syncthreads();   // before the next A tile overwrites sA
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(MTC_N) : "memory");
}

// =================================================================================================
// K8f: the same grid search with the synthesis fused in.  A point source is ONE group, so everything the candidates of a
// (location, receiver) pair have in common is 4 corners x ng rows of one window.  Instead of writing six unit-tensor
// seismograms per pair to HBM (k_synth, 18 rows) and reading them back (k_mt_contract), the CTA
//   1a gathers the rows once and combines the corners bilinearly (gfdb.f90:943-948)            -> U_k, k = g1..g10, shared memory
//   1b applies the tap filter of trace_multiply_add (sparse_trace.f90:639-705, shift table of k_tap_table) to every U_k
//          G_k(x) = sum_m sum_t h_m[t] U_k(x - 4 q_m - t),    U_k = 0 left of the window, its last sample to the right
//   2  and contracts on the tensor cores, per receiver component,
//          syn_j(x) - ref(x) = sum_k a_jk G_k(x) - ref(x),
//      where a_jk are candidate j's coefficients of the GF components: make_weights of its tensor (seismogram.f90:316-336:
//      linear in m; the azimuth factors cos^2, sin^2, sin 2a, cos, sin come with the record of a probe source mxx = mxz = 1),
//      the centroid rotation (:196-203) and the receiver's component sign / rotation (:256-289) -- six non-zero coefficients
//      for a horizontal component (g1 g2 g3 g9 -> radial, g4 g5 -> transverse), four for the vertical (g6 g7 g8 g10).
// The reference trace rides along as a seventh row against a coefficient -1, so the accumulator holds syn - ref and the
// epilogue is one packed square-accumulate (|.| for the L1 norm) per two samples.  Operands are split hi + lo (TF32 keeps
// 10 mantissa bits); K = [hi | hi | lo | lo] x [hi | lo | hi | lo] of the 7 values, padded to 32.
// =================================================================================================
#define MTF_K 32
#define MTF_MAXSTEP 8      // distinct quad shifts of the location's taps the kernel keeps (more: general path)
#define MTF_PAD 64         // samples the filtered strips may be longer than the unshifted ones
#define MTF_MAXRCV 32      // receivers of one CTA (their records and headers are staged in shared memory at the start)

__device__ __forceinline__ float4 f4_scale(float s, const float4& v) { return make_float4(s * v.x, s * v.y, s * v.z, s * v.w); }
__device__ __forceinline__ void f4_axpy(float4& a, float s, const float4& v) {
    a.x = fmaf(s, v.x, a.x); a.y = fmaf(s, v.y, a.y); a.z = fmaf(s, v.z, a.z); a.w = fmaf(s, v.w, a.w);
}
__device__ __forceinline__ float4 ldg_quad(const float* slabs, const NodeInfo& n, int q, int comp) {
    unsigned stride;
    return __ldg(reinterpret_cast<const float4*>(corner_ptr(slabs, n, q, comp, stride)));
}

__global__ void __launch_bounds__(128, 4) k_mt_fused(GfdbDev db, const ReceiverDev* __restrict__ rcv, int nrcv, const MtLoc* __restrict__ locs,
                                                   const float* __restrict__ mts /* [n][6] sorted by location */,
                                                   const int* __restrict__ cand_of /* [n] original candidate index */,
                                                   const GeoRec* __restrict__ recs /* [loc][rcv], one group each */,
                                                   const PairHdr* __restrict__ hdrs, const float4* __restrict__ taprec, int strip_cap,
                                                   const float* __restrict__ refdata, const float* __restrict__ taperdata, int method, float dt,
                                                   float syn_factor, int nmisfits, float* __restrict__ out, int rcv_per_cta,
                                                   int* __restrict__ overflow) {
    extern __shared__ __align__(128) unsigned char mtf_smem[];
    float* sA = reinterpret_cast<float*>(mtf_smem);                 // [128 x 32] K-major core matrices, 16 KiB
    float* sB = sA + MTC_M * MTF_K;                                 // 16 KiB
    const int SU = strip_cap + 4, SF = strip_cap + MTF_PAD + 4;     // strip pitches: one quad of zeros, then the window
    float* sU = sA;                                                 // [10][SU] bilinearly combined GF components, unshifted: lives where the
                                                                    // operand tiles are built later (phase 1 is over by then)
    float* sF = sA + max((MTC_M + MTC_N) * MTF_K, (10 * SU + 31) & ~31);   // [10][SF] the same after the tap filter
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ unsigned s_tmem;
    __shared__ double s_red[2][4];
    __shared__ float s_h[MTF_MAXSTEP][5];
    __shared__ int s_q[MTF_MAXSTEP];
    __shared__ __align__(16) GeoRec s_rec[MTF_MAXRCV];
    __shared__ __align__(16) PairHdr s_hdr[MTF_MAXRCV];

    const int nrblk = (nrcv + rcv_per_cta - 1) / rcv_per_cta;
    const int loc = blockIdx.x / nrblk, ir_begin = (blockIdx.x % nrblk) * rcv_per_cta, ir_end = min(nrcv, ir_begin + rcv_per_cta);
    const MtLoc L = locs[loc];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool l1 = method == 2;

    // records and headers of this CTA's (location, receiver) pairs: one round trip to memory for all of them
    for (int i = tid; i < (ir_end - ir_begin) * 8; i += MTC_M)
        reinterpret_cast<uint4*>(s_rec)[i] = __ldg(reinterpret_cast<const uint4*>(recs + (size_t)loc * nrcv + ir_begin) + i);
    for (int i = tid; i < (ir_end - ir_begin) * 2; i += MTC_M)
        reinterpret_cast<uint4*>(s_hdr)[i] = __ldg(reinterpret_cast<const uint4*>(hdrs + (size_t)loc * nrcv + ir_begin) + i);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(MTC_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = s_tmem;
    unsigned phase = 0;
    const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(MTC_N >> 3) << 17) | ((unsigned)(MTC_M >> 4) << 24);
    const float fa = 1.f, fb = syn_factor;
    unsigned nred = 0;

    for (int ir = ir_begin; ir < ir_end; ir++) {
        const ReceiverDev& R = rcv[ir];
        if (!R.enabled || R.ncomp == 0) continue;
        const PairHdr H = s_hdr[ir - ir_begin];
        const GeoRec* rec = &s_rec[ir - ir_begin];
        const int flags = rec->flags, nstep = rec->nstep, tt = rec->tt_begin;
        NodeInfo nd[4];
#pragma unroll
        for (int c = 0; c < 4; c++) nd[c] = rec->node[c];
        int q_first = 0, q_last = -1;
        if (H.T > 0) window_quads(nd[0], nd[1], nd[2], nd[3], q_first, q_last);
        const int nqw = q_last - q_first + 1;       // quads of the unshifted window; the last one holds the continuation value
        bool usable = H.T > 0 && !(flags & GEO_SKIP);
        if (usable && (4 * nqw > strip_cap || nstep > MTF_MAXSTEP || nstep < 1)) usable = false, (void)(tid == 0 && atomicExch(overflow, 1));
        __syncthreads();   // the previous receiver's strips and shift table are no longer read
        if (usable && tid < nstep) {
            const float4 ta = __ldg(taprec + 2 * ((size_t)tt + tid)), tb = __ldg(taprec + 2 * ((size_t)tt + tid) + 1);
            s_q[tid] = __float_as_int(ta.x);
            s_h[tid][0] = ta.y; s_h[tid][1] = ta.z; s_h[tid][2] = ta.w; s_h[tid][3] = tb.x; s_h[tid][4] = tb.y;
        }
        const bool need_h = (R.ja | R.jr | R.jn | R.je) != 0, need_v = R.jd != 0, ng10 = db.ng == 10;
        if (usable) {
            // ---- phase 1a: corners -> U_k ------------------------------------------------------------------
            if (tid < 10) { *reinterpret_cast<float4*>(sU + (size_t)tid * SU) = f4zero(); *reinterpret_cast<float4*>(sF + (size_t)tid * SF) = f4zero(); }
            const bool single = flags & GEO_SINGLE;
            const float dix = rec->dix, diz = rec->diz;
            const float wc0 = single ? 1.f : (1.f - dix) * (1.f - diz), wc1 = single ? 0.f : (1.f - dix) * diz,
                        wc2 = single ? 0.f : dix * (1.f - diz), wc3 = single ? 0.f : dix * diz;
            for (int qi = tid; qi < nqw; qi += MTC_M) {
                const int q = q_first + qi;
                // five components (twenty 128-bit loads) in flight at a time
#pragma unroll
                for (int kb = 0; kb < 10; kb += 5) {
                    float4 t[5][4];
#pragma unroll
                    for (int kk = 0; kk < 5; kk++) {
                        const int k = kb + kk;
                        const bool wanted = ((k < 5 || k == 8) ? need_h : need_v) && (k < 8 || ng10);
#pragma unroll
                        for (int c = 0; c < 4; c++) t[kk][c] = (wanted && (c == 0 || !single)) ? ldg_quad(db.slabs, nd[c], q, k) : f4zero();
                    }
#pragma unroll
                    for (int kk = 0; kk < 5; kk++) {
                        const int k = kb + kk;
                        const bool wanted = ((k < 5 || k == 8) ? need_h : need_v) && (k < 8 || ng10);
                        if (!wanted) continue;
                        float4 r = f4_scale(wc0, t[kk][0]);
                        f4_axpy(r, wc1, t[kk][1]); f4_axpy(r, wc2, t[kk][2]); f4_axpy(r, wc3, t[kk][3]);
                        *reinterpret_cast<float4*>(sU + (size_t)k * SU + 4 + 4 * qi) = r;
                    }
                }
            }
        }
        __syncthreads();
        // ---- phase 1b: tap filter, one thread per output quad Q = q_first + mmin + Qi:
        //      out(4Q + j) = sum_m sum_t h_m[t] U(4(Q - q_m) + j - t): own quad a and previous quad p of the source --------------------
        int mmin = 0, mmax = 0;
        if (usable) {
            mmin = INT_MAX; mmax = INT_MIN;
            for (int m = 0; m < nstep; m++) { mmin = min(mmin, s_q[m]); mmax = max(mmax, s_q[m]); }
        }
        const int nqf = nqw + 1 + (mmax - mmin);    // quads of the filtered window; the last one is constant (continuation)
        if (usable && 4 * nqf > strip_cap + MTF_PAD) usable = false, (void)(tid == 0 && atomicExch(overflow, 1));
        const int x0f = 4 * (q_first + mmin), Lf = 4 * nqf;
        if (usable) {
            for (int Qi = tid; Qi < nqf; Qi += MTC_M) {
                float4 acc[10];
#pragma unroll
                for (int k = 0; k < 10; k++) acc[k] = f4zero();
                for (int m = 0; m < nstep; m++) {
                    const int si = Qi + mmin - s_q[m];                                 // source quad, relative to q_first
                    const int ia = 4 + 4 * min(max(si, -1), nqw - 1), ip = 4 + 4 * min(max(si - 1, -1), nqw - 1);
                    const float h0 = s_h[m][0], h1 = s_h[m][1], h2 = s_h[m][2], h3 = s_h[m][3], h4 = s_h[m][4];
#pragma unroll
                    for (int k = 0; k < 10; k++) {
                        const bool wanted = ((k < 5 || k == 8) ? need_h : need_v) && (k < 8 || ng10);
                        if (!wanted) continue;
                        const float4 a = *reinterpret_cast<const float4*>(sU + (size_t)k * SU + ia), p = *reinterpret_cast<const float4*>(sU + (size_t)k * SU + ip);
                        float4& o = acc[k];
                        o.x = fmaf(h4, p.x, fmaf(h3, p.y, fmaf(h2, p.z, fmaf(h1, p.w, fmaf(h0, a.x, o.x)))));
                        o.y = fmaf(h4, p.y, fmaf(h3, p.z, fmaf(h2, p.w, fmaf(h1, a.x, fmaf(h0, a.y, o.y)))));
                        o.z = fmaf(h4, p.z, fmaf(h3, p.w, fmaf(h2, a.x, fmaf(h1, a.y, fmaf(h0, a.z, o.z)))));
                        o.w = fmaf(h4, p.w, fmaf(h3, a.x, fmaf(h2, a.y, fmaf(h1, a.z, fmaf(h0, a.w, o.w)))));
                    }
                }
#pragma unroll
                for (int k = 0; k < 10; k++) {
                    const bool wanted = ((k < 5 || k == 8) ? need_h : need_v) && (k < 8 || ng10);
                    if (wanted) *reinterpret_cast<float4*>(sF + (size_t)k * SF + 4 + 4 * Qi) = acc[k];
                }
            }
        }
        __syncthreads();
        if (ir + 1 < ir_end) {   // the next receiver's rows on their way into L2 while this one is contracted: one prefetch per 128-byte line
            const GeoRec* nrec = &s_rec[ir + 1 - ir_begin];
            const PairHdr nH = s_hdr[ir + 1 - ir_begin];
            const ReceiverDev& nR = rcv[ir + 1];
            if (nH.T > 0 && !(nrec->flags & GEO_SKIP) && nR.enabled) {
                const bool nh = (nR.ja | nR.jr | nR.jn | nR.je) != 0, nv = nR.jd != 0;
                const int ncorner = (nrec->flags & GEO_SINGLE) ? 1 : 4;
                for (int c = 0; c < ncorner; c++) {
                    const NodeInfo n = nrec->node[c];
                    const int nq = n.wn >> 2, lines = (nq + 7) >> 3;
                    const char* base = reinterpret_cast<const char*>(db.slabs + n.off);
                    for (int i = tid; i < lines * db.ng; i += MTC_M) {
                        const int k = i / lines, l = i - k * lines;
                        const bool wanted = ((k < 5 || k == 8) ? nh : nv);
                        if (wanted) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + ((size_t)(k * nq + l * 8) << 4)));
                    }
                }
            }
        }
        // azimuth factors of the probe source mxx = mxz = 1: f = {cos^2, cos, 0, -sin(2a)/2, -sin, sin^2}
        const float ca2 = rec->f[0], ca = rec->f[1], s2a = -2.f * rec->f[3], sa = -rec->f[4], sa2 = rec->f[5];
        const float c2a = ca2 - sa2;
        const float cl = rec->cl, sl = rec->sl;
        const float* tp = taperdata + R.taper_off;
        const bool tapered = R.has_taper != 0;
        const int s12lo = min(H.s1lo, H.s2lo), s12hi = max(H.s1hi, H.s2hi);

        for (int m0 = 0; m0 < L.mt_count; m0 += MTC_M) {
            const int j = m0 + tid;
            const bool have = j < L.mt_count;
            float f1 = 0.f, f2 = 0.f, f3 = 0.f, f4 = 0.f, f5 = 0.f, f6 = 0.f;
            if (have) {   // make_weights of this thread's candidate tensor (seismogram.f90:316-336)
                float mt[6];
#pragma unroll
                for (int k = 0; k < 6; k++) mt[k] = __ldg(&mts[(size_t)(L.mt_begin + j) * 6 + k]);
                f1 = mt[0] * ca2 + mt[1] * sa2 + mt[3] * s2a;
                f2 = mt[4] * ca + mt[5] * sa;
                f3 = mt[2];
                f4 = 0.5f * (mt[1] - mt[0]) * s2a + mt[3] * c2a;
                f5 = mt[5] * ca - mt[4] * sa;
                f6 = mt[0] * sa2 + mt[1] * ca2 - mt[3] * s2a;
            }
            for (int ic = 0; ic < R.ncomp; ic++) {
                const int id = R.comp[ic], aid = id < 0 ? -id : id;
                const float sg = id < 0 ? -1.f : 1.f;
                int sds0, sds1;
                if (aid == 1) { sds0 = H.s1lo; sds1 = H.s1hi; }
                else if (aid == 2) { sds0 = H.s2lo; sds1 = H.s2hi; }
                else if (aid == 3) { sds0 = H.s3lo; sds1 = H.s3hi; }
                else { sds0 = s12lo; sds1 = s12hi; }
                float* o = have ? out + ((size_t)cand_of[L.mt_begin + j] * nmisfits + R.misfit_base + ic) * 2 : nullptr;
                if (!usable || sds1 < sds0) { if (o) { o[0] = nanf(""); o[1] = nanf(""); } continue; }
                // ---- this candidate's coefficients of the six (four) GF components of the component, and their strips ------------
                float a7[7];
                const float* row[6];
                if (aid == 3) {
                    a7[0] = R.sd * f1; a7[1] = R.sd * f2; a7[2] = R.sd * f3; a7[3] = R.sd * f6; a7[4] = 0.f; a7[5] = 0.f;
                    row[0] = sF + 5 * (size_t)SF; row[1] = sF + 6 * (size_t)SF; row[2] = sF + 7 * (size_t)SF; row[3] = sF + 9 * (size_t)SF;
                    row[4] = row[0]; row[5] = row[0];
                } else {
                    // value = c1 * A1 + c2 * A2 with A1 = cl R - sl T, A2 = cl T + sl R (seismogram.f90:200-203, 256-289)
                    const float c1 = aid == 1 ? sg : (aid == 2 ? 0.f : (aid == 4 ? sg * R.cl0 : sg * R.sl0));
                    const float c2 = aid == 1 ? 0.f : (aid == 2 ? sg : (aid == 4 ? -sg * R.sl0 : sg * R.cl0));
                    const float al = c1 * cl + c2 * sl, be = c2 * cl - c1 * sl;
                    a7[0] = al * f1; a7[1] = al * f2; a7[2] = al * f3; a7[3] = al * f6; a7[4] = be * f4; a7[5] = be * f5;
                    row[0] = sF; row[1] = sF + (size_t)SF; row[2] = sF + 2 * (size_t)SF; row[3] = sF + 8 * (size_t)SF;
                    row[4] = sF + 3 * (size_t)SF; row[5] = sF + 4 * (size_t)SF;
                }
                if (!ng10) { a7[3] = 0.f; row[3] = row[0]; }   // (no g9 / g10 in an eight-component database: their strips were not written)
                a7[6] = have ? -1.f : 0.f;
                {   // A tile: [hi | hi | lo | lo]
                    float hi[7], lo[7];
#pragma unroll
                    for (int k = 0; k < 7; k++) { hi[k] = tf32_hi(a7[k]); lo[k] = tf32_hi(a7[k] - hi[k]); }
                    char* a = reinterpret_cast<char*>(sA);
#pragma unroll
                    for (int blk = 0; blk < 4; blk++) {
                        const float* s = blk < 2 ? hi : lo;
                        *reinterpret_cast<float4*>(a + operand_off(tid, 8 * blk, MTC_M)) = make_float4(s[0], s[1], s[2], s[3]);
                        *reinterpret_cast<float4*>(a + operand_off(tid, 8 * blk + 4, MTC_M)) = make_float4(s[4], s[5], s[6], 0.f);
                    }
                }
                const int nrow = aid == 3 ? 4 : 6;
                auto strip_at = [&](int k, int x) -> float { return row[k][4 + min(max(x - x0f, -1), Lf - 1)]; };
                const int rds0 = R.ref_ds0[ic], rds1 = R.ref_ds1[ic];
                const float* rdat = refdata + R.ref_off[ic];
                int F0, F1;
                probe_spans(rds0, rds1, R.ref_sp0[ic], R.ref_sp1[ic], sds0, sds1, F0, F1);
                int p0, p1, q0, q1;
                if (tapered) { p0 = max(R.dps0, F0); p1 = min(R.dps1, F1); q0 = p0; q1 = p1; }
                else { p0 = min(rds0, sds0); p1 = max(rds1, sds1); q0 = rds0; q1 = rds1; }
                auto refval = [&](int x) -> float {
                    if (x < rds0) return 0.f;
                    float v = rdat[min(x, rds1) - rds0];
                    if (tapered) v = (x >= R.tp0 && x <= R.tp1) ? v * tp[x - R.tp0] : 0.f;
                    return v;
                };
                auto tapval = [&](int x) -> float { return tapered ? ((x >= R.tp0 && x <= R.tp1) ? tp[x - R.tp0] : 0.f) : 1.f; };
                // (the residuals are squared in fp32: both operand rows scaled by a power of two that brings the reference to order one, see k_mt_contract)
                const float rs = R.ref_rs[ic];
                double acc = 0.;                   // sum of the scaled residuals (or of their squares)
                // (i) left of the synthetic's data span the synthetic is zero: reference only
                for (int x = p0; x <= min(p1, sds0 - 1); x++) { const float a = (fa * refval(x)) * rs; acc += l1 ? (double)fabsf(a) : (double)a * (double)a; }
                // (ii) columns x in [xs, xe] come out of the tensor-core contraction, 128 at a time
                const int xs = max(p0, sds0), xe = min(p1, sds1) < xs ? -1 : ((p1 > sds1) ? sds1 : min(p1, sds1));
                // last sample of this thread's candidate (continued to the right, comparator.f90:264-267)
                float e_last = 0.f;
                if (p1 > sds1 && sds1 >= sds0) {
#pragma unroll
                    for (int k = 0; k < 6; k++) if (k < nrow) e_last = fmaf(a7[k], strip_at(k, sds1), e_last);
                }
                // column x = c0 + tid of the chunk: the GF component samples times taper and synthetics factor, and fa * reference
                float colv[7];
                auto fetch_col = [&](int c0) {
                    const int x = c0 + tid;
                    const bool in = x <= xe;
                    const float scale = in ? (fb * tapval(x)) * rs : 0.f;
#pragma unroll
                    for (int k = 0; k < 6; k++) colv[k] = (in && k < nrow) ? strip_at(k, x) * scale : 0.f;
                    colv[6] = in ? (fa * refval(x)) * rs : 0.f;
                };
                if (xe >= xs) fetch_col(xs);
                for (int c0 = xs; xe >= xs && c0 <= xe; c0 += MTC_N) {
                    {   // ---- B chunk: [hi | lo | hi | lo]
                        char* bsm = reinterpret_cast<char*>(sB);
                        float hi[7], lo[7];
#pragma unroll
                        for (int k = 0; k < 7; k++) { hi[k] = tf32_hi(colv[k]); lo[k] = tf32_hi(colv[k] - hi[k]); }
#pragma unroll
                        for (int blk = 0; blk < 4; blk++) {
                            const float* s = (blk & 1) ? lo : hi;
                            *reinterpret_cast<float4*>(bsm + operand_off(tid, 8 * blk, MTC_N)) = make_float4(s[0], s[1], s[2], s[3]);
                            *reinterpret_cast<float4*>(bsm + operand_off(tid, 8 * blk + 4, MTC_N)) = make_float4(s[4], s[5], s[6], 0.f);
                        }
                    }
                    if (c0 + MTC_N <= xe) fetch_col(c0 + MTC_N);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> async proxy (tensor core)
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncthreads();
                    if (tid == 0) {
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const unsigned a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll
                        for (int ks = 0; ks < MTF_K / 8; ks++) {
                            const unsigned long long da = umma_desc(a0 + ks * 2 * (MTC_M / 8) * 128, (MTC_M / 8) * 128, 128);
                            const unsigned long long dbb = umma_desc(b0 + ks * 2 * (MTC_N / 8) * 128, (MTC_N / 8) * 128, 128);
                            const unsigned accum = ks > 0 ? 1u : 0u;
                            asm volatile(
                                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
                                "l"(da), "l"(dbb), "r"(idesc), "r"(accum)
                                : "memory");
                        }
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
                    }
                    {
                        unsigned done = 0, spins = 0;
                        while (!done) {
                            if (++spins > (1u << 26)) { asm volatile("trap;"); }   // never spin for ever on a lost completion
                            asm volatile(
                                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                                : "=r"(done)
                                : "r"(smem_u32(&s_bar)), "r"(phase)
                                : "memory");
                        }
                        phase ^= 1;
                    }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const int ncol = min(MTC_N, xe - c0 + 1);
                    for (int cc = 0; cc < ncol; cc += 32) {   // (columns beyond ncol: zero rows and zero reference, residual 0)
                        unsigned v[32];
                        const unsigned taddr = tmem + ((unsigned)(warp * 32) << 16) + (unsigned)cc;
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                            : "r"(taddr)
                            : "memory");
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        // 32 residuals in fp32 (interleaved partial sums), then into the fp64 sum (comparator.f90:639-659 sums in double)
                        float pa, pb;
                        if (l1) {
                            pa = 0.f; pb = 0.f;
#pragma unroll
                            for (int u = 0; u < 16; u++) { pa += fabsf(__uint_as_float(v[2 * u])); pb += fabsf(__uint_as_float(v[2 * u + 1])); }
                        } else {
                            u64 p01 = 0ull, p23 = 0ull;
#pragma unroll
                            for (int u4 = 0; u4 < 8; u4++) {
                                const u64 r01 = pk2(__uint_as_float(v[4 * u4]), __uint_as_float(v[4 * u4 + 1])), r23 = pk2(__uint_as_float(v[4 * u4 + 2]), __uint_as_float(v[4 * u4 + 3]));
                                asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(p01) : "l"(r01));
                                asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(p23) : "l"(r23));
                            }
                            float a0, a1, a2, a3;
                            unpk2(p01, a0, a1); unpk2(p23, a2, a3);
                            pa = a0 + a1; pb = a2 + a3;
                        }
                        acc += (double)(pa + pb);
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncthreads();      // accumulator and operand tiles are free again
                }
                // (iii) right of the data span the synthetic repeats its last sample
                for (int x = max(max(p0, sds0), sds1 + 1); x <= p1; x++) {
                    const float r = (fa * refval(x) - fb * (e_last * tapval(x))) * rs;
                    acc += l1 ? (double)fabsf(r) : (double)r * (double)r;
                }
                acc = l1 ? acc / (double)rs : acc / ((double)rs * (double)rs);
                // reference-only norm (the same for all candidates).  Untapered it is the sum over the reference's data span, which the host
                // has formed once per receiver component; with a taper the span depends on the synthetic's: block reduction
                double accn;
                if (!tapered && q0 == rds0 && q1 == rds1) accn = l1 ? R.ref_sa[ic] : R.ref_ss[ic];
                else {
                    accn = 0.;
                    for (int x = q0 + tid; x <= q1; x += MTC_M) { const float a = refval(x); accn += l1 ? (double)fabsf(a) : (double)a * (double)a; }
                    for (int ofs = 16; ofs; ofs >>= 1) accn += __shfl_xor_sync(0xffffffffu, accn, ofs);
                    const unsigned rb = nred++ & 1u;   // alternating buffers: one barrier between the writes and the reads is enough
                    if (lane == 0) s_red[rb][warp] = accn;
                    __syncthreads();
                    accn = s_red[rb][0] + s_red[rb][1] + s_red[rb][2] + s_red[rb][3];
                }
                if (o) {
                    float mis, nf;
                    if (l1) { mis = (float)((double)dt * acc); nf = fa * (float)((double)dt * accn); }
                    else { mis = (float)sqrt((double)dt * acc); nf = fa * (float)sqrt((double)dt * accn); }
                    if (p1 < p0) mis = 0.f;
                    o[0] = mis; o[1] = nf;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(MTC_N) : "memory");
}

// =================================================================================================
// Ground-motion diagnostics of the synthetics (get_peak_amplitudes / get_arias_intensities, minimizer_engine.f90:1174-1245;
// receiver_get_maxabs / receiver_get_arias_intensity receiver.f90:544-596; max_vecnorm_d1/d2_*, arias_intensity_*
// comparator.f90:519-625 through probes_norm_timedomain[_3] :700-765).  One warp per (candidate, receiver):
// out[pair][0..2] = peak vector norm of the velocity, of the acceleration, Arias intensity, over the vertical and/or
// the complete horizontal pair of components (get_component_ids receiver.f90:505-542).  Fresh-state probe spans.
// =================================================================================================
__global__ void __launch_bounds__(128) k_ground_motion(const ReceiverDev* __restrict__ rcv, int nrcv, const CandDev* __restrict__ cands, int ncand,
                                                        const float* __restrict__ seis, size_t seis_stride, const SeisHdr* __restrict__ shdrs,
                                                        const float* __restrict__ taperdata, float dt, float syn_factor,
                                                        float* __restrict__ out /* [ncand][nrcv][3] */) {
    const int lane = threadIdx.x & 31;
    const long long pair = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pair >= (long long)ncand * nrcv) return;
    const int b = (int)(pair / nrcv), ir = (int)(pair % nrcv);
    const ReceiverDev& R = rcv[ir];
    float* o = out + (size_t)pair * 3;
    const CandDev cand = cands[b];
    if (lane == 0) { o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; }
    if (!R.enabled || cand.status != 0) { if (lane == 0 && cand.status != 0) { o[0] = nanf(""); o[1] = nanf(""); o[2] = nanf(""); } return; }
    // components: vertical, and the horizontal pair a/c + r/l or else n/s + e/w (only if complete)
    int iver = -1, ih1 = -1, ih2 = -1;
    for (int ic = 0; ic < R.ncomp; ic++) { const int t = abs(R.comp[ic]); if (t == 1) ih1 = ic; if (t == 2) ih2 = ic; if (t == 3) iver = ic; }
    if (ih1 < 0 || ih2 < 0) for (int ic = 0; ic < R.ncomp; ic++) { const int t = abs(R.comp[ic]); if (t == 4) ih1 = ic; if (t == 5) ih2 = ic; }
    if (ih1 < 0 || ih2 < 0) { ih1 = -1; ih2 = -1; }
    int comp[3], nc = 0;
    if (iver >= 0) comp[nc++] = iver;
    if (ih1 >= 0) { comp[nc++] = ih1; comp[nc++] = ih2; }
    if (nc == 0) return;
    const float* row[3]; int ds0[3], ds1[3], rbase[3];
    for (int k = 0; k < nc; k++) {
        const SeisHdr sh = shdrs[(size_t)pair * KIWI_MAX_COMP + comp[k]];
        if (sh.hi < sh.lo) return;                 // nothing synthesised for a component: leave the zeros
        row[k] = seis + ((size_t)pair * KIWI_MAX_COMP + comp[k]) * seis_stride;
        ds0[k] = sh.lo; ds1[k] = sh.hi; rbase[k] = sh.base;
    }
    // common probe span F (probe_set_array comparator.f90:240-256, probes_adjust_spans[_3] :464-517)
    int u0 = ds0[0], u1 = ds1[0], minlength = 0, F0, F1;
    int sp0[3], sp1[3];
    for (int k = 0; k < nc; k++) {
        allowed_span(ds0[k], ds1[k], ceil_len2(ds1[k] - ds0[k] + 1), sp0[k], sp1[k]);
        u0 = min(u0, ds0[k]); u1 = max(u1, ds1[k]); minlength = max(minlength, ceil_len2(ds1[k] - ds0[k] + 1));
    }
    allowed_span(u0, u1, minlength, F0, F1);
    bool same = (sp1[0] - sp0[0]) == (F1 - F0);
    for (int k = 1; k < nc; k++) same = same && sp0[k] == sp0[0] && sp1[k] == sp1[0];
    for (int k = 0; k < nc; k++) for (int j = 0; j < nc; j++) same = same && sp0[k] <= ds0[j] && ds1[j] <= sp1[k];
    if (same || nc == 1) { F0 = sp0[0]; F1 = sp1[0]; }
    int s0, s1;
    if (R.has_taper) { s0 = max(R.dps0, F0); s1 = min(R.dps1, F1); } else { s0 = u0; s1 = u1; }
    if (s1 < s0) return;                           // "applying timedomain norm to empty region": 0
    const float* tp = taperdata + R.taper_off;
    const float moment = cand.moment;
    auto val = [&](int k, int x) -> float {      // probe array with the continuation rule (:264-267), tapered (:1173-1184)
        if (x < ds0[k]) return 0.f;
        float v = row[k][min(x, ds1[k]) - rbase[k]] * moment;
        if (R.has_taper) v = (x >= R.tp0 && x <= R.tp1) ? v * tp[x - R.tp0] : 0.f;
        return v;
    };
    const double f2 = (double)(syn_factor * syn_factor);
    double m1 = -DBL_MAX, m2 = -DBL_MAX, sum2 = 0.;
    for (int x = s0 + lane; x < s1; x += 32) {
        double a1 = 0., a2 = 0.;
        for (int k = 0; k < nc; k++) {
            const float v0 = val(k, x), v1 = val(k, x + 1);
            const double d1 = (double)(v0 - v1);
            a1 += f2 * (d1 * d1);
            if (x + 2 <= s1) { const double d2 = (double)(v0 - 2.0f * v1 + val(k, x + 2)); a2 += f2 * (d2 * d2); }
        }
        m1 = fmax(m1, a1);
        if (x + 2 <= s1) { m2 = fmax(m2, a2); sum2 += a2; }
    }
    m1 = warp_max_d(m1); m2 = warp_max_d(m2); sum2 = warp_sum_d(sum2);
    if (lane == 0) {
        const int n = s1 - s0 + 1;
        const float pi_f = 3.14159265358979f;
        if (n >= 2) o[0] = (float)(sqrt(m1) / (double)dt);
        if (n >= 3) {
            o[1] = (float)(sqrt(m2) / (double)(dt * dt));
            o[2] = (float)((double)(pi_f / (2.f * 9.81f) * dt) * sum2 / (double)(dt * dt));
        }
    }
}

// =================================================================================================
// Outer misfit (python/tunguska/seismosizer.py:843-922 make_global_misfits): per-receiver norms over
// components, receiver weights, optional "anarchy" normalisation, optional bootstrap re-weighting of
// the receivers, global misfit per candidate; then the best candidate per bootstrap row.  Doing it
// here keeps the [ns, nmisfits, 2] cube on the GPU: only [rows, ns] (or just the minima) goes back.
// The reference does this in numpy float64, so it is done in fp64 here too.
// Linear form: both outer norms are dot products of the bootstrap counts with per-receiver terms
//   l2: a_r = (m_sr w_r)^2, b_r = (n_sr w_r)^2, misfit = sqrt(sum bw_r a_r / sum bw_r b_r)
//   l1: a_r =  m_sr w_r,    b_r =  n_sr w_r,    misfit =      sum bw_r a_r / sum bw_r b_r
// =================================================================================================
struct OuterRcv { int misfit_base, ncomp; };   // enabled receivers only: ncomp > 0; disabled: ncomp = 0

__global__ void __launch_bounds__(128) k_outer_misfits(const float* __restrict__ mis /* [ns][nm][2] */, int nm, const OuterRcv* __restrict__ rc,
                                                        int nr, const double* __restrict__ rweights /* [nr] or null */, int l1, int anarchy,
                                                        int nrows, const double* __restrict__ bweights /* [nrows-1][nr] or null */,
                                                        double* __restrict__ out /* [nrows][ns] */, int ns, int row0) {
    // rows row0 .. row0 + nrows - 1 of the [1 + nboot][ns] matrix (row 0: no bootstrap weights), written to out[0 .. nrows)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* a = reinterpret_cast<double*>(smem_raw);
    double* b = a + nr;
    __shared__ double red[2][4];
    const int s = blockIdx.x;
    const float* m = mis + (size_t)s * nm * 2;
    for (int r = threadIdx.x; r < nr; r += blockDim.x) {
        double ms = 0., ns_ = 0.;
        const OuterRcv R = rc[r];
        for (int c = 0; c < R.ncomp; c++) {
            const double mv = (double)m[(size_t)(R.misfit_base + c) * 2], nv = (double)m[(size_t)(R.misfit_base + c) * 2 + 1];
            if (l1) { ms += mv; ns_ += nv; } else { ms += mv * mv; ns_ += nv * nv; }
        }
        if (!l1) { ms = sqrt(ms); ns_ = sqrt(ns_); }
        double w = rweights ? rweights[r] : 1.0;
        if (anarchy) w = fmax(w / (ns_ != 0. ? ns_ : -1.), 0.);
        const double mw = ms * w, nw = ns_ * w;
        a[r] = l1 ? mw : mw * mw;
        b[r] = l1 ? nw : nw * nw;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int row = 0; row < nrows; row++) {
        const double* bw = (row0 + row == 0 || !bweights) ? nullptr : bweights + (size_t)(row0 + row - 1) * nr;
        double sa = 0., sb = 0.;
        for (int r = threadIdx.x; r < nr; r += blockDim.x) {
            const double k = bw ? bw[r] : 1.0;
            sa += k * a[r]; sb += k * b[r];
        }
        for (int o = 16; o; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
        if (lane == 0) { red[0][warp] = sa; red[1][warp] = sb; }
        __syncthreads();
        if (threadIdx.x == 0) {
            const double A = red[0][0] + red[0][1] + red[0][2] + red[0][3], B = red[1][0] + red[1][1] + red[1][2] + red[1][3];
            double g = B > 0. ? (l1 ? A / B : sqrt(A / B)) : -1.;
            if (g < 0. || !(g == g)) g = nan("");
            out[(size_t)row * ns + s] = g;
        }
        __syncthreads();
    }
}
// nanargmin per row (gridsearch.py:250-266)
__global__ void __launch_bounds__(256) k_row_argmin(const double* __restrict__ v, int ns, int* __restrict__ best, double* __restrict__ bestv) {
    __shared__ double sv[256]; __shared__ int si[256];
    const double* row = v + (size_t)blockIdx.x * ns;
    double mv = INFINITY; int mi = -1;
    for (int i = threadIdx.x; i < ns; i += blockDim.x) { const double x = row[i]; if (x == x && (x < mv || mi < 0)) { mv = x; mi = i; } }
    sv[threadIdx.x] = mv; si[threadIdx.x] = mi;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if ((int)threadIdx.x < o) {
            const int j = threadIdx.x + o;
            if (si[j] >= 0 && (si[threadIdx.x] < 0 || sv[j] < sv[threadIdx.x] || (sv[j] == sv[threadIdx.x] && si[j] < si[threadIdx.x]))) { sv[threadIdx.x] = sv[j]; si[threadIdx.x] = si[j]; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { best[blockIdx.x] = si[0]; bestv[blockIdx.x] = si[0] >= 0 ? sv[0] : nan(""); }
}

// flag[row] = 1 where row `row` of v[nrows][ncols] holds a NaN or an Inf (status 2 of a candidate, minimizer_engine.f90:1163-1166)
__global__ void __launch_bounds__(256) k_flag_nonfinite(const float* __restrict__ v, int nrows, int ncols, int* __restrict__ flag, int* __restrict__ count) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= nrows) return;
    bool bad = false;
    for (int i = lane; i < ncols; i += 32) bad = bad || !isfinite(v[(size_t)row * ncols + i]);
    if (__any_sync(0xffffffffu, bad) && lane == 0) { flag[row] = 1; if (count) atomicAdd(count, 1); }
}
void launch_flag_nonfinite(const float* v, int nrows, int ncols, int* flag, int* count, cudaStream_t st) {
    if (nrows > 0 && ncols > 0) k_flag_nonfinite<<<(nrows + 7) / 8, 256, 0, st>>>(v, nrows, ncols, flag, count);
}

// ---- host-callable launch wrappers ---------------------------------------------------------------
void launch_bilat_groups(const BilatCand* d_cands, int ncand, GroupSoA g, TapSoA taps, float dt, int ngroups_total, cudaStream_t st) {
    if (ncand > 0) k_bilat_groups<<<ncand, 128, 0, st>>>(d_cands, g, taps, dt, ngroups_total);
}
void launch_group_tap_range(GroupSoA g, TapSoA taps, float dt, int gbegin, int gend, cudaStream_t st) {
    int n = gend - gbegin;
    if (n > 0) k_group_tap_range<<<(n + 127) / 128, 128, 0, st>>>(g, taps, dt, gbegin, gend);
}
void launch_tap_table(GroupSoA g, TapSoA taps, float dt, int ngroups, cudaStream_t st) {
    if (ngroups > 0) k_tap_table<<<(ngroups + 7) / 8, 256, 0, st>>>(g, taps, dt, ngroups);
}
void launch_expand_centroids(CandDev cand, GroupSoA g, TapSoA taps, int ngroups_total, float* d_table, int cap, cudaStream_t st) {
    k_expand_centroids<<<(cand.ngroups + 127) / 128, 128, 0, st>>>(cand, g, taps, ngroups_total, d_table, cap);
}
void launch_geometry(GfdbDev db, const ReceiverDev* rcv, int nrcv, const CandDev* cands, int ncand, GroupSoA g, int ngroups_total,
                     int interpolate, int xunder, int zunder, GeoRec* recs, size_t rec_stride, PairHdr* hdrs, int* tmax, cudaStream_t st, int trig_only,
                     float* azf_out) {
    const int npairs = ncand * nrcv;
    if (npairs <= 0) return;
    if (rec_stride == 1) {   // single-group candidates: one thread per pair
        k_geometry<true><<<(npairs + 127) / 128, 128, 0, st>>>(db, rcv, nrcv, cands, g, ngroups_total, interpolate, xunder, zunder, recs, rec_stride, hdrs,
                                                           tmax, npairs, trig_only, azf_out);
        return;
    }
    // one thread per group: a CTA no wider than the longest group list
    const int threads = (int)std::min<size_t>(256, std::max<size_t>(32, (rec_stride + 31) / 32 * 32));
    k_geometry<false><<<npairs, threads, 0, st>>>(db, rcv, nrcv, cands, g, ngroups_total, interpolate, xunder, zunder, recs, rec_stride, hdrs, tmax,
                                                 npairs, trig_only, azf_out);
}
size_t synth_smem_bytes(int nwarps, int nq) {
    return (size_t)nwarps * 3 * nq * (sizeof(float4) + sizeof(float)) + 16 + (size_t)nwarps * 3 * sizeof(GeoRec) +
           (size_t)nwarps * SYN_STAGES * 4 * 32 * sizeof(float4);
}
size_t synth_partial_bytes(int nq) { return (3 * (size_t)nq + (3 * (size_t)nq + 3) / 4) * sizeof(float4); }
cudaError_t launch_synth(GfdbDev db, const ReceiverDev* rcv, int nrcv, const CandDev* cands, int ncand, GroupSoA g, const GeoRec* recs,
                         size_t rec_stride, const PairHdr* hdrs, int nq_alloc, int margin_q, int nwarps, float* seis, size_t seis_stride,
                         SeisHdr* shdrs, int nbands, float* partial, cudaStream_t st) {
    size_t smem = synth_smem_bytes(nwarps, nq_alloc);
    if (const char* ev = getenv("KIWI_SYNTH_SMEM_PAD")) smem += (size_t)atoi(ev);   // occupancy experiments
    cudaError_t e = cudaFuncSetAttribute(k_synth, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    for (int band = 0; band < nbands; band++)
        k_synth<<<ncand * nrcv, nwarps * 32, smem, st>>>(db, rcv, nrcv, cands, g, recs, rec_stride, hdrs, nq_alloc, margin_q, seis, seis_stride, shdrs, -0.0f,
                                                         band, reinterpret_cast<float4*>(partial));
    return cudaGetLastError();
}
void launch_misfit_td(const ReceiverDev* rcv, int nrcv, const CandDev* cands, int ncand, const float* seis, size_t seis_stride,
                      const SeisHdr* shdrs, const float* refdata, const float* taperdata, int method, float dt, float syn_factor,
                      int nmisfits, float* out, int* status, const CandMap* map, cudaStream_t st) {
    long long nitems = (long long)ncand * nrcv * KIWI_MAX_COMP;
    int blocks = (int)((nitems + 3) / 4);
    if (blocks > 0)
        k_misfit_td<<<blocks, 128, 0, st>>>(rcv, nrcv, cands, ncand, seis, seis_stride, shdrs, refdata, taperdata, method, dt, syn_factor,
                                            nmisfits, out, status, map);
}

void launch_ground_motion(const ReceiverDev* rcv, int nrcv, const CandDev* cands, int ncand, const float* seis, size_t seis_stride,
                          const SeisHdr* shdrs, const float* taperdata, float dt, float syn_factor, float* out, cudaStream_t st) {
    const long long npairs = (long long)ncand * nrcv;
    if (npairs > 0) k_ground_motion<<<(int)((npairs + 3) / 4), 128, 0, st>>>(rcv, nrcv, cands, ncand, seis, seis_stride, shdrs, taperdata, dt, syn_factor, out);
}

size_t misfit_general_smem_bytes(int n_alloc, int nshift_alloc) {   // n_alloc = 0: the transform buffer lives in global memory
    return (size_t)n_alloc * sizeof(float2) + (size_t)2 * nshift_alloc * KIWI_MAX_COMP * sizeof(float);
}
cudaError_t launch_misfit_general(const ReceiverDev* rcv, int nrcv, const CandDev* cands, int ncand, const float* seis, size_t seis_stride,
                                  const SeisHdr* shdrs, const float* refdata, const float* taperdata, const float2* tw, int tw_n, int method,
                                  float dt, float syn_factor, int nmisfits, float* out, int* status, int* fshift, int n_alloc,
                                  int nshift_alloc, const CandMap* map, cudaStream_t st, int xs0, int xs1, int premethod, float2* zscratch) {
    const size_t smem = misfit_general_smem_bytes(zscratch ? 0 : n_alloc, nshift_alloc);
    cudaError_t e = cudaFuncSetAttribute(k_misfit_general, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (ncand * nrcv > 0)
        k_misfit_general<<<ncand * nrcv, 256, smem, st>>>(rcv, nrcv, cands, seis, seis_stride, shdrs, refdata, taperdata, tw, tw_n, method, dt,
                                                         syn_factor, nmisfits, out, status, fshift, n_alloc, nshift_alloc, map, xs0, xs1, premethod,
                                                         zscratch);
    return cudaGetLastError();
}

cudaError_t launch_probe_export(const ReceiverDev* rcv, int ir, int ic, const CandDev* cands, const float* seis, size_t seis_stride, const SeisHdr* shdrs, int nrcv,
                                const float* refdata, const float* taperdata, const float2* tw, int tw_n, int which_probe, int processing, int spectrum,
                                float dt, int n_alloc, int* hdr, float* out, cudaStream_t st) {
    const size_t smem = sizeof(float2) * (size_t)n_alloc;
    cudaError_t e = cudaFuncSetAttribute(k_probe_export, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_probe_export<<<1, 256, smem, st>>>(rcv, ir, ic, cands, seis, seis_stride, shdrs, nrcv, refdata, taperdata, tw, tw_n, which_probe, processing, spectrum, dt,
                                         n_alloc, hdr, out);
    return cudaGetLastError();
}

void launch_mt_contract(const ReceiverDev* rcv, int nrcv, const MtLoc* locs, int nloc, const float* mts, const int* cand_of, const float* seis,
                        size_t seis_stride, const SeisHdr* shdrs, const float* refdata, const float* taperdata, int method, float dt,
                        float syn_factor, int nmisfits, float* out, cudaStream_t st) {
    if (nloc * nrcv > 0)
    {
        // receivers per CTA: enough CTAs for a few waves over the 148 SMs x 4 resident CTAs, as few set-ups as possible
        int rpc = 1;
        while (rpc < nrcv && (long long)nloc * ((nrcv + 2 * rpc - 1) / (2 * rpc)) >= 148LL * 4 * 6) rpc *= 2;
        const int nrblk = (nrcv + rpc - 1) / rpc;
        k_mt_contract<<<nloc * nrblk, 128, 0, st>>>(rcv, nrcv, locs, mts, cand_of, seis, seis_stride, shdrs, refdata, taperdata, method, dt,
                                                   syn_factor, nmisfits, out, rpc);
    }
}

size_t mt_fused_smem_bytes(int strip_cap, int) {
    const size_t tiles = (size_t)(MTC_M + MTC_N) * MTF_K, u = ((size_t)10 * (strip_cap + 4) + 31) & ~(size_t)31;
    return (std::max(tiles, u) + (size_t)10 * (strip_cap + MTF_PAD + 4)) * sizeof(float);
}
int mt_fused_max_steps() { return MTF_MAXSTEP; }
cudaError_t launch_mt_fused(GfdbDev db, const ReceiverDev* rcv, int nrcv, const MtLoc* locs, int nloc, const float* mts, const int* cand_of,
                            const GeoRec* recs, const PairHdr* hdrs, const float4* taprec, int strip_cap, int ncomp_max, const float* refdata,
                            const float* taperdata, int method, float dt, float syn_factor, int nmisfits, float* out, int* overflow, cudaStream_t st) {
    const size_t smem = mt_fused_smem_bytes(strip_cap, ncomp_max);
    cudaError_t e = cudaFuncSetAttribute(k_mt_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (nloc * nrcv > 0) {
        // receivers per CTA: enough CTAs for a few waves over the 148 SMs x 3 resident CTAs, as few set-ups as possible
        int rpc = 1;
        while (rpc < nrcv && 2 * rpc <= 8 && (long long)nloc * ((nrcv + 2 * rpc - 1) / (2 * rpc)) >= 148LL * 4 * 8) rpc *= 2;   // (measured flat from 5 to 13, slower beyond)
        if (const char* ev = getenv("KIWI_MTF_RPC")) rpc = std::min(std::max(atoi(ev), 1), MTF_MAXRCV);   // tuning experiments
        int nrblk = (nrcv + rpc - 1) / rpc;
        rpc = (nrcv + nrblk - 1) / nrblk;          // blocks of equal size
        nrblk = (nrcv + rpc - 1) / rpc;
        k_mt_fused<<<nloc * nrblk, 128, smem, st>>>(db, rcv, nrcv, locs, mts, cand_of, recs, hdrs, taprec, strip_cap, refdata, taperdata, method, dt,
                                                   syn_factor, nmisfits, out, rpc, overflow);
    }
    return cudaGetLastError();
}

cudaError_t launch_fold(const ReceiverDev* rcv, int nrcv, const CandDev* cands, int ncand, float* seis, size_t seis_stride, SeisHdr* shdrs,
                        float dt, cudaStream_t st) {
    const size_t smem = 2 * seis_stride * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(k_fold, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long items = (long long)ncand * nrcv * KIWI_MAX_COMP;
    if (items > 0) k_fold<<<(unsigned)items, 256, smem, st>>>(rcv, nrcv, cands, seis, seis_stride, shdrs, dt);
    return cudaGetLastError();
}

cudaError_t launch_outer_misfits(const float* mis, int nm, const void* rc, int nr, const double* rweights, int l1, int anarchy, int nrows,
                                 const double* bweights, double* out, int ns, int* best, double* bestv, cudaStream_t st, int row0) {
    const size_t smem = (size_t)2 * nr * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(k_outer_misfits, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (ns > 0) {
        k_outer_misfits<<<ns, 128, smem, st>>>(mis, nm, (const OuterRcv*)rc, nr, rweights, l1, anarchy, nrows, bweights, out, ns, row0);
        k_row_argmin<<<nrows, 256, 0, st>>>(out, ns, best + row0, bestv + row0);
    }
    return cudaGetLastError();
}
