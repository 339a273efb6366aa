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

// ---- exactly-rounded fp32 helpers: never contracted into FMA by ptxas -------------------------
__device__ __forceinline__ float A_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float S_(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float M_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float D_(float a, float b) { return __fdiv_rn(a, b); }

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
        g.tap_begin[gi] = c.tap_begin;
        g.tap_count[gi] = c.nt;
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
            for (int m = 0; m < 6; m++) t[4 + m] = M_(g.mhat[(size_t)m * ngroups_total + gi], taps.wt[tb + k]);
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
                                            double& new_azimuth, double& new_backazimuth, double& new_dist) {
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
        double lambda = (double)atan2f(delta_y, delta_x);
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

// trace-span union of the GF components of one set at the corners of one group
__device__ __forceinline__ void set_span(const GfdbDev& db, const int inode[4], int ncorner, const int* igs, int nig, int& lo, int& hi) {
    lo = INT_MAX; hi = INT_MIN;
    for (int c = 0; c < ncorner; c++)
        for (int k = 0; k < nig; k++) {
            int2 s = __ldg(&db.tspan[(size_t)inode[c] * db.ng + (igs[k] - 1)]);
            lo = min(lo, s.x); hi = max(hi, s.y);
        }
}

__global__ void __launch_bounds__(256) k_geometry(GfdbDev db, const ReceiverDev* __restrict__ rcv, int nrcv,
                                                   const CandDev* __restrict__ cands, GroupSoA g, int interpolate, int xunder,
                                                   int zunder, GeoRec* __restrict__ recs, size_t rec_stride,
                                                   PairHdr* __restrict__ hdrs, int* __restrict__ tmax) {
    const int pair = blockIdx.x;
    const int b = pair / nrcv, ir = pair % nrcv;
    const ReceiverDev& R = rcv[ir];
    const CandDev cand = cands[b];
    GeoRec* myrecs = recs + (size_t)pair * rec_stride;
    __shared__ int red[8][10];
    __shared__ int s_lastrot;

    const bool need_h = (R.ja | R.jr | R.jn | R.je) != 0;
    const bool need_v = R.jd != 0;
    const int set1[4] = {1, 2, 3, 9}, set2[2] = {4, 5}, set3[4] = {6, 7, 8, 10};
    const int n1 = db.ng == 10 ? 4 : 3, n3 = db.ng == 10 ? 4 : 3;

    SpanAcc urot, n1all, n2all, s3; urot.init(); n1all.init(); n2all.init(); s3.init();
    int last_rot = -1, any_nonrot = 0;

    if (R.enabled && cand.status == 0) {
        for (int ip = threadIdx.x; ip < cand.ngroups; ip += blockDim.x) {
            const int gi = cand.group_begin + ip;
            const float dnorth = g.north[gi], deast = g.east[gi], depth = g.depth[gi];
            double azi, bazi, dist;
            approx_differential_azidist(dnorth, deast, R.azi0, R.bazi0, R.dist0, azi, bazi, dist);
            GeoRec rec;
            rec.azi = (float)azi;
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
                const float fx = ax - floorf(ax), fz = az - floorf(az);
                const float tx = 4.f * 1.1920929e-7f * fmaxf(fabsf(ax), 1.f), tz = 4.f * 1.1920929e-7f * fmaxf(fabsf(az), 1.f);
                if (fx < tx || 1.f - fx < tx || fz < tz || 1.f - fz < tz) flags |= GEO_NEAR;
            } else {            // gfdb_get_indices gfdb.f90:781-792 (nint: half away from zero)
                const float ax = D_(S_(x, db.firstx), db.dx), az = D_(S_(z, db.firstz), db.dz);
                ix1 = (int)roundf(ax) + 1; iz1 = (int)roundf(az) + 1;
                ix2 = ix1 + 1; iz2 = iz1 + 1; dix = 0.f; diz = 0.f;
                const float fx = fabsf(fabsf(ax - floorf(ax)) - 0.5f), fz = fabsf(fabsf(az - floorf(az)) - 0.5f);
                const float tx = 4.f * 1.1920929e-7f * fmaxf(fabsf(ax), 1.f), tz = 4.f * 1.1920929e-7f * fmaxf(fabsf(az), 1.f);
                if (fx < tx || fz < tz) flags |= GEO_NEAR;
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
            for (int c = 0; c < ncorner; c++) {
                if (cx[c] < 1 || cx[c] > db.nx || cz[c] < 1 || cz[c] > db.nz) { ok = false; inode[c] = 0; continue; }
                inode[c] = (cx[c] - 1) * db.nz + (cz[c] - 1);
                if (__ldg(&db.nodes[inode[c]].off) == ~0ull) ok = false;
            }
            if (!ok) flags |= GEO_SKIP;
            rec.ix1 = ix1; rec.iz1 = iz1; rec.dix = dix; rec.diz = diz; rec.flags = flags;
            myrecs[ip] = rec;
            if (ok && (need_h || need_v)) {
                const int smin = g.its_min[gi], smax = g.its_max[gi];
                int lo, hi;
                if (need_h) {
                    int lo1, hi1, lo2, hi2;
                    set_span(db, inode, ncorner, set1, n1, lo1, hi1);
                    set_span(db, inode, ncorner, set2, 2, lo2, hi2);
                    lo1 += smin; hi1 += smax + 1; lo2 += smin; hi2 += smax + 1;   // sparse_trace.f90:649-653
                    if (flags & GEO_ROT) { urot.add(lo1, hi1); urot.add(lo2, hi2); last_rot = max(last_rot, ip); }
                    else { n1all.add(lo1, hi1); n2all.add(lo2, hi2); any_nonrot = 1; }
                }
                if (need_v) { set_span(db, inode, ncorner, set3, n3, lo, hi); s3.add(lo + smin, hi + smax + 1); }
            }
        }
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
            int lo1, hi1, lo2, hi2;
            set_span(db, inode, ncorner, set1, n1, lo1, hi1);
            set_span(db, inode, ncorner, set2, 2, lo2, hi2);
            n1before.add(lo1 + smin, hi1 + smax + 1); n2before.add(lo2 + smin, hi2 + smax + 1);
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
        if (h.T > 0) atomicMax(tmax, h.T);
    }
}

// =================================================================================================
// K3: synthesis.  One CTA per (candidate, receiver); every warp owns a private set of three
// accumulator strips (displacement_ar(1), displacement_ar(2), vertical) in shared memory and works
// through its share of the groups.  Per group and 128-sample chunk a lane owns one aligned sample
// quad: 4 corners x ng rows are fetched with coalesced 128-bit loads from the HBM slabs, combined
// bilinearly (gfdb.f90:943-948), weighted with the moment-tensor/azimuth factors (make_weights
// seismogram.f90:316-336), rotated by the centroid's back-azimuth difference (:196-203), and then
// added nt times with the sample shift and linear sub-sample interpolation of trace_multiply_add
// (sparse_trace.f90:639-705).  The "last sample repeats for ever" rule (:696-703) becomes a step
// per (group, tap) that is prefix-summed once at the end.
// =================================================================================================
#define SYN_MAXTAPS 32

__device__ __forceinline__ NodeInfo ld_node(const NodeInfo* p) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    NodeInfo n; n.off = ((unsigned long long)u.y << 32) | u.x; n.w0 = (int)u.z; n.wn = (int)u.w;
    return n;
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void fma4(float4& a, float s, const float4& v) {
    a.x = fmaf(s, v.x, a.x); a.y = fmaf(s, v.y, a.y); a.z = fmaf(s, v.z, a.z); a.w = fmaf(s, v.w, a.w);
}
__device__ __forceinline__ float4 shfl_up4(const float4& v, int d) {
    float4 r;
    r.x = __shfl_up_sync(0xffffffffu, v.x, d); r.y = __shfl_up_sync(0xffffffffu, v.y, d);
    r.z = __shfl_up_sync(0xffffffffu, v.z, d); r.w = __shfl_up_sync(0xffffffffu, v.w, d);
    return r;
}
__device__ __forceinline__ float4 shfl4(const float4& v, int src) {
    float4 r;
    r.x = __shfl_sync(0xffffffffu, v.x, src); r.y = __shfl_sync(0xffffffffu, v.y, src);
    r.z = __shfl_sync(0xffffffffu, v.z, src); r.w = __shfl_sync(0xffffffffu, v.w, src);
    return r;
}

// out quad += wl * A(y - its) + wr * A(y - its - 1) for the four samples of one output quad;
// P = A(4q-4..4q-1), C = A(4q..4q+3), s = its mod 4 (warp-uniform)
template <int S>
__device__ __forceinline__ void tap_quad(float4& o, const float4& P, const float4& C, float wl, float wr) {
    // E[i] = A(4q-4+i), i=0..7;  o[j] += wl*E[4-S+j] + wr*E[3-S+j]
    const float E[8] = {P.x, P.y, P.z, P.w, C.x, C.y, C.z, C.w};
    o.x = fmaf(wl, E[4 - S], o.x); o.x = fmaf(wr, E[3 - S], o.x);
    o.y = fmaf(wl, E[5 - S], o.y); o.y = fmaf(wr, E[4 - S], o.y);
    o.z = fmaf(wl, E[6 - S], o.z); o.z = fmaf(wr, E[5 - S], o.z);
    o.w = fmaf(wl, E[7 - S], o.w); o.w = fmaf(wr, E[6 - S], o.w);
}
__device__ __forceinline__ void tap_apply(float4* acc, int qrel, int nq, int s, const float4& P, const float4& C, float wl, float wr) {
    if (qrel < 0 || qrel >= nq) return;
    float4 o = acc[qrel];
    switch (s) {
        case 0: tap_quad<0>(o, P, C, wl, wr); break;
        case 1: tap_quad<1>(o, P, C, wl, wr); break;
        case 2: tap_quad<2>(o, P, C, wl, wr); break;
        default: tap_quad<3>(o, P, C, wl, wr); break;
    }
    acc[qrel] = o;
}

__global__ void __launch_bounds__(256, 2) k_synth(GfdbDev db, const ReceiverDev* __restrict__ rcv, int nrcv,
                                                   const CandDev* __restrict__ cands, GroupSoA g, TapSoA taps, int ngroups_total,
                                                   int interpolate, int xunder, int zunder, const GeoRec* __restrict__ recs,
                                                   size_t rec_stride, const PairHdr* __restrict__ hdrs, int nq_alloc,
                                                   float* __restrict__ seis, size_t seis_stride /* floats per component row */,
                                                   SeisHdr* __restrict__ shdrs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int pair = blockIdx.x;
    const int b = pair / nrcv, ir = pair % nrcv;
    const ReceiverDev& R = rcv[ir];
    const CandDev cand = cands[b];
    const PairHdr H = hdrs[pair];
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    SeisHdr* myshdr = shdrs + (size_t)pair * KIWI_MAX_COMP;
    if (H.T <= 0 || !R.enabled) {
        if (threadIdx.x < KIWI_MAX_COMP) { SeisHdr e; e.lo = 0; e.hi = -1; e.base = 0; e.pad = 0; myshdr[threadIdx.x] = e; }
        return;
    }
    const int base = floor4(H.out0);
    const int baseq = base >> 2;
    const int nq = nq_alloc;   // quads per accumulator strip
    // shared memory: per warp 3 strips of nq float4 + 3 step rows of nq floats
    float4* acc_all = reinterpret_cast<float4*>(smem_raw);
    float* step_all = reinterpret_cast<float*>(acc_all + (size_t)nwarps * 3 * nq);
    float4* acc = acc_all + (size_t)warp * 3 * nq;
    float* step = step_all + (size_t)warp * 3 * nq;
    for (int i = lane; i < 3 * nq; i += 32) { acc[i] = f4zero(); step[i] = 0.f; }
    __syncwarp();

    const bool need_h = (R.ja | R.jr | R.jn | R.je) != 0;
    const bool need_v = R.jd != 0;
    const bool ng10 = db.ng == 10;
    const GeoRec* myrecs = recs + (size_t)pair * rec_stride;
    const float dt = db.dt;

    for (int ip = warp; ip < cand.ngroups; ip += nwarps) {
        const GeoRec rec = myrecs[ip];
        if (rec.flags & GEO_SKIP) continue;
        const int gi = cand.group_begin + ip;
        // ---- moment tensor / azimuth weights, make_weights seismogram.f90:316-336 -------------
        float m[6];
#pragma unroll
        for (int k = 0; k < 6; k++) m[k] = g.mhat[(size_t)k * ngroups_total + gi];
        float sa, ca, s2a, c2a;
        sincosf(rec.azi, &sa, &ca);
        sincosf(2.f * rec.azi, &s2a, &c2a);
        const float f1 = m[0] * ca * ca + m[1] * sa * sa + m[3] * s2a;
        const float f2 = m[4] * ca + m[5] * sa;
        const float f3 = m[2];
        const float f4 = 0.5f * (m[1] - m[0]) * s2a + m[3] * c2a;
        const float f5 = m[5] * ca - m[4] * sa;
        const float f6 = m[0] * sa * sa + m[1] * ca * ca - m[3] * s2a;
        const float cl = rec.cl, sl = rec.sl, sd = R.sd;
        // ---- corners ------------------------------------------------------------------------------
        const bool single = rec.flags & GEO_SINGLE;
        const int ncorner = single ? 1 : 4;
        const int ix2 = rec.ix1 + (interpolate ? xunder : 1), iz2 = rec.iz1 + (interpolate ? zunder : 1);
        const float dix = rec.dix, diz = rec.diz;
        // gfdb.f90:943-948 weights, in the reference's association
        float wc[4] = {(1.f - dix) * (1.f - diz), (1.f - dix) * diz, dix * (1.f - diz), dix * diz};
        if (single) wc[0] = 1.f;
        const float* rowbase[4]; int w0[4], wn[4];
        int U0 = INT_MAX, U1 = INT_MIN;
        {
            const int cx[4] = {rec.ix1, rec.ix1, ix2, ix2}, cz[4] = {rec.iz1, iz2, rec.iz1, iz2};
#pragma unroll
            for (int c = 0; c < 4; c++) {
                if (c < ncorner) {
                    const NodeInfo ni = ld_node(&db.nodes[(cx[c] - 1) * db.nz + (cz[c] - 1)]);
                    rowbase[c] = db.slabs + ni.off; w0[c] = ni.w0; wn[c] = ni.wn;
                    U0 = min(U0, ni.w0); U1 = max(U1, ni.w0 + ni.wn);
                } else { rowbase[c] = db.slabs; w0[c] = 0; wn[c] = 4; wc[c] = 0.f; }
            }
        }
        // ---- taps: lane k prepares tap k (sparse_trace.f90:639-646) ---------------------------------
        const int tb = g.tap_begin[gi], tn = min(g.tap_count[gi], SYN_MAXTAPS);
        int my_its = 0; float my_wl = 0.f, my_wr = 0.f;
        if (lane < tn) {
            const float time = A_(g.tbase[gi], taps.toff[tb + lane]);
            const float rshift = D_(time, dt);
            my_its = (int)floorf(rshift);
            const float wr0 = S_(rshift, (float)my_its);
            const float wl0 = S_(1.f, wr0);
            const float wt = taps.wt[tb + lane];
            my_wr = M_(wr0, wt); my_wl = M_(wl0, wt);
        }
        // ---- sample loop ----------------------------------------------------------------------------
        float4 carry1 = f4zero(), carry2 = f4zero(), carry3 = f4zero();   // quad left of the chunk (zeros left of U0)
        const int q_first = U0 >> 2, q_last = U1 >> 2;                    // q_last = first constant quad
        float4 aend1 = f4zero(), aend2 = f4zero(), aend3 = f4zero();
        for (int q0 = q_first; q0 <= q_last; q0 += 32) {
            const int q = q0 + lane;
            const bool active = q <= q_last;
            float4 A1 = f4zero(), A2 = f4zero(), A3 = f4zero();
            if (active) {
                const int x = q << 2;
                int off[4]; float wcl[4];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int o = x - w0[c];
                    wcl[c] = (o < 0) ? 0.f : wc[c];
                    off[c] = min(max(o, 0), wn[c] - 4);
                    off[c] |= (o >= wn[c]) ? 0x40000000 : 0;   // flag: right of the window -> splat last sample
                }
                auto fetch = [&](int igm1) -> float4 {   // bilinear combination of one GF component
                    float4 r = f4zero();
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        if (c < ncorner) {
                            float4 v = ldg4(rowbase[c] + (size_t)igm1 * wn[c] + (off[c] & 0x3fffffff));
                            if (off[c] & 0x40000000) { v.x = v.w; v.y = v.w; v.z = v.w; }
                            fma4(r, wcl[c], v);
                        }
                    }
                    return r;
                };
                if (need_h) {
                    float4 Rr = f4zero(), Tt = f4zero();
                    float4 v;
                    v = fetch(0); fma4(Rr, f1, v);
                    v = fetch(1); fma4(Rr, f2, v);
                    v = fetch(2); fma4(Rr, f3, v);
                    if (ng10) { v = fetch(8); fma4(Rr, f6, v); }
                    v = fetch(3); fma4(Tt, f4, v);
                    v = fetch(4); fma4(Tt, f5, v);
                    // seismogram.f90:200-203: ar1 += cl*temp1 - sl*temp2; ar2 += cl*temp2 + sl*temp1
                    fma4(A1, cl, Rr); fma4(A1, -sl, Tt);
                    fma4(A2, cl, Tt); fma4(A2, sl, Rr);
                }
                if (need_v) {
                    float4 v;
                    v = fetch(5); fma4(A3, f1 * sd, v);
                    v = fetch(6); fma4(A3, f2 * sd, v);
                    v = fetch(7); fma4(A3, f3 * sd, v);
                    if (ng10) { v = fetch(9); fma4(A3, f6 * sd, v); }
                }
            }
            // previous quad: lane-1, lane 0 takes the carry of the previous chunk
            float4 P1 = shfl_up4(A1, 1), P2 = shfl_up4(A2, 1), P3 = shfl_up4(A3, 1);
            if (lane == 0) { P1 = carry1; P2 = carry2; P3 = carry3; }
            carry1 = shfl4(A1, 31); carry2 = shfl4(A2, 31); carry3 = shfl4(A3, 31);
            // the constant tail value lives in the quad q_last
            {
                const int src = q_last - q0;
                if (src >= 0 && src < 32) { aend1 = shfl4(A1, src); aend2 = shfl4(A2, src); aend3 = shfl4(A3, src); }
            }
            for (int k = 0; k < tn; k++) {
                const int its = __shfl_sync(0xffffffffu, my_its, k);
                const float wl = __shfl_sync(0xffffffffu, my_wl, k), wr = __shfl_sync(0xffffffffu, my_wr, k);
                if (active) {
                    const int mq = its >> 2, s = its & 3;
                    const int qrel = q + mq - baseq;
                    if (need_h) { tap_apply(acc, qrel, nq, s, P1, A1, wl, wr); tap_apply(acc + nq, qrel, nq, s, P2, A2, wl, wr); }
                    if (need_v) tap_apply(acc + 2 * nq, qrel, nq, s, P3, A3, wl, wr);
                }
            }
            __syncwarp();
        }
        // ---- end-value repetition (sparse_trace.f90:696-703): step at the first quad after the
        // last processed one, height (wl+wr)*A_end, one lane per tap -------------------------------------
        {   // lane c owns strip c: no two lanes touch the same word, order over taps is fixed
            const float ae = lane == 0 ? aend1.w : (lane == 1 ? aend2.w : aend3.w);
            const bool mine = lane < 3 && (lane < 2 ? need_h : need_v);
            for (int k = 0; k < tn; k++) {
                const int its = __shfl_sync(0xffffffffu, my_its, k);
                const float w = __shfl_sync(0xffffffffu, my_wl, k) + __shfl_sync(0xffffffffu, my_wr, k);
                const int qs = q_last + 1 + (its >> 2) - baseq;
                if (mine && qs >= 0 && qs < nq) step[lane * nq + qs] += w * ae;
            }
        }
        __syncwarp();
    }
    __syncthreads();
    // ---- reduce the warps' strips (fixed order: deterministic) -------------------------------------
    for (int i = threadIdx.x; i < 3 * nq; i += blockDim.x) {
        float4 s = acc_all[i]; float st = step_all[i];
        for (int w = 1; w < nwarps; w++) {
            const float4 v = acc_all[(size_t)w * 3 * nq + i];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            st += step_all[(size_t)w * 3 * nq + i];
        }
        acc_all[i] = s; step_all[i] = st;
    }
    __syncthreads();
    // inclusive prefix sum of the steps over quads, one warp per strip
    if (warp < 3) {
        float* st = step_all + (size_t)warp * nq;
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
    const float* a2 = a1 + (size_t)4 * nq;
    const float* a3 = a2 + (size_t)4 * nq;
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
// K5: scaling + time-domain misfit.  One warp per (candidate, receiver, component).
// =================================================================================================
__device__ __forceinline__ int next_pow2(int n) { int m = 1; while (m < n) m <<= 1; return m; }   // comparator.f90:1111-1118
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
                                                    int* __restrict__ status) {
    const int lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nitems = (long long)ncand * nrcv * KIWI_MAX_COMP;
    if (item >= nitems) return;
    const int ic = (int)(item % KIWI_MAX_COMP);
    const int pair = (int)(item / KIWI_MAX_COMP);
    const int b = pair / nrcv, ir = pair % nrcv;
    const ReceiverDev& R = rcv[ir];
    if (!R.enabled || ic >= R.ncomp) return;
    const CandDev cand = cands[b];
    float* o = out + ((size_t)b * nmisfits + R.misfit_base + ic) * 2;
    const SeisHdr sh = shdrs[(size_t)pair * KIWI_MAX_COMP + ic];
    if (cand.status != 0 || sh.hi < sh.lo) {
        if (lane == 0) { o[0] = nanf(""); o[1] = nanf(""); if (cand.status == 0) atomicMax(&status[b], 1); }
        return;
    }
    const float* srow = seis + ((size_t)pair * KIWI_MAX_COMP + ic) * seis_stride;
    const float moment = cand.moment;
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

// ---- host-callable launch wrappers ---------------------------------------------------------------
void launch_bilat_groups(const BilatCand* d_cands, int ncand, GroupSoA g, TapSoA taps, float dt, int ngroups_total, cudaStream_t st) {
    if (ncand > 0) k_bilat_groups<<<ncand, 128, 0, st>>>(d_cands, g, taps, dt, ngroups_total);
}
void launch_group_tap_range(GroupSoA g, TapSoA taps, float dt, int gbegin, int gend, cudaStream_t st) {
    int n = gend - gbegin;
    if (n > 0) k_group_tap_range<<<(n + 127) / 128, 128, 0, st>>>(g, taps, dt, gbegin, gend);
}
void launch_expand_centroids(CandDev cand, GroupSoA g, TapSoA taps, int ngroups_total, float* d_table, int cap, cudaStream_t st) {
    k_expand_centroids<<<(cand.ngroups + 127) / 128, 128, 0, st>>>(cand, g, taps, ngroups_total, d_table, cap);
}
void launch_geometry(GfdbDev db, const ReceiverDev* rcv, int nrcv, const CandDev* cands, int ncand, GroupSoA g, int interpolate,
                     int xunder, int zunder, GeoRec* recs, size_t rec_stride, PairHdr* hdrs, int* tmax, cudaStream_t st) {
    k_geometry<<<ncand * nrcv, 256, 0, st>>>(db, rcv, nrcv, cands, g, interpolate, xunder, zunder, recs, rec_stride, hdrs, tmax);
}
size_t synth_smem_bytes(int nwarps, int nq) { return (size_t)nwarps * 3 * nq * (sizeof(float4) + sizeof(float)); }
cudaError_t launch_synth(GfdbDev db, const ReceiverDev* rcv, int nrcv, const CandDev* cands, int ncand, GroupSoA g, TapSoA taps,
                         int ngroups_total, int interpolate, int xunder, int zunder, const GeoRec* recs, size_t rec_stride,
                         const PairHdr* hdrs, int nq_alloc, int nwarps, float* seis, size_t seis_stride, SeisHdr* shdrs,
                         cudaStream_t st) {
    size_t smem = synth_smem_bytes(nwarps, nq_alloc);
    cudaError_t e = cudaFuncSetAttribute(k_synth, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_synth<<<ncand * nrcv, nwarps * 32, smem, st>>>(db, rcv, nrcv, cands, g, taps, ngroups_total, interpolate, xunder, zunder, recs,
                                                    rec_stride, hdrs, nq_alloc, seis, seis_stride, shdrs);
    return cudaGetLastError();
}
void launch_misfit_td(const ReceiverDev* rcv, int nrcv, const CandDev* cands, int ncand, const float* seis, size_t seis_stride,
                      const SeisHdr* shdrs, const float* refdata, const float* taperdata, int method, float dt, float syn_factor,
                      int nmisfits, float* out, int* status, cudaStream_t st) {
    long long nitems = (long long)ncand * nrcv * KIWI_MAX_COMP;
    int blocks = (int)((nitems + 3) / 4);
    if (blocks > 0)
        k_misfit_td<<<blocks, 128, 0, st>>>(rcv, nrcv, cands, ncand, seis, seis_stride, shdrs, refdata, taperdata, method, dt, syn_factor,
                                            nmisfits, out, status);
}
