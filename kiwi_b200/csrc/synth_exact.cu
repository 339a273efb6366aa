// Reference-order synthesis (kiwi_set_accumulation(ctx, 1)): make_seismogram (seismogram.f90:131-289) with every floating-point
// operation of the reference applied to every output sample in the reference's order.
//
// k_synth (kernels.cu) sums the same terms in another order -- taps merged per quad shift, groups spread over warps, partial strips
// per depth band -- which is what makes it fast and what leaves it a few 1e-5 of the trace peak away from the reference's own
// sequentially accumulated fp32 result at ~1e4 sub-sources (the reference is that far from the exact sum of its terms itself,
// DESIGN.md section 5).  This kernel is the other trade: one thread per output sample walks all centroids one after the other,
//     per centroid:  temp1 = 0 (+) g1 (+) g2 (+) g3 (+) g9,  temp2 = 0 (+) g4 (+) g5          (trace_multiply_add, sparse_trace.f90:597-707)
//                    ar1 = ar1 + cl*temp1 - sl*temp2,  ar2 = ar2 + cl*temp2 + sl*temp1         (seismogram.f90:196-203)
//                    dz  = dz (+) g6 (+) g7 (+) g8 (+) g10
// where every (+) is the pair of rounded multiply-adds `+ wl*tr(x-its)`, `+ wr*tr(x-1-its)` of trace_multiply_add or, right of the
// trace, its single `+ factor*lastval`, and tr is the bilinear trace of gfdb_get_trace_bilin (gfdb.f90:865-950: four rounded
// products added in corner order).  The strips of the reference grow as centroids arrive (new samples repeat the last one,
// strip_extend sparse_trace.f90:316-345); a sample that enters a strip late holds exactly the sum of the tails it would have
// collected, so the per-sample sequence is the same whatever the growth order.  Compiled with -fmad=false: every operation is
// the IEEE operation the source shows.  The transcendentals whose device versions differ from the host library's often enough to
// show come from the host: atan2f of the sub-source azimuths (engine.cpp, every mode) and sinf / cosf of the per-(receiver,
// sub-source) azimuths (`trig`, this mode).  With them the seismograms equal the fp32 restatement of the Fortran path bit for bit
// (tests/test_reference_order_gpu.py: every sample of C3 at full size).
//
// One CTA per (candidate, receiver); per group the ten bilinear traces are formed once in shared memory (threads over trace samples),
// then every thread applies the group's centroids to its output samples.  The Green's function rows come in as whole node blocks:
// a node's ng rows are one contiguous, 16-byte aligned block (kiwi_dev.cuh), so one elected thread fetches the four corner blocks of
// the NEXT group with four bulk asynchronous copies (cp.async.bulk, completion counted in bytes on an mbarrier) into the second of two
// shared-memory stages while all threads work on the current one: the gather latency is off the critical path and each block is
// read from HBM / L2 once per (candidate, receiver, group) and then used by all 256 threads and all of the group's centroids.  19 ms per C3-sized candidate, 130 ms per C5-sized one:
// a regression / verification mode, 9-12 x the batched kernel's time.
#include "kiwi_dev.cuh"
#include "kernels.cuh"
#include <cfloat>
#include <climits>

#define SX_THREADS 256
#define SX_NS 4            // output samples per thread: windows of up to 1024 samples

namespace {

struct SxTrace {           // one bilinear (or single) trace of the current group in shared memory
    int s0, s1;            // its span (union over the corners), trace%span
    float lastval;         // its last sample
    int ok;
};

__device__ __forceinline__ float slab_at(const float* __restrict__ blk, const NodeInfo& n, int comp, int y) {
    // dense row of the staged node block: zeros left of the trace and in its gaps, the last sample repeated to the right
    const int i = min(max(y - n.w0, 0), n.wn - 1);
    return blk[comp * n.wn + i];
}
__device__ __forceinline__ unsigned sx_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// the corner blocks of the group whose record is *rec -> stage buffer; returns false (nothing issued) for a skipped group
__device__ __forceinline__ bool sx_issue(const GfdbDev& db, const GeoRec* rec, float* stage, int blk_cap /* floats per corner */, unsigned bar) {
    const int flags = rec->flags;
    if (flags & GEO_SKIP) return false;
    const int ncorner = (flags & GEO_SINGLE) ? 1 : 4;
    unsigned bytes = 0;
    for (int c = 0; c < ncorner; c++) bytes += (unsigned)(rec->node[c].wn * db.ng) * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    for (int c = 0; c < ncorner; c++) {
        const NodeInfo n = rec->node[c];
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sx_smem_u32(stage + (size_t)c * blk_cap)),
                     "l"(db.slabs + n.off), "r"((unsigned)(n.wn * db.ng) * 4u), "r"(bar)
                     : "memory");
    }
    return true;
}

}  // namespace

__global__ void __launch_bounds__(SX_THREADS) k_synth_exact(GfdbDev db, const ReceiverDev* __restrict__ rcv, int nrcv, const CandDev* __restrict__ cands,
                                                            GroupSoA g, TapSoA taps, int ngroups_total, const GeoRec* __restrict__ recs, size_t rec_stride,
                                                            const PairHdr* __restrict__ hdrs, int nq_alloc, int margin_q, int interpolate, int xunder,
                                                            int zunder, int wcap /* floats per trace row in shared memory */,
                                                            int blk_cap /* floats per staged node block */, float* __restrict__ seis,
                                                            size_t seis_stride, SeisHdr* __restrict__ shdrs, int* __restrict__ overflow,
                                                            const float4* __restrict__ trig) {
    extern __shared__ __align__(16) unsigned char sx_smem[];
    float* s_tr = reinterpret_cast<float*>(sx_smem);          // [10][wcap]
    float* s_stage = s_tr + (size_t)KIWI_NG_MAX * wcap;       // [2 stages][4 corners][blk_cap]: node blocks, 16-byte aligned
    __shared__ SxTrace s_meta[KIWI_NG_MAX];
    __shared__ __align__(16) GeoRec s_rec2[2];                // records of the current and the next group (with the stage they belong to)
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ int s_issued[2];
    const int pair = blockIdx.x;
    const int b = pair / nrcv, ir = pair % nrcv;
    const ReceiverDev& R = rcv[ir];
    const CandDev cand = cands[b];
    const PairHdr H = hdrs[pair];
    const int tid = threadIdx.x;
    SeisHdr* myshdr = shdrs + (size_t)pair * KIWI_MAX_COMP;
    if (H.T <= 0 || !R.enabled) {
        if (tid < KIWI_MAX_COMP) { SeisHdr e; e.lo = 0; e.hi = -1; e.base = 0; e.pad = 0; myshdr[tid] = e; }
        return;
    }
    const int base = (H.out0 & ~3) - 4 * margin_q;
    const int nsamp = 4 * nq_alloc;
    const bool need_h = (R.ja | R.jr | R.jn | R.je) != 0, need_v = R.jd != 0, ng10 = db.ng == 10;
    const float sd = R.sd, dt = db.dt;
    float ar0[SX_NS], ar1[SX_NS], dz[SX_NS];
#pragma unroll
    for (int j = 0; j < SX_NS; j++) { ar0[j] = 0.f; ar1[j] = 0.f; dz[j] = 0.f; }
    const GeoRec* myrecs = recs + (size_t)pair * rec_stride;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sx_smem_u32(&s_bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sx_smem_u32(&s_bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 8 && cand.ngroups > 0) reinterpret_cast<uint4*>(&s_rec2[0])[tid] = __ldg(reinterpret_cast<const uint4*>(myrecs) + tid);
    __syncthreads();
    if (tid == 0 && cand.ngroups > 0) s_issued[0] = sx_issue(db, &s_rec2[0], s_stage, blk_cap, sx_smem_u32(&s_bar[0])) ? 1 : 0;
    unsigned par0 = 0, par1 = 0;     // phase parities of the two stage barriers

    for (int ip = 0; ip < cand.ngroups; ip++) {
        const int st = ip & 1;
        __syncthreads();             // everybody is done with group ip - 1: its stage and record slot (the other ones) are free
        if (tid < 8 && ip + 1 < cand.ngroups) reinterpret_cast<uint4*>(&s_rec2[st ^ 1])[tid] = __ldg(reinterpret_cast<const uint4*>(myrecs + ip + 1) + tid);
        __syncthreads();
        if (tid == 0 && ip + 1 < cand.ngroups)   // the next group's blocks start travelling now
            s_issued[st ^ 1] = sx_issue(db, &s_rec2[st ^ 1], s_stage + (size_t)(st ^ 1) * 4 * blk_cap, blk_cap, sx_smem_u32(&s_bar[st ^ 1])) ? 1 : 0;
        const GeoRec& s_rec = s_rec2[st];
        const int flags = s_rec.flags;
        if (flags & GEO_SKIP) continue;          // a node is missing: the reference leaves the centroid (seismogram.f90:172)
        {   // this group's blocks have landed
            const unsigned bar = sx_smem_u32(&s_bar[st]), par = st ? par1 : par0;
            unsigned done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(par) : "memory");
            if (st) par1 ^= 1; else par0 ^= 1;
        }
        const float* blk0 = s_stage + (size_t)st * 4 * blk_cap;
        const bool single = flags & GEO_SINGLE;
        const NodeInfo n0 = s_rec.node[0], n1 = s_rec.node[1], n2 = s_rec.node[2], n3 = s_rec.node[3];
        const float dix = s_rec.dix, diz = s_rec.diz;
        const float w00 = (1.f - dix) * (1.f - diz), w01 = (1.f - dix) * diz, w10 = dix * (1.f - diz), w11 = dix * diz;   // gfdb.f90:943-948
        // window all four node windows lie in
        const int wlo = single ? n0.w0 : min(min(n0.w0, n1.w0), min(n2.w0, n3.w0));
        const int whi = single ? n0.w0 + n0.wn : max(max(n0.w0 + n0.wn, n1.w0 + n1.wn), max(n2.w0 + n2.wn, n3.w0 + n3.wn));
        const int wl_ = whi - wlo;
        if (wl_ > wcap || n0.wn * db.ng > blk_cap || (!single && (n1.wn * db.ng > blk_cap || n2.wn * db.ng > blk_cap || n3.wn * db.ng > blk_cap))) {
            if (tid == 0) atomicExch(overflow, 1);
            continue;
        }
        // node indices for the trace spans
        const int ix2 = s_rec.ix1 + (interpolate ? xunder : 1), iz2 = s_rec.iz1 + (interpolate ? zunder : 1);
        const int in0 = (s_rec.ix1 - 1) * db.nz + (s_rec.iz1 - 1), in1 = (s_rec.ix1 - 1) * db.nz + (iz2 - 1), in2 = (ix2 - 1) * db.nz + (s_rec.iz1 - 1),
                  in3 = (ix2 - 1) * db.nz + (iz2 - 1);
        // ---- the group's traces, once (gfdb_get_trace_bilin) ------------------------------------------------------------
        if (tid < db.ng) {
            const int k = tid;
            int2 a = __ldg(&db.tspan[(size_t)in0 * db.ng + k]);
            int s0 = a.x, s1 = a.y;
            if (!single) {
                const int2 b1 = __ldg(&db.tspan[(size_t)in1 * db.ng + k]), b2 = __ldg(&db.tspan[(size_t)in2 * db.ng + k]), b3 = __ldg(&db.tspan[(size_t)in3 * db.ng + k]);
                s0 = min(min(s0, b1.x), min(b2.x, b3.x)); s1 = max(max(s1, b1.y), max(b2.y, b3.y));
            }
            s_meta[k].s0 = s0; s_meta[k].s1 = s1; s_meta[k].ok = 1;
        }
        for (int idx = tid; idx < db.ng * wl_; idx += SX_THREADS) {
            const int k = idx / wl_, i = idx - k * wl_, y = wlo + i;
            float v;
            if (single) v = slab_at(blk0, n0, k, y);
            else {
                v = w00 * slab_at(blk0, n0, k, y);
                v = v + w01 * slab_at(blk0 + blk_cap, n1, k, y);
                v = v + w10 * slab_at(blk0 + 2 * (size_t)blk_cap, n2, k, y);
                v = v + w11 * slab_at(blk0 + 3 * (size_t)blk_cap, n3, k, y);
            }
            s_tr[(size_t)k * wcap + i] = v;
        }
        __syncthreads();
        if (tid < db.ng) s_meta[tid].lastval = s_tr[(size_t)tid * wcap + (s_meta[tid].s1 - wlo)];
        __syncthreads();
        // ---- the group's centroids, one after the other --------------------------------------------------------------------
        const int gi = cand.group_begin + ip;
        const int tb = g.tap_begin[gi], tn = g.tap_count[gi];
        const float tbase = g.tbase[gi], gw = g.gw[gi];
        float mh[6];
#pragma unroll
        for (int q = 0; q < 6; q++) mh[q] = g.mhat[(size_t)q * ngroups_total + gi];
        float ca = s_rec.f[0], sa = s_rec.f[1], s2a = s_rec.f[2], c2a = s_rec.f[3];   // written by k_geometry in this mode ...
        if (trig) { const float4 t4 = __ldg(trig + (size_t)pair * rec_stride + ip); ca = t4.x; sa = t4.y; s2a = t4.z; c2a = t4.w; }   // ... or by the host library
        const float cl = s_rec.cl, sl = s_rec.sl;
        const bool rot = flags & GEO_ROT;
        for (int it = 0; it < tn; it++) {
            const float time = tbase + taps.toff[tb + it];
            const float wt = taps.wt[tb + it];
            float m[6], f[6];
#pragma unroll
            for (int q = 0; q < 6; q++) m[q] = (mh[q] * wt) * gw;
            f[0] = m[0] * (ca * ca) + m[1] * (sa * sa) + m[3] * s2a;         // make_weights seismogram.f90:316-336
            f[1] = m[4] * ca + m[5] * sa;
            f[2] = m[2];
            f[3] = 0.5f * (m[1] - m[0]) * s2a + m[3] * c2a;
            f[4] = m[5] * ca - m[4] * sa;
            f[5] = m[0] * (sa * sa) + m[1] * (ca * ca) - m[3] * s2a;
            const float rshift = time / dt;
            const int its = (int)floorf(rshift);
            const float wr0 = rshift - (float)its, wl0 = 1.f - wr0;
            // trace_multiply_add of trace k with `factor` into the accumulator `acc` of output sample x (sparse_trace.f90:639-705)
            auto madd = [&](float& acc, int k, float factor, int x) {
                const SxTrace mt = s_meta[k];
                const float wr = wr0 * factor, wl = wl0 * factor;
                const float* tr = s_tr + (size_t)k * wcap - wlo;
                const int y0 = x - its;
                if (y0 >= mt.s0 && y0 <= mt.s1) acc = acc + wl * tr[y0];
                if (y0 - 1 >= mt.s0 && y0 - 1 <= mt.s1 - 1) acc = acc + wr * tr[y0 - 1];
                if (y0 > mt.s1 && mt.lastval != 0.f) acc = acc + factor * mt.lastval;
            };
#pragma unroll
            for (int j = 0; j < SX_NS; j++) {
                const int i = tid + j * SX_THREADS;
                if (i >= nsamp) break;
                const int x = base + i;
                if (need_h) {
                    if (rot) {
                        float t1 = 0.f, t2 = 0.f;
                        madd(t1, 0, f[0], x); madd(t1, 1, f[1], x); madd(t1, 2, f[2], x);
                        if (ng10) madd(t1, 8, f[5], x);
                        madd(t2, 3, f[3], x); madd(t2, 4, f[4], x);
                        const float a0 = ar0[j], a1 = ar1[j];
                        ar0[j] = a0 + cl * t1 - sl * t2;
                        ar1[j] = a1 + cl * t2 + sl * t1;
                    } else {
                        madd(ar0[j], 0, f[0], x); madd(ar0[j], 1, f[1], x); madd(ar0[j], 2, f[2], x);
                        if (ng10) madd(ar0[j], 8, f[5], x);
                        madd(ar1[j], 3, f[3], x); madd(ar1[j], 4, f[4], x);
                    }
                }
                if (need_v) {
                    madd(dz[j], 5, f[0] * sd, x); madd(dz[j], 6, f[1] * sd, x); madd(dz[j], 7, f[2] * sd, x);
                    if (ng10) madd(dz[j], 9, f[5] * sd, x);
                }
            }
        }
    }
    // ---- components: signs and the (away, right) -> (north, east) rotation, seismogram.f90:256-289 -----------------------------
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
#pragma unroll
        for (int j = 0; j < SX_NS; j++) {
            const int i = tid + j * SX_THREADS;
            if (i >= nsamp || i >= (int)seis_stride) break;
            float v;
            if (aid == 3) v = dz[j];
            else if (aid == 1) v = ar0[j] * sg;
            else if (aid == 2) v = ar1[j] * sg;
            else if (aid == 4) v = (R.cl0 * ar0[j] - R.sl0 * ar1[j]) * sg;
            else v = (R.cl0 * ar1[j] + R.sl0 * ar0[j]) * sg;
            row[i] = v;
        }
        if (tid == 0) { SeisHdr e; e.lo = lo; e.hi = hi; e.base = base; e.pad = 0; myshdr[ic] = e; }
    }
}

int synth_exact_max_samples() { return SX_THREADS * SX_NS; }
size_t synth_exact_smem_bytes(int wcap, int blk_cap) { return ((size_t)KIWI_NG_MAX * wcap + (size_t)2 * 4 * blk_cap) * sizeof(float); }
cudaError_t launch_synth_exact(GfdbDev db, const ReceiverDev* rcv, int nrcv, const CandDev* cands, int ncand, GroupSoA g, TapSoA taps, int ngroups_total,
                               const GeoRec* recs, size_t rec_stride, const PairHdr* hdrs, int nq_alloc, int margin_q, int interpolate, int xunder,
                               int zunder, int wcap, int blk_cap, float* seis, size_t seis_stride, SeisHdr* shdrs, int* overflow, cudaStream_t st,
                               const float4* trig) {
    const size_t smem = synth_exact_smem_bytes(wcap, blk_cap);
    cudaError_t e = cudaFuncSetAttribute(k_synth_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (ncand * nrcv > 0)
        k_synth_exact<<<ncand * nrcv, SX_THREADS, smem, st>>>(db, rcv, nrcv, cands, g, taps, ngroups_total, recs, rec_stride, hdrs, nq_alloc, margin_q,
                                                             interpolate, xunder, zunder, wcap, blk_cap, seis, seis_stride, shdrs, overflow, trig);
    return cudaGetLastError();
}
