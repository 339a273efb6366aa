// Fast-marching solver of the eikonal sources on the device (eikonal.f90:29-199, heap.f90:48-232).
//
// The reference's solver cannot be reordered: a node is updated from the current values of all four neighbours, tentative
// or final (eikonal.f90:151-155), and the two-sided update is accepted whenever its discriminant is non-negative (:163-166),
// so the table it leaves depends on the exact order in which the binary heap hands out equal and nearly equal keys.  What
// runs in parallel here is therefore one *sequential* solve per candidate: one warp per candidate, the heap of (key, index)
// pairs in shared memory, times / back-pointers / speeds in global memory (the front touches a few rows at a time, which stay
// in L1/L2).  Lane 0 replays heap.f90 operation by operation; the four neighbour stencils of a popped node do not contain
// each other, so lanes 0..3 evaluate them side by side (one memory round trip) before lane 0 applies their heap updates in
// the reference's order (left, right, down, up).  All arithmetic is IEEE fp32 in the reference's operation order
// (__f*_rn intrinsics: no FMA contraction), so the result equals the host solver's (source_eikonal_host.cpp) bit for bit.
//
// A warp walks ~4e3 dependent cycles (2.1 us, ~470 instructions) per node where a host core needs ~100 ns, so one solve is ~20 x
// slower than on the host; the device wins by running thousands of them side by side (0.73-0.85 ns per node over 3848 solves,
// profiles/r02_eikonal_device.txt), i.e. on the batches of a grid search, which is when the engine uses it.
#include "kernels.cuh"
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

// Heap entries kept in shared memory; the rest spills to global memory.  Two builds of the solver: 2040 entries (16 KB per candidate,
// 13 candidates per SM) for batches that fit one such wave, 980 entries (8 KB, 26 per SM) for larger batches -- a warp spends its time
// in dependent fixed-latency instructions (profiles/r02_k_eikonal_fmm_full.md: issue slots 13.5 % at 13 warps per SM), so twice the
// resident solves is close to twice the throughput, at the price of the deepest heap level of the larger fronts living in L2.
#define EIK_HCAP_LARGE 2040
#define EIK_HCAP_SMALL 980

namespace {

// shared-memory accesses by 32-bit shared address (a generic pointer makes the compiler re-derive the shared window per access)
__device__ __forceinline__ EikItem lds_item(unsigned addr) {
    EikItem v;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=f"(v.key), "=r"(v.idx) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void lds_pair(unsigned addr, EikItem& a, EikItem& b) {
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=f"(a.key), "=r"(a.idx), "=f"(b.key), "=r"(b.idx) : "r"(addr) : "memory");
}
__device__ __forceinline__ void sts_item(unsigned addr, const EikItem& v) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "f"(v.key), "r"(v.idx) : "memory");
}
__device__ __forceinline__ void sts_key(unsigned addr, float k) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "f"(k) : "memory"); }

template <int EIK_HCAP>
struct EikHeap {
    unsigned sbase;     // shared address of entry 0 of the shared-memory part (entries 1..EIK_HCAP used, 8 bytes each)
    EikItem* ovf;       // global overflow, entries EIK_HCAP+1 ...
    int* bp;            // back-pointers, 1-based node index
    int n;
    // heap positions of the four neighbours of the node being processed, kept current in registers: their back-pointers were just
    // written to global memory and reading them back would cost a round trip to L2 per neighbour
    int w0, w1, w2, w3, p0, p1, p2, p3;
    __device__ __forceinline__ EikItem get(int i) const { return i <= EIK_HCAP ? lds_item(sbase + 8u * (unsigned)i) : ovf[i - EIK_HCAP]; }
    __device__ __forceinline__ void track(int idx, int i) {
        if (idx == w0) p0 = i;
        if (idx == w1) p1 = i;
        if (idx == w2) p2 = i;
        if (idx == w3) p3 = i;
    }
    __device__ __forceinline__ void put(int i, EikItem v) {
        if (i <= EIK_HCAP) sts_item(sbase + 8u * (unsigned)i, v); else ovf[i - EIK_HCAP] = v;
        bp[v.idx] = i;
        track(v.idx, i);
    }
    // the same store without the position tracking (during a pop none of the tracked nodes is set)
    __device__ __forceinline__ void put_plain(int i, EikItem v) {
        if (i <= EIK_HCAP) sts_item(sbase + 8u * (unsigned)i, v); else ovf[i - EIK_HCAP] = v;
        bp[v.idx] = i;
    }
    __device__ __forceinline__ void setkey(int i, float k) { if (i <= EIK_HCAP) sts_key(sbase + 8u * (unsigned)i, k); else ovf[i - EIK_HCAP].key = k; }
    // heap.f90:210-232.  The part of the path that lies in the global overflow first, then the shared-memory part to the root.
    __device__ void upheap(int v) {
        const EikItem x = get(v);
        const int v0 = v;
        while (v > EIK_HCAP) {                   // (not entered while the whole heap is in shared memory, the usual case)
            const int u = (v - 2) / 2 + 1;
            const EikItem p = get(u);
            if (p.key <= x.key) { if (v != v0) put(v, x); return; }
            put(v, p);
            v = u;
        }
        while (v > 1) {
            const int u = v >> 1;                // (v - 2) / 2 + 1
            const EikItem p = lds_item(sbase + 8u * (unsigned)u);
            if (p.key <= x.key) break;
            sts_item(sbase + 8u * (unsigned)v, p); bp[p.idx] = v; track(p.idx, v);
            v = u;
        }
        if (v != v0) { sts_item(sbase + 8u * (unsigned)v, x); bp[x.idx] = v; track(x.idx, v); }
    }
    // heap.f90:176-208: item x sinks from entry v (x_in_place: it is stored there already).  The children of entry v are entries 2v
    // and 2v+1: one 16-byte read while both are in shared memory.
    template <bool TRACK>
    __device__ void sink(int v, const EikItem x, const bool x_in_place) {
        const int v0 = v;
        int w = 2 * v;                           // 2 * (v - 1) + 2
        const int nboth = min(n - 1, EIK_HCAP - 1);   // w <= nboth: both children exist and are in shared memory
        bool open = true;
        while (w <= nboth) {
            EikItem c, c2;
            lds_pair(sbase + 8u * (unsigned)w, c, c2);   // (w is even: 16-byte aligned)
            if (c2.key < c.key) { c = c2; w = w + 1; }
            if (x.key <= c.key) { open = false; break; }
            sts_item(sbase + 8u * (unsigned)v, c); bp[c.idx] = v;
            if (TRACK) track(c.idx, v);
            v = w;
            w = 2 * v;
        }
        if (open) {
            while (w <= n) {                     // an only child, or the deepest levels of a heap that reaches into the global overflow
                EikItem c = get(w);
                if (w + 1 <= n) { const EikItem c2 = get(w + 1); if (c2.key < c.key) { c = c2; w = w + 1; } }
                if (x.key <= c.key) break;
                if (TRACK) put(v, c); else put_plain(v, c);
                v = w;
                w = 2 * v;
            }
        }
        if (!x_in_place || v != v0) { if (TRACK) put(v, x); else put_plain(v, x); }
    }
    __device__ void downheap(int v) { sink<true>(v, get(v), true); }
};

}  // namespace

template <int EIK_HCAP>
__global__ void __launch_bounds__(32) k_eikonal_fmm(const EikJob* __restrict__ jobs, int njobs, int prefetch_entries) {
    __shared__ __align__(16) EikItem s_heap[EIK_HCAP + 2];
    __shared__ float4 s_x[4];
    const int job = blockIdx.x;
    if (job >= njobs) return;
    const EikJob J = jobs[job];
    const int lane = threadIdx.x;
    const int nx = J.nx, ny = J.ny, nn = nx * ny;
    const float infinity = FLT_MAX * 0.1f;
    const float dx = J.dx, dy = J.dy;
    const float rnx = 1.f / (float)nx;   // (only to guess a quotient that is then corrected in integers)
    const float dx2 = __fmul_rn(dx, dx), dy2 = __fmul_rn(dy, dy), dx2dy2 = __fmul_rn(dx2, dy2), dx2pdy2 = __fadd_rn(dx2, dy2);
    float* T = J.T - 1;            // 1-based views
    int* bp = J.bp - 1;
    float* S = const_cast<float*>(J.S) - 1;
    const int FARAWAY = -1, ALIVE = 0;
    for (int i = 1 + lane; i <= nn; i += 32) { T[i] = infinity; bp[i] = FARAWAY; }
    if (J.invalid_speed > 0.f) for (int i = 1 + lane; i <= nn; i += 32) if (S[i] == 0.f) S[i] = J.invalid_speed;
    __syncwarp();
    EikHeap<EIK_HCAP> H;
    H.sbase = (unsigned)__cvta_generic_to_shared(s_heap); H.ovf = J.ovf; H.bp = bp; H.n = 0;
    H.w0 = H.w1 = H.w2 = H.w3 = 0; H.p0 = H.p1 = H.p2 = H.p3 = 0;
    const int ix0 = J.ix0, iy0 = J.iy0;
    const int i0 = (iy0 - 1) * nx + ix0;
    if (lane == 0) {
        T[i0] = 0.f;
        if (!(nx == 1 && ny == 1)) {
            bp[i0] = ALIVE;
            if (1 < ix0) T[i0 - 1] = __fdiv_rn(dx, S[i0 - 1]);
            if (ix0 < nx) T[i0 + 1] = __fdiv_rn(dx, S[i0 + 1]);
            if (1 < iy0) T[i0 - nx] = __fdiv_rn(dy, S[i0 - nx]);
            if (iy0 < ny) T[i0 + nx] = __fdiv_rn(dy, S[i0 + nx]);
            auto push = [&](int i) {   // heap.f90:70-93
                H.n = H.n + 1;
                EikItem it; it.key = T[i]; it.idx = i;
                H.put(H.n, it);
                H.upheap(H.n);
            };
            if (1 < ix0) push(i0 - 1);
            if (ix0 < nx) push(i0 + 1);
            if (1 < iy0) push(i0 - nx);
            if (iy0 < ny) push(i0 + nx);
        }
    }
    if (nx == 1 && ny == 1) return;
    __syncwarp();
    int nalive = 1;
    int hn = __shfl_sync(0xffffffffu, H.n, 0);
    while (nalive <= nn) {
        if (hn == 0) break;
        // ---- popheap heap.f90:95-124 (lane 0) ------------------------------------------------------------------
        int imin = 0;
        H.w0 = H.w1 = H.w2 = H.w3 = 0;
        if (lane == 0) {
            const EikItem top = H.get(1), last = H.get(H.n);
            imin = top.idx;
            H.n = H.n - 1;
            if (H.n >= 1) H.template sink<false>(1, last, false);   // (the last entry moves to the root and sinks)
            bp[imin] = ALIVE;
        }
        __syncwarp();
        // What the next pops will read, asked for now (to L2): with thousands of solves in flight the rows around a front do not stay
        // in L2 between two visits (profiles/r02_k_eikonal_fmm_wave_full.md: hit rate 19 %, the warps wait for DRAM 40 % of the time).
        // The root after this pop is the next node unless one of the four updates below undercuts it; entries 2 and 3 are its likely
        // successors.  Lanes 0..10 / 11..21 / 22..31 take one heap entry each: T rows -2..2, S rows -1..1, back-pointer rows -1..1.
        {
            const int g = lane / 11, q = lane - g * 11;
            if (g < prefetch_entries && 1 + g <= hn - 1) {
                const int m = lds_item(H.sbase + 8u * (unsigned)(1 + g)).idx;
                const int r = q < 5 ? q - 2 : (q < 8 ? q - 6 : q - 9);
                const int k = max(1, min(nn, m + r * nx));
                const void* ptr = q < 5 ? (const void*)(T + k) : (q < 8 ? (const void*)(S + k) : (const void*)(bp + k));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
            }
        }
        imin = __shfl_sync(0xffffffffu, imin, 0);
        nalive = nalive + 1;
        int iy = __float2int_rz(__fmul_rn(__int2float_rn(imin - 1), rnx)), ix = imin - 1 - iy * nx;   // (imin - 1) / nx and the remainder
        while (ix < 0) { iy = iy - 1; ix = ix + nx; }
        while (ix >= nx) { iy = iy + 1; ix = ix - nx; }
        iy = iy + 1; ix = ix + 1;
        // ---- the four neighbour stencils, lanes 0..3: left, right, down, up (eikonal.f90:134-190) -----------------------
        int i = 0, state = ALIVE;      // state: ALIVE (skip), FARAWAY, or > 0 (in the heap)
        float t = 0.f, told = 0.f;
        if (lane < 4) {
            const bool valid = lane == 0 ? 1 < ix : (lane == 1 ? ix < nx : (lane == 2 ? 1 < iy : iy < ny));
            if (valid) {
                i = lane == 0 ? imin - 1 : (lane == 1 ? imin + 1 : (lane == 2 ? imin - nx : imin + nx));
                const int jx = lane == 0 ? ix - 1 : (lane == 1 ? ix + 1 : ix), jy = lane == 2 ? iy - 1 : (lane == 3 ? iy + 1 : iy);
                // (all loads are issued before the state is looked at: one round trip)
                float a = infinity, b = infinity, c = infinity, d = infinity;
                state = bp[i];
                told = T[i];
                const float sp = S[i];
                if (1 < jx) a = T[i - 1];
                if (jx < nx) b = T[i + 1];
                if (1 < jy) c = T[i - nx];
                if (jy < ny) d = T[i + nx];
                if (state != ALIVE) {
                    const float aa = fminf(a, b), cc = fminf(c, d);
                    if (fmaxf(aa, cc) != infinity) {
                        const float q = __fmul_rn(__fsub_rn(aa, cc), sp);
                        const float s = __fmul_rn(dx2dy2, __fsub_rn(dx2pdy2, __fmul_rn(q, q)));
                        if (s >= 0.f)
                            t = fmaxf(t, __fdiv_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(aa, dy2), __fmul_rn(cc, dx2)), sp), __fsqrt_rn(s)), __fmul_rn(sp, dx2pdy2)));
                    }
                    if (cc == infinity) {
                        if (a < infinity) t = fmaxf(t, __fadd_rn(a, __fdiv_rn(dx, sp)));
                        if (b < infinity) t = fmaxf(t, __fadd_rn(b, __fdiv_rn(dx, sp)));
                    }
                    if (aa == infinity) {
                        if (c < infinity) t = fmaxf(t, __fadd_rn(c, __fdiv_rn(dy, sp)));
                        if (d < infinity) t = fmaxf(t, __fadd_rn(d, __fdiv_rn(dy, sp)));
                    }
                    if (t == 0.f) {   // fallback condition
                        t = infinity;
                        if (a < infinity) t = fminf(t, __fadd_rn(a, __fdiv_rn(dx, sp)));
                        if (b < infinity) t = fminf(t, __fadd_rn(b, __fdiv_rn(dx, sp)));
                        if (c < infinity) t = fminf(t, __fadd_rn(c, __fdiv_rn(dy, sp)));
                        if (d < infinity) t = fminf(t, __fadd_rn(d, __fdiv_rn(dy, sp)));
                    }
                }
            }
        }
        // ---- their heap updates in the reference's order (lane 0) ----------------------------------------------------------
        // (the four results travel to lane 0 through shared memory: one store and four loads instead of two dozen shuffles)
        if (lane < 4) s_x[lane] = make_float4(__int_as_float(i), __int_as_float(state), t, told);
        __syncwarp();
        float4 xs[4];
#pragma unroll
        for (int k = 0; k < 4; k++) xs[k] = s_x[k];
        // the neighbours' heap positions as the stencil lanes read them; put() keeps them current from here on
        H.w0 = __float_as_int(xs[0].x); H.w1 = __float_as_int(xs[1].x); H.w2 = __float_as_int(xs[2].x); H.w3 = __float_as_int(xs[3].x);
        H.p0 = __float_as_int(xs[0].y); H.p1 = __float_as_int(xs[1].y); H.p2 = __float_as_int(xs[2].y); H.p3 = __float_as_int(xs[3].y);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int ik = __float_as_int(xs[k].x), sk = __float_as_int(xs[k].y);
            const float tk = xs[k].z, toldk = xs[k].w;
            if (lane == 0 && sk != ALIVE && ik != 0) {
                if (sk == FARAWAY) {   // pushheap with the node's current (infinite) time
                    H.n = H.n + 1;
                    EikItem it; it.key = toldk; it.idx = ik;
                    H.put(H.n, it);
                    H.upheap(H.n);
                }
                if (tk != 0.f && toldk != tk) {   // updateheap heap.f90:126-150
                    T[ik] = tk;
                    const int pos = k == 0 ? H.p0 : (k == 1 ? H.p1 : (k == 2 ? H.p2 : H.p3));
                    H.setkey(pos, tk);
                    if (tk < toldk) H.upheap(pos);
                    if (tk > toldk) H.downheap(k == 0 ? H.p0 : (k == 1 ? H.p1 : (k == 2 ? H.p2 : H.p3)));
                }
            }
        }
        __syncwarp();
        hn = __shfl_sync(0xffffffffu, H.n, 0);
    }
}

// ---- the rest of the discretiser's fine-grid work (source_eikonal.f90:435-601), so that a candidate whose solve runs on the device
// never sends its grid over PCIe: the speed field before the solve, the down-sampling after it.  -fmad=false: every operation is the
// IEEE operation of the host code (source_eikonal_host.cpp), in its order.
namespace {
__device__ __forceinline__ void eik_rc_to_ned(const EikGeom& G, const float rc[3], float out[3]) {   // :612-617
    for (int i = 0; i < 3; i++) {
        float a = 0.f;
        for (int j = 0; j < 3; j++) a = a + G.rot[i * 3 + j] * rc[j];
        out[i] = a + G.shift[i];
    }
}
__device__ __forceinline__ void eik_ned_to_rc(const EikGeom& G, const float pt[3], float out[3]) {   // :605-610
    const float d[3] = {pt[0] - G.shift[0], pt[1] - G.shift[1], pt[2] - G.shift[2]};
    for (int i = 0; i < 3; i++) {
        float a = 0.f;
        for (int j = 0; j < 3; j++) a = a + G.rot[j * 3 + i] * d[j];
        out[i] = a;
    }
}
__device__ __forceinline__ void eik_fine_point(const EikGeom& G, int ix, int iy, float pt[3]) {
    const float rc[3] = {G.first[0] + ((float)ix - 0.5f) * G.delta[0], G.first[1] + ((float)iy - 0.5f) * G.delta[1], 0.f};
    eik_rc_to_ned(G, rc, pt);
}
}  // namespace

// speed field: one thread per fine node; blockIdx.y = candidate
__global__ void __launch_bounds__(256) k_eik_speed(EikGeom* __restrict__ geoms, float* __restrict__ S) {
    __shared__ EikGeom G;
    __shared__ int s_min;
    EikGeom* gp = geoms + blockIdx.y;
    for (int i = threadIdx.x; i < (int)(sizeof(EikGeom) / 4); i += blockDim.x) reinterpret_cast<int*>(&G)[i] = reinterpret_cast<const int*>(gp)[i];
    if (threadIdx.x == 0) s_min = 0x7f7fffff;
    __syncthreads();
    const int nn = G.fnx * G.fny;
    float* Sc = S + G.node_off;
    int mymin = 0x7f7fffff;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nn; c += gridDim.x * blockDim.x) {
        const int iy = c / G.fnx + 1, ix = c - (iy - 1) * G.fnx + 1;
        float pt[3];
        eik_fine_point(G, ix, iy, pt);
        const float d[3] = {pt[0] - G.center[0], pt[1] - G.center[1], pt[2] - G.center[2]};
        float dd = 0.f;
        for (int i = 0; i < 3; i++) dd = dd + d[i] * d[i];
        bool inside = !(sqrtf(dd) > G.radius);
        for (int k = 0; k < G.ncons && inside; k++) {   // point_in_halfspace geometry.f90:57-71
            const float v[3] = {G.cpoint[k][0] - pt[0], G.cpoint[k][1] - pt[1], G.cpoint[k][2] - pt[2]};
            float a = 0.f;
            for (int i = 0; i < 3; i++) a = a + G.cnormal[k][i] * v[i];
            inside = a >= 0.f;
        }
        float sp = 0.f;
        if (inside) {   // crust2x2_get_at_depth crust2x2.f90:160-193
            float vs = G.vs[5];
            for (int l = 0; l < 5; l++) if (G.thr[l] >= pt[2]) { vs = G.vs[l]; break; }
            sp = vs * G.relv;
            mymin = min(mymin, __float_as_int(sp));       // (positive floats order like their bit patterns)
        }
        Sc[c] = sp;
    }
    atomicMin(&s_min, mymin);
    __syncthreads();
    if (threadIdx.x == 0) atomicMin(&gp->minspeed_bits, s_min);
}

// down-sampling: one thread per sub-fault (coarse cell) walks the fine points that can fall into it, in fine-grid order (iy outer, ix
// inner: the order in which the reference adds them up), twice: means, then durations.  blockIdx.y = candidate.
// coarse output per cell: ntimes, mean time (-1 = no point), duration, mean north / east / depth
__global__ void __launch_bounds__(128) k_eik_down(const EikGeom* __restrict__ geoms, const float* __restrict__ S, const float* __restrict__ T,
                                                  float* __restrict__ coarse) {
    __shared__ EikGeom G;
    for (int i = threadIdx.x; i < (int)(sizeof(EikGeom) / 4); i += blockDim.x) reinterpret_cast<int*>(&G)[i] = reinterpret_cast<const int*>(geoms + blockIdx.y)[i];
    __syncthreads();
    const int nc = G.nxc * G.nyc;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nc) return;
    const int iyc = k / G.nxc + 1, ixc = k - (iyc - 1) * G.nxc + 1;
    const float* Sc = S + G.node_off;
    const float* Tc = T + G.node_off;
    // fine points that can map into this cell: two fine cells of slack on either side (the mapping goes through a rotation and back)
    const float rx = G.cdelta[0] / G.delta[0], ry = G.cdelta[1] / G.delta[1];
    const int ix_lo = max(1, (int)floorf((float)(ixc - 1) * rx) - 2), ix_hi = min(G.fnx, (int)ceilf((float)ixc * rx) + 3);
    const int iy_lo = max(1, (int)floorf((float)(iyc - 1) * ry) - 2), iy_hi = min(G.fny, (int)ceilf((float)iyc * ry) + 3);
    float ntimes = 0.f, ctimes = -1.f, cspeed = 0.f, cp[3] = {0.f, 0.f, 0.f};
    auto mine = [&](int ix, int iy, int c, float pt[3]) -> bool {
        if (Sc[c] == G.invalid_speed || Sc[c] == 0.f) return false;   // outside the rupture area: time set to -1 by the reference (:513-515);
                                                                       // (0: a candidate solved on the host, whose device copy was never filled)
        eik_fine_point(G, ix, iy, pt);
        float rc[3];
        eik_ned_to_rc(G, pt, rc);
        const int jx = (int)floorf((rc[0] - G.first[0]) / G.cdelta[0]) + 1, jy = (int)floorf((rc[1] - G.first[1]) / G.cdelta[1]) + 1;
        return jx == ixc && jy == iyc;
    };
    for (int iy = iy_lo; iy <= iy_hi; iy++)
        for (int ix = ix_lo; ix <= ix_hi; ix++) {
            const int c = (iy - 1) * G.fnx + (ix - 1);
            float pt[3];
            if (!mine(ix, iy, c, pt)) continue;
            ntimes = ntimes + 1.f;
            if (ctimes == -1.f) ctimes = 0.f;
            ctimes = ctimes + Tc[c];
            cspeed = cspeed + 1.f / Sc[c];
            for (int q = 0; q < 3; q++) cp[q] = cp[q] + pt[q];
        }
    float cdur = 0.f;
    if (ntimes > 0.f) {
        ctimes = 1.f / ntimes * ctimes;
        for (int q = 0; q < 3; q++) cp[q] = 1.f / ntimes * cp[q];
        for (int iy = iy_lo; iy <= iy_hi; iy++)
            for (int ix = ix_lo; ix <= ix_hi; ix++) {
                const int c = (iy - 1) * G.fnx + (ix - 1);
                float pt[3];
                if (!mine(ix, iy, c, pt)) continue;
                cdur = cdur + fabsf(Tc[c] - ctimes);
            }
        cdur = 4.f / ntimes * cdur;
    }
    (void)cspeed;
    float* o = coarse + G.coarse_off + (size_t)6 * k;
    o[0] = ntimes; o[1] = ctimes; o[2] = cdur; o[3] = cp[0]; o[4] = cp[1]; o[5] = cp[2];
}

cudaError_t launch_eik_speed(EikGeom* d_geoms, int ncand, int max_nodes, float* S, cudaStream_t st) {
    if (ncand <= 0) return cudaSuccess;
    const int bx = std::max(1, std::min((max_nodes + 255) / 256, 64));
    k_eik_speed<<<dim3(bx, ncand), 256, 0, st>>>(d_geoms, S);
    return cudaGetLastError();
}
cudaError_t launch_eik_down(const EikGeom* d_geoms, int ncand, int max_cells, const float* S, const float* T, float* coarse, cudaStream_t st) {
    if (ncand <= 0) return cudaSuccess;
    k_eik_down<<<dim3((max_cells + 127) / 128, ncand), 128, 0, st>>>(d_geoms, S, T, coarse);
    return cudaGetLastError();
}

int eikonal_heap_smem_entries() { return EIK_HCAP_SMALL; }   // (a job carries overflow room whenever its grid has more nodes than this)
int eikonal_wave_jobs(int small_heap) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms * (small_heap ? 26 : 13);
}
cudaError_t launch_eikonal_fmm(const EikJob* d_jobs, int njobs, cudaStream_t st) {
    if (njobs <= 0) return cudaSuccess;
    const char* env = getenv("KIWI_EIKONAL_HEAP");   // tests and measurements: 980 or 2040
    const int forced = env ? atoi(env) : 0;
    const bool small = forced ? forced < EIK_HCAP_LARGE : njobs > eikonal_wave_jobs(0);
    const char* penv = getenv("KIWI_EIKONAL_PREFETCH");   // heap entries whose stencils are prefetched after a pop (0..3)
    const int prefetch = penv ? atoi(penv) : (njobs > 2 * 148 ? 3 : 0);
    if (small) {
        cudaFuncSetAttribute(k_eikonal_fmm<EIK_HCAP_SMALL>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        k_eikonal_fmm<EIK_HCAP_SMALL><<<njobs, 32, 0, st>>>(d_jobs, njobs, prefetch);
    } else {
        cudaFuncSetAttribute(k_eikonal_fmm<EIK_HCAP_LARGE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        k_eikonal_fmm<EIK_HCAP_LARGE><<<njobs, 32, 0, st>>>(d_jobs, njobs, prefetch);
    }
    return cudaGetLastError();
}
void eikonal_start_node(const float origin[2], const float delta[2], const float initialpoint[2], int nx, int ny, int* ix0, int* iy0) {   // eikonal.f90:60-67
    int ix = (int)((initialpoint[0] - origin[0]) / delta[0]) + 1, iy = (int)((initialpoint[1] - origin[1]) / delta[1]) + 1;
    if (ix < 1) ix = 1;
    if (nx < ix) ix = nx;
    if (iy < 1) iy = 1;
    if (ny < iy) iy = ny;
    *ix0 = ix; *iy0 = iy;
}

// Batch entry point for tests and measurements: njobs grids (host arrays), solved concurrently on the device.
// speed[j] / times[j]: host (nx[j] x ny[j]), ix fastest.  Returns the kernel time in *kernel_ms.
extern "C" int kiwi_eikonal_fmm_device(int njobs, const int* nx, const int* ny, const float* const* speed, const float* origin2,
                                       const float* delta2, const float* initialpoint2, float* const* times, float* kernel_ms) {
    if (njobs <= 0) return 0;
    std::vector<EikJob> jobs(njobs);
    std::vector<void*> allocs;
    auto fail = [&](const char* what) { for (void* p : allocs) cudaFree(p); fprintf(stderr, "kiwi_eikonal_fmm_device: %s\n", what); return 1; };
    for (int j = 0; j < njobs; j++) {
        const size_t nn = (size_t)nx[j] * ny[j];
        if (nx[j] < 1 || ny[j] < 1) return fail("invalid grid");
        float *dS = nullptr, *dT = nullptr; int* dbp = nullptr; EikItem* dovf = nullptr;
        if (cudaMalloc(&dS, nn * 4) != cudaSuccess || cudaMalloc(&dT, nn * 4) != cudaSuccess || cudaMalloc(&dbp, nn * 4) != cudaSuccess) return fail("out of device memory");
        allocs.push_back(dS); allocs.push_back(dT); allocs.push_back(dbp);
        if (nn > EIK_HCAP_SMALL) { if (cudaMalloc(&dovf, (nn - EIK_HCAP_SMALL + 1) * sizeof(EikItem)) != cudaSuccess) return fail("out of device memory"); allocs.push_back(dovf); }
        cudaMemcpy(dS, speed[j], nn * 4, cudaMemcpyHostToDevice);
        EikJob& J = jobs[j];
        J.nx = nx[j]; J.ny = ny[j]; J.dx = delta2[2 * j]; J.dy = delta2[2 * j + 1];
        eikonal_start_node(origin2 + 2 * j, delta2 + 2 * j, initialpoint2 + 2 * j, J.nx, J.ny, &J.ix0, &J.iy0);
        J.S = dS; J.T = dT; J.bp = dbp; J.ovf = dovf; J.invalid_speed = 0.f;
    }
    EikJob* djobs = nullptr;
    if (cudaMalloc(&djobs, sizeof(EikJob) * njobs) != cudaSuccess) return fail("out of device memory");
    allocs.push_back(djobs);
    cudaMemcpy(djobs, jobs.data(), sizeof(EikJob) * njobs, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    launch_eikonal_fmm(djobs, njobs, 0);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (kernel_ms) *kernel_ms = ms;
    if (err != cudaSuccess) return fail(cudaGetErrorString(err));
    for (int j = 0; j < njobs; j++) cudaMemcpy(times[j], jobs[j].T, (size_t)nx[j] * ny[j] * 4, cudaMemcpyDeviceToHost);
    for (void* p : allocs) cudaFree(p);
    return 0;
}
