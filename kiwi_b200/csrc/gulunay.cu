// Gulunay's generalised f-k interpolation of the Green's function database on the GPU:
//   interpolation.f90:29-159 gulunay2d, :161-311 gulunay3d           -> gulunay_batched (kernels below)
//   gfdb.f90:1234-1310 interpolate3d, :1109-1232 gfdb_interpolate_block, :205-246 gfdb_init with nipx / nipz
//                                                                      -> kiwi_gfdb_interpolate (host driver at the end)
// One call interpolates `batch` independent fields (the ng Green's function components of a block, or the slices of the pseudo-3-D
// branch) at once.  A field is (t, s1, s2) in the reference's column-major order (time fastest); the result is (t, s1*l1, s2*l2).
//
// FFTW (absent here) is replaced by radix-2 transforms in shared memory with FFTW's conventions (unnormalised; the time dimension is
// the halved one; the multi-dimensional c2r transforms the trace dimensions first and ignores the imaginary parts of time bins 0 and
// t/2).  The spectra of the zero-interleaved (B), zero-padded (C) and decimated (D) arrays are all built from ONE pass over the
// tapered traces: a zero trace transforms to zeros, and a D trace equals the C trace at the same place, so per input trace two time
// transforms (length t and l*t) feed three spectra, of which only the first t/2+1 frequency rows are ever used.
//
// This file is compiled with -fmad=false and the arithmetic is written out operation by operation (butterflies, Smith's complex division
// as gfortran emits it, |z| through a double square root, taper weights tabulated on the host with glibc's cosf), so that the result is
// reproducible bit for bit by a scalar CPU statement of the same formulas.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <climits>
#include <cstring>
#include <vector>
#include "kiwi_internal.hpp"

namespace {

#define GUL_TW_N 16384      // twiddle table exp(-2 pi i k / GUL_TW_N), k < GUL_TW_N/2: transforms up to this length

__device__ __forceinline__ float2 cmul_rn(float2 a, float2 w) {
    return make_float2(__fsub_rn(__fmul_rn(a.x, w.x), __fmul_rn(a.y, w.y)), __fadd_rn(__fmul_rn(a.x, w.y), __fmul_rn(a.y, w.x)));
}
__device__ __forceinline__ float cabs_d(float2 z) { return (float)sqrt((double)z.x * (double)z.x + (double)z.y * (double)z.y); }
// complex division by Smith's method (what gfortran emits for a / b, -fcx-fortran-rules)
__device__ __forceinline__ float2 cdiv_smith(float2 a, float2 b) {
    if (fabsf(b.x) >= fabsf(b.y)) {
        const float r = __fdiv_rn(b.y, b.x), den = __fadd_rn(b.x, __fmul_rn(b.y, r));
        return make_float2(__fdiv_rn(__fadd_rn(a.x, __fmul_rn(a.y, r)), den), __fdiv_rn(__fsub_rn(a.y, __fmul_rn(a.x, r)), den));
    }
    const float r = __fdiv_rn(b.x, b.y), den = __fadd_rn(__fmul_rn(b.x, r), b.y);
    return make_float2(__fdiv_rn(__fadd_rn(__fmul_rn(a.x, r), a.y), den), __fdiv_rn(__fsub_rn(__fmul_rn(a.y, r), a.x), den));
}

// in-place radix-2 decimation-in-time transform of z[0..n) (already in bit-reversed order) by `nthr` cooperating threads (rank tid);
// sync() separates the stages.  sign = -1 forward, +1 inverse.
template <class Sync>
__device__ __forceinline__ void fft_stages(float2* z, int n, int sign, const float2* __restrict__ tw, int tid, int nthr, Sync sync) {
    for (int len = 2; len <= n; len <<= 1) {
        const int h = len >> 1, tstep = GUL_TW_N / len;
        for (int q = tid; q < (n >> 1); q += nthr) {
            const int k = q & (h - 1), i = ((q - k) << 1) + k;
            float2 w = tw[k * tstep];
            if (sign > 0) w.y = -w.y;
            const float2 u = z[i], v = cmul_rn(z[i + h], w);
            z[i] = make_float2(__fadd_rn(u.x, v.x), __fadd_rn(u.y, v.y));
            z[i + h] = make_float2(__fsub_rn(u.x, v.x), __fsub_rn(u.y, v.y));
        }
        sync();
    }
}
__device__ __forceinline__ int brev_n(int i, int log2n) { return (int)(__brev((unsigned)i) >> (32 - log2n)); }

// taper in place (interpolation.f90:66-83, :207-231): last dimension, then the middle one, then time; weights are 1 outside the margins
__global__ void k_gul_taper(float* __restrict__ A, long long total, int t, int s1, int s2, const float* __restrict__ wt, const float* __restrict__ w1,
                            const float* __restrict__ w2) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int it = (int)(i % t), i1 = (int)((i / t) % s1), i2 = (int)((i / ((long long)t * s1)) % s2);
    float v = A[i];
    v = __fmul_rn(v, w2[i2]);
    v = __fmul_rn(v, w1[i1]);
    v = __fmul_rn(v, wt[it]);
    A[i] = v;
}

// one CTA per input trace (b, i1, i2): time transforms of length t (-> FB at the interleaved place) and l*t (zero padded -> FC, and FD if
// the trace is on the stride-l lattice); rows 0..fny-1 only.  F* are [batch][kk2][kk1][fny] complex, zeroed beforehand.
__global__ void __launch_bounds__(256) k_gul_time_forward(const float* __restrict__ A, int t, int s1, int s2, int l1, int l2, int l, float2* __restrict__ FB,
                                                          float2* __restrict__ FC, float2* __restrict__ FD, const float2* __restrict__ tw) {
    extern __shared__ float2 zsm[];
    const int kk1 = s1 * l1, fny = t / 2 + 1, ff = l * t;
    const int line = blockIdx.x, i1 = line % s1, i2 = (line / s1) % s2, b = line / (s1 * s2);
    const float* a = A + (size_t)line * t;
    const size_t plane = (size_t)kk1 * (s2 * l2);
    int lt = 0; while ((1 << lt) < t) lt++;
    int lf = 0; while ((1 << lf) < ff) lf++;
    auto sync = [] { __syncthreads(); };
    for (int i = threadIdx.x; i < t; i += blockDim.x) zsm[brev_n(i, lt)] = make_float2(a[i], 0.f);
    __syncthreads();
    fft_stages(zsm, t, -1, tw, threadIdx.x, blockDim.x, sync);
    float2* ob = FB + ((size_t)b * plane + (size_t)(i2 * l2) * kk1 + i1 * l1) * fny;
    for (int f = threadIdx.x; f < fny; f += blockDim.x) ob[f] = zsm[f];
    __syncthreads();
    for (int i = threadIdx.x; i < ff; i += blockDim.x) zsm[brev_n(i, lf)] = make_float2(i < t ? a[i] : 0.f, 0.f);
    __syncthreads();
    fft_stages(zsm, ff, -1, tw, threadIdx.x, blockDim.x, sync);
    const size_t oc = ((size_t)b * plane + (size_t)i2 * kk1 + i1) * fny;
    const bool lattice = (i1 % l1 == 0) && (i2 % l2 == 0);
    for (int f = threadIdx.x; f < fny; f += blockDim.x) { FC[oc + f] = zsm[f]; if (lattice) FD[oc + f] = zsm[f]; }
}

// transform along a trace dimension: element k of line (o, r) sits at a[o * ostride + r + k * ninner], r < ninner contiguous.
// A CTA takes 8 neighbouring r (coalesced 64-byte pieces), one warp per line.
#define GUL_TL 8
__global__ void __launch_bounds__(256) k_gul_axis(float2* __restrict__ a, int n, long long ninner, long long nouter, long long ostride, int sign,
                                                  const float2* __restrict__ tw) {
    extern __shared__ float2 zsm[];
    const long long tiles = (ninner + GUL_TL - 1) / GUL_TL;
    const long long o = blockIdx.x / tiles, r0 = (blockIdx.x % tiles) * GUL_TL;
    if (o >= nouter) return;
    int ln = 0; while ((1 << ln) < n) ln++;
    float2* base = a + o * ostride + r0;
    const int nr = (int)min((long long)GUL_TL, ninner - r0);
    for (int e = threadIdx.x; e < n * GUL_TL; e += blockDim.x) {
        const int k = e / GUL_TL, r = e % GUL_TL;
        if (r < nr) zsm[r * n + brev_n(k, ln)] = base[(long long)k * ninner + r];
    }
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (w < nr) fft_stages(zsm + w * n, n, sign, tw, lane, 32, [] { __syncwarp(); });
    __syncthreads();
    for (int e = threadIdx.x; e < n * GUL_TL; e += blockDim.x) {
        const int k = e / GUL_TL, r = e % GUL_TL;
        if (r < nr) base[(long long)k * ninner + r] = zsm[r * n + k];
    }
}

// m = 0.01 * maxval(abs(fD(fny,:,:))) per field (interpolation.f90:120, :269); non-negative floats order like their bit patterns
__global__ void k_gul_max(const float2* __restrict__ FD, long long plane, int fny, unsigned* __restrict__ mx) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (c >= plane) return;
    atomicMax(&mx[b], __float_as_uint(cabs_d(FD[((size_t)b * plane + c) * fny + (fny - 1)])));
}

// white noise, operator, clipping, product (interpolation.f90:117-149, :266-301); FI overwrites FB
__global__ void k_gul_operator(float2* __restrict__ FB, const float2* __restrict__ FC, const float2* __restrict__ FD, long long per_field, long long total,
                               const unsigned* __restrict__ mx, float ls, float lowcut, float norm) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float m = __fmul_rn(0.01f, __uint_as_float(mx[i / per_field]));
    float2 d = FD[i];
    if (cabs_d(d) < __fdiv_rn(m, 1000.f)) d = make_float2(m, d.y);
    { const float ad = cabs_d(d); if (ad < m) { const float s = __fdiv_rn(m, ad); d = make_float2(__fmul_rn(s, d.x), __fmul_rn(s, d.y)); } }
    float2 op = cdiv_smith(FC[i], d);
    { const float ao = cabs_d(op); if (ao > ls) { const float s = __fdiv_rn(ls, ao); op = make_float2(__fmul_rn(s, op.x), __fmul_rn(s, op.y)); } }
    if (cabs_d(op) < lowcut) op = make_float2(0.f, 0.f);
    const float2 p = cmul_rn(FB[i], op);
    FB[i] = make_float2(__fdiv_rn(p.x, norm), __fdiv_rn(p.y, norm));
}

// last step of the c2r transform: Hermitian completion of the time dimension, inverse transform, real part
__global__ void __launch_bounds__(256) k_gul_time_inverse(const float2* __restrict__ FI, int t, float* __restrict__ out, const float2* __restrict__ tw) {
    extern __shared__ float2 zsm[];
    const int fny = t / 2 + 1;
    const float2* in = FI + (size_t)blockIdx.x * fny;
    int lt = 0; while ((1 << lt) < t) lt++;
    for (int k = threadIdx.x; k < t; k += blockDim.x) {
        float2 v;
        if (k <= t / 2) v = in[k]; else { v = in[t - k]; v.y = -v.y; }
        if (k == 0 || k == t / 2) v.y = 0.f;
        zsm[brev_n(k, lt)] = v;
    }
    __syncthreads();
    fft_stages(zsm, t, +1, tw, threadIdx.x, blockDim.x, [] { __syncthreads(); });
    float* o = out + (size_t)blockIdx.x * t;
    for (int i = threadIdx.x; i < t; i += blockDim.x) o[i] = zsm[i].x;
}

#define GCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return kiwi_set_error("CUDA error in the Gulunay interpolation: %s", cudaGetErrorString(e_)); } while (0)

struct DevArr {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    ~DevArr() { if (p) cudaFree(p); }
};
struct GulWork {
    DevArr A, out, FB, FC, FD, tw, wt, w1, w2, mx;
    bool tw_ready = false;
    cudaStream_t st = nullptr;
};

// interpolation.f90:69, :75 ...: weight of the k-th trace from the edge, (1 - cos(2 pi k / (2 margin / l))) / 2
float taper_weight(int k, int margin, int l) {
    const float pi = 3.14159265358979f;   // constants.f90:21
    const float w = 2.f * (float)margin / (float)l;
    return (1.f - cosf(2.f * pi * ((float)k / w))) / 2.f;
}
int make_taper(std::vector<float>& w, int n, int margin, int l, bool active) {
    w.assign((size_t)n, 1.f);
    if (!active) return 0;
    const int m = margin / l;
    if (2 * m > n) return kiwi_set_error("gulunay: taper margins overlap (%d traces, margin %d)", n, m);
    for (int x = 1; x <= m; x++) w[x - 1] = taper_weight(x - 1, margin, l);
    for (int x = n - m + 1; x <= n; x++) w[x - 1] = taper_weight(n - x, margin, l);
    return 0;
}
bool is_pow2(int n) { return n >= 1 && (n & (n - 1)) == 0; }

// gulunay2d / gulunay3d for `batch` fields: h_A [batch][s2][s1][t] (tapered in place, as the reference's intent(inout) A), h_out [batch][s2*l2][s1*l1][t]
int gulunay_batched(GulWork& W, float* h_A, int batch, int t, int s1, int s2, int l1, int l2, int ntmargin, int margin1, int margin2, float* h_out) {
    const int l = std::max(l1, l2), kk1 = s1 * l1, kk2 = s2 * l2, ff = l * t, fny = t / 2 + 1;
    if (!is_pow2(t) || !is_pow2(kk1) || !is_pow2(kk2) || !is_pow2(l) || t < 2) return kiwi_set_error("gulunay: array sizes must be powers of two");
    if (ff > GUL_TW_N || kk1 > 512 || kk2 > 512) return kiwi_set_error("gulunay: block too large (%d samples x %d x %d traces)", ff, kk1, kk2);
    cudaStream_t st = W.st;
    if (!W.tw_ready) {
        std::vector<float2> tw((size_t)GUL_TW_N / 2);
        for (int k = 0; k < GUL_TW_N / 2; k++) {   // rounded from double, the table an fp32 FFT library would hold
            const double a = -2.0 * M_PI * (double)k / (double)GUL_TW_N;
            tw[k] = make_float2((float)cos(a), (float)sin(a));
        }
        GCU(W.tw.ensure(sizeof(float2) * tw.size()));
        GCU(cudaMemcpy(W.tw.p, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice));
        W.tw_ready = true;
    }
    std::vector<float> wt, w1, w2;
    if (make_taper(w2, s2, margin2, l, l2 > 1) || make_taper(w1, s1, margin1, l, l1 > 1) || make_taper(wt, t, ntmargin, l, true)) return 1;
    const size_t nA = (size_t)batch * s2 * s1 * t, nOut = (size_t)batch * kk2 * kk1 * t, nF = (size_t)batch * kk2 * kk1 * fny;
    GCU(W.A.ensure(sizeof(float) * nA)); GCU(W.out.ensure(sizeof(float) * nOut));
    GCU(W.FB.ensure(sizeof(float2) * nF)); GCU(W.FC.ensure(sizeof(float2) * nF)); GCU(W.FD.ensure(sizeof(float2) * nF));
    GCU(W.wt.ensure(sizeof(float) * t)); GCU(W.w1.ensure(sizeof(float) * s1)); GCU(W.w2.ensure(sizeof(float) * s2)); GCU(W.mx.ensure(sizeof(unsigned) * batch));
    GCU(cudaMemcpyAsync(W.A.p, h_A, sizeof(float) * nA, cudaMemcpyHostToDevice, st));
    GCU(cudaMemcpyAsync(W.wt.p, wt.data(), sizeof(float) * t, cudaMemcpyHostToDevice, st));
    GCU(cudaMemcpyAsync(W.w1.p, w1.data(), sizeof(float) * s1, cudaMemcpyHostToDevice, st));
    GCU(cudaMemcpyAsync(W.w2.p, w2.data(), sizeof(float) * s2, cudaMemcpyHostToDevice, st));
    GCU(cudaMemsetAsync(W.FB.p, 0, sizeof(float2) * nF, st)); GCU(cudaMemsetAsync(W.FC.p, 0, sizeof(float2) * nF, st)); GCU(cudaMemsetAsync(W.FD.p, 0, sizeof(float2) * nF, st));
    GCU(cudaMemsetAsync(W.mx.p, 0, sizeof(unsigned) * batch, st));
    const float2* tw = (const float2*)W.tw.p;
    float2 *FB = (float2*)W.FB.p, *FC = (float2*)W.FC.p, *FD = (float2*)W.FD.p;
    k_gul_taper<<<(unsigned)((nA + 255) / 256), 256, 0, st>>>((float*)W.A.p, (long long)nA, t, s1, s2, (const float*)W.wt.p, (const float*)W.w1.p, (const float*)W.w2.p);
    GCU(cudaMemcpyAsync(h_A, W.A.p, sizeof(float) * nA, cudaMemcpyDeviceToHost, st));
    const size_t smem_t = sizeof(float2) * (size_t)ff;
    if (smem_t > 48 * 1024) GCU(cudaFuncSetAttribute(k_gul_time_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
    if (sizeof(float2) * (size_t)t > 48 * 1024) GCU(cudaFuncSetAttribute(k_gul_time_inverse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float2) * t)));
    k_gul_time_forward<<<batch * s1 * s2, 256, smem_t, st>>>((const float*)W.A.p, t, s1, s2, l1, l2, l, FB, FC, FD, tw);
    auto axis = [&](float2* a, int which, int sign) {   // which: 1 = middle dimension (kk1), 2 = last (kk2)
        const int n = which == 1 ? kk1 : kk2;
        if (n <= 1) return;
        const long long ninner = which == 1 ? fny : (long long)kk1 * fny;
        const long long nouter = which == 1 ? (long long)batch * kk2 : batch;
        const long long ostride = which == 1 ? (long long)kk1 * fny : (long long)kk2 * kk1 * fny;
        const long long tiles = (ninner + GUL_TL - 1) / GUL_TL;
        k_gul_axis<<<(unsigned)(nouter * tiles), 256, sizeof(float2) * GUL_TL * n, st>>>(a, n, ninner, nouter, ostride, sign, tw);
    };
    for (float2* a : {FB, FC, FD}) { axis(a, 1, -1); axis(a, 2, -1); }
    const long long plane = (long long)kk1 * kk2;
    k_gul_max<<<dim3((unsigned)((plane + 255) / 256), batch), 256, 0, st>>>(FD, plane, fny, (unsigned*)W.mx.p);
    const float ls = (float)(l1 * l2), lowcut = (l1 > 1 && l2 > 1) ? 0.5f * (float)(l * l) : (float)l * 0.5f;
    k_gul_operator<<<(unsigned)((nF + 255) / 256), 256, 0, st>>>(FB, FC, FD, plane * fny, (long long)nF, (const unsigned*)W.mx.p, ls, lowcut, (float)(t * kk1 * kk2));
    axis(FB, 2, +1); axis(FB, 1, +1);
    k_gul_time_inverse<<<(unsigned)(batch * plane), 256, sizeof(float2) * t, st>>>(FB, t, (float*)W.out.p, tw);
    GCU(cudaMemcpyAsync(h_out, W.out.p, sizeof(float) * nOut, cudaMemcpyDeviceToHost, st));
    GCU(cudaStreamSynchronize(st));
    GCU(cudaGetLastError());
    return 0;
}

// interpolate3d (gfdb.f90:1234-1310) for `batch` fields: fin [batch][nx_in][nz_in][nt] -> fout [batch][nx_out][nz_out][nt]
int interpolate3d_batched(GulWork& W, std::vector<float>& fin, int batch, int nt, int nz_in, int nx_in, std::vector<float>& fout, int nz_out, int nx_out,
                          int ntmargin, int nxmargin, int nzmargin) {
    const int nipx = nx_out / nx_in, nipz = nz_out / nz_in;
    fout.assign((size_t)batch * nx_out * nz_out * nt, 0.f);
    if (nipz == 1) return gulunay_batched(W, fin.data(), batch, nt, nx_in, 1, nipx, 1, ntmargin, nxmargin, 0, fout.data());
    if (nipx == 1) return gulunay_batched(W, fin.data(), batch, nt, nz_in, 1, nipz, 1, ntmargin, nzmargin, 0, fout.data());
    if (nipx == 4 && nipz == 4) {   // two 3-D passes with l = 2 (:1271-1277)
        std::vector<float> mid((size_t)batch * (nx_out / 2) * (nz_out / 2) * nt);
        if (gulunay_batched(W, fin.data(), batch, nt, nz_in, nx_in, 2, 2, ntmargin, nzmargin / 2, nxmargin / 2, mid.data())) return 1;
        return gulunay_batched(W, mid.data(), batch, nt, nz_out / 2, nx_out / 2, 2, 2, ntmargin, nzmargin, nxmargin, fout.data());
    }
    if (nipx == nipz) return gulunay_batched(W, fin.data(), batch, nt, nz_in, nx_in, nipz, nipx, ntmargin, nzmargin, nxmargin, fout.data());
    // pseudo 3-D (:1285-1307): the horizontal pass of all depth rows is one batch, the vertical pass of all output columns another.
    // Kept as the reference has it: the vertical pass takes the ORIGINAL column ix_in where mod(ix_in-1, nipx) == 0 (ix_in =
    // (ix_out-1)/nipx+1), the horizontally interpolated one elsewhere, and tapers depth with the horizontal margin.
    std::vector<float> in((size_t)batch * nz_in * nx_in * nt), out((size_t)batch * nz_in * nx_out * nt);
    for (int b = 0; b < batch; b++)
        for (int iz = 0; iz < nz_in; iz++)
            for (int ix = 0; ix < nx_in; ix++)
                memcpy(&in[(((size_t)b * nz_in + iz) * nx_in + ix) * nt], &fin[(((size_t)b * nx_in + ix) * nz_in + iz) * nt], sizeof(float) * nt);
    if (gulunay_batched(W, in.data(), batch * nz_in, nt, nx_in, 1, nipx, 1, ntmargin, nxmargin, 0, out.data())) return 1;
    for (int b = 0; b < batch; b++)
        for (int iz = 0; iz < nz_in; iz++)
            for (int ix = 0; ix < nx_out; ix++)
                memcpy(&fout[(((size_t)b * nx_out + ix) * nz_out + iz * nipz) * nt], &out[(((size_t)b * nz_in + iz) * nx_out + ix) * nt], sizeof(float) * nt);
    in.assign((size_t)batch * nx_out * nz_in * nt, 0.f);
    for (int b = 0; b < batch; b++)
        for (int ixo = 1; ixo <= nx_out; ixo++) {
            const int ix_in = (ixo - 1) / nipx + 1;
            for (int iz = 0; iz < nz_in; iz++) {
                const float* src = ((ix_in - 1) % nipx == 0) ? &fin[(((size_t)b * nx_in + ix_in - 1) * nz_in + iz) * nt]
                                                             : &fout[(((size_t)b * nx_out + ixo - 1) * nz_out + iz * nipz) * nt];
                memcpy(&in[(((size_t)b * nx_out + ixo - 1) * nz_in + iz) * nt], src, sizeof(float) * nt);
            }
        }
    return gulunay_batched(W, in.data(), batch * nx_out, nt, nz_in, 1, nipz, 1, ntmargin, nxmargin, 0, fout.data());
}

// gfdb.f90:1313-1330 (next_power_of_two through default-real logarithms, :1332-1339)
void allowed_span_gfdb(const int span[2], int minlength, int out[2]) {
    int length = span[1] - span[0] + 1;
    if (length < minlength) length = minlength;
    const int lengthp = 1 << (int)ceilf(logf((float)length) / logf(2.f));
    out[0] = span[0] - (int)floorf((float)(lengthp - length) / 2.f);
    out[1] = out[0] + lengthp - 1;
}

}  // namespace

extern "C" {

// Stand-alone entry point of the batched operator (tests, other callers): a [batch][s2][s1][t] is tapered in place
int kiwi_gulunay(int device, float* a, int batch, int t, int s1, int s2, int l1, int l2, int ntmargin, int margin1, int margin2, float* out) {
    if (cudaSetDevice(device) != cudaSuccess) return kiwi_set_error("no CUDA device %d", device);
    GulWork W;
    return gulunay_batched(W, a, batch, t, s1, s2, l1, l2, ntmargin, margin1, margin2, out);
}

// set_database dbpath nipx nipz (minimizer.f90:89-135, gfdb_init gfdb.f90:205-246): the database that pretends to hold nipx x nipz as
// many traces, with every interpolation block (gfdb_interpolate_block gfdb.f90:1109-1232: 128 x 32 traces, payload 96 x 24) filled
// eagerly on the GPU instead of on first access.  Real traces keep their samples; interpolated ones are single strips over the union of
// the spans of their real neighbours (not re-packed).
kiwi_gfdb* kiwi_gfdb_interpolate(const kiwi_gfdb* src_, int nipx, int nipz, int device) {
    if (!src_) { kiwi_set_error("kiwi_gfdb_interpolate: null database"); return nullptr; }
    auto ok_ip = [](int n) { return n == 1 || n == 2 || n == 4 || n == 8 || n == 16; };
    if (!ok_ip(nipx) || !ok_ip(nipz)) { kiwi_set_error("interpolation factors must be 1, 2, 4, 8 or 16"); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { kiwi_set_error("no CUDA device %d", device); return nullptr; }
    kiwi_gfdb* srcm = const_cast<kiwi_gfdb*>(src_);
    srcm->flatten();
    const kiwi_gfdb& src = *src_;
    kiwi_gfdb* db = kiwi_gfdb_create(src.nx * nipx, src.nz * nipz, src.ng, src.dt, src.dx / (float)nipx, src.dz / (float)nipz, src.firstx, src.firstz);
    if (!db) return nullptr;
    const int ng = src.ng;
    for (int ix = 1; ix <= src.nx; ix++)
        for (int iz = 1; iz <= src.nz; iz++)
            for (int ig = 1; ig <= ng; ig++) {
                const size_t i = src.idx(ix, iz, ig), o = db->idx((ix - 1) * nipx + 1, (iz - 1) * nipz + 1, ig);
                if (src.len[i] <= 0) continue;
                db->pending[o].assign(&src.data[(size_t)src.offset[i]], &src.data[(size_t)src.offset[i]] + src.len[i]);
                db->span0[o] = src.span0[i]; db->len[o] = src.len[i];
            }
    const int nbx = nipx != 1 ? 128 : 1, ovx = nipx != 1 ? 32 : 0, pax = nipx != 1 ? 96 : 1;   // gfdb.f90:31-37
    const int nbz = nipz != 1 ? 32 : 1, ovz = nipz != 1 ? 8 : 0, paz = nipz != 1 ? 24 : 1;
    const int nxo = nbx / nipx, nzo = nbz / nipz;
    GulWork W;
    std::vector<int> spans((size_t)2 * nbz * nbx);
    std::vector<float> fin, fout;
    auto fail = [&](void) -> kiwi_gfdb* { kiwi_gfdb_destroy(db); return nullptr; };
    for (int bx0 = 1; bx0 <= db->nx; bx0 += pax)
        for (int bz0 = 1; bz0 <= db->nz; bz0 += paz) {
            const int ixfirst = bx0 - ovx / 2, izfirst = bz0 - ovz / 2, ixlast = ixfirst + nbx - 1, izlast = izfirst + nbz - 1;
            auto real_ix = [&](int ix) { return (std::min(std::max(ix, 1), db->nx) - 1) / nipx * nipx + 1; };   // end points repeated
            auto real_iz = [&](int iz) { return (std::min(std::max(iz, 1), db->nz) - 1) / nipz * nipz + 1; };
            auto sp = [&](int k, int bz, int bx) -> int& { return spans[((size_t)(bx - 1) * nbz + (bz - 1)) * 2 + k]; };
            int span[2] = {INT_MAX, -INT_MAX};
            for (int ix = ixfirst; ix <= ixlast; ix += nipx)
                for (int iz = izfirst; iz <= izlast; iz += nipz)
                    for (int ig = 1; ig <= ng; ig++) {
                        const size_t i = db->idx(real_ix(ix), real_iz(iz), ig);
                        if (db->len[i] <= 0) { kiwi_set_error("gfdb_interpolate_block(): missing trace in interpolation block"); return fail(); }
                        const int s0 = db->span0[i], s1 = s0 + db->len[i] - 1;
                        span[0] = std::min(span[0], s0); span[1] = std::max(span[1], s1);
                        sp(0, iz - izfirst + 1, ix - ixfirst + 1) = s0; sp(1, iz - izfirst + 1, ix - ixfirst + 1) = s1;
                    }
            const int raw_len = span[1] - span[0];
            { int a[2]; allowed_span_gfdb(span, std::min(64, (int)((float)raw_len * 1.2f)), a); span[0] = a[0]; span[1] = a[1]; }
            const int nt = span[1] - span[0] + 1;
            if (nt <= 1) continue;
            // the real traces of the block on the common window, continued with their last sample (trace_multiply_add_nogrow,
            // sparse_trace.f90:710-793), all components at once: fin [ig][bx][bz][t]
            fin.assign((size_t)ng * nxo * nzo * nt, 0.f);
            for (int ig = 1; ig <= ng; ig++)
                for (int iz = izfirst; iz <= izlast; iz += nipz)
                    for (int ix = ixfirst; ix <= ixlast; ix += nipx) {
                        const size_t i = db->idx(real_ix(ix), real_iz(iz), ig);
                        const float* d = db->pending[i].data();
                        const int s0 = db->span0[i], n = db->len[i];
                        float* dst = &fin[((((size_t)(ig - 1) * nxo + (ix - ixfirst) / nipx) * nzo) + (iz - izfirst) / nipz) * nt];
                        // (a block edge outside the grid repeats the end trace; += onto the zeroed slot is the reference's multiply-add)
                        for (int x = std::max(span[0], s0); x <= span[1]; x++) {
                            const float v = d[std::min(x - s0, n - 1)];
                            if (x - s0 < n || v != 0.f) dst[x - span[0]] += v;
                        }
                    }
            if (interpolate3d_batched(W, fin, ng, nt, nzo, nxo, fout, nbz, nbx, (int)(0.1f * (float)(span[1] - span[0])), ovx / 2, ovz / 2)) return fail();
            for (int ig = 1; ig <= ng; ig++)
                for (int iz = izfirst + ovz / 2; iz <= izlast - ovz / 2; iz++)
                    for (int ix = ixfirst + ovx / 2; ix <= ixlast - ovx / 2; ix++) {
                        if ((ix - 1) % nipx == 0 && (iz - 1) % nipz == 0) continue;   // only the interpolated traces are inserted
                        if (ix < 1 || db->nx < ix || iz < 1 || db->nz < iz) continue;
                        const int bx = ix - ixfirst + 1, bz = iz - izfirst + 1;
                        const int lrx = ((bx - 1) / nipx) * nipx + 1, lrz = ((bz - 1) / nipz) * nipz + 1, nrx = lrx + nipx, nrz = lrz + nipz;
                        int d0 = sp(0, lrz, lrx), d1 = sp(1, lrz, lrx);   // union of the spans of the neighbouring real traces
                        if (nrx <= nbx) { d0 = std::min(d0, sp(0, lrz, nrx)); d1 = std::max(d1, sp(1, lrz, nrx)); }
                        if (nrz <= nbz) { d0 = std::min(d0, sp(0, nrz, lrx)); d1 = std::max(d1, sp(1, nrz, lrx)); }
                        if (nrx <= nbx && nrz <= nbz) { d0 = std::min(d0, sp(0, nrz, nrx)); d1 = std::max(d1, sp(1, nrz, nrx)); }
                        const size_t o = db->idx(ix, iz, ig);
                        if (db->len[o] > 0) continue;
                        const float* srcp = &fout[(((size_t)(ig - 1) * nbx + (bx - 1)) * nbz + (bz - 1)) * nt + (d0 - span[0])];
                        db->pending[o].assign(srcp, srcp + (d1 - d0 + 1));
                        db->span0[o] = d0; db->len[o] = d1 - d0 + 1;
                    }
        }
    db->flatten();
    return db;
}

}  // extern "C"
