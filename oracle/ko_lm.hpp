// TEST INFRASTRUCTURE (oracle): sequential restatement of the Levenberg-Marquardt code the reference links,
// sminpack/{lmdif,fdjac2,lmpar,qrfac,qrsolv,enorm,spmpar}.f (MINPACK-1, More/Garbow/Hillstrom 1980, single
// precision).  One function evaluation at a time, as the reference runs it.  Arrays keep the Fortran's 1-based,
// column-major addressing through the macros below so that every loop reads like the routine it follows.
// Not pinned by reference tests (the reference has none for sminpack): pinned on MINPACK's own published test
// problems in tests/test_lm.py.
#pragma once
#include <algorithm>
#include <cmath>
#include <functional>
#include <vector>

namespace ko {

// fcn(m, n, x, fvec) -> iflag (< 0 stops); x may be changed in place (minimizer_engine.f90:829-848)
typedef std::function<int(int m, int n, float* x, float* fvec)> LmFcn;

static const float lm_epsmch = 1.192091E-07f;   // spmpar(1)
static const float lm_dwarf = 1.175495E-38f;    // spmpar(2)

// sminpack/enorm.f
static inline float lm_enorm(int n, const float* x_) {
#define X(i) x_[(i)-1]
    const float one = 1.0f, zero = 0.0f, rdwarf = 3.834e-20f, rgiant = 1.304e19f;
    float s1 = zero, s2 = zero, s3 = zero, x1max = zero, x3max = zero;
    const float floatn = (float)n;
    const float agiant = rgiant / floatn;
    for (int i = 1; i <= n; i++) {
        const float xabs = fabsf(X(i));
        if (xabs > rdwarf && xabs < agiant) { s2 = s2 + xabs * xabs; continue; }
        if (xabs <= rdwarf) {
            if (xabs <= x3max) { if (xabs != zero) { const float t = xabs / x3max; s3 = s3 + t * t; } }
            else { const float t = x3max / xabs; s3 = one + s3 * (t * t); x3max = xabs; }
        } else {
            if (xabs <= x1max) { const float t = xabs / x1max; s1 = s1 + t * t; }
            else { const float t = x1max / xabs; s1 = one + s1 * (t * t); x1max = xabs; }
        }
    }
    if (s1 != zero) return x1max * sqrtf(s1 + (s2 / x1max) / x1max);
    if (s2 != zero) {
        if (s2 >= x3max) return sqrtf(s2 * (one + (x3max / s2) * (x3max * s3)));
        return sqrtf(x3max * ((s2 / x3max) + (x3max * s3)));
    }
    return x3max * sqrtf(s3);
#undef X
}

// sminpack/qrfac.f
static inline void lm_qrfac(int m, int n, float* a_, int lda, bool pivot, int* ipvt_, float* rdiag_, float* acnorm_, float* wa_) {
#define A(i, j) a_[((i)-1) + (size_t)((j)-1) * lda]
#define IPVT(j) ipvt_[(j)-1]
#define RDIAG(j) rdiag_[(j)-1]
#define ACNORM(j) acnorm_[(j)-1]
#define WA(j) wa_[(j)-1]
    const float one = 1.0f, p05 = 5.0e-2f, zero = 0.0f;
    for (int j = 1; j <= n; j++) {
        ACNORM(j) = lm_enorm(m, &A(1, j));
        RDIAG(j) = ACNORM(j);
        WA(j) = RDIAG(j);
        if (pivot) IPVT(j) = j;
    }
    const int minmn = std::min(m, n);
    for (int j = 1; j <= minmn; j++) {
        if (pivot) {
            int kmax = j;
            for (int k = j; k <= n; k++) if (RDIAG(k) > RDIAG(kmax)) kmax = k;
            if (kmax != j) {
                for (int i = 1; i <= m; i++) { const float temp = A(i, j); A(i, j) = A(i, kmax); A(i, kmax) = temp; }
                RDIAG(kmax) = RDIAG(j);
                WA(kmax) = WA(j);
                const int k = IPVT(j); IPVT(j) = IPVT(kmax); IPVT(kmax) = k;
            }
        }
        float ajnorm = lm_enorm(m - j + 1, &A(j, j));
        if (ajnorm != zero) {
            if (A(j, j) < zero) ajnorm = -ajnorm;
            for (int i = j; i <= m; i++) A(i, j) = A(i, j) / ajnorm;
            A(j, j) = A(j, j) + one;
            const int jp1 = j + 1;
            for (int k = jp1; k <= n; k++) {
                float sum = zero;
                for (int i = j; i <= m; i++) sum = sum + A(i, j) * A(i, k);
                float temp = sum / A(j, j);
                for (int i = j; i <= m; i++) A(i, k) = A(i, k) - temp * A(i, j);
                if (!pivot || RDIAG(k) == zero) continue;
                temp = A(j, k) / RDIAG(k);
                RDIAG(k) = RDIAG(k) * sqrtf(std::max(zero, one - temp * temp));
                const float q = RDIAG(k) / WA(k);
                if (p05 * (q * q) > lm_epsmch) continue;
                RDIAG(k) = lm_enorm(m - j, &A(jp1, k));
                WA(k) = RDIAG(k);
            }
        }
        RDIAG(j) = -ajnorm;
    }
#undef A
#undef IPVT
#undef RDIAG
#undef ACNORM
#undef WA
}

// sminpack/qrsolv.f
static inline void lm_qrsolv(int n, float* r_, int ldr, const int* ipvt_, const float* diag_, const float* qtb_, float* x_, float* sdiag_, float* wa_) {
#define R(i, j) r_[((i)-1) + (size_t)((j)-1) * ldr]
#define IPVT(j) ipvt_[(j)-1]
#define DIAG(j) diag_[(j)-1]
#define QTB(j) qtb_[(j)-1]
#define X(j) x_[(j)-1]
#define SDIAG(j) sdiag_[(j)-1]
#define WA(j) wa_[(j)-1]
    const float p5 = 5.0e-1f, p25 = 2.5e-1f, zero = 0.0f;
    for (int j = 1; j <= n; j++) {
        for (int i = j; i <= n; i++) R(i, j) = R(j, i);
        X(j) = R(j, j);
        WA(j) = QTB(j);
    }
    for (int j = 1; j <= n; j++) {
        const int l = IPVT(j) ;
        if (DIAG(l) != zero) {
            for (int k = j; k <= n; k++) SDIAG(k) = zero;
            SDIAG(j) = DIAG(l);
            float qtbpj = zero;
            for (int k = j; k <= n; k++) {
                if (SDIAG(k) == zero) continue;
                float cos_, sin_;
                if (fabsf(R(k, k)) >= fabsf(SDIAG(k))) {
                    const float tan_ = SDIAG(k) / R(k, k);
                    cos_ = p5 / sqrtf(p25 + p25 * (tan_ * tan_));
                    sin_ = cos_ * tan_;
                } else {
                    const float cotan = R(k, k) / SDIAG(k);
                    sin_ = p5 / sqrtf(p25 + p25 * (cotan * cotan));
                    cos_ = sin_ * cotan;
                }
                R(k, k) = cos_ * R(k, k) + sin_ * SDIAG(k);
                float temp = cos_ * WA(k) + sin_ * qtbpj;
                qtbpj = -sin_ * WA(k) + cos_ * qtbpj;
                WA(k) = temp;
                for (int i = k + 1; i <= n; i++) {
                    temp = cos_ * R(i, k) + sin_ * SDIAG(i);
                    SDIAG(i) = -sin_ * R(i, k) + cos_ * SDIAG(i);
                    R(i, k) = temp;
                }
            }
        }
        SDIAG(j) = R(j, j);
        R(j, j) = X(j);
    }
    int nsing = n;
    for (int j = 1; j <= n; j++) {
        if (SDIAG(j) == zero && nsing == n) nsing = j - 1;
        if (nsing < n) WA(j) = zero;
    }
    for (int k = 1; k <= nsing; k++) {
        const int j = nsing - k + 1;
        float sum = zero;
        for (int i = j + 1; i <= nsing; i++) sum = sum + R(i, j) * WA(i);
        WA(j) = (WA(j) - sum) / SDIAG(j);
    }
    for (int j = 1; j <= n; j++) X(IPVT(j)) = WA(j);
#undef R
#undef IPVT
#undef DIAG
#undef QTB
#undef X
#undef SDIAG
#undef WA
}

// sminpack/lmpar.f
static inline void lm_lmpar(int n, float* r_, int ldr, const int* ipvt_, const float* diag_, const float* qtb_, float delta, float& par, float* x_,
                            float* sdiag_, float* wa1_, float* wa2_) {
#define R(i, j) r_[((i)-1) + (size_t)((j)-1) * ldr]
#define IPVT(j) ipvt_[(j)-1]
#define DIAG(j) diag_[(j)-1]
#define QTB(j) qtb_[(j)-1]
#define X(j) x_[(j)-1]
#define SDIAG(j) sdiag_[(j)-1]
#define WA1(j) wa1_[(j)-1]
#define WA2(j) wa2_[(j)-1]
    const float p1 = 1.0e-1f, p001 = 1.0e-3f, zero = 0.0f;
    const float dwarf = lm_dwarf;
    int nsing = n;
    for (int j = 1; j <= n; j++) {
        WA1(j) = QTB(j);
        if (R(j, j) == zero && nsing == n) nsing = j - 1;
        if (nsing < n) WA1(j) = zero;
    }
    for (int k = 1; k <= nsing; k++) {
        const int j = nsing - k + 1;
        WA1(j) = WA1(j) / R(j, j);
        const float temp = WA1(j);
        for (int i = 1; i <= j - 1; i++) WA1(i) = WA1(i) - R(i, j) * temp;
    }
    for (int j = 1; j <= n; j++) X(IPVT(j)) = WA1(j);
    int iter = 0;
    for (int j = 1; j <= n; j++) WA2(j) = DIAG(j) * X(j);
    float dxnorm = lm_enorm(n, wa2_);
    float fp = dxnorm - delta;
    if (fp > p1 * delta) {
        float parl = zero;
        if (nsing >= n) {
            for (int j = 1; j <= n; j++) { const int l = IPVT(j); WA1(j) = DIAG(l) * (WA2(l) / dxnorm); }
            for (int j = 1; j <= n; j++) {
                float sum = zero;
                for (int i = 1; i <= j - 1; i++) sum = sum + R(i, j) * WA1(i);
                WA1(j) = (WA1(j) - sum) / R(j, j);
            }
            const float temp = lm_enorm(n, wa1_);
            parl = ((fp / delta) / temp) / temp;
        }
        for (int j = 1; j <= n; j++) {
            float sum = zero;
            for (int i = 1; i <= j; i++) sum = sum + R(i, j) * QTB(i);
            const int l = IPVT(j);
            WA1(j) = sum / DIAG(l);
        }
        const float gnorm = lm_enorm(n, wa1_);
        float paru = gnorm / delta;
        if (paru == zero) paru = dwarf / std::min(delta, p1);
        par = std::max(par, parl);
        par = std::min(par, paru);
        if (par == zero) par = gnorm / dxnorm;
        for (;;) {
            iter = iter + 1;
            if (par == zero) par = std::max(dwarf, p001 * paru);
            float temp = sqrtf(par);
            for (int j = 1; j <= n; j++) WA1(j) = temp * DIAG(j);
            lm_qrsolv(n, r_, ldr, ipvt_, wa1_, qtb_, x_, sdiag_, wa2_);
            for (int j = 1; j <= n; j++) WA2(j) = DIAG(j) * X(j);
            dxnorm = lm_enorm(n, wa2_);
            temp = fp;
            fp = dxnorm - delta;
            if (fabsf(fp) <= p1 * delta || (parl == zero && fp <= temp && temp < zero) || iter == 10) break;
            for (int j = 1; j <= n; j++) { const int l = IPVT(j); WA1(j) = DIAG(l) * (WA2(l) / dxnorm); }
            for (int j = 1; j <= n; j++) {
                WA1(j) = WA1(j) / SDIAG(j);
                temp = WA1(j);
                for (int i = j + 1; i <= n; i++) WA1(i) = WA1(i) - R(i, j) * temp;
            }
            temp = lm_enorm(n, wa1_);
            const float parc = ((fp / delta) / temp) / temp;
            if (fp > zero) parl = std::max(parl, par);
            if (fp < zero) paru = std::min(paru, par);
            par = std::max(parl, par + parc);
        }
    }
    if (iter == 0) par = zero;
#undef R
#undef IPVT
#undef DIAG
#undef QTB
#undef X
#undef SDIAG
#undef WA1
#undef WA2
}

// sminpack/fdjac2.f
static inline void lm_fdjac2(const LmFcn& fcn, int m, int n, float* x_, const float* fvec_, float* fjac_, int ldfjac, int& iflag, float epsfcn, float* wa_) {
    const float zero = 0.0f;
    const float eps = sqrtf(std::max(epsfcn, lm_epsmch));
    for (int j = 1; j <= n; j++) {
        const float temp = x_[j - 1];
        float h = eps * fabsf(temp);
        if (h == zero) h = eps;
        x_[j - 1] = temp + h;
        iflag = fcn(m, n, x_, wa_);
        if (iflag < 0) return;
        x_[j - 1] = temp;
        for (int i = 1; i <= m; i++) fjac_[(i - 1) + (size_t)(j - 1) * ldfjac] = (wa_[i - 1] - fvec_[i - 1]) / h;
    }
}

// sminpack/lmdif.f (nprint = 0).  fcn returns the iflag it leaves (>= 0 to go on).
static inline void lm_lmdif(const LmFcn& fcn, int m, int n, float* x_, float* fvec_, float ftol, float xtol, float gtol, int maxfev, float epsfcn,
                            float* diag_, int mode, float factor, int& info, int& nfev) {
#define X(j) x_[(j)-1]
#define FVEC(i) fvec_[(i)-1]
#define DIAG(j) diag_[(j)-1]
#define FJAC(i, j) fjac[((i)-1) + (size_t)((j)-1) * ldfjac]
#define IPVT(j) ipvt[(j)-1]
#define QTF(j) qtf[(j)-1]
#define WA1(j) wa1[(j)-1]
#define WA2(j) wa2[(j)-1]
#define WA3(j) wa3[(j)-1]
#define WA4(i) wa4[(i)-1]
    const float one = 1.0f, p1 = 1.0e-1f, p5 = 5.0e-1f, p25 = 2.5e-1f, p75 = 7.5e-1f, p0001 = 1.0e-4f, zero = 0.0f;
    const float epsmch = lm_epsmch;
    const int ldfjac = m;
    info = 0; nfev = 0;
    int iflag = 0;
    if (n <= 0 || m < n || ftol < zero || xtol < zero || gtol < zero || maxfev <= 0 || factor <= zero) return;
    if (mode == 2) for (int j = 1; j <= n; j++) if (DIAG(j) <= zero) return;
    std::vector<float> fjac((size_t)m * n), qtf(n), wa1(n), wa2(n), wa3(n), wa4(m);
    std::vector<int> ipvt(n);
    float actred, delta = zero, dirder, fnorm, fnorm1, gnorm = zero, par, pnorm, prered, ratio, sum, temp = zero, temp1, temp2, xnorm = zero;
    iflag = fcn(m, n, x_, fvec_);
    nfev = 1;
    if (iflag < 0) { info = iflag; return; }
    fnorm = lm_enorm(m, fvec_);
    par = zero;
    int iter = 1;
    for (;;) {   // 30: outer loop
        iflag = 2;
        lm_fdjac2(fcn, m, n, x_, fvec_, fjac.data(), ldfjac, iflag, epsfcn, wa4.data());
        nfev = nfev + n;
        if (iflag < 0) break;
        lm_qrfac(m, n, fjac.data(), ldfjac, true, ipvt.data(), wa1.data(), wa2.data(), wa3.data());
        if (iter == 1) {
            if (mode != 2) for (int j = 1; j <= n; j++) { DIAG(j) = WA2(j); if (WA2(j) == zero) DIAG(j) = one; }
            for (int j = 1; j <= n; j++) WA3(j) = DIAG(j) * X(j);
            xnorm = lm_enorm(n, wa3.data());
            delta = factor * xnorm;
            if (delta == zero) delta = factor;
        }
        for (int i = 1; i <= m; i++) WA4(i) = FVEC(i);
        for (int j = 1; j <= n; j++) {
            if (FJAC(j, j) != zero) {
                sum = zero;
                for (int i = j; i <= m; i++) sum = sum + FJAC(i, j) * WA4(i);
                temp = -sum / FJAC(j, j);
                for (int i = j; i <= m; i++) WA4(i) = WA4(i) + FJAC(i, j) * temp;
            }
            FJAC(j, j) = WA1(j);
            QTF(j) = WA4(j);
        }
        gnorm = zero;
        if (fnorm != zero)
            for (int j = 1; j <= n; j++) {
                const int l = IPVT(j);
                if (WA2(l) == zero) continue;
                sum = zero;
                for (int i = 1; i <= j; i++) sum = sum + FJAC(i, j) * (QTF(i) / fnorm);
                gnorm = std::max(gnorm, fabsf(sum / WA2(l)));
            }
        if (gnorm <= gtol) info = 4;
        if (info != 0) break;
        if (mode != 2) for (int j = 1; j <= n; j++) DIAG(j) = std::max(DIAG(j), WA2(j));
        bool stop = false;
        for (;;) {   // 200: inner loop
            lm_lmpar(n, fjac.data(), ldfjac, ipvt.data(), diag_, qtf.data(), delta, par, wa1.data(), wa2.data(), wa3.data(), wa4.data());
            for (int j = 1; j <= n; j++) {
                WA1(j) = -WA1(j);
                WA2(j) = X(j) + WA1(j);
                WA3(j) = DIAG(j) * WA1(j);
            }
            pnorm = lm_enorm(n, wa3.data());
            if (iter == 1) delta = std::min(delta, pnorm);
            iflag = fcn(m, n, wa2.data(), wa4.data());
            nfev = nfev + 1;
            if (iflag < 0) { stop = true; break; }
            fnorm1 = lm_enorm(m, wa4.data());
            actred = -one;
            if (p1 * fnorm1 < fnorm) { const float q = fnorm1 / fnorm; actred = one - q * q; }
            for (int j = 1; j <= n; j++) {
                WA3(j) = zero;
                const int l = IPVT(j);
                temp = WA1(l);
                for (int i = 1; i <= j; i++) WA3(i) = WA3(i) + FJAC(i, j) * temp;
            }
            temp1 = lm_enorm(n, wa3.data()) / fnorm;
            temp2 = (sqrtf(par) * pnorm) / fnorm;
            prered = temp1 * temp1 + (temp2 * temp2) / p5;
            dirder = -(temp1 * temp1 + temp2 * temp2);
            ratio = zero;
            if (prered != zero) ratio = actred / prered;
            if (ratio <= p25) {
                if (actred >= zero) temp = p5;
                if (actred < zero) temp = p5 * dirder / (dirder + p5 * actred);
                if (p1 * fnorm1 >= fnorm || temp < p1) temp = p1;
                delta = temp * std::min(delta, pnorm / p1);
                par = par / temp;
            } else if (!(par != zero && ratio < p75)) {
                delta = pnorm / p5;
                par = p5 * par;
            }
            if (!(ratio < p0001)) {
                for (int j = 1; j <= n; j++) { X(j) = WA2(j); WA2(j) = DIAG(j) * X(j); }
                for (int i = 1; i <= m; i++) FVEC(i) = WA4(i);
                xnorm = lm_enorm(n, wa2.data());
                fnorm = fnorm1;
                iter = iter + 1;
            }
            if (fabsf(actred) <= ftol && prered <= ftol && p5 * ratio <= one) info = 1;
            if (delta <= xtol * xnorm) info = 2;
            if (fabsf(actred) <= ftol && prered <= ftol && p5 * ratio <= one && info == 2) info = 3;
            if (info != 0) { stop = true; break; }
            if (nfev >= maxfev) info = 5;
            if (fabsf(actred) <= epsmch && prered <= epsmch && p5 * ratio <= one) info = 6;
            if (delta <= epsmch * xnorm) info = 7;
            if (gnorm <= epsmch) info = 8;
            if (info != 0) { stop = true; break; }
            if (!(ratio < p0001)) break;
        }
        if (stop) break;
    }
    if (iflag < 0) info = iflag;   // 300
#undef X
#undef FVEC
#undef DIAG
#undef FJAC
#undef IPVT
#undef QTF
#undef WA1
#undef WA2
#undef WA3
#undef WA4
}

}  // namespace ko
