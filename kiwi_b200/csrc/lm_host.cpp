// MINPACK lmdif in single precision with a batched finite-difference Jacobian; see lm_host.hpp.
// Algorithm and constants: J. J. More, B. S. Garbow, K. E. Hillstrom, MINPACK-1 (Argonne 1980), as shipped with the
// reference in sminpack/ (lmdif.f, fdjac2.f, lmpar.f, qrfac.f, qrsolv.f, enorm.f, spmpar.f).  Arrays are 0-based and
// column-major here (a[i + j*lda]); the order of every floating-point operation is the reference's (this file is
// built with -ffp-contract=off like the rest of the host code).
#include "lm_host.hpp"

#include <algorithm>
#include <cmath>

namespace klm {

namespace {
const float EPSMCH = 1.192091E-07f;   // spmpar(1), sminpack/spmpar.f
const float DWARF = 1.175495E-38f;    // spmpar(2)
inline float sq(float v) { return v * v; }
}  // namespace

float enorm(int n, const float* x) {
    const float rdwarf = 3.834e-20f, rgiant = 1.304e19f;
    float s1 = 0.f, s2 = 0.f, s3 = 0.f, x1max = 0.f, x3max = 0.f;
    const float agiant = rgiant / (float)n;
    for (int i = 0; i < n; i++) {
        const float xabs = fabsf(x[i]);
        if (xabs > rdwarf && xabs < agiant) {
            s2 = s2 + sq(xabs);                                   // intermediate components
        } else if (xabs <= rdwarf) {                              // small components
            if (xabs > x3max) { s3 = 1.f + s3 * sq(x3max / xabs); x3max = xabs; }
            else if (xabs != 0.f) s3 = s3 + sq(xabs / x3max);
        } else {                                                  // large components
            if (xabs > x1max) { s1 = 1.f + s1 * sq(x1max / xabs); x1max = xabs; }
            else s1 = s1 + sq(xabs / x1max);
        }
    }
    if (s1 != 0.f) return x1max * sqrtf(s1 + (s2 / x1max) / x1max);
    if (s2 != 0.f) {
        if (s2 >= x3max) return sqrtf(s2 * (1.f + (x3max / s2) * (x3max * s3)));
        return sqrtf(x3max * ((s2 / x3max) + (x3max * s3)));
    }
    return x3max * sqrtf(s3);
}

// Householder QR with column pivoting: a*p = q*r; the strict upper triangle of r and the Householder vectors are
// left in `a`, the diagonal of r in rdiag
void qrfac(int m, int n, float* a, int lda, bool pivot, int* ipvt, float* rdiag, float* acnorm, float* wa) {
#define A(i, j) a[(size_t)(i) + (size_t)(j) * lda]
    for (int j = 0; j < n; j++) {
        acnorm[j] = enorm(m, &A(0, j));
        rdiag[j] = acnorm[j];
        wa[j] = rdiag[j];
        if (pivot) ipvt[j] = j;
    }
    const int minmn = std::min(m, n);
    for (int j = 0; j < minmn; j++) {
        if (pivot) {   // bring the column of largest norm into the pivot position
            int kmax = j;
            for (int k = j; k < n; k++) if (rdiag[k] > rdiag[kmax]) kmax = k;
            if (kmax != j) {
                for (int i = 0; i < m; i++) std::swap(A(i, j), A(i, kmax));
                rdiag[kmax] = rdiag[j];
                wa[kmax] = wa[j];
                std::swap(ipvt[j], ipvt[kmax]);
            }
        }
        float ajnorm = enorm(m - j, &A(j, j));
        if (ajnorm != 0.f) {
            if (A(j, j) < 0.f) ajnorm = -ajnorm;
            for (int i = j; i < m; i++) A(i, j) = A(i, j) / ajnorm;
            A(j, j) = A(j, j) + 1.f;
            for (int k = j + 1; k < n; k++) {   // apply the transformation to the remaining columns, update the norms
                float sum = 0.f;
                for (int i = j; i < m; i++) sum = sum + A(i, j) * A(i, k);
                const float temp = sum / A(j, j);
                for (int i = j; i < m; i++) A(i, k) = A(i, k) - temp * A(i, j);
                if (pivot && rdiag[k] != 0.f) {
                    const float t = A(j, k) / rdiag[k];
                    rdiag[k] = rdiag[k] * sqrtf(std::max(0.f, 1.f - sq(t)));
                    if (0.05f * sq(rdiag[k] / wa[k]) <= EPSMCH) {
                        rdiag[k] = enorm(m - j - 1, &A(j + 1, k));
                        wa[k] = rdiag[k];
                    }
                }
            }
        }
        rdiag[j] = -ajnorm;
    }
#undef A
}

// least squares solution of a*x = b, d*x = 0 given the QR factorisation of a (Givens rotations eliminate d)
void qrsolv(int n, float* r, int ldr, const int* ipvt, const float* diag, const float* qtb, float* x, float* sdiag, float* wa) {
#define R(i, j) r[(size_t)(i) + (size_t)(j) * ldr]
    for (int j = 0; j < n; j++) {
        for (int i = j; i < n; i++) R(i, j) = R(j, i);
        x[j] = R(j, j);
        wa[j] = qtb[j];
    }
    for (int j = 0; j < n; j++) {
        const int l = ipvt[j];
        if (diag[l] != 0.f) {
            for (int k = j; k < n; k++) sdiag[k] = 0.f;
            sdiag[j] = diag[l];
            float qtbpj = 0.f;
            for (int k = j; k < n; k++) {
                if (sdiag[k] == 0.f) continue;
                float cs, sn;
                if (fabsf(R(k, k)) >= fabsf(sdiag[k])) {
                    const float tn = sdiag[k] / R(k, k);
                    cs = 0.5f / sqrtf(0.25f + 0.25f * sq(tn));
                    sn = cs * tn;
                } else {
                    const float ct = R(k, k) / sdiag[k];
                    sn = 0.5f / sqrtf(0.25f + 0.25f * sq(ct));
                    cs = sn * ct;
                }
                R(k, k) = cs * R(k, k) + sn * sdiag[k];
                const float temp = cs * wa[k] + sn * qtbpj;
                qtbpj = -sn * wa[k] + cs * qtbpj;
                wa[k] = temp;
                for (int i = k + 1; i < n; i++) {
                    const float t = cs * R(i, k) + sn * sdiag[i];
                    sdiag[i] = -sn * R(i, k) + cs * sdiag[i];
                    R(i, k) = t;
                }
            }
        }
        sdiag[j] = R(j, j);
        R(j, j) = x[j];
    }
    int nsing = n;
    for (int j = 0; j < n; j++) {
        if (sdiag[j] == 0.f && nsing == n) nsing = j;
        if (nsing < n) wa[j] = 0.f;
    }
    for (int j = nsing - 1; j >= 0; j--) {
        float sum = 0.f;
        for (int i = j + 1; i < nsing; i++) sum = sum + R(i, j) * wa[i];
        wa[j] = (wa[j] - sum) / sdiag[j];
    }
    for (int j = 0; j < n; j++) x[ipvt[j]] = wa[j];
#undef R
}

// Levenberg-Marquardt parameter: par such that the step solves the problem with |d*x| within 10 % of delta
void lmpar(int n, float* r, int ldr, const int* ipvt, const float* diag, const float* qtb, float delta, float& par, float* x, float* sdiag,
           float* wa1, float* wa2) {
#define R(i, j) r[(size_t)(i) + (size_t)(j) * ldr]
    // Gauss-Newton direction
    int nsing = n;
    for (int j = 0; j < n; j++) {
        wa1[j] = qtb[j];
        if (R(j, j) == 0.f && nsing == n) nsing = j;
        if (nsing < n) wa1[j] = 0.f;
    }
    for (int j = nsing - 1; j >= 0; j--) {
        wa1[j] = wa1[j] / R(j, j);
        const float temp = wa1[j];
        for (int i = 0; i < j; i++) wa1[i] = wa1[i] - R(i, j) * temp;
    }
    for (int j = 0; j < n; j++) x[ipvt[j]] = wa1[j];
    int iter = 0;
    for (int j = 0; j < n; j++) wa2[j] = diag[j] * x[j];
    float dxnorm = enorm(n, wa2);
    float fp = dxnorm - delta;
    if (fp <= 0.1f * delta) { par = 0.f; return; }   // (iter == 0: par = zero)
    // lower bound parl from the Newton step of the secular function (full rank only)
    float parl = 0.f;
    if (nsing >= n) {
        for (int j = 0; j < n; j++) { const int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
        for (int j = 0; j < n; j++) {
            float sum = 0.f;
            for (int i = 0; i < j; i++) sum = sum + R(i, j) * wa1[i];
            wa1[j] = (wa1[j] - sum) / R(j, j);
        }
        const float temp = enorm(n, wa1);
        parl = ((fp / delta) / temp) / temp;
    }
    // upper bound paru
    for (int j = 0; j < n; j++) {
        float sum = 0.f;
        for (int i = 0; i <= j; i++) sum = sum + R(i, j) * qtb[i];
        wa1[j] = sum / diag[ipvt[j]];
    }
    const float gnorm = enorm(n, wa1);
    float paru = gnorm / delta;
    if (paru == 0.f) paru = DWARF / std::min(delta, 0.1f);
    par = std::max(par, parl);
    par = std::min(par, paru);
    if (par == 0.f) par = gnorm / dxnorm;
    for (;;) {
        iter++;
        if (par == 0.f) par = std::max(DWARF, 0.001f * paru);
        float temp = sqrtf(par);
        for (int j = 0; j < n; j++) wa1[j] = temp * diag[j];
        qrsolv(n, r, ldr, ipvt, wa1, qtb, x, sdiag, wa2);
        for (int j = 0; j < n; j++) wa2[j] = diag[j] * x[j];
        dxnorm = enorm(n, wa2);
        temp = fp;
        fp = dxnorm - delta;
        if (fabsf(fp) <= 0.1f * delta || (parl == 0.f && fp <= temp && temp < 0.f) || iter == 10) break;
        // Newton correction
        for (int j = 0; j < n; j++) { const int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
        for (int j = 0; j < n; j++) {
            wa1[j] = wa1[j] / sdiag[j];
            const float t = wa1[j];
            for (int i = j + 1; i < n; i++) wa1[i] = wa1[i] - R(i, j) * t;
        }
        temp = enorm(n, wa1);
        const float parc = ((fp / delta) / temp) / temp;
        if (fp > 0.f) parl = std::max(parl, par);
        if (fp < 0.f) paru = std::min(paru, par);
        par = std::max(parl, par + parc);
    }
#undef R
}

Result lmdif_batched(const BatchFcn& fcn, int m, int n, float* x, float* fvec, float ftol, float xtol, float gtol, int maxfev, float epsfcn,
                     float* diag, int mode, float factor) {
    Result res;
    if (n <= 0 || m < n || ftol < 0.f || xtol < 0.f || gtol < 0.f || maxfev <= 0 || factor <= 0.f) return res;
    if (mode == 2) for (int j = 0; j < n; j++) if (diag[j] <= 0.f) return res;
    const int ldfjac = m;
    std::vector<float> fjac((size_t)m * n), qtf(n), wa1(n), wa2(n), wa3(n), wa4(m), xs((size_t)n * n), fs((size_t)n * m), hs(n);
    std::vector<int> ipvt(n);
#define FJ(i, j) fjac[(size_t)(i) + (size_t)(j) * ldfjac]
    // function at the starting point
    if (fcn(1, x, fvec) < 1) { res.nfev = 1; res.info = -2; return res; }
    res.nfev = 1;
    float fnorm = enorm(m, fvec);
    float par = 0.f, delta = 0.f, xnorm = 0.f, gnorm = 0.f;
    int iter = 1;
    const float eps = sqrtf(std::max(epsfcn, EPSMCH));
    for (;;) {   // outer loop
        // ---- forward-difference Jacobian (fdjac2): all n perturbed vectors in one batch -------------------------------
        for (int j = 0; j < n; j++) {
            float* xj = &xs[(size_t)j * n];
            std::copy(x, x + n, xj);
            const float temp = x[j];
            float h = eps * fabsf(temp);
            if (h == 0.f) h = eps;
            xj[j] = temp + h;
            hs[j] = h;
        }
        const int nok = fcn(n, xs.data(), fs.data());
        res.nfev += n;
        if (nok < n) {   // fdjac2 returns with x(j) still perturbed when the function stops it (sminpack/fdjac2.f: go to 30)
            x[nok] = xs[(size_t)nok * n + nok];
            res.info = -2;
            return res;
        }
        for (int j = 0; j < n; j++)
            for (int i = 0; i < m; i++) FJ(i, j) = (fs[(size_t)j * m + i] - fvec[i]) / hs[j];
        // ---- QR factorisation of the Jacobian -----------------------------------------------------------------------------
        qrfac(m, n, fjac.data(), ldfjac, true, ipvt.data(), wa1.data(), wa2.data(), wa3.data());
        if (iter == 1) {   // scale according to the norms of the columns of the initial Jacobian, initial step bound
            if (mode != 2) for (int j = 0; j < n; j++) { diag[j] = wa2[j]; if (wa2[j] == 0.f) diag[j] = 1.f; }
            for (int j = 0; j < n; j++) wa3[j] = diag[j] * x[j];
            xnorm = enorm(n, wa3.data());
            delta = factor * xnorm;
            if (delta == 0.f) delta = factor;
        }
        // ---- (q transpose)*fvec, first n components in qtf -----------------------------------------------------------------
        for (int i = 0; i < m; i++) wa4[i] = fvec[i];
        for (int j = 0; j < n; j++) {
            if (FJ(j, j) != 0.f) {
                float sum = 0.f;
                for (int i = j; i < m; i++) sum = sum + FJ(i, j) * wa4[i];
                const float temp = -sum / FJ(j, j);
                for (int i = j; i < m; i++) wa4[i] = wa4[i] + FJ(i, j) * temp;
            }
            FJ(j, j) = wa1[j];
            qtf[j] = wa4[j];
        }
        // ---- norm of the scaled gradient ----------------------------------------------------------------------------------------
        gnorm = 0.f;
        if (fnorm != 0.f) {
            for (int j = 0; j < n; j++) {
                const int l = ipvt[j];
                if (wa2[l] == 0.f) continue;
                float sum = 0.f;
                for (int i = 0; i <= j; i++) sum = sum + FJ(i, j) * (qtf[i] / fnorm);
                gnorm = std::max(gnorm, fabsf(sum / wa2[l]));
            }
        }
        if (gnorm <= gtol) { res.info = 4; return res; }
        if (mode != 2) for (int j = 0; j < n; j++) diag[j] = std::max(diag[j], wa2[j]);
        // ---- inner loop: trial steps until one is accepted ---------------------------------------------------------------------
        for (;;) {
            lmpar(n, fjac.data(), ldfjac, ipvt.data(), diag, qtf.data(), delta, par, wa1.data(), wa2.data(), wa3.data(), wa4.data());
            for (int j = 0; j < n; j++) {
                wa1[j] = -wa1[j];
                wa2[j] = x[j] + wa1[j];
                wa3[j] = diag[j] * wa1[j];
            }
            const float pnorm = enorm(n, wa3.data());
            if (iter == 1) delta = std::min(delta, pnorm);
            const int ok = fcn(1, wa2.data(), wa4.data());
            res.nfev += 1;
            if (ok < 1) { res.info = -2; return res; }
            const float fnorm1 = enorm(m, wa4.data());
            float actred = -1.f;
            if (0.1f * fnorm1 < fnorm) actred = 1.f - sq(fnorm1 / fnorm);
            for (int j = 0; j < n; j++) {
                wa3[j] = 0.f;
                const float temp = wa1[ipvt[j]];
                for (int i = 0; i <= j; i++) wa3[i] = wa3[i] + FJ(i, j) * temp;
            }
            const float temp1 = enorm(n, wa3.data()) / fnorm;
            const float temp2 = (sqrtf(par) * pnorm) / fnorm;
            const float prered = sq(temp1) + sq(temp2) / 0.5f;
            const float dirder = -(sq(temp1) + sq(temp2));
            float ratio = 0.f;
            if (prered != 0.f) ratio = actred / prered;
            if (ratio <= 0.25f) {
                float temp = 0.f;   // (the reference leaves temp at its previous value when actred is NaN)
                if (actred >= 0.f) temp = 0.5f;
                if (actred < 0.f) temp = 0.5f * dirder / (dirder + 0.5f * actred);
                if (0.1f * fnorm1 >= fnorm || temp < 0.1f) temp = 0.1f;
                delta = temp * std::min(delta, pnorm / 0.1f);
                par = par / temp;
            } else if (par == 0.f || ratio >= 0.75f) {
                delta = pnorm / 0.5f;
                par = 0.5f * par;
            }
            if (ratio >= 1.0e-4f) {   // successful iteration
                for (int j = 0; j < n; j++) { x[j] = wa2[j]; wa2[j] = diag[j] * x[j]; }
                for (int i = 0; i < m; i++) fvec[i] = wa4[i];
                xnorm = enorm(n, wa2.data());
                fnorm = fnorm1;
                iter++;
            }
            // convergence tests
            if (fabsf(actred) <= ftol && prered <= ftol && 0.5f * ratio <= 1.f) res.info = 1;
            if (delta <= xtol * xnorm) res.info = 2;
            if (fabsf(actred) <= ftol && prered <= ftol && 0.5f * ratio <= 1.f && res.info == 2) res.info = 3;
            if (res.info != 0) return res;
            // termination and stringent tolerances
            if (res.nfev >= maxfev) res.info = 5;
            if (fabsf(actred) <= EPSMCH && prered <= EPSMCH && 0.5f * ratio <= 1.f) res.info = 6;
            if (delta <= EPSMCH * xnorm) res.info = 7;
            if (gnorm <= EPSMCH) res.info = 8;
            if (res.info != 0) return res;
            if (ratio >= 1.0e-4f) break;
        }
    }
#undef FJ
}

}  // namespace klm
