// Levenberg-Marquardt driver of the inversion (minimizer_engine.f90:729-874) with the MINPACK algorithm the
// reference links (sminpack/lmdif.f and the routines it calls, single precision), restructured so that
// the n finite-difference columns of the Jacobian (sminpack/fdjac2.f) are ONE batched evaluation.
#pragma once
#include <functional>
#include <vector>

namespace klm {

// Evaluates `ncols` parameter vectors at once.  xs: [ncols][n], may be changed in place (the reference's forward
// step clips parameters to their limits, minimizer_engine.f90:829-848); fvecs: [ncols][m].  Returns the number
// of leading columns that were evaluated successfully (ncols if all were): the sequential reference stops at
// the first failure (iflag < 0).
typedef std::function<int(int ncols, float* xs, float* fvecs)> BatchFcn;

struct Result {
    int info = 0;   // as lmdif: 0 improper input, 1-8 convergence / termination, < 0 stopped by the function (iflag)
    int nfev = 0;   // function evaluations requested (every Jacobian column counts as one, sminpack/lmdif.f:294)
};

// lmdif (sminpack/lmdif.f); x [n] and fvec [m] are updated to the final iterate, diag [n] is input for mode 2
Result lmdif_batched(const BatchFcn& fcn, int m, int n, float* x, float* fvec, float ftol, float xtol, float gtol, int maxfev, float epsfcn,
                     float* diag, int mode, float factor);

// building blocks (exported for the known-answer tests)
float enorm(int n, const float* x);                                                                             // sminpack/enorm.f
void qrfac(int m, int n, float* a, int lda, bool pivot, int* ipvt, float* rdiag, float* acnorm, float* wa);     // sminpack/qrfac.f
void qrsolv(int n, float* r, int ldr, const int* ipvt, const float* diag, const float* qtb, float* x, float* sdiag, float* wa);   // qrsolv.f
void lmpar(int n, float* r, int ldr, const int* ipvt, const float* diag, const float* qtb, float delta, float& par, float* x, float* sdiag,
           float* wa1, float* wa2);                                                                            // sminpack/lmpar.f

}  // namespace klm
