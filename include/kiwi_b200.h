/* kiwi_b200.h -- C ABI of the B200-native forward-modelling + misfit engine for Kiwi's
 * source-inversion inner loop.
 *
 * The reference (emolch/kiwi, Fortran 90) has no plugin/FFI interface.  Its de-facto boundary is
 * the module API of minimizer_engine.f90 (38 public subroutines `(args..., ok)` + global error
 * string, minimizer_engine.f90:39-76, util.f90:106-116), driven by the text protocol of the
 * `minimizer` program (minimizer.f90:1676-1813).  Every entry point below replaces one of those
 * subroutines (cited as file:line of /root/reference) for the hot path
 *     set_source_params -> discretise -> make_seismogram (all receivers) -> scale -> misfits.
 * A Fortran host binds them with iso_c_binding (fortran/kiwi_b200_binding.f90).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message (same wording as the
 *     reference's error() calls where one exists) is available from kiwi_last_error();
 *   - plain pointers and sizes only; caller-provided output buffers with an explicit capacity;
 *   - receiver / component / GF indices are 1-based as in the reference;
 *   - one context = the module-level singletons of minimizer_engine.f90:78-108; calls on one
 *     context must be serialised by the caller (the reference is single-threaded at this level);
 *   - there is NO CPU fallback: any call that computes needs a CUDA device and fails loudly
 *     without one.
 */
#ifndef KIWI_B200_H
#define KIWI_B200_H

#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct kiwi_ctx kiwi_ctx;
typedef struct kiwi_gfdb kiwi_gfdb;

/* source type ids (parameterized_source.f90 / source_all.f90:58-60) */
#define KIWI_SOURCE_BILATERAL 1      /* source_bilat.f90, 14 parameters          */
#define KIWI_SOURCE_CIRCULAR 2       /* source_circular.f90, 11 parameters        */
#define KIWI_SOURCE_POINT_LP 3       /* source_point_lp.f90, 13 parameters        */
#define KIWI_SOURCE_EIKONAL 4        /* source_eikonal.f90, 15 parameters         */
#define KIWI_SOURCE_MT_EIKONAL 5     /* source_mt_eikonal.f90, 20 parameters      */
#define KIWI_SOURCE_MOMENT_TENSOR 6  /* source_moment_tensor.f90, 11 parameters   */

/* misfit norm ids (comparator.f90:33-42) */
#define KIWI_L2NORM 1
#define KIWI_L1NORM 2
#define KIWI_AMPSPEC_L2NORM 3
#define KIWI_AMPSPEC_L1NORM 4
#define KIWI_SCALAR_PRODUCT 5
#define KIWI_PEAK 6
#define KIWI_FLOATING_L2NORM 7
#define KIWI_FLOATING_L1NORM 8

/* per-candidate status of kiwi_eval_sources (the reference reports per-source failures through
 * `ok=.false.` + g_errstr and the Python driver collects them as `failings`,
 * python/tunguska/seismosizer.py:703-716) */
#define KIWI_STATUS_OK 0
#define KIWI_STATUS_BAD_PARAMS 1     /* discretisation failed                        */
#define KIWI_STATUS_NONFINITE 2      /* NaN/Inf misfit (minimizer_engine.f90:1163-1166) */

/* util.f90:106-116 error()/g_errstr.  Thread-local, valid until the next failing call. */
const char* kiwi_last_error(void);
/* library version string */
const char* kiwi_version(void);

/* ---------------------------------------------------------------------------------------------
 * Green's function database, host side.  Replaces the storage half of gfdb.f90 (t_gfdb :93-146,
 * gfdb_init :163-264, gfdb_save_trace) and gfdb_io_hdf.f90 (HDF5 is not available): a flat
 * container with, for every (ix,iz,ig), one dense fp32 sample array that starts at sample index
 * span0.  Packing follows trace_pack (sparse_trace.f90:443-555): leading zeros are dropped, of the
 * trailing zeros exactly one is kept, inter-strip gaps are stored as explicit zeros.
 * --------------------------------------------------------------------------------------------- */
/* gfdb_build (gfdb_build.f90) / gfdb_init: create an empty database grid */
kiwi_gfdb* kiwi_gfdb_create(int nx, int nz, int ng, float dt, float dx, float dz, float firstx, float firstz);
void kiwi_gfdb_destroy(kiwi_gfdb* db);
/* gfdb_build_ahfull.f90:193-216 gfdb_save_array: store samples data[0..n) of GF (ix,iz,ig), first
 * sample at index span0 = nint(tbegin/dt) */
int kiwi_gfdb_save_array(kiwi_gfdb* db, int ix, int iz, int ig, int span0, int n, const float* data);
/* gfdb_build_ahfull.f90:70-191 addentry for EVERY grid node: analytical homogeneous full space
 * (elseis.f90), material rho/alpha/beta, source time function stf[nstf] sampled at dt.
 * nthreads <= 0: all host cores. */
int kiwi_gfdb_build_ahfull(kiwi_gfdb* db, float rho, float alpha, float beta, const float* stf, int nstf,
                           int nfflag, int ffflag, int nthreads);
/* flat binary file "KGF1" (this repo's own format) */
int kiwi_gfdb_write(const kiwi_gfdb* db, const char* path);
kiwi_gfdb* kiwi_gfdb_read(const char* path);
/* Kiwi's own HDF5 database <basepath>.index + <basepath>.<i>.chunk (gfdb_io_hdf.f90:119-180, 429-605; gfdb.f90:1437-1467)
 * through a minimal parser of the published HDF5 file format (no libhdf5): files as HDF5 1.6/1.8 write them by default. */
kiwi_gfdb* kiwi_gfdb_read_hdf(const char* basepath);
/* One dataset of the root group of an HDF5 file through the same parser (the index file of a database is a set of these,
 * h5_open_scalar gfdb_io_hdf.f90:63-83).  name == NULL: buf receives the NUL-separated names of the root group's members.
 * dtype_class: 0 integer, 1 float, 7 reference ...; dims8: up to 8 extents; *nbytes: size of the data / of the name list. */
int kiwi_h5_read_root_dataset(const char* path, const char* name, int* dtype_class, int* dtype_size, int* rank, long long* dims8, void* buf,
                              long long cap, long long* nbytes, int* nattrs);
/* set_database dbpath nipx nipz (minimizer.f90:89-135; gfdb_init gfdb.f90:205-246): Gulunay's generalised f-k interpolation of
 * the database (gfdb_interpolate_block gfdb.f90:1109-1232, interpolate3d :1234-1310, interpolation.f90) on GPU `device`.  Returns a
 * new database with nx*nipx x nz*nipz traces at dx/nipx, dz/nipz; every interpolation block is filled eagerly (the reference fills
 * a block on first access).  nipx, nipz in {1, 2, 4, 8, 16}; every trace of the source database must be present. */
kiwi_gfdb* kiwi_gfdb_interpolate(const kiwi_gfdb* db, int nipx, int nipz, int device);
/* the interpolation operator itself (gulunay2d / gulunay3d, interpolation.f90:29-311) on `batch` fields a[batch][s2][s1][t]
 * (tapered in place like the reference's A) -> out[batch][s2*l2][s1*l1][t]; l1, l2 in {1, l}; margins as in the reference */
int kiwi_gulunay(int device, float* a, int batch, int t, int s1, int s2, int l1, int l2, int ntmargin, int margin1, int margin2, float* out);
/* gfdb_info: grid metadata */
int kiwi_gfdb_meta(const kiwi_gfdb* db, int* nx, int* nz, int* ng, float* dt, float* dx, float* dz, float* firstx,
                   float* firstz, long long* ntraces, long long* nsamples);
/* borrowed pointers to the flat arrays, index = ((ix-1)*nz + (iz-1))*ng + (ig-1); len==0: no trace */
int kiwi_gfdb_view(kiwi_gfdb* db, const int** span0, const int** len, const long long** offset, const float** data);

/* ---------------------------------------------------------------------------------------------
 * Engine context = module state of minimizer_engine.f90:78-108, resident on one GPU
 * --------------------------------------------------------------------------------------------- */
kiwi_ctx* kiwi_create(int device);
void kiwi_destroy(kiwi_ctx* ctx);

/* set_database (minimizer_engine.f90:114-139): packs the database into 16-byte aligned slabs and
 * uploads it to HBM once.  For nipx/nipz (Gulunay interpolation) pass the result of kiwi_gfdb_interpolate. */
int kiwi_set_database(kiwi_ctx* ctx, kiwi_gfdb* db);
/* set_local_interpolation (minimizer_engine.f90:140-145): 0 nearest neighbour, 1 bilinear */
int kiwi_set_local_interpolation(kiwi_ctx* ctx, int bilinear);
/* set_spacial_undersampling (minimizer_engine.f90:147-163) */
int kiwi_set_spacial_undersampling(kiwi_ctx* ctx, int xunder, int zunder);
/* set_receivers (minimizer_engine.f90:165-286) with the file parsing stripped: coordinates in
 * DEGREES as in the receivers file, components = string of distinct letters from "acrlduneswe"
 * (receiver.f90:136-209) */
int kiwi_set_receivers(kiwi_ctx* ctx, int n, const double* lat_deg, const double* lon_deg, const float* depth,
                       const char* const* components);
/* switch_receiver (minimizer_engine.f90:288-311) */
int kiwi_switch_receiver(kiwi_ctx* ctx, int ireceiver, int state);
/* set_source_location (minimizer.f90:485-519 + minimizer_engine.f90:453-467): degrees, seconds */
int kiwi_set_source_location(kiwi_ctx* ctx, float lat_deg, float lon_deg, double ref_time);
/* crust2x2_load (minimizer.f90:1669-1674, crust2x2.f90:68-74): the CRUST2.0 model the eikonal sources take
 * their rupture velocity and default depth constraints from; `path` is the binary table written by
 * tools/make_crust2x2_table.py (kiwi_b200/data/crust2x2.kcr) */
int kiwi_set_crust2x2(kiwi_ctx* ctx, const char* path);
/* set_source_constraints (minimizer.f90:521-579): n half-spaces, points[n][3] and outward normals[n][3];
 * set_source_location re-derives the two default constraints (parameterized_source.f90:183-196) */
int kiwi_set_source_constraints(kiwi_ctx* ctx, int n, const float* points, const float* normals);
/* set_source_crustal_thickness_limit (minimizer.f90:581-612) */
int kiwi_set_source_crustal_thickness_limit(kiwi_ctx* ctx, float limit);
/* set_effective_dt (minimizer_engine.f90:610-618) */
int kiwi_set_effective_dt(kiwi_ctx* ctx, float effective_dt);
/* set_ref_seismograms (minimizer_engine.f90:313-352, receiver.f90:746-851) with the file reading
 * stripped: n samples of component icomponent at receiver ireceiver, first sample at time tbegin
 * [s] relative to the source reference time */
int kiwi_set_ref_seismogram(kiwi_ctx* ctx, int ireceiver, int icomponent, float tbegin, int n, const float* data);
/* shift_ref_seismogram (minimizer_engine.f90:354-378): move the reference traces of one receiver by nint(shift/dt) samples */
int kiwi_shift_ref_seismogram(kiwi_ctx* ctx, int ireceiver, float shift);
/* autoshift_ref_seismogram (minimizer_engine.f90:380-416, receiver.f90:816-832): cross-correlate the synthetics of the current
 * source with the references pulled through [shift_lo, shift_hi] seconds (probes_windowed_cross_corr comparator.f90:1061-1090,
 * taper / filter applied as set) and move the references of receiver ireceiver (0 = all) to the best shift.  shifts: the
 * applied shifts in seconds, one per receiver addressed (disabled receivers: 0). */
int kiwi_autoshift_ref_seismogram(kiwi_ctx* ctx, int ireceiver, float shift_lo, float shift_hi, float* shifts, int cap, int* n);
/* In-memory replacement of output_cross_correlations (minimizer_engine.f90:1283-1306, receiver.f90:597-616; the reference
 * only writes files): cc[component][shift] of receiver ireceiver for the source set by kiwi_set_source_params, shifts
 * nint(shift_lo/dt)..nint(shift_hi/dt) samples. */
int kiwi_get_cross_correlations(kiwi_ctx* ctx, int ireceiver, float shift_lo, float shift_hi, float* cc, int cap, int* ncomp, int* nshift);
/* set_misfit_method (minimizer_engine.f90:620-628) */
int kiwi_set_misfit_method(kiwi_ctx* ctx, int norm_id);
/* set_misfit_taper (minimizer_engine.f90:668-698): ireceiver in 1..n */
int kiwi_set_misfit_taper(kiwi_ctx* ctx, int ireceiver, int n, const float* x, const float* y);
/* set_misfit_filter (minimizer_engine.f90:632-666): ireceiver 0 = all receivers */
int kiwi_set_misfit_filter(kiwi_ctx* ctx, int ireceiver, int n, const float* x, const float* y);
/* set_synthetics_factor (minimizer_engine.f90:700-711) */
int kiwi_set_synthetics_factor(kiwi_ctx* ctx, float factor);
/* set_floating_shiftrange (minimizer_engine.f90:418-451): seconds; ireceiver 0 = all */
int kiwi_set_floating_shiftrange(kiwi_ctx* ctx, int ireceiver, float shift_lo, float shift_hi);

/* Point moment-tensor grid searches (candidates of kiwi_eval_sources that share time, position and rise
 * time and differ only in the tensor; the grids of python/tunguska/gridsearch.py:114-139 over a
 * moment_tensor source): by default such batches are evaluated per (location, receiver) by one kernel that
 * gathers and time-shifts the Green's function components of the location once and contracts the make_weights
 * coefficients of all its tensors against them on the tensor cores (tcgen05), the misfit norm as epilogue.
 * enabled = 0 forces the direct per-candidate path, enabled = 2 the unfused variant (six unit-tensor basis
 * seismograms per location through the general synthesis, then the contraction), which is also what windows
 * too long for the fused kernel fall back to (same results within rounding). */
int kiwi_set_mt_grid(kiwi_ctx* ctx, int enabled);

/* Order of the floating-point operations of the synthesis.  0 (default): the batched kernel, which sums the reference's terms in its own
 * order (taps merged per sample shift, sub-sources spread over warps): within 1e-5 of the trace peak of the exactly accumulated sum at
 * any size, and as far from the reference's result as the reference's own sequential fp32 accumulation is (5e-5 at 1e4 sub-sources).
 * 1: every operation of make_seismogram / trace_multiply_add / gfdb_get_trace_bilin (seismogram.f90:131-289, sparse_trace.f90:597-707,
 * gfdb.f90:865-950) per output sample in the reference's order, with atan2f / sinf / cosf of the azimuths from the host library: the
 * seismograms equal the restated reference path bit for bit (checked on every sample of config C3 at full size).  ~10 x slower: for
 * regression against results of the Fortran code and as the checker of the batched kernel's operands.  Point moment-tensor grids take
 * the candidate-by-candidate path in this mode. */
int kiwi_set_accumulation(kiwi_ctx* ctx, int reference_order);

/* Eikonal / mt_eikonal sources: the fast-marching solve of the rupture front (eikonal.f90:29-199), sequential by construction and 70 % of the
 * discretiser's time, runs per candidate on a host thread or on the device (one warp per candidate, bit-identical results, up to 3848
 * solves side by side, each ~25 x slower than on a host core; a wave lasts as long as its largest grid).  min_batch < 0 (default): the
 * engine shares the solves of a batch between host threads and device where that is faster (the device takes the small grids while the
 * host threads work through the large ones; the split follows a cost model that every batch corrects with what it took, the results do
 * not depend on it); 0: host threads only; k > 0: all solves of batches of k candidates or more on the device. */
int kiwi_set_eikonal_device(kiwi_ctx* ctx, int min_batch);

/* Candidates of one kiwi_eval_sources batch that differ only in the scalar moment (bilateral, eikonal,
 * mt_eikonal: parameter 5) share one synthesis and differ only in the scaling + misfit stage -- the batched
 * form of the reference's `only_moment_changed` shortcut (minimizer_engine.f90:511-521).  On by default;
 * enabled = 0 synthesises every candidate separately (same results). */
int kiwi_set_share_syntheses(kiwi_ctx* ctx, int enabled);

/* number of (misfit, norm-factor) pairs get_misfits returns: components of enabled receivers
 * (minimizer_engine.f90:1141-1148) */
int kiwi_get_nmisfits(kiwi_ctx* ctx);
/* number of parameters of a source type (source_all.f90:97-121), 0 if unknown */
int kiwi_get_n_source_params(int sourcetype);

/* THE batched hot path.  For each of the ns candidates: set_source_params + get_misfits
 * (minimizer_engine.f90:500-523, 876-945, 1130-1172), every candidate evaluated with fresh-state
 * semantics.  params: [ns][nparams] host floats.  misfits: [ns][nmisfits][2] host
 * floats, (misfit, norm factor) interleaved, receiver-major, enabled receivers only.
 * status: [ns] (may be NULL); misfits may be NULL when only kiwi_outer_misfits is wanted.  This replaces the candidate loop of
 * python/tunguska/seismosizer.py:682-722 (make_misfits_for_sources). */
int kiwi_eval_sources(kiwi_ctx* ctx, int sourcetype, int ns, int nparams, const float* params, float* misfits,
                      int* status);
/* same, results left in DEVICE memory (d_misfits: [ns][nmisfits][2] floats on the context's GPU,
 * e.g. an NCCL send buffer); returns after the work is enqueued and finished on the stream */
int kiwi_eval_sources_device(kiwi_ctx* ctx, int sourcetype, int ns, int nparams, const float* params,
                             float* d_misfits, int* status);
/* global misfit per candidate, sqrt(sum m^2)/sqrt(sum n^2) (minimizer_engine.f90:939-942), from a
 * host misfit block as returned by kiwi_eval_sources */
int kiwi_global_misfits(int ns, int nmisfits, const float* misfits, float* global_misfits);

/* Outer misfit and best candidate on the device: make_global_misfits (python/tunguska/seismosizer.py:843-922)
 * followed by nanargmin (gridsearch.py:250-266).  d_misfits: device block [ns][nmisfits][2] as written by
 * kiwi_eval_sources_device, or NULL for the block the last kiwi_eval_sources call left on the GPU (pass
 * misfits = NULL there to skip the download of the cube altogether).  receiver_weights: [nreceivers] (all
 * receivers, disabled ones ignored) or NULL for 1; outer_norm: KIWI_L2NORM or KIWI_L1NORM; anarchy: weights
 * divided by each receiver's norm; bweights: [nboot][nreceivers] bootstrap multiplicities (the caller draws
 * them: numpy bincount of randint, seismosizer.py:871-873) or NULL.  Row 0 of the outputs is the plain result,
 * rows 1..nboot the bootstrap realisations: misfits_by_s [nboot+1][ns] (may be NULL), best / best_value
 * [nboot+1] index and value of the smallest non-NaN misfit (-1 / NaN if none).  All in double like numpy. */
int kiwi_outer_misfits(kiwi_ctx* ctx, int ns, const float* d_misfits, const double* receiver_weights, int outer_norm, int anarchy,
                       int nboot, const double* bweights, double* misfits_by_s, int* best, double* best_value);

/* ns = 1 convenience pair with the reference's names and change detection
 * (minimizer_engine.f90:500-523: identical parameters are a no-op) */
int kiwi_set_source_params(kiwi_ctx* ctx, int sourcetype, int nparams, const float* params);
int kiwi_get_misfits(kiwi_ctx* ctx, float* misfits, int cap_pairs, int* nmisfits);   /* minimizer_engine.f90:1130-1172 */
int kiwi_get_global_misfit(kiwi_ctx* ctx, float* misfit);                            /* minimizer_engine.f90:1083-1093 */

/* ---- ground-motion diagnostics of the synthetics -----------------------------------
 * get_peak_amplitudes (minimizer_engine.f90:1174-1212, differentiate 1 = velocity, 2 = acceleration) and
 * get_arias_intensities (:1214-1245): one value per enabled receiver, over its vertical and/or complete
 * horizontal pair of components (receiver.f90:505-596), for the source set by kiwi_set_source_params;
 * kiwi_eval_ground_motion does all three for a batch: out[ns][enabled receivers][3] = peak velocity, peak
 * acceleration, Arias intensity.  Probe spans as in a fresh process; not available with a misfit filter set. */
int kiwi_get_peak_amplitudes(kiwi_ctx* ctx, int differentiate, float* maxabs, int cap, int* n);
int kiwi_get_arias_intensities(kiwi_ctx* ctx, float* intensities, int cap, int* n);
int kiwi_eval_ground_motion(kiwi_ctx* ctx, int sourcetype, int ns, int nparams, const float* params, float* out, int* status);

/* ---- Levenberg-Marquardt inversion ----------------------------------------
 * The sub-parameter machinery of the parameterised source (parameterized_source.f90:244-309,
 * source_all.f90:377-428) and minimize_lm (minimizer_engine.f90:729-874): MINPACK lmdif on the
 * per-trace misfits of the masked, normalised source parameters.  The n finite-difference sources
 * of every Jacobian (sminpack/fdjac2.f) are evaluated as ONE batch on the GPU; everything else of
 * lmdif/lmpar/qrfac/qrsolv/enorm runs on the host in the reference's single precision.
 * As in the reference the source is left at the LAST model evaluated, `iterations` counts the
 * forward evaluations and `misfit` is the global misfit of that last model. */
int kiwi_set_source_params_mask(kiwi_ctx* ctx, const int* mask, int n);                   /* minimizer_engine.f90:525-543 */
int kiwi_set_source_subparams(kiwi_ctx* ctx, const float* subparams, int n);              /* minimizer_engine.f90:545-566 */
int kiwi_set_source_subparams_limits(kiwi_ctx* ctx, const float* mins, const float* maxs, int n);   /* :580-610 */
int kiwi_get_source_subparams(kiwi_ctx* ctx, float* subparams, int cap, int* n);          /* minimizer_engine.f90:1069-1081 */
int kiwi_minimize_lm(kiwi_ctx* ctx, int* info, int* iterations, float* misfit);           /* minimizer_engine.f90:729-874 */
/* lmdif on a caller-supplied function, Jacobian columns in one batch (also what the known-answer tests drive).
 * fcn(user, ncols, n, m, xs[ncols][n] (may be changed in place), fvecs[ncols][m]) returns the number of leading
 * columns evaluated successfully.  mode/diag/factor/tolerances as in sminpack/lmdif.f; x, fvec, diag are in/out. */
typedef int (*kiwi_lm_fcn)(void* user, int ncols, int n, int m, float* xs, float* fvecs);
int kiwi_lmdif_batched(kiwi_lm_fcn fcn, void* user, int m, int n, float* x, float* fvec, float ftol, float xtol, float gtol, int maxfev,
                       float epsfcn, float* diag, int mode, float factor, int* info, int* nfev);
int kiwi_get_floating_shifts(kiwi_ctx* ctx, int* shifts, int cap, int* n);           /* minimizer_engine.f90:1095-1128, in samples */
/* In-memory replacement for output_seismograms (minimizer_engine.f90:947-1012; the reference only
 * writes files): synthetic trace of the source set by kiwi_set_source_params.
 * which: 0 = raw displacement (receiver%displacement), 1 = scaled (moment, rise time) as put into
 * the syn probe.  first_index: sample index of buf[0] (sample i <-> time (i-1)*dt, receiver.f90:649) */
int kiwi_get_seismogram(kiwi_ctx* ctx, int ireceiver, int icomponent, int which, int* first_index, int* n,
                        float* buf, int cap);

/* get_distances (minimizer_engine.f90:1260-1281; command output_distances): distance [m] and azimuth [rad] of every receiver */
int kiwi_get_distances(kiwi_ctx* ctx, double* distances, double* azimuths, int cap, int* n);
/* get_source_crustal_thickness (minimizer_engine.f90:488-498): crust2x2 thickness [m] under the source location, limited by
 * set_source_crustal_thickness_limit */
int kiwi_get_source_crustal_thickness(kiwi_ctx* ctx, float* thickness);
/* get_principal_axes (minimizer_engine.f90:1248-1258): (azimuth, polar angle) in degrees of the p- and the t-axis of the source set by
 * kiwi_set_source_params (bilateral, circular, eikonal; zeros for the source types without a slip direction) */
int kiwi_get_principal_axes(kiwi_ctx* ctx, float* pax2, float* tax2);

/* In-memory replacements of output_seismograms in all its variants (minimizer_engine.f90:947-1012; probe_get comparator.f90:356-433) and of
 * output_seismogram_spectra (:1014-1067; probe_get_amp_spectrum comparator.f90:332-354) for the source set by kiwi_set_source_params:
 * which_probe 0 synthetics / 1 references; which_processing 0 plain / 1 tapered / 2 filtered.  The probes carry the spans of a fresh
 * process (no misfit has been evaluated before).  Trace: first_index = sample index of buf[0]; spectrum: n amplitudes at k * df. */
int kiwi_get_probe(kiwi_ctx* ctx, int ireceiver, int icomponent, int which_probe, int which_processing, int* first_index, int* n, float* buf, int cap);
int kiwi_get_probe_spectrum(kiwi_ctx* ctx, int ireceiver, int icomponent, int which_probe, int which_processing, float* df, int* n, float* buf, int cap);

/* ---- inspection entry points used by the parity tests (bit-exact integer contract) ---------- */
/* discretise one source on the device; table: [cap][10] floats north east depth time m(6) in the
 * reference's centroid order (discrete_source.f90:27-30); grid3: nx,ny,nt.  Returns through
 * *ncentroids. */
int kiwi_discretize_source(kiwi_ctx* ctx, int sourcetype, int nparams, const float* params, float* table, int cap,
                           int* ncentroids, int* grid3);
/* per-centroid GF indices and sample shifts for receiver ireceiver of the source set by
 * kiwi_set_source_params (gfdb_get_indices[_bilin] gfdb.f90:781-815; floor(time/dt)
 * sparse_trace.f90:640).  near_boundary[i] != 0 flags centroids whose scaled grid coordinate lies
 * within 4 fp32 ulps of an integer (device libm vs glibc may then legitimately differ). */
int kiwi_get_indices(kiwi_ctx* ctx, int ireceiver, int* ix, int* iz, int* its, float* dix, float* diz,
                     int* near_boundary, int cap, int* n);
/* output spans [first,last] of the three internal strips (displacement_ar(1), displacement_ar(2),
 * vertical) for receiver ireceiver (seismogram.f90:131-289) */
int kiwi_get_spans(kiwi_ctx* ctx, int ireceiver, int* spans6);
/* stored span of one GF trace in the HBM slab layout */
int kiwi_trace_span(kiwi_ctx* ctx, int ix, int iz, int ig, int* span2);

/* the fast-marching solver of the eikonal sources (eikonal.f90:29-199, heap.f90) as the host runs it per candidate: arrival times on a
 * grid (nx, ny), ix fastest, of a front starting at initialpoint.  Host only, no GPU needed. */
int kiwi_eikonal_fmm(int nx, int ny, const float* speed, const float* origin2, const float* delta2, const float* initialpoint2, float* times);
/* the same solver on the device, njobs grids at a time (one warp per grid replaying heap.f90 operation by operation: results equal the
 * host solver's bit for bit; csrc/eikonal.cu).  nx, ny: [njobs]; speed, times: [njobs] host arrays of nx*ny; origin2, delta2,
 * initialpoint2: [njobs][2].  *kernel_ms (may be NULL) receives the kernel time. */
int kiwi_eikonal_fmm_device(int njobs, const int* nx, const int* ny, const float* const* speed, const float* origin2, const float* delta2,
                            const float* initialpoint2, float* const* times, float* kernel_ms);

/* ---- measurement support -------------------------------------------------------------------- */
/* algorithmic / logical bytes per evaluation of the last kiwi_eval_sources batch (SURVEY.md
 * section 8d), averaged over the first min(max_candidates, chunk) candidates of its last chunk:
 * b_alg = distinct GF nodes per (candidate, receiver) x components used x stored window length x 4 B
 *         + one write of each synthetic + one read of each reference;
 * b_log = every (centroid, corner, component) trace counted once per use (what the reference streams).
 * nskipped: centroids of the sampled candidates dropped because a GF node was outside the database
 * (seismogram.f90:172). */
int kiwi_last_batch_bytes(kiwi_ctx* ctx, int max_candidates, double* b_alg, double* b_log, int* nsampled,
                          long long* nskipped);
/* device time [ms] of the stages of the last kiwi_eval_sources call, measured with CUDA events on
 * the engine's stream: [0] discretise, [1] geometry/index pre-pass, [2] synthesis, [3] misfit,
 * [4] whole call incl. H2D/D2H; launches[0..3]: kernel launches per stage */
int kiwi_last_timing(kiwi_ctx* ctx, float* ms5, int* launches4);

/* Page-locked host memory for the buffers handed to kiwi_eval_sources (parameters in, misfit block out): the
 * device-to-host copy of a large misfit block (10^5 candidates x 300 traces = 240 MB) then runs at the speed of
 * the bus instead of through the driver's bounce buffers.  Plain malloc'ed buffers keep working. */
void* kiwi_host_alloc(size_t bytes);
void kiwi_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* KIWI_B200_H */
