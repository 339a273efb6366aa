// Launch wrappers of kernels.cu (host-callable) and the host->device parameter blocks.
#pragma once
#include "kiwi_dev.cuh"

// per-candidate parameter block of the bilateral discretiser; the host evaluates everything that
// needs libm (d2r, init_euler euler.f90:28-67) or is sequential and tiny (STF taps
// source_bilat.f90:379-411, m_rot :426-428), the device lays out the sub-fault grid
struct BilatCand {
    float time, north, east, depth;
    float length_a, length_b, width, rupvel;
    float rot_rup[9];   // row-major rotmat_rup
    float mhat[6];      // m_rot/np: (1,1) (2,2) (3,3) (1,2) (1,3) (2,3)
    int nx, ny, nt;
    int group_begin, tap_begin;
    int tt_begin;       // first shift-table entry of the candidate (ngroups x nt entries)
};

void launch_bilat_groups(const BilatCand* d_cands, int ncand, GroupSoA g, TapSoA taps, float dt, int ngroups_total, cudaStream_t st);
void launch_group_tap_range(GroupSoA g, TapSoA taps, float dt, int gbegin, int gend, cudaStream_t st);
void launch_tap_table(GroupSoA g, TapSoA taps, float dt, int ngroups, cudaStream_t st);
void launch_expand_centroids(CandDev cand, GroupSoA g, TapSoA taps, int ngroups_total, float* d_table, int cap, cudaStream_t st);
void launch_geometry(GfdbDev db, const ReceiverDev* rcv, int nrcv, const CandDev* cands, int ncand, GroupSoA g, int ngroups_total,
                     int interpolate, int xunder, int zunder, GeoRec* recs, size_t rec_stride, PairHdr* hdrs, int* tmax, cudaStream_t st,
                     int trig_only = 0 /* records carry cos, sin, sin 2a, cos 2a of the azimuth instead of make_weights (synth_exact.cu) */,
                     float* azf_out = nullptr /* [pair][rec_stride]: the fp32 azimuth of every (pair, group), for the host library's sinf / cosf */);
// reference-order synthesis (synth_exact.cu): every operation of make_seismogram per output sample in the reference's order
int synth_exact_max_samples();
size_t synth_exact_smem_bytes(int wcap, int blk_cap);
cudaError_t launch_synth_exact(GfdbDev db, const ReceiverDev* rcv, int nrcv, const CandDev* cands, int ncand, GroupSoA g, TapSoA taps, int ngroups_total,
                               const GeoRec* recs, size_t rec_stride, const PairHdr* hdrs, int nq_alloc, int margin_q, int interpolate, int xunder,
                               int zunder, int wcap, int blk_cap /* floats of the largest node block, multiple of 4 */, float* seis, size_t seis_stride,
                               SeisHdr* shdrs, int* overflow, cudaStream_t st,
                               const float4* trig = nullptr /* [pair][rec_stride] cos, sin, sin 2a, cos 2a from the host library, or null */);
size_t synth_smem_bytes(int nwarps, int nq);
// point moment-tensor grid search with the basis synthesis fused in (one group per location): recs / hdrs of the probe sources
// [loc][rcv]; *overflow is set where a window or shift table does not fit (the caller then takes the general path)
size_t mt_fused_smem_bytes(int strip_cap, int ncomp);
int mt_fused_max_steps();
cudaError_t launch_mt_fused(GfdbDev db, const ReceiverDev* rcv, int nrcv, const MtLoc* locs, int nloc, const float* mts, const int* cand_of,
                            const GeoRec* recs, const PairHdr* hdrs, const float4* taprec, int strip_cap, int ncomp_max, const float* refdata,
                            const float* taperdata, int method, float dt, float syn_factor, int nmisfits, float* out, int* overflow, cudaStream_t st);
size_t synth_partial_bytes(int nq);   // per (candidate, receiver): running sums of the depth bands (nbands > 1)
cudaError_t launch_synth(GfdbDev db, const ReceiverDev* rcv, int nrcv, const CandDev* cands, int ncand, GroupSoA g, const GeoRec* recs,
                         size_t rec_stride, const PairHdr* hdrs, int nq_alloc, int margin_q, int nwarps, float* seis, size_t seis_stride,
                         SeisHdr* shdrs, int nbands, float* partial, cudaStream_t st);
cudaError_t launch_fold(const ReceiverDev* rcv, int nrcv, const CandDev* cands, int ncand, float* seis, size_t seis_stride, SeisHdr* shdrs,
                        float dt, cudaStream_t st);
void launch_misfit_td(const ReceiverDev* rcv, int nrcv, const CandDev* cands, int ncand, const float* seis, size_t seis_stride,
                      const SeisHdr* shdrs, const float* refdata, const float* taperdata, int method, float dt, float syn_factor,
                      int nmisfits, float* out, int* status, const CandMap* map, cudaStream_t st);
void launch_ground_motion(const ReceiverDev* rcv, int nrcv, const CandDev* cands, int ncand, const float* seis, size_t seis_stride,
                          const SeisHdr* shdrs, const float* taperdata, float dt, float syn_factor, float* out, cudaStream_t st);
size_t misfit_general_smem_bytes(int n_alloc, int nshift_alloc);
cudaError_t launch_misfit_general(const ReceiverDev* rcv, int nrcv, const CandDev* cands, int ncand, const float* seis, size_t seis_stride,
                                  const SeisHdr* shdrs, const float* refdata, const float* taperdata, const float2* tw, int tw_n, int method,
                                  float dt, float syn_factor, int nmisfits, float* out, int* status, int* fshift, int n_alloc,
                                  int nshift_alloc, const CandMap* map, cudaStream_t st, int xs0 = 0, int xs1 = 0, int premethod = 0,
                                  float2* zscratch = nullptr);
cudaError_t launch_probe_export(const ReceiverDev* rcv, int ir, int ic, const CandDev* cands, const float* seis, size_t seis_stride, const SeisHdr* shdrs, int nrcv,
                                const float* refdata, const float* taperdata, const float2* tw, int tw_n, int which_probe, int processing, int spectrum,
                                float dt, int n_alloc, int* hdr, float* out, cudaStream_t st);
void launch_mt_contract(const ReceiverDev* rcv, int nrcv, const MtLoc* locs, int nloc, const float* mts, const int* cand_of, const float* seis,
                        size_t seis_stride, const SeisHdr* shdrs, const float* refdata, const float* taperdata, int method, float dt,
                        float syn_factor, int nmisfits, float* out, cudaStream_t st);
void launch_flag_nonfinite(const float* v, int nrows, int ncols, int* flag, int* count /* may be null */, cudaStream_t st);
cudaError_t launch_outer_misfits(const float* mis, int nm, const void* rc /* {misfit_base, ncomp}[nr] */, int nr, const double* rweights, int l1,
                                 int anarchy, int nrows, const double* bweights, double* out, int ns, int* best, double* bestv, cudaStream_t st,
                                 int row0 = 0 /* rows row0 .. row0 + nrows - 1 of the [1 + nboot][ns] matrix, written to out[0 .. nrows) */);

// ---- fast-marching solver of the eikonal sources on the device (eikonal.cu) ----------------------------------------------------------
struct EikItem { float key; int idx; };
// one solve: T, bp, S are 0-based device arrays of nx*ny (node i of the reference = element i-1)
struct EikJob {
    int nx, ny;
    float dx, dy;
    int ix0, iy0;           // 1-based start node (eikonal_start_node)
    const float* S;
    float* T;
    int* bp;
    EikItem* ovf;           // heap entries beyond eikonal_heap_smem_entries(): room for nx*ny - that many, or null if none are needed
    float invalid_speed;    // > 0: nodes of speed 0 (outside the rupture area, k_eik_speed) are given this speed first (source_eikonal.f90:497-507)
};
// the fine grid of one eikonal candidate as the device sees it (psm_make_eikonal_grid / psm_downsample_grid, source_eikonal.f90:435-601)
struct EikGeom {
    int fnx, fny;
    float first[2], delta[2];
    float shift[3];         // north, east, depth of the source (params 2..4)
    float rot[9];           // rotmat_rup, row-major
    float center[3], radius, relv;
    int ncons;
    float cpoint[4][3], cnormal[4][3];   // half-spaces the rupture area is clipped to (parameterized_source.f90:127-181)
    float thr[5], vs[6];    // crust2x2_get_at_depth as a table (eikonal_layer_table)
    unsigned long long node_off;         // the candidate's nodes in the S / T arenas
    int minspeed_bits;      // written by k_eik_speed: bits of the smallest rupture speed inside the area (0x7f7fffff = none)
    // set by the host once the smallest speed is known
    int nxc, nyc;
    float cdelta[2], invalid_speed;
    unsigned long long coarse_off;       // the candidate's cells in the coarse output arena (6 floats per cell)
};
cudaError_t launch_eik_speed(EikGeom* d_geoms, int ncand, int max_nodes, float* S, cudaStream_t st);
cudaError_t launch_eik_down(const EikGeom* d_geoms, int ncand, int max_cells, const float* S, const float* T, float* coarse, cudaStream_t st);
int eikonal_heap_smem_entries();
int eikonal_wave_jobs(int small_heap);   // solves resident at a time with the 16 KB (0) or the 8 KB (1) heap
void eikonal_start_node(const float origin[2], const float delta[2], const float initialpoint[2], int nx, int ny, int* ix0, int* iy0);
cudaError_t launch_eikonal_fmm(const EikJob* d_jobs, int njobs, cudaStream_t st);
