// Engine context and the C ABI of the hot path (include/kiwi_b200.h).
//
// This file is the B200-native counterpart of the module state and drivers of
// minimizer_engine.f90 (state singletons :78-108, setters :114-711, calculate_seismograms /
// scale_seismograms / calculate_misfits :885-945, get_misfits :1130-1172): it owns the HBM-resident
// database, receivers, reference traces and tapers, prepares candidate sources on the host
// (the few libm calls per candidate, host_math.cpp), and enqueues the kernels of kernels.cu on one
// CUDA stream.  There is no CPU evaluation path: without a CUDA device kiwi_create() fails.
// Citations are file:line of /root/reference.
#include "kiwi_internal.hpp"
#include "kernels.cuh"
#include "host_math.hpp"
#include "lm_host.hpp"
#include <algorithm>
#include <atomic>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <chrono>
#include <cmath>
#include <cstring>
#include <climits>
#include <unordered_set>
#include <unordered_map>
#include <string>
#include <functional>

#define CU_OK(call)                                                                                          \
    do {                                                                                                     \
        cudaError_t e__ = (call);                                                                            \
        if (e__ != cudaSuccess) return kiwi_set_error("CUDA error: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

namespace {

// Host worker threads that live as long as the context: starting a thread inside a process that holds a CUDA context was measured at
// 2-3 ms, which is the whole host preparation of a small batch.  run(count, body) executes body(0..count-1), dealing the items out in
// order (long items first is the caller's business), and returns when all are done; the calling thread works too.
class WorkerPool {
public:
    explicit WorkerPool(int nthreads) {
        for (int t = 0; t < nthreads; t++) threads_.emplace_back([this]() { loop(); });
    }
    ~WorkerPool() {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; }
        cv_.notify_all();
        for (std::thread& t : threads_) t.join();
    }
    int size() const { return (int)threads_.size() + 1; }
    void run(size_t count, const std::function<void(size_t)>& body) {
        if (count == 0) return;
        if (count == 1 || threads_.empty()) { for (size_t k = 0; k < count; k++) body(k); return; }
        {
            std::lock_guard<std::mutex> lk(m_);
            body_ = &body; count_ = count; next_.store(0); pending_ = (int)threads_.size(); generation_++;
        }
        cv_.notify_all();
        for (size_t k = next_.fetch_add(1); k < count; k = next_.fetch_add(1)) body(k);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this]() { return pending_ == 0; });
        body_ = nullptr;
    }
private:
    void loop() {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<void(size_t)>* body; size_t count;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&]() { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_; body = body_; count = count_;
            }
            for (size_t k = next_.fetch_add(1); k < count; k = next_.fetch_add(1)) (*body)(k);
            {
                std::lock_guard<std::mutex> lk(m_);
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(size_t)>* body_ = nullptr;
    size_t count_ = 0;
    std::atomic<size_t> next_{0};
    int pending_ = 0;
    unsigned long long generation_ = 0;
    bool stop_ = false;
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes + bytes / 4 + 256);
        if (e == cudaSuccess) cap = bytes + bytes / 4 + 256;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

// receiver.f90:56 component_names(-5:5) = w s u l c ? a r d n e
int character_to_id(char ch) {
    static const char names[11] = {'w', 's', 'u', 'l', 'c', '?', 'a', 'r', 'd', 'n', 'e'};
    for (int i = -5; i <= 5; i++) if (i != 0 && ch == names[i + 5]) return i;
    return 0;
}

struct HostReceiver {   // t_receiver, receiver.f90:58-99 (host mirror)
    double lat = 0., lon = 0.;   // radians
    float depth = 0.f;
    bool enabled = true;
    int ncomp = 0;
    int comp[KIWI_MAX_COMP] = {0, 0, 0, 0, 0};
    std::vector<float> ref[KIWI_MAX_COMP];
    int ref_ds0[KIWI_MAX_COMP] = {0}, ref_ds1[KIWI_MAX_COMP] = {0};
    bool has_ref[KIWI_MAX_COMP] = {false, false, false, false, false};
    std::vector<float> taper_x, taper_y, filter_x, filter_y;
    int fs0 = 0, fs1 = 0;
};

}  // namespace

struct kiwi_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // [6], [7]: around the fast-marching wave
    // database (set_database)
    bool db_set = false;
    GfdbDev db{};
    DevBuf d_slabs, d_nodes, d_tspan, d_nspan;
    std::vector<NodeInfo> h_nodes;
    std::vector<int2> h_tspan;
    int db_tmin = 0, db_tmax = 0;
    // engine settings
    bool interpolate = false;                 // minimizer_engine.f90:104
    int xunder = 1, zunder = 1;               // :105
    float effective_dt = 1.f;                 // :79
    int misfit_method = KIWI_L2NORM;          // :100
    float syn_factor = 1.f;
    // receivers
    bool receivers_set = false, receivers_dirty = true;
    std::vector<HostReceiver> rcv;
    DevBuf d_rcv, d_refdata, d_taper;
    std::vector<ReceiverDev> h_rcvdev;
    int nmisfits = 0;
    // source location (minimizer_engine.f90:453-467)
    bool loc_set = false;
    double olat = 0., olon = 0., ref_time = 0.;
    // workspace
    DevBuf d_cands, d_bilat, d_gf, d_gi, d_tf, d_recs, d_hdrs, d_seis, d_shdrs, d_out, d_status, d_tmax, d_table, d_tw, d_fshift;
    int tw_n = 0;                            // twiddle table exp(-2 pi i k / tw_n), k < tw_n/2
    std::vector<int> last_fshift;            // floating shifts of the last ns = 1 evaluation
    PinBuf h_stage, h_out, h_mt, h_eik;
    std::unique_ptr<WorkerPool> pool;       // host worker threads (created on first use)
    size_t work_budget = 0;
    kh::Crust2x2 crust;                      // crust2x2 model (minimizer.f90:1669-1674), needed by the eikonal sources
    std::vector<kh::Halfspace> constraints;  // psm%constraints (parameterized_source.f90:127-166)
    bool user_constraints = false;
    float thickness_limit = 0.f;
    std::string prep_error;                  // message of the last failed discretisation
    bool mt_grid_enabled = true;             // point moment-tensor grid searches go through the tcgen05 contraction
    // where the fast-marching solves of a batch of eikonal sources run: -1 (default) = shared between the host threads and the device
    // (prep_eikonal_batch_device), 0 = host threads only, k > 0 = all on the device for batches of k candidates or more.  All on the
    // device loses to 16 host threads (C4, 1024 distinct candidates: 184 against 211 evaluations/s: a wave lasts as long as its largest
    // grid, 2.4 s); sharing gives the device the small grids while the host threads work through the large ones.
    int eikonal_device_min = getenv("KIWI_EIKONAL_DEVICE_MIN") ? atoi(getenv("KIWI_EIKONAL_DEVICE_MIN")) : -1;
    int eikonal_last_device_solves = 0;      // solves of the last batch that ran on the device
    // cost model of the shared fast-marching solves, seconds per fine node: a device solve at up to 13 / at 26 solves per SM (the wave
    // lasts as long as its largest grid), a host core.  Start values measured on a B200 with 16 host cores
    // (profiles/r02_eikonal_device.txt); every shared batch corrects them with what it took (prep_eikonal_batch_device)
    double eik_dev_node[2] = {2.4e-6, [] { const char* e = getenv("KIWI_EIKONAL_DEV_NODE_SMALL"); return e ? atof(e) : 3.2e-6; }()};
    double eik_host_node = 1.05e-7;
    bool eik_adapt = [] { const char* e = getenv("KIWI_EIKONAL_ADAPT"); return !e || atoi(e) != 0; }();
    // kiwi_set_accumulation: synthesis in the reference's order of operations (synth_exact.cu); KIWI_ACCUMULATION=reference makes it the default,
    // for drivers that talk to the command front-end and are not to be touched
    bool accum_reference = getenv("KIWI_ACCUMULATION") && std::string(getenv("KIWI_ACCUMULATION")) == "reference";
    bool mt_grid_fused = true;               // ... with the synthesis fused into it where the windows fit (k_mt_fused)
    DevBuf d_map, d_status_out;
    DevBuf d_taprec;              // shift table of the current batch (k_tap_table)
    DevBuf d_partial;             // running strip sums of the depth bands of k_synth
    DevBuf d_fftz;                // transform buffers of k_misfit_general for spans beyond 16384 samples
    size_t l2_bytes = 0;          // cudaDevAttrL2CacheSize (choose_bands)
    double rcv_dmin = 0., rcv_dmax = 0., rcv_depmin = 0., rcv_depmax = 0.;   // distance / depth range of the enabled receivers (upload_receivers)
    size_t slab_floats = 0;       // floats of the database slabs in HBM
    int db_blk_floats = 0;        // floats of the largest node block (ng rows x window), computed on first use
    DevBuf d_gm;                  // ground-motion values [cand][rcv][3]
    DevBuf d_xcorr;               // cross-correlations [rcv][component][shift] (autoshift_ref_seismogram)
    bool dedup_enabled = true;               // candidates that differ only in the moment share one synthesis
    DevBuf d_azf, d_trig;   // reference-order mode: azimuths of the (pair, group)s and their host-library sinf / cosf
    DevBuf d_eik_s, d_eik_t, d_eik_bp, d_eik_ovf, d_eik_jobs, d_eik_geoms, d_eik_coarse;   // fast-marching solves of a wave of eikonal candidates
    DevBuf d_mtlocs, d_mts, d_candof, d_orc, d_orw, d_obw, d_oout, d_obest, d_obestv;
    int last_eval_ns = 0;                    // candidates whose misfit block sits in d_out (kiwi_eval_sources)
    // description of the last chunk evaluated (inspection entry points, accounting)
    struct Last {
        bool valid = false;
        int sourcetype = 0, n = 0, nrcv = 0;
        size_t rec_stride = 0, seis_stride = 0;
        int ngroups_total = 0;
        std::vector<CandDev> cands;
        GroupSoA g{};
        TapSoA taps{};
        std::vector<float> toff, wt;
        std::vector<int> g0_tap_begin, g0_tap_count;   // taps of the groups of candidate 0
        bool seis_valid = false;
        int syn_lo = 0, syn_hi = 0, tmax = 0;   // bounds of the synthetic spans of the chunk
    } last;
    float ms[5] = {0, 0, 0, 0, 0};
    int launches[4] = {0, 0, 0, 0};
    // ns = 1 state (set_source_params / get_misfits pair)
    bool src_set = false, src_dirty = true;
    int src_type = 0;
    std::vector<float> src_params, src_misfits;
    int src_status = 0;
    // sub-parameter view of the source (psm%params_mask, g_subparam_mins/maxs minimizer_engine.f90:89-90)
    std::vector<char> src_mask;               // empty = all true (source_all.f90:251)
    std::vector<float> sub_mins, sub_maxs;    // empty = no limits
};

namespace {

int require_db(kiwi_ctx* c) { return c->db_set ? 0 : kiwi_set_error("no database set"); }               // minimizer_engine.f90:1346
int require_receivers(kiwi_ctx* c) { return c->receivers_set ? 0 : kiwi_set_error("no receivers set"); }  // :1362

// ---- receivers -> device --------------------------------------------------------------------------
int upload_receivers(kiwi_ctx* c) {
    if (!c->receivers_dirty) return 0;
    const int n = (int)c->rcv.size();
    c->h_rcvdev.assign(n, ReceiverDev());
    std::vector<float> refdata, taperdata;
    int nm = 0;
    const float dt = c->db.dt;
    for (int i = 0; i < n; i++) {
        const HostReceiver& h = c->rcv[i];
        ReceiverDev& r = c->h_rcvdev[i];
        memset(&r, 0, sizeof r);
        if (c->loc_set) {   // seismogram.f90:99-100
            kh::azibazi(c->olat, c->olon, h.lat, h.lon, &r.azi0, &r.bazi0);
            r.dist0 = kh::distance_accurate50m(c->olat, c->olon, h.lat, h.lon);
            kh::final_rotation(r.bazi0, &r.cl0, &r.sl0);
        }
        r.depth = h.depth;
        r.enabled = h.enabled ? 1 : 0;
        r.ncomp = h.ncomp;
        r.misfit_base = nm;
        if (h.enabled) nm += h.ncomp;
        for (int k = 0; k < h.ncomp; k++) {
            const int id = h.comp[k], aid = id < 0 ? -id : id;
            const float sg = id < 0 ? -1.f : 1.f;
            r.comp[k] = id;
            if (aid == 1) { r.ja = k + 1; r.sa = sg; }
            if (aid == 2) { r.jr = k + 1; r.sr = sg; }
            if (aid == 3) { r.jd = k + 1; r.sd = sg; }
            if (aid == 4) { r.jn = k + 1; r.sn = sg; }
            if (aid == 5) { r.je = k + 1; r.se = sg; }
            if (h.has_ref[k]) {
                r.ref_ds0[k] = h.ref_ds0[k]; r.ref_ds1[k] = h.ref_ds1[k];
                kh::initial_probe_span(h.ref_ds0[k], h.ref_ds1[k], &r.ref_sp0[k], &r.ref_sp1[k]);   // comparator.f90:222-271
                r.ref_off[k] = (long long)refdata.size();
                refdata.insert(refdata.end(), h.ref[k].begin(), h.ref[k].end());
                {
                    double ss = 0.;
                    for (float v : h.ref[k]) ss += (double)v * (double)v;
                    const double rms = h.ref[k].empty() ? 0. : std::sqrt(ss / (double)h.ref[k].size());
                    int ex = 0;
                    if (rms > 0. && std::isfinite(rms)) std::frexp(rms, &ex);
                    r.ref_rs[k] = std::ldexp(1.f, -std::min(std::max(ex, -100), 100));
                    double sa = 0.;
                    for (float v : h.ref[k]) sa += (double)fabsf(v);
                    r.ref_ss[k] = ss; r.ref_sa[k] = sa;
                }
            } else {
                r.ref_ds0[k] = 0; r.ref_ds1[k] = -1; r.ref_off[k] = 0; r.ref_rs[k] = 1.f; r.ref_ss[k] = 0.; r.ref_sa[k] = 0.;
            }
        }
        if (!h.taper_x.empty()) {
            std::vector<float> tab;
            r.has_taper = 1;
            kh::taper_table(h.taper_x, h.taper_y, dt, &r.tp0, &r.tp1, &tab);
            kh::discrete_plf_span(h.taper_x, dt, &r.dps0, &r.dps1);
            r.taper_off = (long long)taperdata.size();
            taperdata.insert(taperdata.end(), tab.begin(), tab.end());
        }
        r.has_filter = h.filter_x.empty() ? 0 : 1;
        r.fs0 = h.fs0; r.fs1 = h.fs1;
        if (h.taper_x.size() > KIWI_MAX_PLF || h.filter_x.size() > KIWI_MAX_PLF)
            return kiwi_set_error("tapers and filters are limited to %d points", KIWI_MAX_PLF);
        r.ntp = (int)h.taper_x.size(); r.nfp = (int)h.filter_x.size();
        for (int k = 0; k < r.ntp; k++) { r.tpx[k] = h.taper_x[k]; r.tpy[k] = h.taper_y[k]; }
        for (int k = 0; k < r.nfp; k++) { r.fpx[k] = h.filter_x[k]; r.fpy[k] = h.filter_y[k]; }
    }
    c->nmisfits = nm;
    c->rcv_dmin = c->rcv_depmin = 1e300; c->rcv_dmax = c->rcv_depmax = -1e300;   // (all receivers, enabled or not: see choose_bands)
    for (int i = 0; i < n; i++) {
        c->rcv_dmin = std::min(c->rcv_dmin, c->h_rcvdev[i].dist0); c->rcv_dmax = std::max(c->rcv_dmax, c->h_rcvdev[i].dist0);
        c->rcv_depmin = std::min(c->rcv_depmin, (double)c->h_rcvdev[i].depth); c->rcv_depmax = std::max(c->rcv_depmax, (double)c->h_rcvdev[i].depth);
    }
    if (refdata.empty()) refdata.push_back(0.f);
    if (taperdata.empty()) taperdata.push_back(0.f);
    CU_OK(c->d_rcv.ensure(sizeof(ReceiverDev) * std::max(n, 1)));
    CU_OK(c->d_refdata.ensure(sizeof(float) * refdata.size()));
    CU_OK(c->d_taper.ensure(sizeof(float) * taperdata.size()));
    CU_OK(cudaMemcpyAsync(c->d_rcv.p, c->h_rcvdev.data(), sizeof(ReceiverDev) * n, cudaMemcpyHostToDevice, c->stream));
    CU_OK(cudaMemcpyAsync(c->d_refdata.p, refdata.data(), sizeof(float) * refdata.size(), cudaMemcpyHostToDevice, c->stream));
    CU_OK(cudaMemcpyAsync(c->d_taper.p, taperdata.data(), sizeof(float) * taperdata.size(), cudaMemcpyHostToDevice, c->stream));
    CU_OK(cudaStreamSynchronize(c->stream));
    c->receivers_dirty = false;
    c->last_eval_ns = 0;   // receivers, references, tapers or filters changed: a misfit block kept on the device is stale
    return 0;
}

// references are read for the enabled receivers only (receiver_set_ref_seismogram returns at once for a disabled one, receiver.f90:764),
// so they are required of those only
bool all_refs_set(const kiwi_ctx* c) {
    for (const HostReceiver& h : c->rcv)
        if (h.enabled)
            for (int k = 0; k < h.ncomp; k++) if (!h.has_ref[k]) return false;
    return true;
}

WorkerPool& workers(kiwi_ctx* c) {
    if (!c->pool) c->pool.reset(new WorkerPool((int)std::max(1u, std::thread::hardware_concurrency()) - 1));
    return *c->pool;
}

void eikonal_to_prep(const kh::EikonalPrep& e, kh::SourcePrep* sp) {
    kh::SourcePrep& o = *sp;
    o = kh::SourcePrep();
    o.explicit_groups = true;
    o.nx = e.nx; o.ny = e.ny; o.ngroups = (int)e.groups.size();
    o.moment = e.moment; o.risetime = e.risetime;
    memcpy(o.mhat, e.mhat, sizeof o.mhat);
    o.toff = e.tap_time; o.wt = e.tap_wt;
    o.nt = 0;
    for (const kh::EikonalGroup& g : e.groups) {
        o.g_north.push_back(g.north); o.g_east.push_back(g.east); o.g_depth.push_back(g.depth); o.g_gw.push_back(g.gw);
        o.g_tap_begin.push_back(g.tap_begin); o.g_tap_count.push_back(g.tap_count); o.g_tbase.push_back(0.f);
        o.nt = std::max(o.nt, g.tap_count);
    }
}

int prep_candidate(kiwi_ctx* c, int sourcetype, const float* p, float effective_dt, kh::SourcePrep* sp, std::string* err) {
    if (sourcetype == KIWI_SOURCE_BILATERAL) return kh::prep_bilateral(p, effective_dt, sp) ? 0 : 1;
    if (sourcetype == KIWI_SOURCE_MOMENT_TENSOR) return kh::prep_moment_tensor(p, effective_dt, sp) ? 0 : 1;
    if (sourcetype == KIWI_SOURCE_CIRCULAR) return kh::prep_circular(p, effective_dt, sp) ? 0 : 1;
    if (sourcetype == KIWI_SOURCE_POINT_LP) return kh::prep_point_lp(p, effective_dt, sp) ? 0 : 1;
    if (sourcetype == KIWI_SOURCE_EIKONAL || sourcetype == KIWI_SOURCE_MT_EIKONAL) {
        kh::EikonalPrep e;
        if (!kh::prep_eikonal(p, sourcetype == KIWI_SOURCE_MT_EIKONAL, effective_dt, c->olat, c->olon, c->crust, c->constraints, &e)) {
            if (err) *err = e.err;
            return 1;
        }
        eikonal_to_prep(e, sp);
        return 0;
    }
    return 1;
}

// Eikonal sources of a large batch, optionally (kiwi_set_eikonal_device): the fast-marching solves (70 % of the host discretiser,
// sequential by construction) run on the device, one warp per candidate (csrc/eikonal.cu: the host solver's results bit for bit),
// between the two host parts of the discretiser, which are spread over the host cores.  One solve is ~25 x slower on the device than
// on a host core (2.5 us against 0.1 us per node) but up to 1924 of them run side by side.
// `share`: true = the engine decides which solves go to the device (the small grids, as long as their wave ends before the host threads
// are through with the large ones -- they run at the same time); false = all of them.
// The rest of the fine-grid work of EVERY candidate runs on the device: the speed field before the solve (k_eik_speed) and the
// down-sampling after it (k_eik_down).  A candidate solved on the device sends its geometry up and gets its sub-fault table back; one
// solved on the host gets its speed field through page-locked memory, solves (kh::eikonal_solver_fmm) and sends the times back.
int prep_eikonal_batch_device(kiwi_ctx* c, int sourcetype, int n, int nparams, const float* params, std::vector<kh::SourcePrep>& prep,
                              std::vector<int>& bad, std::vector<std::string>& errs, bool share) {
    const bool mt = sourcetype == KIWI_SOURCE_MT_EIKONAL;
    std::vector<kh::EikonalWork> works(n);
    std::vector<kh::EikonalPrep> eps(n);
    const int ncores = (int)std::max(1u, std::thread::hardware_concurrency());
    auto parallel_over = [&](size_t count, const std::function<void(size_t)>& body) { workers(c).run(count, body); };
    for (int i = 0; i < n; i++) {   // geometry of the fine grids (cheap)
        bad[i] = kh::prep_eikonal_setup(params + (size_t)i * nparams, mt, c->effective_dt, c->olat, c->olon, c->crust, c->constraints, &works[i], &eps[i]) ? 0 : 1;
        if (bad[i]) errs[i] = eps[i].err;
    }
    auto nodes_of = [&](int i) { return (size_t)works[i].fnx * works[i].fny; };
    size_t fr = 0, tot = 0;
    CU_OK(cudaMemGetInfo(&fr, &tot));
    const size_t round_nodes = std::max<size_t>((size_t)1 << 22, std::min<size_t>(fr / 2, (size_t)80 << 30) / 20);   // up to 20 bytes per node on the device
    const int wave_large = eikonal_wave_jobs(0), wave_jobs = eikonal_wave_jobs(1);   // solves resident at a time (16 KB / 8 KB heaps, see eikonal.cu)
    const int hcap = eikonal_heap_smem_entries();
    c->eikonal_last_device_solves = 0;
    std::vector<int> valid;
    for (int i = 0; i < n; i++) if (!bad[i]) valid.push_back(i);
    size_t vat = 0;
    while (vat < valid.size()) {
        // ---- one round: as many candidates as the device arenas hold ------------------------------------------------------
        size_t vend = vat, rnodes = 0;
        while (vend < valid.size() && (vend == vat || rnodes + nodes_of(valid[vend]) <= round_nodes)) rnodes += nodes_of(valid[vend++]);
        std::vector<int> order(valid.begin() + vat, valid.begin() + vend);      // small grids first
        std::sort(order.begin(), order.end(), [&](int a, int b) { return nodes_of(a) < nodes_of(b); });
        const int nr = (int)order.size();
        // which solves go to the device: a prefix of the size-ordered list (its wave lasts as long as its largest grid).  Measured rates
        // (profiles/r02_eikonal_device.txt): a warp 2.1-2.5 us per node, a host core 0.1 us per node for the solve alone
        int ndev = std::min(nr, wave_jobs);
        if (share) {
            const double dev_node_large = c->eik_dev_node[0], dev_node_small = c->eik_dev_node[1], host_node = c->eik_host_node;
            double all_host = 0.;
            for (int i : order) all_host += host_node * nodes_of(i);
            double best = all_host / ncores, dev_host = 0.;
            ndev = 0;
            for (int k = 0; k < nr && k < wave_jobs; k++) {
                const double nn = (double)nodes_of(order[k]);
                dev_host += host_node * nn;
                const double total = std::max((k < wave_large ? dev_node_large : dev_node_small) * nn, (all_host - dev_host) / ncores);   // (nn = the largest grid so far)
                if (total < 0.97 * best) { best = total; ndev = k + 1; }
            }
        }
        // device order: the solves of the device, largest grid first (it sets the wave's length), then the candidates solved on the host
        std::vector<int> cand(nr);
        for (int j = 0; j < ndev; j++) cand[j] = order[ndev - 1 - j];
        for (int j = ndev; j < nr; j++) cand[j] = order[nr - 1 - (j - ndev)];     // host solves: long ones first
        size_t dev_nodes = 0, host_nodes = 0;
        for (int j = 0; j < nr; j++) (j < ndev ? dev_nodes : host_nodes) += nodes_of(cand[j]);
        const size_t all_nodes = dev_nodes + host_nodes;
        CU_OK(c->d_eik_s.ensure(all_nodes * 4)); CU_OK(c->d_eik_t.ensure(all_nodes * 4));
        CU_OK(c->d_eik_bp.ensure(std::max<size_t>(dev_nodes, 1) * 4)); CU_OK(c->d_eik_ovf.ensure(std::max<size_t>(dev_nodes, 1) * sizeof(EikItem)));
        CU_OK(c->d_eik_jobs.ensure(sizeof(EikJob) * std::max(ndev, 1))); CU_OK(c->d_eik_geoms.ensure(sizeof(EikGeom) * nr));
        CU_OK(c->h_eik.ensure(std::max<size_t>(host_nodes, 1) * 8));            // page-locked: speeds, then times of the host solves
        float* h_speed = c->h_eik.as<float>();
        float* h_times = h_speed + host_nodes;
        // ---- geometry up, speed fields, smallest speeds down ------------------------------------------------------------------
        std::vector<EikGeom> geoms(nr);
        size_t off = 0;
        int max_nodes = 1;
        for (int j = 0; j < nr; j++) {
            const kh::EikonalWork& w = works[cand[j]];
            EikGeom& G = geoms[j];
            memset(&G, 0, sizeof G);
            G.fnx = w.fnx; G.fny = w.fny;
            for (int k = 0; k < 2; k++) { G.first[k] = w.first[k]; G.delta[k] = w.delta[k]; }
            for (int k = 0; k < 3; k++) { G.shift[k] = w.p[1 + k]; G.center[k] = w.center[k]; }
            memcpy(G.rot, w.rot_rup, sizeof G.rot);
            G.radius = w.bord_radius; G.relv = w.relv;
            G.ncons = (int)w.constraints->size();
            for (int k = 0; k < G.ncons; k++) for (int q = 0; q < 3; q++) { G.cpoint[k][q] = (*w.constraints)[k].point[q]; G.cnormal[k][q] = (*w.constraints)[k].normal[q]; }
            kh::eikonal_layer_table(w.profile, G.thr, G.vs);
            G.node_off = off; G.minspeed_bits = 0x7f7fffff;
            off += (size_t)w.fnx * w.fny;
            max_nodes = std::max(max_nodes, w.fnx * w.fny);
        }
        CU_OK(cudaMemcpyAsync(c->d_eik_geoms.p, geoms.data(), sizeof(EikGeom) * nr, cudaMemcpyHostToDevice, c->stream));
        cudaError_t e = launch_eik_speed(c->d_eik_geoms.as<EikGeom>(), nr, max_nodes, c->d_eik_s.as<float>(), c->stream);
        if (e != cudaSuccess) return kiwi_set_error("CUDA error launching the speed-field kernel: %s", cudaGetErrorString(e));
        std::vector<EikGeom> back(nr);
        CU_OK(cudaMemcpyAsync(back.data(), c->d_eik_geoms.p, sizeof(EikGeom) * nr, cudaMemcpyDeviceToHost, c->stream));
        if (host_nodes > 0)   // the speed fields of the host solves (they are the tail of the arena)
            CU_OK(cudaMemcpyAsync(h_speed, c->d_eik_s.as<float>() + dev_nodes, host_nodes * 4, cudaMemcpyDeviceToHost, c->stream));
        CU_OK(cudaStreamSynchronize(c->stream));
        // ---- coarse grids; solves of the device and their down-sampling ---------------------------------------------------------------
        std::vector<EikJob> jobs(std::max(ndev, 1));
        std::vector<kh::EikonalCoarse> cgs(nr);
        size_t coff = 0;
        int max_cells = 1;
        for (int j = 0; j < nr; j++) {
            const int i = cand[j];
            kh::EikonalWork& w = works[i];
            EikGeom& G = geoms[j];
            float minspeed;
            memcpy(&minspeed, &back[j].minspeed_bits, 4);
            G.minspeed_bits = back[j].minspeed_bits;
            w.minspeed = minspeed; w.invalid_speed = minspeed * 0.5f;
            bool ok = true;
            if (!(minspeed > 0.f) || back[j].minspeed_bits == 0x7f7fffff) { ok = false; errs[i] = "no valid point in the rupture area"; }
            if (ok && !kh::prep_eikonal_coarse_dims(w, &cgs[j], &errs[i])) ok = false;
            if (!ok) { bad[i] = 1; cgs[j].nxc = cgs[j].nyc = 0; w.invalid_speed = 1.f; }   // (a device solve runs all the same, its result is ignored)
            G.nxc = cgs[j].nxc; G.nyc = cgs[j].nyc; G.cdelta[0] = cgs[j].cdelta[0]; G.cdelta[1] = cgs[j].cdelta[1];
            G.invalid_speed = w.invalid_speed; G.coarse_off = coff;
            coff += (size_t)6 * cgs[j].nxc * cgs[j].nyc;
            max_cells = std::max(max_cells, cgs[j].nxc * cgs[j].nyc);
            if (j < ndev) {
                EikJob& J = jobs[j];
                const size_t nn = (size_t)w.fnx * w.fny;
                J.nx = w.fnx; J.ny = w.fny; J.dx = w.delta[0]; J.dy = w.delta[1];
                eikonal_start_node(w.first, w.delta, w.initialpoint, w.fnx, w.fny, &J.ix0, &J.iy0);
                J.S = c->d_eik_s.as<float>() + G.node_off; J.T = c->d_eik_t.as<float>() + G.node_off; J.bp = c->d_eik_bp.as<int>() + G.node_off;
                J.ovf = nn > (size_t)hcap ? c->d_eik_ovf.as<EikItem>() + G.node_off : nullptr;
                J.invalid_speed = w.invalid_speed;
            }
        }
        CU_OK(c->d_eik_coarse.ensure(sizeof(float) * std::max<size_t>(coff, 6)));
        CU_OK(cudaMemcpyAsync(c->d_eik_geoms.p, geoms.data(), sizeof(EikGeom) * nr, cudaMemcpyHostToDevice, c->stream));
        if (ndev > 0) CU_OK(cudaMemcpyAsync(c->d_eik_jobs.p, jobs.data(), sizeof(EikJob) * ndev, cudaMemcpyHostToDevice, c->stream));
        CU_OK(cudaStreamSynchronize(c->stream));   // (geoms and jobs are stack-lifetime staging vectors)
        if (ndev > 0) {
            cudaEventRecord(c->ev[6], c->stream);
            e = launch_eikonal_fmm(c->d_eik_jobs.as<EikJob>(), ndev, c->stream);
            cudaEventRecord(c->ev[7], c->stream);
            if (e != cudaSuccess) return kiwi_set_error("CUDA error launching the fast-marching solver: %s", cudaGetErrorString(e));
            e = launch_eik_down(c->d_eik_geoms.as<EikGeom>(), ndev, max_cells, c->d_eik_s.as<float>(), c->d_eik_t.as<float>(), c->d_eik_coarse.as<float>(), c->stream);
            if (e != cudaSuccess) return kiwi_set_error("CUDA error launching the down-sampling kernel: %s", cudaGetErrorString(e));
            c->launches[0] += 2;
        }
        c->launches[0] += 1;
        // ---- solves of the host, while the wave runs: speeds and times in page-locked memory ----------------------------------------------
        const auto th0 = std::chrono::steady_clock::now();
        parallel_over((size_t)(nr - ndev), [&](size_t k) {
            const int j = ndev + (int)k, i = cand[j];
            if (bad[i]) return;
            const kh::EikonalWork& w = works[i];
            const size_t nn = (size_t)w.fnx * w.fny, o = geoms[j].node_off - dev_nodes;
            float* sp = h_speed + o;
            for (size_t q = 0; q < nn; q++) if (sp[q] == 0.f) sp[q] = w.invalid_speed;        // source_eikonal.f90:497-507
            kh::eikonal_solver_fmm(sp, w.fnx, w.fny, w.first, w.delta, w.initialpoint, h_times + o);
        });
        const double host_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - th0).count();
        if (nr > ndev) {
            CU_OK(cudaMemcpyAsync(c->d_eik_t.as<float>() + dev_nodes, h_times, host_nodes * 4, cudaMemcpyHostToDevice, c->stream));
            e = launch_eik_down(c->d_eik_geoms.as<EikGeom>() + ndev, nr - ndev, max_cells, c->d_eik_s.as<float>(), c->d_eik_t.as<float>(), c->d_eik_coarse.as<float>(),
                                c->stream);
            if (e != cudaSuccess) return kiwi_set_error("CUDA error launching the down-sampling kernel: %s", cudaGetErrorString(e));
            c->launches[0] += 1;
        }
        std::vector<float> coarse(std::max<size_t>(coff, 6));
        CU_OK(cudaMemcpyAsync(coarse.data(), c->d_eik_coarse.p, sizeof(float) * coff, cudaMemcpyDeviceToHost, c->stream));
        CU_OK(cudaStreamSynchronize(c->stream));
        CU_OK(cudaGetLastError());
        if (share && c->eik_adapt && ndev > 0) {
            // what this round took corrects the cost model of the next one (a grid search sends batch after batch of the same kind):
            // the wave's length over its largest grid, the host threads' time over their nodes
            float wave_ms = 0.f;
            if (cudaEventElapsedTime(&wave_ms, c->ev[6], c->ev[7]) == cudaSuccess && wave_ms > 1.f) {
                const double dn = wave_ms * 1e-3 / (double)nodes_of(cand[0]);
                double& m = c->eik_dev_node[ndev > wave_large ? 1 : 0];
                m = std::min(1e-5, std::max(1e-6, 0.5 * (m + dn)));
            }
            if (nr - ndev >= ncores && host_s > 1e-3) {
                const double hn = host_s * ncores / (double)host_nodes;
                c->eik_host_node = std::min(1e-6, std::max(5e-8, 0.5 * (c->eik_host_node + hn)));
            }
            if (getenv("KIWI_TRACE")) fprintf(stderr, "[kiwi trace] fast-marching round: %d solves on the device (wave %.1f ms, largest grid %zu nodes), %d on %d host threads (%.1f ms); "
                                              "model now %.2f / %.2f us per node on the device, %.0f ns on a host core\n", ndev, wave_ms, nodes_of(cand[0]), nr - ndev, ncores,
                                              host_s * 1e3, c->eik_dev_node[0] * 1e6, c->eik_dev_node[1] * 1e6, c->eik_host_node * 1e9);
        }
        // ---- sub-fault tables ---------------------------------------------------------------------------------------------------------------
        parallel_over((size_t)nr, [&](size_t j) {
            const int i = cand[j];
            if (bad[i]) return;
            kh::EikonalCoarse& cg = cgs[j];
            const size_t ncell = (size_t)cg.nxc * cg.nyc;
            cg.ntimes.resize(ncell); cg.ctimes.resize(ncell); cg.cdur.resize(ncell); cg.cpoints.resize(3 * ncell);
            const float* o = coarse.data() + geoms[j].coarse_off;
            for (size_t k = 0; k < ncell; k++) {
                cg.ntimes[k] = o[6 * k]; cg.ctimes[k] = o[6 * k + 1]; cg.cdur[k] = o[6 * k + 2];
                cg.cpoints[3 * k] = o[6 * k + 3]; cg.cpoints[3 * k + 1] = o[6 * k + 4]; cg.cpoints[3 * k + 2] = o[6 * k + 5];
            }
            if (!kh::prep_eikonal_table(works[i], cg, &eps[i])) { bad[i] = 1; errs[i] = eps[i].err; }
            else eikonal_to_prep(eps[i], &prep[i]);
        });
        c->eikonal_last_device_solves += ndev;
        vat = vend;
    }
    return 0;
}

// Depth bands of k_synth.  A launch over all groups of its candidates gathers from the whole depth range of the sources: for a
// database larger than L2 every (candidate, receiver) pair then streams its node blocks from HBM although the launch as a whole
// touches each of them hundreds of times (2000 receivers x 1470 sub-faults x 4 corners over ~4e4 distinct nodes at config C5).
// Launching band by band -- the sub-faults of a few depth rows at a time -- keeps the slice of the database a launch touches
// resident in L2.  The band count of a candidate is the smallest one whose largest slice (distinct depth rows of the database its
// groups touch x bytes per depth row x fraction of the distance range the receivers span) fits a third of L2 (candidates of a search
// that run side by side differ in depth, so two or three such slices are live at a time), with at least 16 groups per warp and
// band.  It depends on the candidate, the database and the receivers only: a candidate's result (the order of its partial sums)
// does not depend on what else is in the batch.  KIWI_SYNTH_BANDS=n forces n.
int choose_bands(kiwi_ctx* c, const kh::SourcePrep& sp, int nwarps) {
    if (sp.ngroups <= 0) return 1;
    if (const char* ev = getenv("KIWI_SYNTH_BANDS")) { const int v = atoi(ev); if (v > 0) return v; }
    const GfdbDev& db = c->db;
    if (c->l2_bytes == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, c->device) != cudaSuccess || v <= 0) v = 32 << 20;
        c->l2_bytes = (size_t)v;
    }
    const double slab_bytes = (double)c->slab_floats * 4.0;
    if (slab_bytes < 0.8 * (double)c->l2_bytes || db.nz <= 0) return 1;
    const bool lattice = !sp.explicit_groups && sp.ny > 1 && sp.nx * sp.ny == sp.ngroups;
    const int rows = lattice ? sp.ny : sp.ngroups;
    const int maxbands = std::min(rows, sp.ngroups / std::max(1, 16 * nwarps));
    if (maxbands < 2) return 1;
    // depth of every group, horizontal reach of the source
    std::vector<float> depth((size_t)sp.ngroups);
    double rmax = 0.;
    if (sp.explicit_groups) {
        for (int k = 0; k < sp.ngroups; k++) { depth[k] = sp.g_depth[k]; rmax = std::max(rmax, (double)hypotf(sp.g_north[k], sp.g_east[k])); }
    } else if (lattice) {   // source_bilat.f90:349-377 (positions only; exactness is not needed here)
        const float length = sp.p[9] + sp.p[10], width = sp.p[11];
        for (int ix = 0; ix < sp.nx; ix++)
            for (int iy = 0; iy < sp.ny; iy++) {
                const float g0 = (2.f * ix - sp.nx + 1.f) / (2.f * sp.nx) * length, g1 = (2.f * iy - sp.ny + 1.f) / (2.f * sp.ny) * width;
                depth[(size_t)ix * sp.ny + iy] = sp.rot_rup[6] * g0 + sp.rot_rup[7] * g1 + sp.p[3];
                rmax = std::max(rmax, (double)hypotf(sp.rot_rup[0] * g0 + sp.rot_rup[1] * g1 + sp.p[1], sp.rot_rup[3] * g0 + sp.rot_rup[4] * g1 + sp.p[2]));
            }
    } else return 1;
    const double xlo = std::max((double)db.firstx, c->rcv_dmin - rmax), xhi = std::min((double)db.firstx + (double)db.dx * db.nx, c->rcv_dmax + rmax);
    const double xfrac = std::min(1.0, std::max(0.05, (xhi - xlo) / ((double)db.dx * db.nx)));
    const double row_bytes = slab_bytes / db.nz * xfrac;
    const double budget = 0.33 * (double)c->l2_bytes;
    std::vector<char> mark((size_t)db.nz);
    int best = 1;
    for (int nb : {1, 2, 3, 4, 5, 6, 8, 10, 12, 15, 16, 20, 24, 32}) {
        if (nb > maxbands) break;
        best = nb;
        double worst = 0.;
        for (int b = 0; b < nb; b++) {
            std::fill(mark.begin(), mark.end(), 0);
            auto touch = [&](float d) {
                const int z0 = (int)floor(((double)d - c->rcv_depmax - db.firstz) / db.dz), z1 = (int)floor(((double)d - c->rcv_depmin - db.firstz) / db.dz) + 1;
                for (int z = std::max(z0, 0); z <= std::min(z1, db.nz - 1); z++) mark[z] = 1;
            };
            if (lattice) {
                const int r0 = (int)((long long)rows * b / nb), r1 = (int)((long long)rows * (b + 1) / nb);
                for (int iy = r0; iy < r1; iy++) { touch(depth[iy]); touch(depth[(size_t)(sp.nx - 1) * rows + iy]); }
            } else {
                const int k0 = (int)((long long)sp.ngroups * b / nb), k1 = (int)((long long)sp.ngroups * (b + 1) / nb);
                for (int k = k0; k < k1; k++) touch(depth[k]);
            }
            int cnt = 0;
            for (char m : mark) cnt += m;
            worst = std::max(worst, cnt * row_bytes);
        }
        if (worst <= budget) break;
    }
    return best;
}

// called after the synthesis of every sub-chunk when the caller consumes the synthetics itself
// (point moment-tensor grid search): candidates [cand0, cand0+ncand) of the batch, their rows in `seis`
struct SynthHook {
    int align = 1;   // sub-chunks hold a multiple of `align` candidates
    std::function<int(int cand0, int ncand, const float* seis, size_t seis_stride, const SeisHdr* shdrs, const CandDev* d_cands)> fn;
    // fused: every candidate is a single-group probe source (one per grid location); instead of the synthesis launch the hook gets the
    // geometry records and pair headers of candidates [cand0, cand0+ncand) and produces synthetics and misfits itself (k_mt_fused).
    // Returns 0, 1 (error) or 2 (the batch does not fit that kernel: eval_batch returns 2 and the caller takes the general path).
    bool fused = false;
    std::function<int(int cand0, int ncand, const GeoRec* recs, const PairHdr* hdrs, const float4* taprec, int nq)> fused_fn;
};

int eval_mt_grid(kiwi_ctx* c, int n, const float* params, float* d_out, int* h_status, bool* used);

// Candidates that differ only in the scalar moment share one synthesis (minimizer_engine.f90:511-521
// `only_moment_changed`: the reference then reruns only scale_seismograms + calculate_misfits): the pipeline
// runs over the distinct syntheses, the misfit stage over all candidates ("slots", sorted by synthesis).
struct Dedup {
    int n_out = 0;
    std::vector<int> first;       // [nu+1]: slots of synthesis u are [first[u], first[u+1])
    std::vector<int> out_of;      // [n_out]: original candidate of a slot
    std::vector<float> moment;    // [n_out]: its moment
};

// twiddle table of the FFTs: exp(-2 pi i k / N), N = 32768 or the longest transform asked for so far, rounded from double as an
// fp32 FFT library tabulates them
int ensure_twiddles(kiwi_ctx* c, int n_needed = 0) {
    int N = 32768;
    while (N < n_needed) N <<= 1;
    if (c->tw_n >= N) return 0;
    std::vector<float> twh((size_t)N);
    for (int k = 0; k < N / 2; k++) {
        const double a = -2.0 * M_PI * (double)k / (double)N;
        twh[2 * (size_t)k] = (float)cos(a); twh[2 * (size_t)k + 1] = (float)sin(a);
    }
    CU_OK(cudaStreamSynchronize(c->stream));
    CU_OK(c->d_tw.ensure(sizeof(float) * N));
    // on the engine's stream and waited for: a plain cudaMemcpy from pageable memory may return before the DMA has landed, and the
    // non-blocking stream the kernels run on is not ordered against the default stream
    CU_OK(cudaMemcpyAsync(c->d_tw.p, twh.data(), sizeof(float) * N, cudaMemcpyHostToDevice, c->stream));
    CU_OK(cudaStreamSynchronize(c->stream));
    c->tw_n = N;
    return 0;
}

// k_misfit_general on nslots (candidate, receiver-set) slots: sizes the shared-memory FFT from the bound of the padded probe span
// (comparator.f90:1092-1109) over the chunk -- union of all synthetic and (shifted) reference spans, at least twice the longest data
// span, next power of two.  method = KIWI_INTERNAL_XCORR: cross-correlation over the shifts xs0..xs1 (autoshift_ref_seismogram).
int run_misfit_general(kiwi_ctx* c, int method, int xs0, int xs1, int syn_lo, int syn_hi, int tmax, const CandDev* d_cands, int nslots, size_t seis_stride,
                       const SeisHdr* d_shdrs, int nm, float* out_base, int* status_base, int* d_fshift, const CandMap* d_map, int premethod = 0) {
    const int nrcv = (int)c->rcv.size();
    int lo = syn_lo, hi = syn_hi, rlen = 1, nshift = 1;
    const bool xcorr = method == KIWI_INTERNAL_XCORR;
    const bool floating = xcorr ? premethod >= KIWI_FLOATING_L2NORM : c->misfit_method >= KIWI_FLOATING_L2NORM;   // span history in cross-correlation mode
    bool need_fft = (method == KIWI_AMPSPEC_L2NORM || method == KIWI_AMPSPEC_L1NORM);
    if (xcorr) nshift = xs1 - xs0 + 1;
    for (const ReceiverDev& r : c->h_rcvdev) {
        if (!r.enabled) continue;
        if (r.has_filter) need_fft = true;
        if (floating && !xcorr) nshift = std::max(nshift, r.fs1 - r.fs0 + 1);
        int s_lo = 0, s_hi = 0;
        if (floating) { s_lo = std::min(s_lo, r.fs0); s_hi = std::max(s_hi, r.fs1); }
        if (xcorr) { s_lo = std::min(s_lo, xs0); s_hi = std::max(s_hi, xs1); }
        for (int k = 0; k < r.ncomp; k++) {
            lo = std::min(lo, std::min(r.ref_sp0[k], r.ref_ds0[k] + s_lo));
            hi = std::max(hi, std::max(r.ref_sp1[k], r.ref_ds1[k] + s_hi));
            rlen = std::max(rlen, r.ref_ds1[k] - r.ref_ds0[k] + 1);
        }
    }
    if (nshift < 1) return kiwi_set_error("empty shift range");
    int n_alloc = 2;
    if (need_fft) {
        const long long want = std::max<long long>((long long)hi - lo + 1, 2LL * std::max(tmax, rlen));
        while (n_alloc < want) n_alloc <<= 1;
        n_alloc <<= 1;   // head room for re-centred unions
        if (n_alloc > (1 << 22)) return kiwi_set_error("probe span of %d samples is too long", n_alloc);
        if (ensure_twiddles(c, n_alloc)) return 1;
    }
    // transforms of up to 16384 points run in shared memory (128 KiB); longer ones (comparator.f90:1092-1118 puts no bound on the
    // padded span) in a global-memory buffer per CTA
    float2* zscratch = nullptr;
    if (n_alloc > 16384) {
        CU_OK(c->d_fftz.ensure(sizeof(float2) * (size_t)n_alloc * nslots * nrcv));
        zscratch = c->d_fftz.as<float2>();
    }
    if (misfit_general_smem_bytes(zscratch ? 0 : n_alloc, nshift) > (size_t)200 * 1024) return kiwi_set_error("floating shift range too large");
    cudaError_t e = launch_misfit_general(c->d_rcv.as<ReceiverDev>(), nrcv, d_cands, nslots, c->d_seis.as<float>(), seis_stride, d_shdrs,
                                          c->d_refdata.as<float>(), c->d_taper.as<float>(), (const float2*)c->d_tw.p, c->tw_n > 0 ? c->tw_n : 2, method,
                                          c->db.dt, c->syn_factor, nm, out_base, status_base, d_fshift, n_alloc, nshift, d_map, c->stream, xs0, xs1,
                                          premethod, zscratch);
    if (e != cudaSuccess) return kiwi_set_error("CUDA error launching the misfit kernel: %s", cudaGetErrorString(e));
    return 0;
}

// Evaluate candidates [0,n) of `params`; d_out: device [n][nmisfits][2]; h_status: host [n] or null.
// want_misfits = false stops after synthesis (used by the seismogram getters).
int eval_batch(kiwi_ctx* c, int sourcetype, int n, int nparams, const float* params, float* d_out, int* h_status, bool want_misfits,
               const SynthHook* hook = nullptr, const Dedup* dd = nullptr) {
    if (require_db(c)) return 1;
    if (require_receivers(c)) return 1;
    if (!c->loc_set) return kiwi_set_error("no source location set");                   // minimizer_engine.f90:1378
    if (nparams != kiwi_get_n_source_params(sourcetype) || nparams == 0) return kiwi_set_error("wrong number of source parameters or source type not available");
    if (sourcetype < KIWI_SOURCE_BILATERAL || sourcetype > KIWI_SOURCE_MOMENT_TENSOR) return kiwi_set_error("unknown source type");
    if ((sourcetype == KIWI_SOURCE_EIKONAL || sourcetype == KIWI_SOURCE_MT_EIKONAL) && !c->crust.loaded) return kiwi_set_error("crust2x2 model not loaded");
    if (want_misfits && !all_refs_set(c)) return kiwi_set_error("no reference seismograms set");   // :1428
    bool general = false;   // anything beyond plain time-domain norms goes through k_misfit_general
    if (want_misfits) {
        const int m = c->misfit_method;
        general = (m == KIWI_AMPSPEC_L2NORM || m == KIWI_AMPSPEC_L1NORM || m == KIWI_FLOATING_L2NORM || m == KIWI_FLOATING_L1NORM);
        for (const HostReceiver& h : c->rcv) if (h.enabled && !h.filter_x.empty()) general = true;
    }
    if (upload_receivers(c)) return 1;
    if (!hook && want_misfits && !general && sourcetype == KIWI_SOURCE_MOMENT_TENSOR &&
        (c->misfit_method == KIWI_L2NORM || c->misfit_method == KIWI_L1NORM) && c->mt_grid_enabled && !c->accum_reference) {
        bool used = false;
        const int rc = eval_mt_grid(c, n, params, d_out, h_status, &used);
        if (rc || used) return rc;
    }
    if (!hook && !dd && want_misfits && n >= 2 && c->dedup_enabled &&
        (sourcetype == KIWI_SOURCE_BILATERAL || sourcetype == KIWI_SOURCE_EIKONAL || sourcetype == KIWI_SOURCE_MT_EIKONAL ||
         sourcetype == KIWI_SOURCE_CIRCULAR || sourcetype == KIWI_SOURCE_POINT_LP)) {
        // key = all parameters except the moment (index 4 in all three layouts: source_bilat.f90:206, source_eikonal.f90:219-224)
        struct KeyHash { size_t operator()(const std::string& k) const { return std::hash<std::string>()(k); } };
        std::unordered_map<std::string, int, KeyHash> ids;
        std::vector<int> syn_of(n), first_of;
        std::string key((size_t)nparams * 4, '\0');
        for (int i = 0; i < n; i++) {
            memcpy(&key[0], params + (size_t)i * nparams, (size_t)nparams * 4);
            memset(&key[16], 0, 4);
            auto it = ids.find(key);
            if (it == ids.end()) { it = ids.emplace(key, (int)first_of.size()).first; first_of.push_back(i); }
            syn_of[i] = it->second;
        }
        const int nu = (int)first_of.size();
        if (nu < n) {
            Dedup d;
            d.n_out = n;
            d.first.assign(nu + 1, 0);
            for (int i = 0; i < n; i++) d.first[syn_of[i] + 1]++;
            for (int u = 0; u < nu; u++) d.first[u + 1] += d.first[u];
            d.out_of.assign(n, 0); d.moment.assign(n, 0.f);
            std::vector<int> fill(d.first.begin(), d.first.end() - 1);
            for (int i = 0; i < n; i++) { const int j = fill[syn_of[i]]++; d.out_of[j] = i; d.moment[j] = params[(size_t)i * nparams + 4]; }
            // the shared synthesis is made with unit moment (the moment scales the synthetics afterwards, receiver.f90:853-904, slot by
            // slot), so that its validity does not depend on which member of the group comes first in the batch
            std::vector<float> up((size_t)nu * nparams);
            for (int u = 0; u < nu; u++) { memcpy(&up[(size_t)u * nparams], params + (size_t)first_of[u] * nparams, (size_t)nparams * 4); up[(size_t)u * nparams + 4] = 1.f; }
            std::vector<int> ustatus(nu, 0), ostatus(n, 0);
            CU_OK(c->d_status_out.ensure(sizeof(int) * n));
            // (on the engine's stream: it is a non-blocking stream, work on the default stream is not ordered against its kernels)
            CU_OK(cudaMemsetAsync(c->d_status_out.p, 0, sizeof(int) * n, c->stream));
            if (eval_batch(c, sourcetype, nu, nparams, up.data(), d_out, ustatus.data(), true, nullptr, &d)) return 1;
            CU_OK(cudaMemcpy(ostatus.data(), c->d_status_out.p, sizeof(int) * n, cudaMemcpyDeviceToHost));
            // a member whose own parameters the direct path would refuse: circular and point_lp sources check every parameter, the
            // moment included (source_circular / source_point_lp set-up); the others take any moment and report what comes out
            const bool moment_checked = sourcetype == KIWI_SOURCE_CIRCULAR || sourcetype == KIWI_SOURCE_POINT_LP;
            if (h_status) for (int i = 0; i < n; i++) {
                h_status[i] = std::max(ostatus[i], ustatus[syn_of[i]]);
                if (moment_checked && !std::isfinite(params[(size_t)i * nparams + 4])) h_status[i] = KIWI_STATUS_BAD_PARAMS;
            }
            c->last.valid = false;   // the tables describe the distinct syntheses, not the candidates
            return 0;
        }
    }
    const int nrcv = (int)c->rcv.size();
    const int nm = c->nmisfits;
    cudaStream_t st = c->stream;
    for (int i = 0; i < 5; i++) c->ms[i] = 0.f;
    for (int i = 0; i < 4; i++) c->launches[i] = 0;
    if (n == 0) return 0;

    // ---- host preparation of all candidates -------------------------------------------------------
    static const bool trace = getenv("KIWI_TRACE") != nullptr;   // host-side wall-clock split of a batch on stderr
    const auto tw0 = std::chrono::steady_clock::now();
    // the evaluation's clock starts here: source discretisation on the host (the fast-marching solve of the eikonal sources) is part
    // of an evaluation (discretise -> synthesise -> scale -> misfit); the stream is idle, so the event marks this instant
    cudaEventRecord(c->ev[0], st);
    std::vector<kh::SourcePrep> prep(n);
    std::vector<int> bad(n, 0);
    size_t max_groups = 1;
    {
        // the fast-marching discretiser of the eikonal sources costs 0.05-0.3 s per candidate (25 m eikonal grid,
        // source_eikonal.f90:435-517): candidates are independent, so the batch is spread over the host cores
        std::vector<std::string> errs(n);
        auto work = [&](int i) {
            bad[i] = prep_candidate(c, sourcetype, params + (size_t)i * nparams, c->effective_dt, &prep[i], &errs[i]);
            if (bad[i]) prep[i] = kh::SourcePrep();
        };
        const bool heavy = sourcetype == KIWI_SOURCE_EIKONAL || sourcetype == KIWI_SOURCE_MT_EIKONAL;
        const int nthreads = heavy ? (int)std::min<size_t>((size_t)n, std::max(1u, std::thread::hardware_concurrency())) : 1;
        // eikonal_device_min: k > 0 = batches of k candidates or more solve on the device, all of them; -1 = the engine shares the solves
        // of a batch between the host threads and the device where that is faster (the default); 0 = host only
        const int device_min = c->eikonal_device_min;
        c->eikonal_last_device_solves = 0;
        if (heavy && c->constraints.size() <= 4 && ((device_min > 0 && n >= device_min) || (device_min < 0 && n >= 2))) {   // (EikGeom holds four half-spaces)
            if (prep_eikonal_batch_device(c, sourcetype, n, nparams, params, prep, bad, errs, device_min < 0)) return 1;
            for (int i = 0; i < n; i++) if (bad[i]) prep[i] = kh::SourcePrep();
        } else if (nthreads > 1) {
            workers(c).run((size_t)n, [&](size_t i) { work((int)i); });
        } else {
            for (int i = 0; i < n; i++) work(i);
        }
        for (int i = 0; i < n; i++) {
            if (!errs[i].empty()) c->prep_error = errs[i];
            max_groups = std::max(max_groups, (size_t)prep[i].ngroups);
        }
    }
    const auto tw1 = std::chrono::steady_clock::now();
    c->ms[0] += (float)std::chrono::duration<double, std::milli>(tw1 - tw0).count();   // host part of the discretisation stage
    // ---- chunking by workspace budget ---------------------------------------------------------------
    if (c->work_budget == 0) {
        size_t fr = 0, tot = 0;
        CU_OK(cudaMemGetInfo(&fr, &tot));
        c->work_budget = std::min<size_t>(fr / 2, (size_t)64 << 30);
        for (DevBuf* b : {&c->d_recs, &c->d_seis}) c->work_budget += b->cap;   // already ours
    }
    const size_t per_cand_geo = (size_t)nrcv * max_groups * sizeof(GeoRec) + (size_t)nrcv * (sizeof(PairHdr) + KIWI_MAX_COMP * sizeof(SeisHdr));
    int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)n, (c->work_budget / 2) / std::max<size_t>(per_cand_geo, 1)));
    const int align = hook ? std::max(1, hook->align) : 1;
    if (chunk < n) chunk = std::max(align, chunk / align * align);

    for (int b0 = 0; b0 < n; b0 += chunk) {
        const int nc = std::min(chunk, n - b0);
        // ---- candidate / group / tap tables ---------------------------------------------------------
        std::vector<CandDev> cands(nc);
        int G = 0, Tp = 0;
        long long TT = 0;               // entries of the shift table: one per (group, tap)
        std::vector<int> tt_of(nc, 0);
        size_t rec_stride = 1;
        for (int i = 0; i < nc; i++) {
            const kh::SourcePrep& sp = prep[b0 + i];
            CandDev& cd = cands[i];
            cd.group_begin = G; cd.ngroups = sp.ngroups; cd.tap_begin = Tp;
            cd.ntaps_total = (int)sp.toff.size();
            cd.moment = sp.moment; cd.risetime = sp.risetime; cd.nx = sp.nx; cd.ny = sp.ny; cd.nt = sp.nt;
            cd.walk_ny = (!sp.explicit_groups && sp.ny > 1 && sp.nx * sp.ny == sp.ngroups) ? sp.ny : 0; cd.nbands = 1;
            cd.status = bad[b0 + i] ? KIWI_STATUS_BAD_PARAMS : KIWI_STATUS_OK;
            G += sp.ngroups; Tp += (int)sp.toff.size();
            rec_stride = std::max(rec_stride, (size_t)sp.ngroups);
            tt_of[i] = (int)TT;
            if (sp.explicit_groups) for (int k = 0; k < sp.ngroups; k++) TT += sp.g_tap_count[k];
            else TT += (long long)sp.ngroups * sp.nt;
        }
        if (TT > 0x7fffff00LL) return kiwi_set_error("too many (sub-source, time) pairs in one batch");
        const int Galloc = std::max(G, 1), Talloc = std::max(Tp, 1);
        CU_OK(c->d_cands.ensure(sizeof(CandDev) * nc));
        CU_OK(c->d_gf.ensure(sizeof(float) * 12 * (size_t)Galloc));
        CU_OK(c->d_gi.ensure(sizeof(int) * 6 * (size_t)Galloc));
        CU_OK(c->d_taprec.ensure(sizeof(float4) * 2 * ((size_t)TT + 1)));
        CU_OK(c->d_tf.ensure(sizeof(float) * 2 * (size_t)Talloc));
        GroupSoA g;
        float* gf = c->d_gf.as<float>();
        g.north = gf; g.east = gf + Galloc; g.depth = gf + 2 * (size_t)Galloc; g.tbase = gf + 3 * (size_t)Galloc; g.mhat = gf + 4 * (size_t)Galloc;
        g.gw = gf + 10 * (size_t)Galloc;
        g.lam = nullptr;
        int* gi = c->d_gi.as<int>();
        g.tap_begin = gi; g.tap_count = gi + Galloc; g.its_min = gi + 2 * (size_t)Galloc; g.its_max = gi + 3 * (size_t)Galloc;
        g.tt_begin = gi + 4 * (size_t)Galloc; g.nstep = gi + 5 * (size_t)Galloc;
        g.taprec = c->d_taprec.as<float4>();
        TapSoA taps; taps.toff = c->d_tf.as<float>(); taps.wt = taps.toff + Talloc;
        std::vector<float> toff(Talloc, 0.f), wt(Talloc, 0.f);
        for (int i = 0; i < nc; i++) {
            const kh::SourcePrep& sp = prep[b0 + i];
            std::copy(sp.toff.begin(), sp.toff.end(), toff.begin() + cands[i].tap_begin);
            std::copy(sp.wt.begin(), sp.wt.end(), wt.begin() + cands[i].tap_begin);
        }
        CU_OK(cudaMemcpyAsync(c->d_cands.p, cands.data(), sizeof(CandDev) * nc, cudaMemcpyHostToDevice, st));
        CU_OK(cudaMemcpyAsync(taps.toff, toff.data(), sizeof(float) * Talloc, cudaMemcpyHostToDevice, st));
        CU_OK(cudaMemcpyAsync(taps.wt, wt.data(), sizeof(float) * Talloc, cudaMemcpyHostToDevice, st));

        cudaEventRecord(c->ev[1], st);
        // ---- K1: sub-source groups ------------------------------------------------------------------
        if (sourcetype == KIWI_SOURCE_BILATERAL) {
            std::vector<BilatCand> bc(nc);
            for (int i = 0; i < nc; i++) {
                const kh::SourcePrep& sp = prep[b0 + i];
                BilatCand& b = bc[i];
                memset(&b, 0, sizeof b);
                b.time = sp.p[0]; b.north = sp.p[1]; b.east = sp.p[2]; b.depth = sp.p[3];
                b.length_a = sp.p[9]; b.length_b = sp.p[10]; b.width = sp.p[11]; b.rupvel = sp.p[12];
                memcpy(b.rot_rup, sp.rot_rup, sizeof b.rot_rup);
                memcpy(b.mhat, sp.mhat, sizeof b.mhat);
                b.nx = sp.ngroups ? sp.nx : 0; b.ny = sp.ngroups ? sp.ny : 0; b.nt = sp.nt;
                b.group_begin = cands[i].group_begin; b.tap_begin = cands[i].tap_begin;
                b.tt_begin = tt_of[i];
            }
            CU_OK(c->d_bilat.ensure(sizeof(BilatCand) * nc));
            CU_OK(cudaMemcpyAsync(c->d_bilat.p, bc.data(), sizeof(BilatCand) * nc, cudaMemcpyHostToDevice, st));
            CU_OK(cudaStreamSynchronize(st));   // bc is a stack-lifetime staging vector
            launch_bilat_groups(c->d_bilat.as<BilatCand>(), nc, g, taps, c->db.dt, Galloc, st);
            c->launches[0] += 1;
        } else {   // groups defined on the host: the single group of a point moment tensor
                   // (source_moment_tensor.f90:256-263), the sub-faults of an eikonal source (source_eikonal.f90:684-707)
            std::vector<float> hf((size_t)12 * Galloc, 0.f);
            std::vector<int> hi((size_t)6 * Galloc, 0);
            for (int i = 0; i < nc; i++) {
                const kh::SourcePrep& sp = prep[b0 + i];
                if (sp.ngroups == 0) continue;
                const int gi0 = cands[i].group_begin;
                int tt = tt_of[i];
                for (int k = 0; k < sp.ngroups; k++) {
                    const size_t gi = (size_t)gi0 + k;
                    hi[4 * (size_t)Galloc + gi] = tt;
                    tt += sp.explicit_groups ? sp.g_tap_count[k] : sp.nt;
                    if (sp.explicit_groups) {
                        hf[gi] = sp.g_north[k]; hf[(size_t)Galloc + gi] = sp.g_east[k]; hf[2 * (size_t)Galloc + gi] = sp.g_depth[k];
                        hf[3 * (size_t)Galloc + gi] = sp.g_tbase[k];             // 0 where the taps carry the complete centroid time
                        hf[10 * (size_t)Galloc + gi] = sp.g_gw[k];
                        hi[gi] = cands[i].tap_begin + sp.g_tap_begin[k]; hi[(size_t)Galloc + gi] = sp.g_tap_count[k];
                    } else {
                        hf[gi] = sp.point[0]; hf[(size_t)Galloc + gi] = sp.point[1]; hf[2 * (size_t)Galloc + gi] = sp.point[2];
                        hf[3 * (size_t)Galloc + gi] = sp.time;
                        hf[10 * (size_t)Galloc + gi] = 1.f;
                        hi[gi] = cands[i].tap_begin; hi[(size_t)Galloc + gi] = sp.nt;
                    }
                    for (int q = 0; q < 6; q++) hf[(4 + q) * (size_t)Galloc + gi] = sp.mhat[q];
                    hf[11 * (size_t)Galloc + gi] = atan2f(hf[(size_t)Galloc + gi], hf[gi]);   // the host library's atan2f(east, north), see below
                }
            }
            g.lam = gf + 11 * (size_t)Galloc;
            CU_OK(cudaMemcpyAsync(c->d_gf.p, hf.data(), sizeof(float) * hf.size(), cudaMemcpyHostToDevice, st));
            CU_OK(cudaMemcpyAsync(c->d_gi.p, hi.data(), sizeof(int) * hi.size(), cudaMemcpyHostToDevice, st));
            CU_OK(cudaStreamSynchronize(st));
        }
        launch_tap_table(g, taps, c->db.dt, G, st);   // also the groups' sample-shift ranges
        c->launches[0] += 1;
        cudaEventRecord(c->ev[2], st);
        // ---- K2: geometry + indices + spans ---------------------------------------------------------
        const size_t npairs = (size_t)nc * nrcv;
        CU_OK(c->d_recs.ensure(sizeof(GeoRec) * npairs * rec_stride));
        CU_OK(c->d_hdrs.ensure(sizeof(PairHdr) * npairs));
        const bool fused = hook && hook->fused;
        CU_OK(c->d_shdrs.ensure(sizeof(SeisHdr) * npairs * KIWI_MAX_COMP));
        CU_OK(c->d_tmax.ensure(sizeof(int) * 4));
        {
            static const int init3[4] = {0, INT_MAX, INT_MIN, 0};   // max window length, min first sample, max last sample
            CU_OK(cudaMemcpyAsync(c->d_tmax.p, init3, sizeof init3, cudaMemcpyHostToDevice, st));
        }
        const bool exact = c->accum_reference && !hook;
        if (!g.lam) {   // atan2f of the sub-source positions from the host library (see approx_differential_azidist in kernels.cu); sources whose
                        // sub-sources were laid out on the host have it already
            std::vector<float> ne((size_t)2 * Galloc), lam((size_t)Galloc, 0.f);
            CU_OK(cudaMemcpyAsync(ne.data(), g.north, sizeof(float) * 2 * (size_t)Galloc, cudaMemcpyDeviceToHost, st));   // north, east: adjacent
            CU_OK(cudaStreamSynchronize(st));
            // (a fifth of a millisecond per 1e4 sub-sources on one core: spread over the context's worker threads from a few thousand on)
            if (G >= 4096) {
                const size_t nblk = ((size_t)G + 2047) / 2048;
                workers(c).run(nblk, [&](size_t b) { for (int k = (int)(b * 2048); k < std::min(G, (int)((b + 1) * 2048)); k++) lam[k] = atan2f(ne[(size_t)Galloc + k], ne[k]); });
            } else {
                for (int k = 0; k < G; k++) lam[k] = atan2f(ne[(size_t)Galloc + k], ne[k]);
            }
            g.lam = gf + 11 * (size_t)Galloc;
            CU_OK(cudaMemcpyAsync(g.lam, lam.data(), sizeof(float) * (size_t)Galloc, cudaMemcpyHostToDevice, st));
            CU_OK(cudaStreamSynchronize(st));
        }
        launch_geometry(c->db, c->d_rcv.as<ReceiverDev>(), nrcv, c->d_cands.as<CandDev>(), nc, g, Galloc, c->interpolate ? 1 : 0, c->xunder, c->zunder,
                        c->d_recs.as<GeoRec>(), rec_stride, c->d_hdrs.as<PairHdr>(), c->d_tmax.as<int>(), st, exact ? 1 : 0,
                        exact ? (c->d_azf.ensure(sizeof(float) * npairs * rec_stride) == cudaSuccess ? c->d_azf.as<float>() : nullptr) : nullptr);
        if (exact && c->d_azf.p) {
            // sinf / cosf of the azimuths from the host library (make_weights seismogram.f90:316-336 calls them per centroid): the device's own
            // are an ulp off often enough to show in a few samples per trace
            const size_t na = npairs * rec_stride;
            std::vector<float> azf(na);
            std::vector<float4> trig(na);
            CU_OK(cudaMemcpyAsync(azf.data(), c->d_azf.p, sizeof(float) * na, cudaMemcpyDeviceToHost, st));
            CU_OK(cudaStreamSynchronize(st));
            const size_t nblk = (na + 16383) / 16384;
            workers(c).run(nblk, [&](size_t b) {
                for (size_t k = b * 16384; k < std::min(na, (b + 1) * 16384); k++) {
                    const float a = azf[k];
                    trig[k] = make_float4(cosf(a), sinf(a), sinf(2.f * a), cosf(2.f * a));
                }
            });
            CU_OK(c->d_trig.ensure(sizeof(float4) * na));
            CU_OK(cudaMemcpyAsync(c->d_trig.p, trig.data(), sizeof(float4) * na, cudaMemcpyHostToDevice, st));
            CU_OK(cudaStreamSynchronize(st));
        }
        c->launches[1] += 1;
        int tm3[4] = {0, 0, 0, 0};
        CU_OK(cudaMemcpyAsync(tm3, c->d_tmax.p, sizeof tm3, cudaMemcpyDeviceToHost, st));
        cudaEventRecord(c->ev[3], st);
        CU_OK(cudaStreamSynchronize(st));
        CU_OK(cudaGetLastError());
        float ms01 = 0.f, ms12 = 0.f, ms23 = 0.f;
        cudaEventElapsedTime(&ms12, c->ev[1], c->ev[2]);
        cudaEventElapsedTime(&ms23, c->ev[2], c->ev[3]);
        (void)ms01;
        c->ms[0] += ms12; c->ms[1] += ms23;
        // ---- K3 + K5: synthesis and misfit, in sub-chunks sized by the seismogram buffer ----------------
        const int tmax = tm3[0];
        // rise-time fold (receiver.f90:853-904): the folded strips grow by about half the boxcar on either side
        float max_rise = 0.f;
        for (int i = 0; i < nc; i++) max_rise = std::max(max_rise, cands[i].risetime);
        int margin_q = 0;
        if (max_rise > 0.f) {
            const int nshifts = 1 + 2 * (int)lroundf(0.5f * max_rise / c->db.dt);
            margin_q = ((nshifts + 1) / 2 + 2 + 3) / 4 + 1;
        }
        const int nq = (tmax + 6) / 4 + 1 + 2 * margin_q;
        const size_t seis_stride = (size_t)4 * nq;
        int nwarps = (int)std::min<size_t>(8, std::max<size_t>(1, rec_stride));   // no more warps than groups per candidate
        while (nwarps > 1 && synth_smem_bytes(nwarps, nq) > (size_t)112 * 1024) nwarps--;   // two CTAs per SM
        if (synth_smem_bytes(nwarps, nq) > (size_t)220 * 1024)
            return kiwi_set_error("synthetic window of %d samples does not fit the shared-memory accumulators", tmax);
        int nbands = 1;   // launches: the largest band count of a candidate of this chunk
        {
            bool changed = false;
            for (int i = 0; i < nc; i++) {
                const int nb = cands[i].status == KIWI_STATUS_OK ? choose_bands(c, prep[b0 + i], 8) : 1;
                if (nb != cands[i].nbands) { cands[i].nbands = nb; changed = true; }
                nbands = std::max(nbands, nb);
            }
            if (changed) CU_OK(cudaMemcpyAsync(c->d_cands.p, cands.data(), sizeof(CandDev) * nc, cudaMemcpyHostToDevice, st));
        }
        if (fused && (rec_stride != 1 || max_rise > 0.f)) return 2;
        const size_t per_cand_seis = fused ? 1 : (size_t)nrcv * KIWI_MAX_COMP * seis_stride * sizeof(float);
        int sub = (int)std::max<size_t>(1, std::min<size_t>((size_t)nc, (c->work_budget / 2) / std::max<size_t>(per_cand_seis, 1)));
        if (sub < nc) sub = std::max(align, sub / align * align);
        CU_OK(c->d_seis.ensure(per_cand_seis * sub));
        CU_OK(c->d_status.ensure(sizeof(int) * nc));
        {
            std::vector<int> stt(nc);
            for (int i = 0; i < nc; i++) stt[i] = cands[i].status;
            CU_OK(cudaMemcpyAsync(c->d_status.p, stt.data(), sizeof(int) * nc, cudaMemcpyHostToDevice, st));
            CU_OK(cudaStreamSynchronize(st));
        }
        for (int s0 = 0; s0 < nc; s0 += sub) {
            const int ns_ = std::min(sub, nc - s0);
            const size_t poff = (size_t)s0 * nrcv;
            cudaEventRecord(c->ev[3], st);
            SeisHdr* shdrs_sub = c->d_shdrs.as<SeisHdr>() + poff * KIWI_MAX_COMP;
            if (fused) {
                cudaEventRecord(c->ev[4], st);   // (no synthesis stage of its own: the hook's kernel is booked as the misfit stage)
                const int rc = hook->fused_fn(b0 + s0, ns_, c->d_recs.as<GeoRec>() + poff, c->d_hdrs.as<PairHdr>() + poff, g.taprec, nq);
                if (rc) return rc;
                c->launches[3] += 1;
                cudaEventRecord(c->ev[5], st);
                CU_OK(cudaStreamSynchronize(st));
                CU_OK(cudaGetLastError());
                float b = 0.f;
                cudaEventElapsedTime(&b, c->ev[4], c->ev[5]);
                c->ms[3] += b;
                continue;
            }
            if (tmax > 0 && exact) {
                // reference-order synthesis: one thread per output sample walks the centroids one after the other
                const int wcap = 4 * nq + 32;
                if (c->db_blk_floats == 0) for (const NodeInfo& ni : c->h_nodes) if (ni.off != ~0ull) c->db_blk_floats = std::max(c->db_blk_floats, ni.wn * c->db.ng);
                const int blk_cap = (c->db_blk_floats + 3) & ~3;     // the largest node block of the database
                if (4 * nq > synth_exact_max_samples() || synth_exact_smem_bytes(wcap, blk_cap) > (size_t)220 * 1024)
                    return kiwi_set_error("synthetic window of %d samples is too long for the reference-order synthesis", tmax);
                CU_OK(c->d_tmax.ensure(sizeof(int) * 8));
                int* d_overflow = c->d_tmax.as<int>() + 5;
                CU_OK(cudaMemsetAsync(d_overflow, 0, sizeof(int), st));
                cudaError_t e = launch_synth_exact(c->db, c->d_rcv.as<ReceiverDev>(), nrcv, c->d_cands.as<CandDev>() + s0, ns_, g, taps, Galloc,
                                                   c->d_recs.as<GeoRec>() + poff * rec_stride, rec_stride, c->d_hdrs.as<PairHdr>() + poff, nq, margin_q,
                                                   c->interpolate ? 1 : 0, c->xunder, c->zunder, wcap, blk_cap, c->d_seis.as<float>(), seis_stride,
                                                   c->d_shdrs.as<SeisHdr>() + poff * KIWI_MAX_COMP, d_overflow, st,
                                                   c->d_trig.p ? c->d_trig.as<float4>() + poff * rec_stride : nullptr);
                if (e != cudaSuccess) return kiwi_set_error("CUDA error launching the reference-order synthesis: %s", cudaGetErrorString(e));
                c->launches[2] += 1;
                int overflow = 0;
                CU_OK(cudaMemcpyAsync(&overflow, d_overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
                CU_OK(cudaStreamSynchronize(st));
                if (overflow) return kiwi_set_error("a Green's function window is too long for the reference-order synthesis");
                if (max_rise > 0.f) {
                    e = launch_fold(c->d_rcv.as<ReceiverDev>(), nrcv, c->d_cands.as<CandDev>() + s0, ns_, c->d_seis.as<float>(), seis_stride,
                                    c->d_shdrs.as<SeisHdr>() + poff * KIWI_MAX_COMP, c->db.dt, st);
                    if (e != cudaSuccess) return kiwi_set_error("CUDA error launching the rise-time fold: %s", cudaGetErrorString(e));
                    c->launches[2] += 1;
                }
            } else if (tmax > 0) {
                if (nbands > 1) CU_OK(c->d_partial.ensure(synth_partial_bytes(nq) * (size_t)ns_ * nrcv));
                cudaError_t e = launch_synth(c->db, c->d_rcv.as<ReceiverDev>(), nrcv, c->d_cands.as<CandDev>() + s0, ns_, g,
                                             c->d_recs.as<GeoRec>() + poff * rec_stride, rec_stride, c->d_hdrs.as<PairHdr>() + poff, nq, margin_q,
                                             nwarps, c->d_seis.as<float>(), seis_stride, c->d_shdrs.as<SeisHdr>() + poff * KIWI_MAX_COMP, nbands,
                                             c->d_partial.as<float>(), st);
                if (e != cudaSuccess) return kiwi_set_error("CUDA error launching synthesis: %s", cudaGetErrorString(e));
                c->launches[2] += nbands;   // one launch of k_synth per depth band
                if (max_rise > 0.f) {
                    e = launch_fold(c->d_rcv.as<ReceiverDev>(), nrcv, c->d_cands.as<CandDev>() + s0, ns_, c->d_seis.as<float>(), seis_stride,
                                    c->d_shdrs.as<SeisHdr>() + poff * KIWI_MAX_COMP, c->db.dt, st);
                    if (e != cudaSuccess) return kiwi_set_error("CUDA error launching the rise-time fold: %s", cudaGetErrorString(e));
                    c->launches[2] += 1;
                }
            } else {
                CU_OK(cudaMemsetAsync(shdrs_sub, 0xff, sizeof(SeisHdr) * (size_t)ns_ * nrcv * KIWI_MAX_COMP, st));
            }
            cudaEventRecord(c->ev[4], st);
            if (hook && hook->fn) {
                if (hook->fn(b0 + s0, ns_, c->d_seis.as<float>(), seis_stride, shdrs_sub, c->d_cands.as<CandDev>() + s0)) return 1;
                c->launches[3] += 1;
            }
            // misfit stage: one slot per candidate; with shared syntheses the slots of this sub-chunk come with a map
            int nslots = ns_;
            const CandMap* d_map = nullptr;
            float* out_base = d_out + ((size_t)(b0 + s0) * nm) * 2;
            int* status_base = c->d_status.as<int>() + s0;
            size_t fshift_off = poff;
            if (dd && want_misfits) {
                const int j0 = dd->first[b0 + s0], j1 = dd->first[b0 + s0 + ns_];
                nslots = j1 - j0;
                std::vector<CandMap> hm(std::max(nslots, 1));
                for (int j = j0; j < j1; j++) {
                    int u = (int)(std::upper_bound(dd->first.begin(), dd->first.end(), j) - dd->first.begin()) - 1;
                    hm[j - j0] = CandMap{dd->out_of[j], u - (b0 + s0), dd->moment[j], 0};
                }
                CU_OK(c->d_map.ensure(sizeof(CandMap) * hm.size()));
                CU_OK(cudaMemcpyAsync(c->d_map.p, hm.data(), sizeof(CandMap) * hm.size(), cudaMemcpyHostToDevice, st));
                CU_OK(cudaStreamSynchronize(st));
                d_map = c->d_map.as<CandMap>();
                out_base = d_out; status_base = c->d_status_out.as<int>(); fshift_off = 0;
            }
            if (want_misfits && nm > 0 && !general) {
                launch_misfit_td(c->d_rcv.as<ReceiverDev>(), nrcv, c->d_cands.as<CandDev>() + s0, nslots, c->d_seis.as<float>(), seis_stride,
                                 c->d_shdrs.as<SeisHdr>() + poff * KIWI_MAX_COMP, c->d_refdata.as<float>(), c->d_taper.as<float>(),
                                 c->misfit_method, c->db.dt, c->syn_factor, nm, out_base, status_base, d_map, st);
                c->launches[3] += 1;
            } else if (want_misfits && nm > 0) {
                CU_OK(c->d_fshift.ensure(sizeof(int) * (size_t)(dd ? dd->n_out : nc) * nrcv));
                if (run_misfit_general(c, c->misfit_method, 0, 0, tm3[1], tm3[2], tmax, c->d_cands.as<CandDev>() + s0, nslots, seis_stride,
                                       c->d_shdrs.as<SeisHdr>() + poff * KIWI_MAX_COMP, nm, out_base, status_base, c->d_fshift.as<int>() + fshift_off, d_map))
                    return 1;
                c->launches[3] += 1;
            }
            cudaEventRecord(c->ev[5], st);
            CU_OK(cudaStreamSynchronize(st));
            CU_OK(cudaGetLastError());
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, c->ev[3], c->ev[4]);
            cudaEventElapsedTime(&b, c->ev[4], c->ev[5]);
            c->ms[2] += a; c->ms[3] += b;
        }
        if (h_status) {
            CU_OK(cudaMemcpy(h_status + b0, c->d_status.p, sizeof(int) * nc, cudaMemcpyDeviceToHost));
        }
        // remember this chunk for the inspection entry points
        kiwi_ctx::Last& L = c->last;
        L.valid = true; L.sourcetype = sourcetype; L.n = nc; L.nrcv = nrcv; L.rec_stride = rec_stride; L.seis_stride = seis_stride;
        L.ngroups_total = Galloc; L.cands = cands; L.g = g; L.taps = taps; L.toff = toff; L.wt = wt;
        L.seis_valid = (sub >= nc) && tmax > 0 && !fused;
        L.syn_lo = tm3[1]; L.syn_hi = tm3[2]; L.tmax = tmax;
        {
            const kh::SourcePrep& sp0 = prep[b0];
            L.g0_tap_begin.assign(sp0.ngroups, cands[0].tap_begin); L.g0_tap_count.assign(sp0.ngroups, sp0.nt);
            if (sp0.explicit_groups) for (int k = 0; k < sp0.ngroups; k++) { L.g0_tap_begin[k] = cands[0].tap_begin + sp0.g_tap_begin[k]; L.g0_tap_count[k] = sp0.g_tap_count[k]; }
        }
    }
    cudaEventRecord(c->ev[1], st);
    CU_OK(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&c->ms[4], c->ev[0], c->ev[1]);
    if (trace) {
        const auto tw2 = std::chrono::steady_clock::now();
        fprintf(stderr, "[kiwi trace] batch of %d: host prep %.3f ms (%d fast-marching solves on the device), rest (uploads, launches, device) %.3f ms, device stages %.3f ms\n", n,
                std::chrono::duration<double, std::milli>(tw1 - tw0).count(), c->eikonal_last_device_solves,
                std::chrono::duration<double, std::milli>(tw2 - tw1).count(), c->ms[0] + c->ms[1] + c->ms[2] + c->ms[3]);
    }
    return 0;
}

typedef MtLoc MtLocHost;

// Point moment-tensor grid search (python/tunguska/gridsearch.py:159-197 over source.py:119-164 grids of a
// moment_tensor source): candidates that share (time, north, east, depth, rise-time) differ only in the six
// tensor components, in which the synthetics are linear.  Per distinct location the six unit-tensor basis
// seismograms are synthesised by the normal path; all candidates of the location are then contracted against
// them on the tensor cores with the misfit as epilogue (k_mt_contract).  *used = false: not worth it / not a grid.
int eval_mt_grid(kiwi_ctx* c, int n, const float* params, float* d_out, int* h_status, bool* used) {
    *used = false;
    if (n < 64) return 0;
    static const bool trace = getenv("KIWI_TRACE") != nullptr;
    const auto tg0 = std::chrono::steady_clock::now();
    struct Key { unsigned v[5]; bool operator==(const Key& o) const { return memcmp(v, o.v, sizeof v) == 0; } };
    struct KeyHash { size_t operator()(const Key& k) const { size_t h = 1469598103934665603ull; for (unsigned x : k.v) { h ^= x; h *= 1099511628211ull; } return h; } };
    std::unordered_map<Key, int, KeyHash> ids;
    std::vector<int> loc_of(n);
    std::vector<int> first_of;
    // grids usually list their candidates location by location: the table is asked once per run, and if no location comes back after
    // another one has started (`runs`), the candidates are already sorted by location
    bool runs = true;
    const unsigned* up0 = reinterpret_cast<const unsigned*>(params);
    {
        // where a run of equal locations starts (worker threads), the table only at the starts, the runs' members again on the workers
        std::vector<char> starts(n);
        const size_t blk = 8192, nblk = ((size_t)n + blk - 1) / blk;
        workers(c).run(nblk, [&](size_t b) {
            for (size_t i = b * blk; i < std::min((size_t)n, (b + 1) * blk); i++) {
                const unsigned* u = up0 + i * 11;
                starts[i] = i == 0 || u[0] != u[-11] || u[1] != u[-10] || u[2] != u[-9] || u[3] != u[-8] || u[10] != u[-1];
            }
        });
        std::vector<int> run_begin, run_loc;
        for (int i = 0; i < n; i++) {
            if (!starts[i]) continue;
            const unsigned* u = up0 + (size_t)i * 11;
            Key k;
            k.v[0] = u[0]; k.v[1] = u[1]; k.v[2] = u[2]; k.v[3] = u[3]; k.v[4] = u[10];
            auto it = ids.find(k);
            if (it == ids.end()) { it = ids.emplace(k, (int)first_of.size()).first; first_of.push_back(i); }
            else runs = false;
            run_begin.push_back(i); run_loc.push_back(it->second);
        }
        run_begin.push_back(n);
        workers(c).run(run_loc.size(), [&](size_t r) { for (int i = run_begin[r]; i < run_begin[r + 1]; i++) loc_of[i] = run_loc[r]; });
    }
    const int nloc = (int)first_of.size();
    if ((long long)nloc * 8 > n) return 0;   // fewer than 8 tensors per location on average: the direct path is as good
    // candidates sorted by location, laid out in page-locked memory: the copies below are DMA transfers that run while eval_batch prepares
    // the probe sources (the buffer is rewritten only after the stream has been waited for at the end of a batch)
    const size_t off_mts = ((size_t)nloc * sizeof(MtLocHost) + 255) & ~(size_t)255, off_cand = (off_mts + sizeof(float) * 6 * (size_t)n + 255) & ~(size_t)255;
    CU_OK(c->h_mt.ensure(off_cand + sizeof(int) * (size_t)n));
    MtLocHost* locs = c->h_mt.as<MtLocHost>();
    float* mts = reinterpret_cast<float*>(c->h_mt.as<char>() + off_mts);
    int* cand_of = reinterpret_cast<int*>(c->h_mt.as<char>() + off_cand);
    if (runs) {   // sorted already: location l is the run [first_of[l], first_of[l+1])
        for (int l = 0; l < nloc; l++) locs[l] = MtLocHost{first_of[l], (l + 1 < nloc ? first_of[l + 1] : n) - first_of[l]};
        const size_t blk = 8192, nblk = ((size_t)n + blk - 1) / blk;
        workers(c).run(nblk, [&](size_t b) {
            for (size_t i = b * blk; i < std::min((size_t)n, (b + 1) * blk); i++) {
                cand_of[i] = (int)i;
                memcpy(&mts[i * 6], params + i * 11 + 4, sizeof(float) * 6);
            }
        });
    } else {
        for (int l = 0; l < nloc; l++) locs[l] = MtLocHost{0, 0};
        for (int i = 0; i < n; i++) locs[loc_of[i]].mt_count++;
        for (int l = 1; l < nloc; l++) locs[l].mt_begin = locs[l - 1].mt_begin + locs[l - 1].mt_count;
        std::vector<int> fill(nloc, 0);
        for (int i = 0; i < n; i++) {
            const int l = loc_of[i], at = locs[l].mt_begin + fill[l]++;
            cand_of[at] = i;
            memcpy(&mts[(size_t)at * 6], params + (size_t)i * 11 + 4, sizeof(float) * 6);
        }
    }
    CU_OK(c->d_mtlocs.ensure(sizeof(MtLocHost) * nloc));
    CU_OK(c->d_mts.ensure(sizeof(float) * 6 * (size_t)n));
    CU_OK(c->d_candof.ensure(sizeof(int) * (size_t)n));
    CU_OK(cudaMemcpyAsync(c->d_mtlocs.p, locs, sizeof(MtLocHost) * nloc, cudaMemcpyHostToDevice, c->stream));
    CU_OK(cudaMemcpyAsync(c->d_mts.p, mts, sizeof(float) * 6 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CU_OK(cudaMemcpyAsync(c->d_candof.p, cand_of, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    const auto tg1 = std::chrono::steady_clock::now();
    const int nrcv = (int)c->rcv.size();
    SynthHook hook;
    hook.fn = [&](int cand0, int ncand, const float* seis, size_t seis_stride, const SeisHdr* shdrs, const CandDev*) -> int {
        const int l0 = cand0 / 6, nl = ncand / 6;
        launch_mt_contract(c->d_rcv.as<ReceiverDev>(), nrcv, reinterpret_cast<const MtLoc*>(c->d_mtlocs.p) + l0, nl, c->d_mts.as<float>(),
                           c->d_candof.as<int>(), seis, seis_stride, shdrs, c->d_refdata.as<float>(), c->d_taper.as<float>(), c->misfit_method,
                           c->db.dt, c->syn_factor, c->nmisfits, d_out, c->stream);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return kiwi_set_error("CUDA error launching the moment-tensor contraction: %s", cudaGetErrorString(e));
        return 0;
    };
    // one probe source per location (mxx = mxz = 1: its azimuth factors give those of all six unit tensors), synthesis fused into the
    // contraction kernel ...
    std::vector<int> bstatus((size_t)nloc * 6, 0);
    {
        std::vector<float> probe((size_t)nloc * 11, 0.f);
        for (int l = 0; l < nloc; l++) {
            float* b = &probe[(size_t)l * 11];
            const float* p = params + (size_t)first_of[l] * 11;
            b[0] = p[0]; b[1] = p[1]; b[2] = p[2]; b[3] = p[3]; b[10] = p[10];
            b[4] = 1.f; b[8] = 1.f;
        }
        std::vector<int> pstatus((size_t)nloc, 0);
        int ncomp_max = 1;
        for (const ReceiverDev& r : c->h_rcvdev) if (r.enabled) ncomp_max = std::max(ncomp_max, r.ncomp);
        CU_OK(c->d_tmax.ensure(sizeof(int) * 8));
        int* d_overflow = c->d_tmax.as<int>() + 4;
        CU_OK(cudaMemsetAsync(d_overflow, 0, sizeof(int), c->stream));
        hook.fused = true; hook.align = 1;
        hook.fused_fn = [&](int cand0, int ncand, const GeoRec* recs, const PairHdr* hdrs, const float4* taprec, int nq) -> int {
            const int strip_cap = 4 * nq + 16;
            if (mt_fused_smem_bytes(strip_cap, ncomp_max) > (size_t)100 * 1024) return 2;   // (two CTAs per SM at least)
            cudaError_t e = launch_mt_fused(c->db, c->d_rcv.as<ReceiverDev>(), nrcv, reinterpret_cast<const MtLoc*>(c->d_mtlocs.p) + cand0, ncand,
                                            c->d_mts.as<float>(), c->d_candof.as<int>(), recs, hdrs, taprec, strip_cap, ncomp_max,
                                            c->d_refdata.as<float>(), c->d_taper.as<float>(), c->misfit_method, c->db.dt, c->syn_factor, c->nmisfits,
                                            d_out, d_overflow, c->stream);
            if (e != cudaSuccess) return kiwi_set_error("CUDA error launching the fused moment-tensor contraction: %s", cudaGetErrorString(e));
            return 0;
        };
        static const bool no_fused = getenv("KIWI_NO_MT_FUSED") != nullptr;
        int rc = (no_fused || !c->mt_grid_fused) ? 2 : eval_batch(c, KIWI_SOURCE_MOMENT_TENSOR, nloc, 11, probe.data(), nullptr, pstatus.data(), false, &hook);
        if (rc == 1) return 1;
        if (rc == 0) {
            int overflow = 0;
            CU_OK(cudaMemcpyAsync(&overflow, d_overflow, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            CU_OK(cudaStreamSynchronize(c->stream));
            if (overflow) rc = 2;
        }
        if (rc == 0) for (int l = 0; l < nloc; l++) for (int k = 0; k < 6; k++) bstatus[(size_t)l * 6 + k] = pstatus[l];
        else {   // ... or, where a window or shift table does not fit that kernel, the six unit tensors through the general synthesis
            std::vector<float> basis((size_t)nloc * 6 * 11, 0.f);
            for (int l = 0; l < nloc; l++)
                for (int k = 0; k < 6; k++) {
                    float* b = &basis[((size_t)l * 6 + k) * 11];
                    const float* p = params + (size_t)first_of[l] * 11;
                    b[0] = p[0]; b[1] = p[1]; b[2] = p[2]; b[3] = p[3]; b[10] = p[10];
                    b[4 + k] = 1.f;
                }
            hook.fused = false; hook.align = 6;
            if (eval_batch(c, KIWI_SOURCE_MOMENT_TENSOR, nloc * 6, 11, basis.data(), nullptr, bstatus.data(), false, &hook)) return 1;
        }
    }
    const auto tg2 = std::chrono::steady_clock::now();
    if (h_status) {   // status of the basis, or 2 where a misfit came out NaN/Inf (as k_misfit_td reports it on the direct path)
        CU_OK(c->d_status_out.ensure(sizeof(int) * ((size_t)n + 1)));
        CU_OK(cudaMemsetAsync(c->d_status_out.p, 0, sizeof(int) * ((size_t)n + 1), c->stream));
        launch_flag_nonfinite(d_out, n, c->nmisfits * 2, c->d_status_out.as<int>(), c->d_status_out.as<int>() + n, c->stream);
        int nbad = 0;   // (the flags themselves only where there are any)
        CU_OK(cudaMemcpyAsync(&nbad, c->d_status_out.as<int>() + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CU_OK(cudaStreamSynchronize(c->stream));
        std::vector<int> nonfinite;
        if (nbad > 0) {
            nonfinite.assign((size_t)n, 0);
            CU_OK(cudaMemcpyAsync(nonfinite.data(), c->d_status_out.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
            CU_OK(cudaStreamSynchronize(c->stream));
        }
        const size_t blk = 16384, nblk = ((size_t)n + blk - 1) / blk;
        workers(c).run(nblk, [&](size_t b) {
            for (size_t i = b * blk; i < std::min((size_t)n, (b + 1) * blk); i++) {
                const int bs = bstatus[(size_t)loc_of[i] * 6];
                h_status[i] = bs != KIWI_STATUS_OK ? bs : ((nbad > 0 && nonfinite[i]) ? KIWI_STATUS_NONFINITE : KIWI_STATUS_OK);
            }
        });
    }
    if (trace) {
        const auto tg3 = std::chrono::steady_clock::now();
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "[kiwi trace] moment-tensor grid of %d candidates at %d locations: grouping + staging %.3f ms, evaluation %.3f ms, status %.3f ms\n", n, nloc,
                ms(tg0, tg1), ms(tg1, tg2), ms(tg2, tg3));
    }
    *used = true;
    return 0;
}

int ensure_single(kiwi_ctx* c, bool want_misfits) {
    if (!c->src_set) return kiwi_set_error("no source parameters set");   // minimizer_engine.f90:1394
    // anything set on the receiver side since the last evaluation (references, tapers, filters, switches, shifts) has not reached the
    // device yet: evaluate again rather than let a getter work on the receivers of the previous evaluation
    if (c->receivers_dirty) c->src_dirty = true;
    if (!c->src_dirty && (!want_misfits || !c->src_misfits.empty()) && c->last.valid && c->last.n == 1) return 0;
    if (upload_receivers(c)) return 1;
    const int nm = c->nmisfits;
    CU_OK(c->d_out.ensure(sizeof(float) * 2 * std::max(nm, 1)));
    int status = 0;
    c->src_misfits.clear();
    c->last_eval_ns = 0;   // d_out is rewritten: the block of an earlier kiwi_eval_sources is gone (kiwi_outer_misfits must not use it)
    if (eval_batch(c, c->src_type, 1, (int)c->src_params.size(), c->src_params.data(), c->d_out.as<float>(), &status, want_misfits)) return 1;
    if (want_misfits) {
        c->src_misfits.assign((size_t)2 * nm, 0.f);
        if (nm > 0) CU_OK(cudaMemcpy(c->src_misfits.data(), c->d_out.p, sizeof(float) * 2 * nm, cudaMemcpyDeviceToHost));
    }
    c->last_fshift.clear();
    if (want_misfits && c->misfit_method >= KIWI_FLOATING_L2NORM && c->d_fshift.p) {
        c->last_fshift.assign(c->rcv.size(), 0);
        CU_OK(cudaMemcpy(c->last_fshift.data(), c->d_fshift.p, sizeof(int) * c->rcv.size(), cudaMemcpyDeviceToHost));
    }
    c->src_status = status;
    c->src_dirty = false;
    if (status == KIWI_STATUS_BAD_PARAMS) return kiwi_set_error("%s", c->prep_error.empty() ? "discretisation of the source failed" : c->prep_error.c_str());
    return 0;
}

}  // namespace

extern "C" {

int kiwi_get_n_source_params(int sourcetype) {   // source_all.f90:97-121
    switch (sourcetype) {
        case KIWI_SOURCE_BILATERAL: return 14;
        case KIWI_SOURCE_MOMENT_TENSOR: return 11;
        case KIWI_SOURCE_CIRCULAR: return 11;      // source_circular.f90:32
        case KIWI_SOURCE_POINT_LP: return 13;      // source_point_lp.f90:14
        case KIWI_SOURCE_EIKONAL: return 15;       // source_eikonal.f90:36
        case KIWI_SOURCE_MT_EIKONAL: return 20;    // source_mt_eikonal.f90:36
        default: return 0;
    }
}

kiwi_ctx* kiwi_create(int device) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { kiwi_set_error("kiwi_create: no CUDA device available (%s); this engine has no CPU path", cudaGetErrorString(e)); return nullptr; }
    if (device < 0 || device >= ndev) { kiwi_set_error("kiwi_create: device %d out of range (0..%d)", device, ndev - 1); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { kiwi_set_error("kiwi_create: cudaSetDevice failed"); return nullptr; }
    kiwi_ctx* c = new kiwi_ctx();
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; kiwi_set_error("kiwi_create: stream creation failed"); return nullptr; }
    for (auto& ev : c->ev) cudaEventCreate(&ev);
    return c;
}

void kiwi_destroy(kiwi_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (DevBuf* b : {&c->d_slabs, &c->d_nodes, &c->d_tspan, &c->d_nspan, &c->d_rcv, &c->d_refdata, &c->d_taper, &c->d_cands, &c->d_bilat, &c->d_gf, &c->d_gi,
                      &c->d_tf, &c->d_recs, &c->d_hdrs, &c->d_seis, &c->d_shdrs, &c->d_out, &c->d_status, &c->d_tmax, &c->d_table, &c->d_tw, &c->d_fshift, &c->d_map, &c->d_status_out, &c->d_taprec, &c->d_partial, &c->d_fftz, &c->d_gm, &c->d_xcorr, &c->d_azf, &c->d_trig, &c->d_eik_s, &c->d_eik_t, &c->d_eik_bp, &c->d_eik_ovf, &c->d_eik_jobs, &c->d_eik_geoms, &c->d_eik_coarse, &c->d_mtlocs, &c->d_mts, &c->d_candof, &c->d_orc, &c->d_orw, &c->d_obw, &c->d_oout, &c->d_obest, &c->d_obestv})
        b->release();
    c->h_stage.release(); c->h_out.release(); c->h_mt.release(); c->h_eik.release();
    for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
    cudaStreamDestroy(c->stream);
    delete c;
}

// set_database (minimizer_engine.f90:114-139): the whole database goes to HBM once, as one slab per
// grid node (layout: kiwi_dev.cuh NodeInfo).
int kiwi_set_database(kiwi_ctx* c, kiwi_gfdb* db) {
    if (!c || !db) return kiwi_set_error("kiwi_set_database: null argument");
    CU_OK(cudaSetDevice(c->device));
    db->flatten();
    const size_t nnodes = (size_t)db->nx * db->nz;
    const int ng = db->ng;
    c->h_nodes.assign(nnodes, NodeInfo());
    c->h_tspan.assign(nnodes * ng, make_int2(0, -1));
    unsigned long long total = 0;
    int tmin = INT_MAX, tmax = INT_MIN;
    for (size_t inode = 0; inode < nnodes; inode++) {
        int lo = INT_MAX, hi = INT_MIN; bool all = true;
        for (int k = 0; k < ng; k++) {
            const size_t it = inode * ng + k;
            if (db->len[it] <= 0) { all = false; continue; }
            lo = std::min(lo, db->span0[it]); hi = std::max(hi, db->span0[it] + db->len[it] - 1);
            c->h_tspan[it] = make_int2(db->span0[it], db->span0[it] + db->len[it] - 1);
        }
        NodeInfo& ni = c->h_nodes[inode];
        if (!all) { ni.off = ~0ull; ni.w0 = 0; ni.wn = 0; continue; }   // node unusable: the centroid is skipped (seismogram.f90:172)
        const int w0 = (int)(floor((double)lo / 4.0)) * 4 - 4;             // first quad = zeros only (left continuation)
        const int wend = (int)(floor((double)hi / 4.0)) * 4 + 8;      // exclusive, multiple of 4; last quad = continuation only
        ni.off = total; ni.w0 = w0; ni.wn = wend - w0;
        total += (unsigned long long)ni.wn * ng;
        tmin = std::min(tmin, lo); tmax = std::max(tmax, hi);
    }
    c->db_tmin = tmin; c->db_tmax = tmax; c->slab_floats = (size_t)total; c->db_blk_floats = 0;
    // fill the slabs through a bounded pinned staging buffer
    CU_OK(c->d_slabs.ensure(sizeof(float) * std::max<unsigned long long>(total, 4)));
    CU_OK(c->d_nodes.ensure(sizeof(NodeInfo) * nnodes));
    CU_OK(c->d_tspan.ensure(sizeof(int2) * nnodes * ng));
    const size_t stage_floats = (size_t)16 << 20;   // 64 MiB
    CU_OK(c->h_stage.ensure(stage_floats * sizeof(float)));
    float* stage = c->h_stage.as<float>();
    size_t inode = 0;
    while (inode < nnodes) {
        // take as many whole nodes as fit the staging buffer
        size_t first = inode; unsigned long long base = ~0ull; size_t used = 0;
        for (; inode < nnodes; inode++) {
            const NodeInfo& ni = c->h_nodes[inode];
            if (ni.off == ~0ull) continue;
            const size_t need = (size_t)ni.wn * ng;
            if (need > stage_floats) return kiwi_set_error("kiwi_set_database: node slab larger than the staging buffer");
            if (base == ~0ull) base = ni.off;
            if (used + need > stage_floats) break;
            float* dst = stage + (ni.off - base);
            for (int k = 0; k < ng; k++) {
                const size_t it = inode * ng + k;
                const float* src = &db->data[(size_t)db->offset[it]];
                const int s0 = db->span0[it], len = db->len[it];
                float* row = dst + (size_t)k * ni.wn;
                const int lead = s0 - ni.w0;
                for (int j = 0; j < lead; j++) row[j] = 0.f;                        // zeros left of the trace span
                memcpy(row + lead, src, sizeof(float) * len);
                const float last = src[len - 1];
                for (int j = lead + len; j < ni.wn; j++) row[j] = last;            // last sample repeats (sparse_trace.f90:29-50)
            }
            used += need;
        }
        (void)first;
        if (used > 0) {
            CU_OK(cudaMemcpyAsync(c->d_slabs.as<float>() + base, stage, sizeof(float) * used, cudaMemcpyHostToDevice, c->stream));
            CU_OK(cudaStreamSynchronize(c->stream));
        }
    }
    {   // span unions of the component sets of every node (what make_seismogram's strips grow by, seismogram.f90:167-250)
        std::vector<int4> ns(2 * nnodes);
        const int set_of[10] = {0, 0, 0, 1, 1, 2, 2, 2, 0, 2};   // g1 g2 g3 | g4 g5 | g6 g7 g8 | g9 | g10
        for (size_t i = 0; i < nnodes; i++) {
            int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
            for (int k = 0; k < ng; k++) {
                const int2 sp = c->h_tspan[i * ng + k];
                lo[set_of[k]] = std::min(lo[set_of[k]], sp.x); hi[set_of[k]] = std::max(hi[set_of[k]], sp.y);
            }
            ns[2 * i] = make_int4(lo[0], hi[0], lo[1], hi[1]); ns[2 * i + 1] = make_int4(lo[2], hi[2], 0, 0);
        }
        CU_OK(c->d_nspan.ensure(sizeof(int4) * ns.size()));
        CU_OK(cudaMemcpyAsync(c->d_nspan.p, ns.data(), sizeof(int4) * ns.size(), cudaMemcpyHostToDevice, c->stream));
        CU_OK(cudaStreamSynchronize(c->stream));
    }
    CU_OK(cudaMemcpyAsync(c->d_nodes.p, c->h_nodes.data(), sizeof(NodeInfo) * nnodes, cudaMemcpyHostToDevice, c->stream));
    CU_OK(cudaMemcpyAsync(c->d_tspan.p, c->h_tspan.data(), sizeof(int2) * nnodes * ng, cudaMemcpyHostToDevice, c->stream));
    CU_OK(cudaStreamSynchronize(c->stream));
    GfdbDev& d = c->db;
    d.dt = db->dt; d.dx = db->dx; d.dz = db->dz; d.firstx = db->firstx; d.firstz = db->firstz;
    d.nx = db->nx; d.nz = db->nz; d.ng = db->ng;
    d.slabs = c->d_slabs.as<float>(); d.nodes = c->d_nodes.as<NodeInfo>(); d.tspan = c->d_tspan.as<int2>(); d.nspan = c->d_nspan.as<int4>(); d.lastval = nullptr;
    c->db_set = true;
    c->receivers_dirty = true; c->src_dirty = true; c->last.valid = false; c->work_budget = 0;
    return 0;
}

int kiwi_set_local_interpolation(kiwi_ctx* c, int bilinear) {
    if (!c) return kiwi_set_error("null context");
    c->interpolate = bilinear != 0; c->src_dirty = true;
    return 0;
}
int kiwi_set_spacial_undersampling(kiwi_ctx* c, int xunder, int zunder) {
    if (!c) return kiwi_set_error("null context");
    if (xunder < 1 || zunder < 1) return kiwi_set_error("invalid undersampling value");   // minimizer_engine.f90:155-158
    c->xunder = xunder; c->zunder = zunder; c->src_dirty = true;
    return 0;
}

int kiwi_set_receivers(kiwi_ctx* c, int n, const double* lat_deg, const double* lon_deg, const float* depth, const char* const* components) {
    if (!c) return kiwi_set_error("null context");
    if (require_db(c)) return 1;
    std::vector<HostReceiver> v(n);
    for (int i = 0; i < n; i++) {
        HostReceiver& h = v[i];
        h.lat = kh::d2r_d(lat_deg[i]); h.lon = kh::d2r_d(lon_deg[i]);   // minimizer_engine.f90:262 d2r(origin)
        h.depth = depth ? depth[i] : 0.f;
        const char* cs = components[i] ? components[i] : "";
        const int nc = (int)strlen(cs);
        if (nc > KIWI_MAX_COMP) return kiwi_set_error("initializing receiver failed: too many components at receiver %d", i + 1);
        for (int k = 0; k < nc; k++) {   // receiver.f90:160-187
            const int id = character_to_id(cs[k]);
            if (id == 0) return kiwi_set_error("initializing receiver failed: unknown component '%c' at receiver %d", cs[k], i + 1);
            for (int j = 0; j < k; j++) if (abs(h.comp[j]) == abs(id)) return kiwi_set_error("initializing receiver failed: conflicting components at receiver %d", i + 1);
            h.comp[k] = id;
        }
        h.ncomp = nc;
        h.enabled = nc > 0;   // receiver.f90:155-157
    }
    c->rcv.swap(v);
    c->receivers_set = true; c->receivers_dirty = true; c->src_dirty = true; c->last.valid = false;
    return 0;
}

int kiwi_switch_receiver(kiwi_ctx* c, int ireceiver, int state) {
    if (!c) return kiwi_set_error("null context");
    if (require_receivers(c)) return 1;
    if (ireceiver < 1 || ireceiver > (int)c->rcv.size()) return kiwi_set_error("receiver index out of range");
    c->rcv[ireceiver - 1].enabled = state != 0;
    c->receivers_dirty = true; c->src_dirty = true;
    return 0;
}

int kiwi_set_source_location(kiwi_ctx* c, float lat_deg, float lon_deg, double ref_time) {
    if (!c) return kiwi_set_error("null context");
    c->olat = (double)kh::d2r_r(lat_deg);   // minimizer.f90:512 d2r(lat) on default reals, widened at the call (:453-456)
    c->olon = (double)kh::d2r_r(lon_deg);
    c->ref_time = ref_time;
    c->loc_set = true; c->receivers_dirty = true; c->src_dirty = true;
    if (c->crust.loaded) {   // psm_set_origin_and_time -> psm_set_default_constraints (parameterized_source.f90:183-196): replaces any user constraints
        kh::default_constraints(c->crust, c->olat, c->olon, c->thickness_limit, &c->constraints);
        c->user_constraints = false;
    }
    return 0;
}

int kiwi_set_effective_dt(kiwi_ctx* c, float effective_dt) {
    if (!c) return kiwi_set_error("null context");
    if (!(effective_dt > 0.f)) return kiwi_set_error("effective dt must be positive");
    c->effective_dt = effective_dt; c->src_dirty = true;
    return 0;
}

int kiwi_set_ref_seismogram(kiwi_ctx* c, int ireceiver, int icomponent, float tbegin, int n, const float* data) {
    if (!c) return kiwi_set_error("null context");
    if (require_receivers(c)) return 1;
    if (ireceiver < 1 || ireceiver > (int)c->rcv.size()) return kiwi_set_error("receiver index out of range");
    HostReceiver& h = c->rcv[ireceiver - 1];
    if (icomponent < 1 || icomponent > h.ncomp) return kiwi_set_error("component index out of range");
    if (n < 1) return kiwi_set_error("empty reference seismogram");
    const int k = icomponent - 1;
    const int ibeg = (int)lroundf(tbegin / c->db.dt);   // receiver.f90:843-848 seismogram_to_strip
    h.ref[k].assign(data, data + n);
    h.ref_ds0[k] = ibeg + 1; h.ref_ds1[k] = ibeg + n;
    h.has_ref[k] = true;
    c->receivers_dirty = true; c->src_misfits.clear();
    return 0;
}

int kiwi_set_misfit_method(kiwi_ctx* c, int norm_id) {
    if (!c) return kiwi_set_error("null context");
    if (norm_id < 1 || norm_id > 8) return kiwi_set_error("unknown norm method");
    c->misfit_method = norm_id; c->src_misfits.clear();
    return 0;
}

int kiwi_set_misfit_taper(kiwi_ctx* c, int ireceiver, int n, const float* x, const float* y) {
    if (!c) return kiwi_set_error("null context");
    if (require_receivers(c)) return 1;
    if (ireceiver < 1 || ireceiver > (int)c->rcv.size()) return kiwi_set_error("receiver index out of range");
    if (n < 2) return kiwi_set_error("a taper needs at least two points");
    HostReceiver& h = c->rcv[ireceiver - 1];
    h.taper_x.assign(x, x + n); h.taper_y.assign(y, y + n);
    c->receivers_dirty = true; c->src_misfits.clear();
    return 0;
}

int kiwi_set_misfit_filter(kiwi_ctx* c, int ireceiver, int n, const float* x, const float* y) {
    if (!c) return kiwi_set_error("null context");
    if (require_receivers(c)) return 1;
    if (ireceiver < 0 || ireceiver > (int)c->rcv.size()) return kiwi_set_error("receiver index out of range");
    if (n < 2) return kiwi_set_error("a filter needs at least two points");
    for (int i = 0; i < (int)c->rcv.size(); i++) {
        if (ireceiver != 0 && i != ireceiver - 1) continue;
        c->rcv[i].filter_x.assign(x, x + n); c->rcv[i].filter_y.assign(y, y + n);
    }
    c->receivers_dirty = true; c->src_misfits.clear();
    return 0;
}

int kiwi_set_crust2x2(kiwi_ctx* c, const char* path) {   // crust2x2_load, minimizer.f90:1669-1674
    if (!c) return kiwi_set_error("null context");
    std::string err;
    if (!kh::crust2x2_load(path, &c->crust, &err)) return kiwi_set_error("%s", err.c_str());
    if (c->loc_set && !c->user_constraints) kh::default_constraints(c->crust, c->olat, c->olon, c->thickness_limit, &c->constraints);
    c->src_dirty = true;
    return 0;
}

int kiwi_set_source_constraints(kiwi_ctx* c, int n, const float* points, const float* normals) {   // minimizer_engine.f90 set_source_constraints
    if (!c) return kiwi_set_error("null context");
    if (n < 0) return kiwi_set_error("negative number of constraints");
    c->constraints.assign(n, kh::Halfspace());
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) { c->constraints[i].point[k] = points[3 * i + k]; c->constraints[i].normal[k] = normals[3 * i + k]; }
    c->user_constraints = true; c->src_dirty = true;
    return 0;
}

int kiwi_set_source_crustal_thickness_limit(kiwi_ctx* c, float limit) {   // parameterized_source.f90:198-207
    if (!c) return kiwi_set_error("null context");
    c->thickness_limit = limit;
    if (c->crust.loaded && c->loc_set) { kh::default_constraints(c->crust, c->olat, c->olon, c->thickness_limit, &c->constraints); c->user_constraints = false; }
    c->src_dirty = true;
    return 0;
}

int kiwi_set_share_syntheses(kiwi_ctx* c, int enabled) {
    if (!c) return kiwi_set_error("null context");
    c->dedup_enabled = enabled != 0;
    return 0;
}

int kiwi_set_accumulation(kiwi_ctx* c, int reference_order) {
    if (!c) return kiwi_set_error("null context");
    c->accum_reference = reference_order != 0;
    c->src_dirty = true;
    return 0;
}

int kiwi_set_eikonal_device(kiwi_ctx* c, int min_batch) {
    if (!c) return kiwi_set_error("null context");
    c->eikonal_device_min = min_batch < 0 ? -1 : min_batch;
    return 0;
}

int kiwi_set_mt_grid(kiwi_ctx* c, int enabled) {
    if (!c) return kiwi_set_error("null context");
    c->mt_grid_enabled = enabled != 0;
    c->mt_grid_fused = enabled != 2;
    return 0;
}

int kiwi_set_synthetics_factor(kiwi_ctx* c, float factor) {
    if (!c) return kiwi_set_error("null context");
    c->syn_factor = factor; c->src_misfits.clear();
    return 0;
}

int kiwi_set_floating_shiftrange(kiwi_ctx* c, int ireceiver, float lo, float hi) {
    if (!c) return kiwi_set_error("null context");
    if (require_receivers(c)) return 1;
    if (ireceiver < 0 || ireceiver > (int)c->rcv.size()) return kiwi_set_error("receiver index out of range");
    const int r0 = (int)lroundf(lo / c->db.dt), r1 = (int)lroundf(hi / c->db.dt);   // minimizer_engine.f90:436-437
    for (int i = 0; i < (int)c->rcv.size(); i++) {
        if (ireceiver != 0 && i != ireceiver - 1) continue;
        c->rcv[i].fs0 = r0; c->rcv[i].fs1 = r1;
    }
    c->receivers_dirty = true; c->src_misfits.clear();
    return 0;
}

int kiwi_get_nmisfits(kiwi_ctx* c) {
    if (!c) return 0;
    int n = 0;
    for (const HostReceiver& h : c->rcv) if (h.enabled) n += h.ncomp;
    return n;
}

int kiwi_eval_sources_device(kiwi_ctx* c, int sourcetype, int ns, int nparams, const float* params, float* d_misfits, int* status) {
    if (!c) return kiwi_set_error("null context");
    CU_OK(cudaSetDevice(c->device));
    if (ns < 0) return kiwi_set_error("negative number of sources");
    c->src_dirty = true;
    return eval_batch(c, sourcetype, ns, nparams, params, d_misfits, status, true);
}

int kiwi_eval_sources(kiwi_ctx* c, int sourcetype, int ns, int nparams, const float* params, float* misfits, int* status) {
    if (!c) return kiwi_set_error("null context");
    CU_OK(cudaSetDevice(c->device));
    if (ns < 0) return kiwi_set_error("negative number of sources");
    if (upload_receivers(c)) return 1;
    const size_t nfl = (size_t)ns * std::max(c->nmisfits, 1) * 2;
    CU_OK(c->d_out.ensure(sizeof(float) * std::max<size_t>(nfl, 2)));
    c->src_dirty = true;
    c->last_eval_ns = 0;
    if (eval_batch(c, sourcetype, ns, nparams, params, c->d_out.as<float>(), status, true)) return 1;
    c->last_eval_ns = ns;
    if (ns > 0 && c->nmisfits > 0 && misfits)
        CU_OK(cudaMemcpy(misfits, c->d_out.p, sizeof(float) * (size_t)ns * c->nmisfits * 2, cudaMemcpyDeviceToHost));
    return 0;
}

int kiwi_global_misfits(int ns, int nmisfits, const float* misfits, float* global_misfits) {
    // minimizer_engine.f90:937-942: fp32 sums in receiver order, sqrt(sum m^2)/sqrt(sum n^2)
    for (int s = 0; s < ns; s++) {
        float m = 0.f, nf = 0.f;
        const float* p = misfits + (size_t)s * nmisfits * 2;
        for (int i = 0; i < nmisfits; i++) { m = m + p[2 * i] * p[2 * i]; nf = nf + p[2 * i + 1] * p[2 * i + 1]; }
        global_misfits[s] = sqrtf(m) / sqrtf(nf);
    }
    return 0;
}

// make_global_misfits (python/tunguska/seismosizer.py:843-922) + nanargmin (gridsearch.py:250-266) on the device
int kiwi_outer_misfits(kiwi_ctx* c, int ns, const float* d_misfits, const double* receiver_weights, int outer_norm, int anarchy, int nboot,
                       const double* bweights, double* misfits_by_s, int* best, double* best_value) {
    if (!c) return kiwi_set_error("null context");
    CU_OK(cudaSetDevice(c->device));
    if (outer_norm != KIWI_L2NORM && outer_norm != KIWI_L1NORM) return kiwi_set_error("unknown norm method");
    if (nboot < 0 || (nboot > 0 && !bweights)) return kiwi_set_error("bootstrap weights missing");
    if (upload_receivers(c)) return 1;
    if (!d_misfits) {
        if (ns != c->last_eval_ns || ns == 0) return kiwi_set_error("no misfit block of %d candidates from kiwi_eval_sources on the device", ns);
        d_misfits = c->d_out.as<float>();
    }
    const int nr = (int)c->rcv.size(), nrows = nboot + 1;
    std::vector<int> rc((size_t)2 * std::max(nr, 1), 0);
    for (int i = 0; i < nr; i++) { rc[2 * i] = c->h_rcvdev[i].misfit_base; rc[2 * i + 1] = c->h_rcvdev[i].enabled ? c->h_rcvdev[i].ncomp : 0; }
    CU_OK(c->d_orc.ensure(sizeof(int) * rc.size()));
    CU_OK(cudaMemcpyAsync(c->d_orc.p, rc.data(), sizeof(int) * rc.size(), cudaMemcpyHostToDevice, c->stream));
    if (receiver_weights) {
        CU_OK(c->d_orw.ensure(sizeof(double) * nr));
        CU_OK(cudaMemcpyAsync(c->d_orw.p, receiver_weights, sizeof(double) * nr, cudaMemcpyHostToDevice, c->stream));
    }
    if (nboot > 0) {
        CU_OK(c->d_obw.ensure(sizeof(double) * (size_t)nboot * nr));
        CU_OK(cudaMemcpyAsync(c->d_obw.p, bweights, sizeof(double) * (size_t)nboot * nr, cudaMemcpyHostToDevice, c->stream));
    }
    // the [1 + nboot][ns] matrix lives on the device as a whole only if the caller wants it; otherwise blocks of rows are reduced to
    // their minima one after the other (1000 bootstrap rows of a 10^6-candidate grid would be 8 GB)
    const size_t row_bytes = sizeof(double) * (size_t)std::max(ns, 1);
    const char* pass_env = getenv("KIWI_OUTER_PASS_BYTES");   // (tests: a small pass)
    const size_t pass_bytes = pass_env ? (size_t)atoll(pass_env) : (size_t)256 << 20;
    const int rows_per_pass = misfits_by_s ? nrows : (int)std::min<size_t>((size_t)nrows, std::max<size_t>(1, pass_bytes / row_bytes));
    CU_OK(c->d_oout.ensure(row_bytes * rows_per_pass));
    CU_OK(c->d_obest.ensure(sizeof(int) * nrows));
    CU_OK(c->d_obestv.ensure(sizeof(double) * nrows));
    if ((size_t)2 * nr * sizeof(double) > (size_t)200 * 1024) return kiwi_set_error("too many receivers for the outer-misfit kernel");
    for (int row0 = 0; row0 < nrows; row0 += rows_per_pass) {
        cudaError_t e = launch_outer_misfits(d_misfits, c->nmisfits, c->d_orc.p, nr, receiver_weights ? c->d_orw.as<double>() : nullptr,
                                             outer_norm == KIWI_L1NORM, anarchy != 0, std::min(rows_per_pass, nrows - row0),
                                             nboot > 0 ? c->d_obw.as<double>() : nullptr, c->d_oout.as<double>(), ns, c->d_obest.as<int>(),
                                             c->d_obestv.as<double>(), c->stream, row0);
        if (e != cudaSuccess) return kiwi_set_error("CUDA error in the outer-misfit kernel: %s", cudaGetErrorString(e));
    }
    if (misfits_by_s && ns > 0) CU_OK(cudaMemcpyAsync(misfits_by_s, c->d_oout.p, sizeof(double) * (size_t)nrows * ns, cudaMemcpyDeviceToHost, c->stream));
    if (best && ns > 0) CU_OK(cudaMemcpyAsync(best, c->d_obest.p, sizeof(int) * nrows, cudaMemcpyDeviceToHost, c->stream));
    if (best_value && ns > 0) CU_OK(cudaMemcpyAsync(best_value, c->d_obestv.p, sizeof(double) * nrows, cudaMemcpyDeviceToHost, c->stream));
    CU_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

int kiwi_set_source_params(kiwi_ctx* c, int sourcetype, int nparams, const float* params) {
    if (!c) return kiwi_set_error("null context");
    if (!c->loc_set) return kiwi_set_error("no source location set");
    if (nparams != kiwi_get_n_source_params(sourcetype) || nparams == 0) return kiwi_set_error("wrong number of source parameters or source type not available");
    // minimizer_engine.f90:511-513: identical parameters are a no-op
    if (c->src_set && c->src_type == sourcetype && (int)c->src_params.size() == nparams &&
        memcmp(c->src_params.data(), params, sizeof(float) * nparams) == 0) return 0;
    if (c->src_type != sourcetype) c->src_mask.clear();   // psm_set: a new source type selects every parameter (source_all.f90:249-253)
    c->src_type = sourcetype; c->src_params.assign(params, params + nparams);
    c->src_set = true; c->src_dirty = true;
    return 0;
}

int kiwi_get_misfits(kiwi_ctx* c, float* misfits, int cap_pairs, int* nmisfits) {
    if (!c) return kiwi_set_error("null context");
    CU_OK(cudaSetDevice(c->device));
    if (ensure_single(c, true)) return 1;
    const int nm = c->nmisfits;
    if (nmisfits) *nmisfits = nm;
    if (cap_pairs < nm) return kiwi_set_error("misfit buffer too small: need %d pairs", nm);
    memcpy(misfits, c->src_misfits.data(), sizeof(float) * 2 * nm);
    for (int i = 0; i < 2 * nm; i++)
        if (!std::isfinite(misfits[i])) return kiwi_set_error("misfit is nan or very big at receiver component %d", i / 2 + 1);   // minimizer_engine.f90:1163-1166
    return 0;
}

int kiwi_get_global_misfit(kiwi_ctx* c, float* misfit) {
    if (!c) return kiwi_set_error("null context");
    CU_OK(cudaSetDevice(c->device));
    if (ensure_single(c, true)) return 1;
    return kiwi_global_misfits(1, c->nmisfits, c->src_misfits.data(), misfit);
}

// ---- ground-motion diagnostics: get_peak_amplitudes / get_arias_intensities ------------------------
namespace {
int require_no_filter(kiwi_ctx* c) {
    for (const HostReceiver& h : c->rcv)
        if (h.enabled && !h.filter_x.empty()) return kiwi_set_error("peak amplitudes / Arias intensities of band-pass filtered synthetics are not available");
    return 0;
}
// values of the enabled receivers, in order, from a [nrcv][3] block
int pick_enabled(kiwi_ctx* c, const float* gm3, int which, float* out, int cap, int* n) {
    int k = 0;
    for (size_t ir = 0; ir < c->rcv.size(); ir++) {
        if (!c->rcv[ir].enabled) continue;
        if (k < cap) out[k] = gm3[3 * ir + (which - 1)];
        k++;
    }
    if (n) *n = k;
    return k <= cap ? 0 : kiwi_set_error("buffer too small: need %d values", k);
}
int single_ground_motion(kiwi_ctx* c, int which, float* out, int cap, int* n) {
    if (!c) return kiwi_set_error("null context");
    CU_OK(cudaSetDevice(c->device));
    if (require_no_filter(c)) return 1;
    if (ensure_single(c, false)) return 1;   // update_syn_probes
    const int nrcv = (int)c->rcv.size();
    std::vector<float> gm((size_t)3 * std::max(nrcv, 1), 0.f);
    if (c->last.seis_valid) {
        CU_OK(c->d_gm.ensure(sizeof(float) * 3 * (size_t)nrcv));
        launch_ground_motion(c->d_rcv.as<ReceiverDev>(), nrcv, c->d_cands.as<CandDev>(), 1, c->d_seis.as<float>(), c->last.seis_stride,
                             c->d_shdrs.as<SeisHdr>(), c->d_taper.as<float>(), c->db.dt, c->syn_factor, c->d_gm.as<float>(), c->stream);
        CU_OK(cudaMemcpyAsync(gm.data(), c->d_gm.p, sizeof(float) * 3 * (size_t)nrcv, cudaMemcpyDeviceToHost, c->stream));
        CU_OK(cudaStreamSynchronize(c->stream));
    }
    return pick_enabled(c, gm.data(), which, out, cap, n);
}
}  // namespace

int kiwi_get_peak_amplitudes(kiwi_ctx* c, int differentiate, float* maxabs, int cap, int* n) {
    if (differentiate != 1 && differentiate != 2)
        return kiwi_set_error("differentiate argument must be 1 for velocity or 2 for acceleration");   // minimizer_engine.f90:1185-1189
    return single_ground_motion(c, differentiate, maxabs, cap, n);
}

int kiwi_get_arias_intensities(kiwi_ctx* c, float* intensities, int cap, int* n) { return single_ground_motion(c, 3, intensities, cap, n); }

// batched: [ns][enabled receivers][3] = peak velocity, peak acceleration, Arias intensity of every candidate (no references needed)
int kiwi_eval_ground_motion(kiwi_ctx* c, int sourcetype, int ns, int nparams, const float* params, float* out, int* status) {
    if (!c) return kiwi_set_error("null context");
    CU_OK(cudaSetDevice(c->device));
    if (ns < 0) return kiwi_set_error("negative number of sources");
    if (require_no_filter(c)) return 1;
    if (upload_receivers(c)) return 1;
    const int nrcv = (int)c->rcv.size();
    CU_OK(c->d_gm.ensure(sizeof(float) * 3 * (size_t)std::max(nrcv, 1) * std::max(ns, 1)));
    SynthHook hook;
    hook.fn = [&](int cand0, int ncand, const float* seis, size_t seis_stride, const SeisHdr* shdrs, const CandDev* d_cands) -> int {
        launch_ground_motion(c->d_rcv.as<ReceiverDev>(), nrcv, d_cands, ncand, seis, seis_stride, shdrs, c->d_taper.as<float>(), c->db.dt, c->syn_factor,
                             c->d_gm.as<float>() + (size_t)cand0 * nrcv * 3, c->stream);
        return 0;
    };
    c->src_dirty = true;
    if (eval_batch(c, sourcetype, ns, nparams, params, nullptr, status, false, &hook)) return 1;
    if (ns == 0) return 0;
    std::vector<float> gm((size_t)3 * nrcv * ns);
    CU_OK(cudaMemcpy(gm.data(), c->d_gm.p, sizeof(float) * gm.size(), cudaMemcpyDeviceToHost));
    int nen = 0;
    for (const HostReceiver& h : c->rcv) nen += h.enabled ? 1 : 0;
    for (int s = 0; s < ns; s++) {
        int k = 0;
        for (int ir = 0; ir < nrcv; ir++) {
            if (!c->rcv[ir].enabled) continue;
            for (int q = 0; q < 3; q++) out[((size_t)s * nen + k) * 3 + q] = gm[((size_t)s * nrcv + ir) * 3 + q];
            k++;
        }
    }
    c->last.valid = false;
    return 0;
}

// ---- sub-parameters and Levenberg-Marquardt -------------------------------------------------
namespace {
// psm_params_norm_* (source_bilat.f90:45-46, source_circular.f90:44-45, source_point_lp.f90:54-55, source_eikonal.f90:48-49,
// source_mt_eikonal.f90:48-50, source_moment_tensor.f90:42-43)
const std::vector<float>& params_norm(int sourcetype) {
    static const std::vector<float> bilat = {1.f, 10000.f, 10000.f, 10000.f, 7e18f, 360.f, 90.f, 360.f, 360.f, 10000.f, 10000.f, 10000.f, 3000.f, 1.f};
    static const std::vector<float> circular = {1.f, 10000.f, 10000.f, 10000.f, 7e18f, 360.f, 90.f, 360.f, 10000.f, 3000.f, 1.f};
    static const std::vector<float> point_lp = {1.f, 10000.f, 10000.f, 10000.f, 7e18f, 1.f, 0.f, -1.f, 1.f, 1.f, 1.f, 20.f, 1.f};
    static const std::vector<float> eikonal = {1.f, 10000.f, 10000.f, 10000.f, 7e18f, 360.f, 90.f, 360.f, 10000.f, 10000.f, 10000.f, 360.f, 10000.f, 1.f, 1.f};
    static const std::vector<float> mt_eikonal = {1.f, 10000.f, 10000.f, 10000.f, 7e18f, 360.f, 90.f, 10000.f, 10000.f, 10000.f, 360.f, 10000.f, 1.f, 7e18f,
                                                  7e18f, 7e18f, 7e18f, 7e18f, 7e18f, 1.f};
    static const std::vector<float> mt = {1.f, 10000.f, 10000.f, 10000.f, 7e18f, 7e18f, 7e18f, 7e18f, 7e18f, 7e18f, 1.f};
    static const std::vector<float> none;
    switch (sourcetype) {
        case KIWI_SOURCE_BILATERAL: return bilat;
        case KIWI_SOURCE_CIRCULAR: return circular;
        case KIWI_SOURCE_POINT_LP: return point_lp;
        case KIWI_SOURCE_EIKONAL: return eikonal;
        case KIWI_SOURCE_MT_EIKONAL: return mt_eikonal;
        case KIWI_SOURCE_MOMENT_TENSOR: return mt;
    }
    return none;
}
bool masked(const kiwi_ctx* c, size_t i) { return c->src_mask.empty() || c->src_mask[i]; }
int count_mask(const kiwi_ctx* c) {
    int k = 0;
    for (size_t i = 0; i < c->src_params.size(); i++) k += masked(c, i) ? 1 : 0;
    return k;
}
// psm_set_subparams (source_all.f90:377-428): ALL parameters make the round trip through the normalised
// representation when `normalized` (params/norm, masked ones replaced, times norm), as in the reference
std::vector<float> apply_subparams(const kiwi_ctx* c, const std::vector<float>& params, const float* sub, bool normalized) {
    const std::vector<float>& norm = params_norm(c->src_type);
    std::vector<float> copy(params.size());
    for (size_t i = 0; i < params.size(); i++) copy[i] = normalized ? params[i] / norm[i] : params[i];
    size_t isub = 0;
    for (size_t i = 0; i < params.size(); i++) if (masked(c, i)) copy[i] = sub[isub++];
    if (normalized) for (size_t i = 0; i < params.size(); i++) copy[i] = copy[i] * norm[i];
    return copy;
}
}  // namespace

int kiwi_set_source_params_mask(kiwi_ctx* c, const int* mask, int n) {
    if (!c) return kiwi_set_error("null context");
    if (!c->src_set) return kiwi_set_error("no source parameters set");
    if (n != (int)c->src_params.size()) return kiwi_set_error("wrong number of elements in source params mask");   // minimizer_engine.f90:533-537
    c->src_mask.assign(n, 0);
    for (int i = 0; i < n; i++) c->src_mask[i] = mask[i] ? 1 : 0;
    c->sub_mins.clear(); c->sub_maxs.clear();   // reset_subparam_limits, :541
    return 0;
}

int kiwi_set_source_subparams(kiwi_ctx* c, const float* sub, int n) {
    if (!c) return kiwi_set_error("null context");
    if (!c->src_set) return kiwi_set_error("no source parameters set");
    if (n != count_mask(c)) return kiwi_set_error("wrong number of subparams");   // minimizer_engine.f90:557-561
    const std::vector<float> p = apply_subparams(c, c->src_params, sub, false);
    return kiwi_set_source_params(c, c->src_type, (int)p.size(), p.data());
}

int kiwi_set_source_subparams_limits(kiwi_ctx* c, const float* mins, const float* maxs, int n) {
    if (!c) return kiwi_set_error("null context");
    if (!c->src_set) return kiwi_set_error("no source parameters set");
    if (n != count_mask(c)) return kiwi_set_error("wrong number of subparam_mins");   // minimizer_engine.f90:590-600
    c->sub_mins.assign(mins, mins + n); c->sub_maxs.assign(maxs, maxs + n);
    return 0;
}

int kiwi_get_source_subparams(kiwi_ctx* c, float* sub, int cap, int* n) {
    if (!c) return kiwi_set_error("null context");
    if (!c->src_set) return kiwi_set_error("no source parameters set");
    const int k = count_mask(c);
    if (n) *n = k;
    if (cap < k) return kiwi_set_error("subparams buffer too small: need %d", k);
    int isub = 0;
    for (size_t i = 0; i < c->src_params.size(); i++) if (masked(c, i)) sub[isub++] = c->src_params[i];
    return 0;
}

int kiwi_lmdif_batched(kiwi_lm_fcn fcn, void* user, int m, int n, float* x, float* fvec, float ftol, float xtol, float gtol, int maxfev, float epsfcn,
                       float* diag, int mode, float factor, int* info, int* nfev) {
    if (!fcn || !x || !fvec || !diag) return kiwi_set_error("null argument");
    const klm::Result r = klm::lmdif_batched([&](int ncols, float* xs, float* fs) { return fcn(user, ncols, n, m, xs, fs); }, m, n, x, fvec, ftol,
                                             xtol, gtol, maxfev, epsfcn, diag, mode, factor);
    if (info) *info = r.info;
    if (nfev) *nfev = r.nfev;
    return 0;
}

int kiwi_minimize_lm(kiwi_ctx* c, int* info_out, int* iterations_out, float* misfit_out) {
    if (!c) return kiwi_set_error("null context");
    CU_OK(cudaSetDevice(c->device));
    if (ensure_single(c, true)) return 1;   // update_misfits, minimizer_engine.f90:735
    const int n = count_mask(c);
    int m = 0;                              // get_nmisfits: components of ALL receivers (disabled ones hold zeros, receiver.f90:430)
    for (const HostReceiver& h : c->rcv) m += h.ncomp;
    if (n <= 0 || m < n) return kiwi_set_error("something went wrong in minimize_lm");   // the reference dies here (:778-780)
    const std::vector<float>& norm = params_norm(c->src_type);
    const size_t np = c->src_params.size();
    std::vector<float> snorm;               // psm_get_subparams_norm
    for (size_t i = 0; i < np; i++) if (masked(c, i)) snorm.push_back(norm[i]);
    std::vector<float> x(n), fvec(m), diag(n, 1.f);
    { int k = 0; for (size_t i = 0; i < np; i++) if (masked(c, i)) x[k++] = c->src_params[i] / norm[i]; }   // psm_get_subparams(normalized)
    const bool limits = !c->sub_mins.empty() && !c->sub_maxs.empty();
    const int nm = c->nmisfits;
    int iterations = 0;
    bool failed_hard = false;
    // lm_forward_step (minimizer_engine.f90:808-874) for `ncols` vectors at once
    auto forward = [&](int ncols, float* xs, float* fs) -> int {
        std::vector<float> penalty(ncols, 0.f), params((size_t)ncols * np);
        std::vector<float> cur = c->src_params;   // psm%params as the sequential reference would hold them before each step
        for (int col = 0; col < ncols; col++) {
            float* xc = xs + (size_t)col * n;
            if (limits)
                for (int i = 0; i < n; i++) {
                    if (xc[i] * snorm[i] < c->sub_mins[i]) {
                        penalty[col] = penalty[col] + fabsf(xc[i] * snorm[i] - c->sub_mins[i]) / fabsf(c->sub_maxs[i] - c->sub_mins[i]);
                        xc[i] = c->sub_mins[i] / snorm[i];
                    }
                    if (xc[i] * snorm[i] > c->sub_maxs[i]) {
                        penalty[col] = penalty[col] + fabsf(xc[i] * snorm[i] - c->sub_maxs[i]) / fabsf(c->sub_maxs[i] - c->sub_mins[i]);
                        xc[i] = c->sub_maxs[i] / snorm[i];
                    }
                }
            cur = apply_subparams(c, cur, xc, true);   // every step starts from the parameters the previous one left
            std::copy(cur.begin(), cur.end(), params.begin() + (size_t)col * np);
        }
        std::vector<int> status(ncols, 0);
        CU_OK(c->d_out.ensure(sizeof(float) * 2 * (size_t)std::max(nm, 1) * ncols));
        c->last_eval_ns = 0;   // (d_out is rewritten)
        if (eval_batch(c, c->src_type, ncols, (int)np, params.data(), c->d_out.as<float>(), status.data(), true)) { failed_hard = true; return 0; }
        std::vector<float> mis((size_t)2 * nm * ncols, 0.f);
        if (nm > 0) CU_OK(cudaMemcpy(mis.data(), c->d_out.p, sizeof(float) * mis.size(), cudaMemcpyDeviceToHost));
        int nok = 0;
        for (; nok < ncols; nok++) {
            // update_misfits fails (discretisation) or a misfit is not finite (get_misfits :1163): the forward step reports iflag = -2
            if (status[nok] != KIWI_STATUS_OK) break;
            float* f = fs + (size_t)nok * m;
            const float* mm = mis.data() + (size_t)nok * 2 * nm;
            int k = 0;
            for (size_t ir = 0; ir < c->rcv.size(); ir++) {
                const HostReceiver& h = c->rcv[ir];
                for (int ic = 0; ic < h.ncomp; ic++) f[k++] = (h.enabled ? mm[2 * (c->h_rcvdev[ir].misfit_base + ic)] : 0.f) * (1.0f + penalty[nok]);
            }
            iterations++;
        }
        // the source and its misfits stay at the last model evaluated (the sequential reference stops at the first failure)
        const int last = std::min(nok, ncols - 1);
        c->src_params.assign(params.begin() + (size_t)last * np, params.begin() + (size_t)(last + 1) * np);
        c->src_misfits.assign(mis.begin() + (size_t)last * 2 * nm, mis.begin() + (size_t)(last + 1) * 2 * nm);
        c->src_status = status[last];
        c->src_dirty = ncols > 1 || nok < ncols;   // the single-source tables (seismograms, indices) describe a batch: rebuild on demand
        c->last.valid = c->last.valid && ncols == 1;
        return nok;
    };
    const float tol = sqrtf(1.192091E-07f);   // sqrt(spmpar(1)), minimizer_engine.f90:773
    const int maxfev = 500 * (n + 1);
    const klm::Result r = klm::lmdif_batched(forward, m, n, x.data(), fvec.data(), tol, tol, 0.f, maxfev, 0.f, diag.data(), 2, 0.01f);
    if (failed_hard) return 1;
    int info = r.info;
    if (info == 8) info = 4;                  // :799
    if (info_out) *info_out = info;
    if (iterations_out) *iterations_out = iterations;
    if (misfit_out) {
        if (c->src_misfits.empty()) *misfit_out = NAN;
        else kiwi_global_misfits(1, nm, c->src_misfits.data(), misfit_out);
    }
    return 0;
}

int kiwi_get_floating_shifts(kiwi_ctx* c, int* shifts, int cap, int* n) {   // minimizer_engine.f90:1095-1128
    if (!c) return kiwi_set_error("null context");
    CU_OK(cudaSetDevice(c->device));
    if (c->misfit_method < KIWI_FLOATING_L2NORM) return kiwi_set_error("floating shifts need a floating misfit method");
    if (ensure_single(c, true)) return 1;
    int k = 0;
    for (size_t i = 0; i < c->rcv.size(); i++) {
        if (!c->rcv[i].enabled) continue;
        if (k < cap) shifts[k] = i < c->last_fshift.size() ? c->last_fshift[i] : 0;
        k++;
    }
    if (n) *n = k;
    return 0;
}

namespace {
// receiver_shift_ref_seismogram (receiver.f90:802-814) under the fresh-state semantics: the data span of every component moves
void shift_refs(kiwi_ctx* c, int ir, int ishift) {
    HostReceiver& h = c->rcv[ir];
    for (int k = 0; k < h.ncomp; k++) if (h.has_ref[k]) { h.ref_ds0[k] += ishift; h.ref_ds1[k] += ishift; }
    c->receivers_dirty = true; c->src_dirty = true;   // dirtyfy_ref_probes
}
}  // namespace

int kiwi_shift_ref_seismogram(kiwi_ctx* c, int ireceiver, float shift) {   // minimizer_engine.f90:354-378
    if (!c) return kiwi_set_error("null context");
    if (require_receivers(c)) return 1;
    if (!all_refs_set(c)) return kiwi_set_error("no reference seismograms set");
    if (ireceiver < 1 || ireceiver > (int)c->rcv.size()) return kiwi_set_error("receiver index out of range");
    shift_refs(c, ireceiver - 1, (int)lroundf(shift / c->db.dt));
    return 0;
}

namespace {
// receiver_calculate_cross_correlations (receiver.f90:597-616) for all receivers of the ns = 1 source on the device; ish: best shift per
// receiver by the rule of receiver.f90:827, cc: [receiver][KIWI_MAX_COMP][nshift] (may be null)
int cross_correlate(kiwi_ctx* c, int xs0, int xs1, int premethod, std::vector<int>* ish, std::vector<float>* cc) {
    const int nrcv = (int)c->rcv.size();
    const int nshift = xs1 - xs0 + 1;
    if (nshift < 1) return kiwi_set_error("empty shift range");
    if (ish) ish->assign((size_t)nrcv, 0);
    if (cc) cc->assign((size_t)nrcv * KIWI_MAX_COMP * nshift, 0.f);
    if (!c->last.seis_valid || nrcv == 0) return 0;
    const size_t ncc = (size_t)nrcv * KIWI_MAX_COMP * nshift;
    CU_OK(c->d_fshift.ensure(sizeof(int) * (size_t)nrcv));
    CU_OK(c->d_xcorr.ensure(sizeof(float) * ncc));
    CU_OK(cudaMemsetAsync(c->d_fshift.p, 0, sizeof(int) * (size_t)nrcv, c->stream));
    CU_OK(cudaMemsetAsync(c->d_xcorr.p, 0, sizeof(float) * ncc, c->stream));
    if (run_misfit_general(c, KIWI_INTERNAL_XCORR, xs0, xs1, c->last.syn_lo, c->last.syn_hi, c->last.tmax, c->d_cands.as<CandDev>(), 1,
                           c->last.seis_stride, c->d_shdrs.as<SeisHdr>(), c->nmisfits, c->d_xcorr.as<float>(), c->d_status.as<int>(),
                           c->d_fshift.as<int>(), nullptr, premethod))
        return 1;
    if (ish) CU_OK(cudaMemcpyAsync(ish->data(), c->d_fshift.p, sizeof(int) * (size_t)nrcv, cudaMemcpyDeviceToHost, c->stream));
    if (cc) CU_OK(cudaMemcpyAsync(cc->data(), c->d_xcorr.p, sizeof(float) * ncc, cudaMemcpyDeviceToHost, c->stream));
    CU_OK(cudaStreamSynchronize(c->stream));
    CU_OK(cudaGetLastError());
    return 0;
}
}  // namespace

int kiwi_autoshift_ref_seismogram(kiwi_ctx* c, int ireceiver, float shift_lo, float shift_hi, float* shifts, int cap, int* n) {   // :380-416
    if (!c) return kiwi_set_error("null context");
    CU_OK(cudaSetDevice(c->device));
    if (ensure_single(c, true)) return 1;   // update_misfits
    const int nrcv = (int)c->rcv.size();
    if (ireceiver != 0 && (ireceiver < 1 || ireceiver > nrcv)) return kiwi_set_error("receiver index out of range");
    std::vector<int> ish;
    if (cross_correlate(c, (int)lroundf(shift_lo / c->db.dt), (int)lroundf(shift_hi / c->db.dt), c->misfit_method, &ish, nullptr)) return 1;
    const int i0 = ireceiver == 0 ? 0 : ireceiver - 1, i1 = ireceiver == 0 ? nrcv : ireceiver;
    int k = 0;
    for (int i = i0; i < i1; i++) {
        const int ishift = c->rcv[i].enabled ? ish[i] : 0;   // receiver.f90:823-824
        if (k < cap && shifts) shifts[k] = (float)ishift * c->db.dt;
        k++;
        if (c->rcv[i].enabled) shift_refs(c, i, ishift);
    }
    if (n) *n = k;
    c->receivers_dirty = true; c->src_dirty = true;
    return 0;
}

// In-memory replacement of output_cross_correlations (minimizer_engine.f90:1283-1306; the reference only writes files)
int kiwi_get_cross_correlations(kiwi_ctx* c, int ireceiver, float shift_lo, float shift_hi, float* cc_out, int cap, int* ncomp, int* nshift) {
    if (!c) return kiwi_set_error("null context");
    CU_OK(cudaSetDevice(c->device));
    if (ensure_single(c, false)) return 1;   // update_syn_probes
    if (!all_refs_set(c)) return kiwi_set_error("no reference seismograms set");
    const int nrcv = (int)c->rcv.size();
    if (ireceiver < 1 || ireceiver > nrcv) return kiwi_set_error("receiver index out of range");
    const int xs0 = (int)lroundf(shift_lo / c->db.dt), xs1 = (int)lroundf(shift_hi / c->db.dt);
    std::vector<float> cc;
    if (cross_correlate(c, xs0, xs1, -1, nullptr, &cc)) return 1;
    const int ns = xs1 - xs0 + 1, nc = c->rcv[ireceiver - 1].enabled ? c->rcv[ireceiver - 1].ncomp : 0;
    if (cap < nc * ns) return kiwi_set_error("buffer too small for the cross-correlations");
    for (int k = 0; k < nc * ns; k++) cc_out[k] = cc[(size_t)(ireceiver - 1) * KIWI_MAX_COMP * ns + k];
    if (ncomp) *ncomp = nc;
    if (nshift) *nshift = ns;
    return 0;
}

namespace {
int export_probe(kiwi_ctx* c, int ireceiver, int icomponent, int which_probe, int which_processing, int spectrum, int* first_index, int* n, float* df,
                 float* buf, int cap) {
    if (!c) return kiwi_set_error("null context");
    CU_OK(cudaSetDevice(c->device));
    if (which_probe != 0 && which_probe != 1) return kiwi_set_error("unknown probe: use 0 (synthetics) or 1 (references)");
    if (which_processing < 0 || which_processing > 2) return kiwi_set_error("unknown processing: use 0 (plain), 1 (tapered) or 2 (filtered)");
    if (ensure_single(c, false)) return 1;   // update_syn_probes
    const int nrcv = (int)c->rcv.size();
    if (ireceiver < 1 || ireceiver > nrcv) return kiwi_set_error("receiver index out of range");
    const HostReceiver& h = c->rcv[ireceiver - 1];
    if (icomponent < 1 || icomponent > h.ncomp) return kiwi_set_error("component index out of range");
    if (which_probe == 1 && !h.has_ref[icomponent - 1]) return kiwi_set_error("no reference seismograms set");
    if (which_probe == 0 && !c->last.seis_valid) return kiwi_set_error("no synthetic seismogram available");
    if (ensure_twiddles(c)) return 1;
    const int n_alloc = 16384;   // 128 KiB of shared memory
    CU_OK(c->d_xcorr.ensure(sizeof(float) * (size_t)n_alloc));
    CU_OK(c->d_fshift.ensure(sizeof(int) * 4));
    cudaError_t e = launch_probe_export(c->d_rcv.as<ReceiverDev>(), ireceiver - 1, icomponent - 1, c->d_cands.as<CandDev>(), c->d_seis.as<float>(),
                                        c->last.seis_stride, c->d_shdrs.as<SeisHdr>(), nrcv, c->d_refdata.as<float>(), c->d_taper.as<float>(),
                                        (const float2*)c->d_tw.p, c->tw_n, which_probe, which_processing, spectrum, c->db.dt, n_alloc,
                                        c->d_fshift.as<int>(), c->d_xcorr.as<float>(), c->stream);
    if (e != cudaSuccess) return kiwi_set_error("CUDA error launching the probe export: %s", cudaGetErrorString(e));
    int hdr[4] = {0, 0, 0, 0};
    CU_OK(cudaMemcpyAsync(hdr, c->d_fshift.p, sizeof hdr, cudaMemcpyDeviceToHost, c->stream));
    CU_OK(cudaStreamSynchronize(c->stream));
    CU_OK(cudaGetLastError());
    if (hdr[3] == 1) return kiwi_set_error("no synthetic seismogram available");
    if (hdr[3] == 2) return kiwi_set_error("probe span does not fit the shared-memory transform");
    if (first_index) *first_index = hdr[0];
    if (n) *n = hdr[1];
    if (df) memcpy(df, &hdr[2], sizeof(float));
    if (hdr[1] > cap) return kiwi_set_error("buffer too small: need %d samples", hdr[1]);
    if (hdr[1] > 0) CU_OK(cudaMemcpy(buf, c->d_xcorr.p, sizeof(float) * (size_t)hdr[1], cudaMemcpyDeviceToHost));
    return 0;
}
}  // namespace

// In-memory replacement of output_seismograms for all its variants (minimizer_engine.f90:947-1012, receiver.f90:616-661, probe_get
// comparator.f90:356-433): which_probe 0 synthetics / 1 references, which_processing 0 plain / 1 tapered / 2 filtered
int kiwi_get_probe(kiwi_ctx* c, int ireceiver, int icomponent, int which_probe, int which_processing, int* first_index, int* n, float* buf, int cap) {
    return export_probe(c, ireceiver, icomponent, which_probe, which_processing, 0, first_index, n, nullptr, buf, cap);
}
// ... and of output_seismogram_spectra (minimizer_engine.f90:1014-1067, probe_get_amp_spectrum comparator.f90:332-354): n amplitudes at k * df
int kiwi_get_probe_spectrum(kiwi_ctx* c, int ireceiver, int icomponent, int which_probe, int which_processing, float* df, int* n, float* buf, int cap) {
    return export_probe(c, ireceiver, icomponent, which_probe, which_processing, 1, nullptr, n, df, buf, cap);
}

// eikonal_solver_fmm (eikonal.f90:29-199) as the eikonal sources run it on the host; speed and times are (nx, ny) with ix fastest.
// Needs no GPU (used by the CPU tests: the reference's own test_eikonal.f90 and bit-exactness against the restatement).
int kiwi_eikonal_fmm(int nx, int ny, const float* speed, const float* origin2, const float* delta2, const float* initialpoint2, float* times) {
    if (nx < 1 || ny < 1 || !speed || !times) return kiwi_set_error("kiwi_eikonal_fmm: invalid grid");
    kh::eikonal_solver_fmm(speed, nx, ny, origin2, delta2, initialpoint2, times);
    return 0;
}

// get_distances (minimizer_engine.f90:1260-1281): epicentral distance [m] and azimuth [rad] of every receiver, in double
int kiwi_get_distances(kiwi_ctx* c, double* distances, double* azimuths, int cap, int* n) {
    if (!c) return kiwi_set_error("null context");
    if (!c->loc_set) return kiwi_set_error("no source location set");
    if (require_receivers(c)) return 1;
    const int nr = (int)c->rcv.size();
    for (int i = 0; i < nr && i < cap; i++) {
        double azi, bazi;
        kh::azibazi(c->olat, c->olon, c->rcv[i].lat, c->rcv[i].lon, &azi, &bazi);
        if (azimuths) azimuths[i] = azi;
        if (distances) distances[i] = kh::distance_accurate50m(c->olat, c->olon, c->rcv[i].lat, c->rcv[i].lon);
    }
    if (n) *n = nr;
    return 0;
}

// get_source_crustal_thickness (minimizer_engine.f90:488-498, parameterized_source.f90:207-221)
int kiwi_get_source_crustal_thickness(kiwi_ctx* c, float* thickness) {
    if (!c) return kiwi_set_error("null context");
    if (!c->loc_set) return kiwi_set_error("no source location set");
    if (!c->crust.loaded) return kiwi_set_error("crust2x2 model not loaded");
    std::vector<kh::Halfspace> hs;
    kh::default_constraints(c->crust, c->olat, c->olon, c->thickness_limit, &hs);   // the second half-space sits at that depth
    *thickness = hs[1].point[2];
    return 0;
}

// get_principal_axes (minimizer_engine.f90:1248-1258): p- and t-axis (azimuth, polar angle; degrees) of the source set by
// kiwi_set_source_params.  Only the sources with a slip direction have them (bilateral, circular, eikonal); the others answer zeros,
// as psm%pax / psm%tax are never set for them.
int kiwi_get_principal_axes(kiwi_ctx* c, float* pax2, float* tax2) {
    if (!c) return kiwi_set_error("null context");
    if (!c->src_set) return kiwi_set_error("no source parameters set");
    pax2[0] = pax2[1] = tax2[0] = tax2[1] = 0.f;
    const int t = c->src_type;
    if (t == KIWI_SOURCE_BILATERAL || t == KIWI_SOURCE_CIRCULAR || t == KIWI_SOURCE_EIKONAL)
        kh::principal_axes(c->src_params[5], c->src_params[6], c->src_params[7], pax2, tax2);
    return 0;
}

int kiwi_get_seismogram(kiwi_ctx* c, int ireceiver, int icomponent, int which, int* first_index, int* n, float* buf, int cap) {
    if (!c) return kiwi_set_error("null context");
    CU_OK(cudaSetDevice(c->device));
    if (ensure_single(c, false)) return 1;
    if (ireceiver < 1 || ireceiver > (int)c->rcv.size()) return kiwi_set_error("receiver index out of range");
    const HostReceiver& h = c->rcv[ireceiver - 1];
    if (icomponent < 1 || icomponent > h.ncomp) return kiwi_set_error("component index out of range");
    if (!c->last.seis_valid) return kiwi_set_error("no synthetic seismogram available");
    const size_t item = (size_t)(ireceiver - 1) * KIWI_MAX_COMP + (icomponent - 1);
    SeisHdr sh;
    CU_OK(cudaMemcpy(&sh, c->d_shdrs.as<SeisHdr>() + item, sizeof sh, cudaMemcpyDeviceToHost));
    if (sh.hi < sh.lo) { *first_index = 0; *n = 0; return 0; }
    const int len = sh.hi - sh.lo + 1;
    *first_index = sh.lo; *n = len;
    const int m = std::min(len, cap);
    if (m > 0) {
        CU_OK(cudaMemcpy(buf, c->d_seis.as<float>() + item * c->last.seis_stride + (sh.lo - sh.base), sizeof(float) * m, cudaMemcpyDeviceToHost));
        if (which == 1) {   // probe_set_array(..., factor_=moment) comparator.f90:265
            const float moment = c->last.cands[0].moment;
            for (int i = 0; i < m; i++) buf[i] = buf[i] * moment;
        }
    }
    return 0;
}

int kiwi_discretize_source(kiwi_ctx* c, int sourcetype, int nparams, const float* params, float* table, int cap, int* ncentroids, int* grid3) {
    if (!c) return kiwi_set_error("null context");
    CU_OK(cudaSetDevice(c->device));
    if (kiwi_set_source_params(c, sourcetype, nparams, params)) return 1;
    if (ensure_single(c, false)) return 1;
    const CandDev& cd = c->last.cands[0];
    int ncent = 0;
    for (int v : c->last.g0_tap_count) ncent += v;
    if (ncentroids) *ncentroids = ncent;
    if (grid3) { grid3[0] = cd.nx; grid3[1] = cd.ny; grid3[2] = cd.nt; }
    const int m = std::min(ncent, cap);
    if (m > 0) {
        CU_OK(c->d_table.ensure(sizeof(float) * 10 * (size_t)m));
        launch_expand_centroids(cd, c->last.g, c->last.taps, c->last.ngroups_total, c->d_table.as<float>(), m, c->stream);
        CU_OK(cudaMemcpyAsync(table, c->d_table.p, sizeof(float) * 10 * (size_t)m, cudaMemcpyDeviceToHost, c->stream));
        CU_OK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int kiwi_get_indices(kiwi_ctx* c, int ireceiver, int* ix, int* iz, int* its, float* dix, float* diz, int* near_boundary, int cap, int* n) {
    if (!c) return kiwi_set_error("null context");
    CU_OK(cudaSetDevice(c->device));
    if (ensure_single(c, false)) return 1;
    if (ireceiver < 1 || ireceiver > (int)c->rcv.size()) return kiwi_set_error("receiver index out of range");
    const kiwi_ctx::Last& L = c->last;
    const CandDev& cd = L.cands[0];
    const int ng_ = cd.ngroups;
    std::vector<GeoRec> recs(std::max(ng_, 1));
    std::vector<float> tbase(std::max(ng_, 1));
    if (ng_ > 0) {
        CU_OK(cudaMemcpy(recs.data(), c->d_recs.as<GeoRec>() + (size_t)(ireceiver - 1) * L.rec_stride, sizeof(GeoRec) * ng_, cudaMemcpyDeviceToHost));
        CU_OK(cudaMemcpy(tbase.data(), L.g.tbase + cd.group_begin, sizeof(float) * ng_, cudaMemcpyDeviceToHost));
    }
    int k = 0;
    const float dt = c->db.dt;
    for (int ig = 0; ig < ng_; ig++)
        for (int it = 0; it < L.g0_tap_count[ig]; it++, k++) {
            if (k >= cap) continue;
            const float time = tbase[ig] + L.toff[L.g0_tap_begin[ig] + it];
            ix[k] = recs[ig].ix1; iz[k] = recs[ig].iz1; dix[k] = recs[ig].dix; diz[k] = recs[ig].diz;
            its[k] = (int)floorf(time / dt);   // sparse_trace.f90:640 on rshift = time/dt (seismogram.f90:139)
            if (near_boundary) near_boundary[k] = (recs[ig].flags & GEO_NEAR) ? 1 : 0;
        }
    if (n) *n = k;
    return 0;
}

int kiwi_get_spans(kiwi_ctx* c, int ireceiver, int* spans6) {
    if (!c) return kiwi_set_error("null context");
    CU_OK(cudaSetDevice(c->device));
    if (ensure_single(c, false)) return 1;
    if (ireceiver < 1 || ireceiver > (int)c->rcv.size()) return kiwi_set_error("receiver index out of range");
    PairHdr h;
    CU_OK(cudaMemcpy(&h, c->d_hdrs.as<PairHdr>() + (ireceiver - 1), sizeof h, cudaMemcpyDeviceToHost));
    spans6[0] = h.s1lo; spans6[1] = h.s1hi; spans6[2] = h.s2lo; spans6[3] = h.s2hi; spans6[4] = h.s3lo; spans6[5] = h.s3hi;
    return 0;
}

int kiwi_trace_span(kiwi_ctx* c, int ix, int iz, int ig, int* span2) {
    if (!c) return kiwi_set_error("null context");
    if (require_db(c)) return 1;
    if (ix < 1 || ix > c->db.nx || iz < 1 || iz > c->db.nz || ig < 1 || ig > c->db.ng) return kiwi_set_error("gfdb: invalid request: out of bounds");
    const int2 s = c->h_tspan[((size_t)(ix - 1) * c->db.nz + (iz - 1)) * c->db.ng + (ig - 1)];
    if (s.y < s.x) return kiwi_set_error("no trace available for index");
    span2[0] = s.x; span2[1] = s.y;
    return 0;
}

int kiwi_last_batch_bytes(kiwi_ctx* c, int max_candidates, double* b_alg, double* b_log, int* nsampled, long long* nskipped) {
    if (!c) return kiwi_set_error("null context");
    CU_OK(cudaSetDevice(c->device));
    const kiwi_ctx::Last& L = c->last;
    if (!L.valid) return kiwi_set_error("no batch has been evaluated");
    const int ns = std::max(1, std::min(max_candidates, L.n));
    const int ng = c->db.ng;
    double alg = 0., logi = 0.;
    long long skipped = 0;
    std::vector<GeoRec> recs(L.rec_stride);
    std::vector<PairHdr> hdrs((size_t)ns * L.nrcv);
    CU_OK(cudaMemcpy(hdrs.data(), c->d_hdrs.p, sizeof(PairHdr) * hdrs.size(), cudaMemcpyDeviceToHost));
    std::unordered_set<int> seen;
    for (int b = 0; b < ns; b++) {
        const CandDev& cd = L.cands[b];
        for (int ir = 0; ir < L.nrcv; ir++) {
            const ReceiverDev& R = c->h_rcvdev[ir];
            if (!R.enabled || cd.ngroups == 0) continue;
            const bool need_h = (R.ja | R.jr | R.jn | R.je) != 0, need_v = R.jd != 0;
            CU_OK(cudaMemcpy(recs.data(), c->d_recs.as<GeoRec>() + ((size_t)b * L.nrcv + ir) * L.rec_stride, sizeof(GeoRec) * cd.ngroups, cudaMemcpyDeviceToHost));
            seen.clear();
            for (int ig = 0; ig < cd.ngroups; ig++) {
                const GeoRec& r = recs[ig];
                if (r.flags & GEO_SKIP) { skipped += cd.nt; continue; }
                const int nco = (r.flags & GEO_SINGLE) ? 1 : 4;
                const int ix2 = r.ix1 + (c->interpolate ? c->xunder : 1), iz2 = r.iz1 + (c->interpolate ? c->zunder : 1);
                const int cx[4] = {r.ix1, r.ix1, ix2, ix2}, cz[4] = {r.iz1, iz2, r.iz1, iz2};
                for (int k = 0; k < nco; k++) {
                    const int inode = (cx[k] - 1) * c->db.nz + (cz[k] - 1);
                    double bytes = 0.;   // stored samples of the GF components this receiver uses
                    for (int j = 0; j < ng; j++) {
                        const bool horiz = (j < 5) || j == 8;
                        if (horiz ? !need_h : !need_v) continue;
                        const int2 sp = c->h_tspan[(size_t)inode * ng + j];
                        bytes += 4.0 * (sp.y - sp.x + 1);
                    }
                    logi += bytes * cd.nt;          // the reference fetches per centroid (seismogram.f90:131-254)
                    if (seen.insert(inode).second) alg += bytes;
                }
            }
            const PairHdr& h = hdrs[(size_t)b * L.nrcv + ir];
            const int s12lo = std::min(h.s1lo, h.s2lo), s12hi = std::max(h.s1hi, h.s2hi);
            for (int k = 0; k < R.ncomp; k++) {
                const int aid = abs(R.comp[k]);
                int lo, hi;
                if (aid == 1) { lo = h.s1lo; hi = h.s1hi; } else if (aid == 2) { lo = h.s2lo; hi = h.s2hi; }
                else if (aid == 3) { lo = h.s3lo; hi = h.s3hi; } else { lo = s12lo; hi = s12hi; }
                if (hi >= lo) { alg += 4.0 * (hi - lo + 1); logi += 4.0 * (hi - lo + 1); }
                if (R.ref_ds1[k] >= R.ref_ds0[k]) { alg += 4.0 * (R.ref_ds1[k] - R.ref_ds0[k] + 1); logi += 4.0 * (R.ref_ds1[k] - R.ref_ds0[k] + 1); }
            }
        }
    }
    if (b_alg) *b_alg = alg / ns;
    if (b_log) *b_log = logi / ns;
    if (nsampled) *nsampled = ns;
    if (nskipped) *nskipped = skipped;
    return 0;
}

void* kiwi_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { kiwi_set_error("cudaHostAlloc of %zu bytes failed", bytes); return nullptr; }
    return p;
}
void kiwi_host_free(void* p) { if (p) cudaFreeHost(p); }

int kiwi_last_timing(kiwi_ctx* c, float* ms5, int* launches4) {
    if (!c) return kiwi_set_error("null context");
    if (ms5) for (int i = 0; i < 5; i++) ms5[i] = c->ms[i];
    if (launches4) for (int i = 0; i < 4; i++) launches4[i] = c->launches[i];
    return 0;
}

}  // extern "C"
