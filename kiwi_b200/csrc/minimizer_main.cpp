// kiwi_minimizer -- command front-end speaking the text protocol of the reference's `minimizer`
// program (minimizer.f90:1676-1813) on top of the C ABI of libkiwi_b200.
//
// One command per stdin line, `#` comments and repeated blanks stripped (reduce_whitespace,
// minimizer.f90:1815-1846); the reply is "<cmd>: ok", "<cmd>: ok >" + answer line, "<cmd>: nok" or
// "<cmd>: nok >" + error line, flushed per command (:1682-1699), so the Python drivers of the
// reference (python/tunguska/seismosizer.py:306-338) can talk to it unchanged.  Only the commands on
// the hot path are implemented; the others answer "nok > unknown command".
// Additions: `set_database` takes a KGF1 file (HDF5 is not available here), and
// `eval_sources <type> <file>` evaluates a whole table of candidates in one call.
#include "../../include/kiwi_b200.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace {

std::string reduce_whitespace(const std::string& in) {   // minimizer.f90:1815-1846
    std::string out;
    bool ws = true;
    for (char ch : in) {
        if (ch != ' ' && ch != '\t') {
            if (ch == '#') break;
            out.push_back(ch); ws = false;
        } else if (!ws) { out.push_back(' '); ws = true; }
    }
    while (!out.empty() && out.back() == ' ') out.pop_back();
    return out;
}
std::vector<std::string> words(const std::string& s) {
    std::vector<std::string> w; std::istringstream is(s); std::string t;
    while (is >> t) w.push_back(t);
    return w;
}
bool to_floats(const std::vector<std::string>& w, size_t from, std::vector<float>* out) {
    out->clear();
    for (size_t i = from; i < w.size(); i++) {
        char* end = nullptr;
        const float v = strtof(w[i].c_str(), &end);
        if (end == w[i].c_str() || *end != 0) return false;
        out->push_back(v);
    }
    return true;
}
int source_id(const std::string& n) {   // source_all.f90:92-97
    if (n == "bilateral") return KIWI_SOURCE_BILATERAL;
    if (n == "circular") return KIWI_SOURCE_CIRCULAR;
    if (n == "point_lp") return KIWI_SOURCE_POINT_LP;
    if (n == "eikonal") return KIWI_SOURCE_EIKONAL;
    if (n == "mt_eikonal") return KIWI_SOURCE_MT_EIKONAL;
    if (n == "moment_tensor") return KIWI_SOURCE_MOMENT_TENSOR;
    return 0;
}
int norm_id(const std::string& n) {   // comparator.f90:33-42, comparator_get_norm_id
    static const char* names[] = {"l2norm", "l1norm", "ampspec_l2norm", "ampspec_l1norm", "scalar_product", "peak", "floating_l2norm", "floating_l1norm"};
    for (int i = 0; i < 8; i++) if (n == names[i]) return i + 1;
    return 0;
}
std::string fmt_floats(const float* v, size_t n) {   // list-directed output: blank separated reals
    std::string s; char buf[40];
    for (size_t i = 0; i < n; i++) { snprintf(buf, sizeof buf, "%s%.9g", i ? " " : " ", v[i]); s += buf; }
    return s;
}
// `table` seismogram file: two columns time value (seismogram_io.f90:123-136, 231-245)
bool read_table(const std::string& fn, float* tbegin, float* dt, std::vector<float>* data) {
    std::ifstream f(fn);
    if (!f) return false;
    std::vector<double> t; data->clear();
    double a, b;
    while (f >> a >> b) { t.push_back(a); data->push_back((float)b); }
    if (t.empty()) return false;
    *tbegin = (float)t[0];
    *dt = t.size() > 1 ? (float)((t.back() - t[0]) / (double)(t.size() - 1)) : 0.f;
    return true;
}

struct State {
    kiwi_ctx* ctx = nullptr;
    kiwi_gfdb* db = nullptr;
    std::vector<std::string> comps;   // component strings of the receivers
    float dt = 0.f;
    double ref_time = 0.;
    float olat = 0.f, olon = 0.f;     // source location as given (degrees)
    int src_type = 0;                 // the source set by set_source_params
    std::vector<float> src_params;
};

// returns ok; answer / err filled
bool do_command(State& S, const std::string& cmd, const std::vector<std::string>& w, std::string* answer, std::string* err) {
    auto fail = [&](const std::string& m) { *err = m; return false; };
    auto cfail = [&]() { *err = kiwi_last_error(); return false; };
    static const char* known[] = {"set_database", "set_local_interpolation", "set_spacial_undersampling", "set_receivers", "switch_receiver",
                                  "set_source_location", "set_source_constraints", "set_source_crustal_thickness_limit", "set_source_params",
                                  "set_effective_dt", "set_ref_seismograms", "set_misfit_method", "set_misfit_taper", "set_misfit_filter",
                                  "set_synthetics_factor", "set_floating_shiftrange", "get_misfits", "get_global_misfit", "get_floating_shifts",
                                  "output_seismograms", "eval_sources", "set_source_params_mask", "set_source_subparams",
                                  "set_source_subparams_limits", "get_source_subparams", "minimize_lm", "get_peak_amplitudes", "get_arias_intensities",
                                  "shift_ref_seismogram", "autoshift_ref_seismogram", "set_misfit_filter_1", "output_cross_correlations",
                                  "get_cached_traces_memory", "set_cached_traces_memory_limit", "set_verbose", "set_ignore_sigint",
                                  "get_principal_axes", "get_source_crustal_thickness", "output_distances", "output_seismogram_spectra",
                                  "output_source_model", "set_accumulation"};
    bool is_known = false;
    for (const char* k : known) if (cmd == k) is_known = true;
    if (!is_known) return fail("unknown command: " + cmd);   // minimizer.f90:1809-1811
    if (!S.ctx) {
        S.ctx = kiwi_create(0);
        if (!S.ctx) return cfail();
        const char* table = getenv("KIWI_CRUST2X2");
        if (table && *table) kiwi_set_crust2x2(S.ctx, table);   // crust2x2_load at start-up, minimizer.f90:1669-1674
    }
    std::vector<float> v;
    if (cmd == "set_database") {
        if (w.size() < 2) return fail("usage: set_database dbpath [ nipx nipz ]");
        // a Kiwi database is <dbpath>.index + <dbpath>.<i>.chunk (HDF5, gfdb.f90:209-211); a single file is this library's KGF1 dump
        kiwi_gfdb* db = nullptr;
        if (FILE* probe = fopen((w[1] + ".index").c_str(), "rb")) { fclose(probe); db = kiwi_gfdb_read_hdf(w[1].c_str()); }
        else db = kiwi_gfdb_read(w[1].c_str());
        if (!db) return cfail();
        if (w.size() >= 4 && (atoi(w[2].c_str()) != 1 || atoi(w[3].c_str()) != 1)) {   // Gulunay interpolation (minimizer.f90:98-112)
            kiwi_gfdb* ip = kiwi_gfdb_interpolate(db, atoi(w[2].c_str()), atoi(w[3].c_str()), 0);
            kiwi_gfdb_destroy(db);
            if (!ip) return cfail();
            db = ip;
        }
        if (kiwi_set_database(S.ctx, db)) { kiwi_gfdb_destroy(db); return cfail(); }
        if (S.db) kiwi_gfdb_destroy(S.db);
        S.db = db;
        kiwi_gfdb_meta(db, nullptr, nullptr, nullptr, &S.dt, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        return true;
    }
    if (cmd == "set_local_interpolation") {
        if (w.size() != 2 || (w[1] != "nearest_neighbor" && w[1] != "bilinear")) return fail("unknown interpolation method: " + (w.size() > 1 ? w[1] : std::string()));
        return kiwi_set_local_interpolation(S.ctx, w[1] == "bilinear") ? cfail() : true;
    }
    if (cmd == "set_spacial_undersampling") {
        if (w.size() != 3) return fail("usage: set_spacial_undersampling xunder zunder");
        return kiwi_set_spacial_undersampling(S.ctx, atoi(w[1].c_str()), atoi(w[2].c_str())) ? cfail() : true;
    }
    if (cmd == "set_receivers") {   // minimizer_engine.f90:165-286
        if (w.size() < 2) return fail("usage: set_receivers filename [ has_depth ]");
        const bool has_depth = w.size() > 2 && w[2] == "has_depth";
        std::ifstream f(w[1]);
        if (!f) return fail("can't open file: " + w[1]);
        std::vector<double> lat, lon; std::vector<float> dep; std::vector<std::string> comps;
        std::string line;
        while (std::getline(f, line)) {
            std::vector<std::string> t = words(reduce_whitespace(line));
            if (t.empty()) continue;
            const size_t need = has_depth ? 3 : 2;
            if (t.size() < need) return fail("failed to parse receivers file: " + w[1]);
            lat.push_back(atof(t[0].c_str())); lon.push_back(atof(t[1].c_str()));
            dep.push_back(has_depth ? (float)atof(t[2].c_str()) : 0.f);
            comps.push_back(t.size() > need ? t[need] : "ned");
        }
        std::vector<const char*> cp;
        for (auto& s : comps) cp.push_back(s.c_str());
        if (kiwi_set_receivers(S.ctx, (int)lat.size(), lat.data(), lon.data(), dep.data(), cp.data())) return cfail();
        S.comps = comps;
        return true;
    }
    if (cmd == "switch_receiver") {
        if (w.size() != 3 || (w[2] != "on" && w[2] != "off")) return fail("usage: switch_receiver ireceiver ( on | off )");
        return kiwi_switch_receiver(S.ctx, atoi(w[1].c_str()), w[2] == "on") ? cfail() : true;
    }
    if (cmd == "set_source_location") {
        if (w.size() != 4) return fail("usage: set_source_location latitude longitude reference-time");
        S.ref_time = atof(w[3].c_str());
        S.olat = (float)atof(w[1].c_str()); S.olon = (float)atof(w[2].c_str());
        return kiwi_set_source_location(S.ctx, (float)atof(w[1].c_str()), (float)atof(w[2].c_str()), S.ref_time) ? cfail() : true;
    }
    if (cmd == "set_source_constraints") {
        if (!to_floats(w, 1, &v) || v.size() % 6 != 0) return fail("usage: set_source_constraints px1 py1 pz1 nx1 ny1 nz1 ...");
        std::vector<float> p, n;
        for (size_t i = 0; i < v.size(); i += 6) { p.insert(p.end(), &v[i], &v[i] + 3); n.insert(n.end(), &v[i + 3], &v[i + 3] + 3); }
        return kiwi_set_source_constraints(S.ctx, (int)(v.size() / 6), p.data(), n.data()) ? cfail() : true;
    }
    if (cmd == "set_source_crustal_thickness_limit") {
        if (!to_floats(w, 1, &v) || v.size() != 1) return fail("usage: set_source_crustal_thickness_limit thickness-limit");
        return kiwi_set_source_crustal_thickness_limit(S.ctx, v[0]) ? cfail() : true;
    }
    if (cmd == "set_source_params") {   // minimizer.f90:636-692
        if (w.size() < 2) return fail("usage: set_source_params source-type source-params ...");
        const int st = source_id(w[1]);
        if (!st) return fail("unknown source type name: " + w[1]);
        const int np = kiwi_get_n_source_params(st);
        if (!to_floats(w, 2, &v)) return fail("failed to parse source params");
        if ((int)v.size() != np) return fail("source of type '" + w[1] + "' requires " + std::to_string(np) + " parameters.");
        if (kiwi_set_source_params(S.ctx, st, np, v.data())) return cfail();
        S.src_type = st; S.src_params = v;
        return true;
    }
    if (cmd == "output_source_model") {   // minimizer.f90:1085-1098, minimizer_engine.f90:948-978
        if (w.size() != 2) return fail("usage: output_source_model filenamebase");
        if (!S.src_type) return fail("no source parameters set");
        std::vector<float> table((size_t)10 << 20);
        int n = 0, grid[3] = {0, 0, 0};
        if (kiwi_discretize_source(S.ctx, S.src_type, (int)S.src_params.size(), S.src_params.data(), table.data(), (int)(table.size() / 10), &n, grid)) return cfail();
        // <base>-tdsm.info (discrete_source.f90:52-74)
        FILE* f = fopen((w[1] + "-tdsm.info").c_str(), "w");
        if (!f) return fail("failed to open output file: " + w[1] + "-tdsm.info");
        fprintf(f, "ncentroids\n %d\n\n", n);
        fclose(f);
        // <base>-dsm.table: north east depth time m(1:6) of every centroid (minimizer_engine.f90:965-976)
        f = fopen((w[1] + "-dsm.table").c_str(), "w");
        if (!f) return fail("failed to open output file: " + w[1] + "-dsm.table");
        for (int i = 0; i < n && (size_t)i < table.size() / 10; i++) fprintf(f, "%s\n", fmt_floats(&table[(size_t)i * 10], 10).c_str());
        fclose(f);
        // <base>-psm.info: the sections every source type writes first (origin in radians as psm%origin holds it, centre of the
        // source: e.g. source_bilat.f90:490-496); the type-specific drawing aids that follow there (outline, rupture and slip
        // arrows, eikonal grids) are not written
        f = fopen((w[1] + "-psm.info").c_str(), "w");
        if (!f) return fail("failed to open output file: " + w[1] + "-psm.info");
        const double d2r = (double)(2.f / 360.f * 3.14159265358979f);   // orthodrome.f90:334-350
        fprintf(f, "origin\n %.17g %.17g\n\ncenter\n%s\n\n", (double)S.olat * d2r, (double)S.olon * d2r, fmt_floats(&S.src_params[1], 3).c_str());
        fclose(f);
        return true;
    }
    if (cmd == "set_effective_dt") {
        if (!to_floats(w, 1, &v) || v.size() != 1) return fail("usage: set_effective_dt effective_dt");
        return kiwi_set_effective_dt(S.ctx, v[0]) ? cfail() : true;
    }
    if (cmd == "set_ref_seismograms") {   // minimizer_engine.f90:313-352, receiver.f90:746-801
        if (w.size() != 3) return fail("usage: set_ref_seismograms filenamebase fileformat");
        if (w[2] != "table") return fail("file format not available: " + w[2]);
        static const char names[11] = {'w', 's', 'u', 'l', 'c', '?', 'a', 'r', 'd', 'n', 'e'};
        (void)names;
        for (size_t ir = 0; ir < S.comps.size(); ir++)
            for (size_t ic = 0; ic < S.comps[ir].size(); ic++) {
                const std::string fn = w[1] + "-" + std::to_string(ir + 1) + "-" + S.comps[ir][ic] + "." + w[2];
                float tb, dtf; std::vector<float> data;
                if (!read_table(fn, &tb, &dtf, &data)) return fail("can't open file: " + fn);
                if (data.size() > 1 && S.dt > 0.f && fabs(dtf - S.dt) > 1e-4f * S.dt) return fail("sampling rate of seismogram does not match gfdb: " + fn);   // receiver.f90:776-781
                if (kiwi_set_ref_seismogram(S.ctx, (int)ir + 1, (int)ic + 1, (float)((double)tb - S.ref_time), (int)data.size(), data.data())) return cfail();
            }
        return true;
    }
    if (cmd == "set_misfit_method") {
        if (w.size() != 2 || !norm_id(w[1])) return fail("unknown norm method: " + (w.size() > 1 ? w[1] : std::string()));
        return kiwi_set_misfit_method(S.ctx, norm_id(w[1])) ? cfail() : true;
    }
    if (cmd == "set_misfit_taper") {
        if (!to_floats(w, 1, &v) || v.size() < 5 || v.size() % 2 != 1) return fail("failed to parse values");
        std::vector<float> x, y;
        for (size_t i = 1; i + 1 < v.size(); i += 2) { x.push_back(v[i]); y.push_back(v[i + 1]); }
        return kiwi_set_misfit_taper(S.ctx, (int)v[0], (int)x.size(), x.data(), y.data()) ? cfail() : true;
    }
    if (cmd == "set_misfit_filter") {
        if (!to_floats(w, 1, &v) || v.size() < 4 || v.size() % 2 != 0) return fail("failed to parse coordinates");
        std::vector<float> x, y;
        for (size_t i = 0; i + 1 < v.size(); i += 2) { x.push_back(v[i]); y.push_back(v[i + 1]); }
        return kiwi_set_misfit_filter(S.ctx, 0, (int)x.size(), x.data(), y.data()) ? cfail() : true;
    }
    if (cmd == "set_misfit_filter_1") {   // minimizer.f90:922-969: per-receiver filter, 0 = all
        if (!to_floats(w, 1, &v) || v.size() < 5 || v.size() % 2 != 1) return fail("failed to parse values");
        std::vector<float> x, y;
        for (size_t i = 1; i + 1 < v.size(); i += 2) { x.push_back(v[i]); y.push_back(v[i + 1]); }
        return kiwi_set_misfit_filter(S.ctx, (int)v[0], (int)x.size(), x.data(), y.data()) ? cfail() : true;
    }
    if (cmd == "output_cross_correlations") {   // minimizer.f90:1442-1482, minimizer_engine.f90:1283-1306: <base>-<ireceiver>-<component>.table
        if (!to_floats(w, 2, &v) || v.size() != 2) return fail("usage: output_cross_correlations filenamebase shift-min shift-max");
        std::vector<float> cc(5 * 8192);
        for (size_t ir = 0; ir < S.comps.size(); ir++) {
            int nc = 0, ns = 0;
            if (kiwi_get_cross_correlations(S.ctx, (int)ir + 1, v[0], v[1], cc.data(), (int)cc.size(), &nc, &ns)) return cfail();
            const long s0 = lroundf(v[0] / S.dt);
            for (int ic = 0; ic < nc; ic++) {   // disabled receivers write nothing (receiver.f90:721)
                const std::string fn = w[1] + "-" + std::to_string(ir + 1) + "-" + S.comps[ir][ic] + ".table";
                FILE* f = fopen(fn.c_str(), "w");
                if (!f) return fail("failed to write output file: " + fn);
                for (int i = 0; i < ns; i++) fprintf(f, "%.9g %.9g\n", (double)(s0 + i) * (double)S.dt, cc[(size_t)ic * ns + i]);   // receiver.f90:731-733
                fclose(f);
            }
        }
        return true;
    }
    if (cmd == "get_principal_axes") {   // minimizer.f90:1374-1402: pax(1) pax(2) tax(1) tax(2)
        float pax[2], tax[2];
        if (kiwi_get_principal_axes(S.ctx, pax, tax)) return cfail();
        const float v4[4] = {pax[0], pax[1], tax[0], tax[1]};
        *answer = fmt_floats(v4, 4);
        return true;
    }
    if (cmd == "get_source_crustal_thickness") {   // minimizer.f90 do_get_source_crustal_thickness
        float t = 0.f;
        if (kiwi_get_source_crustal_thickness(S.ctx, &t)) return cfail();
        *answer = fmt_floats(&t, 1);
        return true;
    }
    if (cmd == "output_distances") {   // minimizer.f90:1404-1440: distance [deg], distance [m], azimuth [deg] per receiver
        if (w.size() != 2) return fail("usage: output_distances filename");
        std::vector<double> d(S.comps.size() + 1), a(S.comps.size() + 1);
        int n = 0;
        if (kiwi_get_distances(S.ctx, d.data(), a.data(), (int)d.size(), &n)) return cfail();
        FILE* f = fopen(w[1].c_str(), "w");
        if (!f) return fail("failed to open file for output: " + w[1]);
        const double r2d = (double)(360.f / 2.f / 3.14159265358979f), earthradius = (double)(6371.f * 1000.f);   // orthodrome.f90:343-350, constants.f90:24
        for (int i = 0; i < n; i++) fprintf(f, "%.17g %.17g %.17g\n", r2d * (d[i] / earthradius), d[i], r2d * a[i]);
        fclose(f);
        return true;
    }
    if (cmd == "get_cached_traces_memory") {   // minimizer.f90:1484-1508: here the whole database is resident, in HBM
        long long nsamples = 0;
        if (S.db) kiwi_gfdb_meta(S.db, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &nsamples);
        *answer = std::to_string(nsamples * 4);
        return true;
    }
    if (cmd == "set_cached_traces_memory_limit" || cmd == "set_verbose" || cmd == "set_ignore_sigint") return true;   // nothing to steer here
    if (cmd == "set_accumulation") {   // extension: order of the floating-point operations of the synthesis (kiwi_set_accumulation)
        if (w.size() != 2 || (w[1] != "reference" && w[1] != "batched")) return fail("usage: set_accumulation reference|batched");
        return kiwi_set_accumulation(S.ctx, w[1] == "reference" ? 1 : 0) ? cfail() : true;
    }
    if (cmd == "set_synthetics_factor") {
        if (!to_floats(w, 1, &v) || v.size() != 1) return fail("usage: set_synthetics_factor factor");
        return kiwi_set_synthetics_factor(S.ctx, v[0]) ? cfail() : true;
    }
    if (cmd == "set_floating_shiftrange") {
        if (!to_floats(w, 1, &v) || v.size() != 3) return fail("usage: set_floating_shiftrange ireceiver min-shift max-shift");
        return kiwi_set_floating_shiftrange(S.ctx, (int)v[0], v[1], v[2]) ? cfail() : true;
    }
    if (cmd == "get_misfits") {
        const int nm = kiwi_get_nmisfits(S.ctx);
        std::vector<float> m((size_t)2 * (nm > 0 ? nm : 1));
        int n = 0;
        if (kiwi_get_misfits(S.ctx, m.data(), nm, &n)) return cfail();
        *answer = fmt_floats(m.data(), (size_t)2 * n);
        return true;
    }
    if (cmd == "get_global_misfit") {
        float g;
        if (kiwi_get_global_misfit(S.ctx, &g)) return cfail();
        *answer = fmt_floats(&g, 1);
        return true;
    }
    if (cmd == "set_source_params_mask") {   // minimizer.f90:694-736: logicals T/F (list-directed input also takes .true. / .false.)
        std::vector<int> mask;
        for (size_t i = 1; i < w.size(); i++) {
            std::string t = w[i];
            while (!t.empty() && t[0] == '.') t.erase(0, 1);
            if (t.empty() || !strchr("TtFf", t[0])) return fail("failed to parse source params mask");
            mask.push_back(t[0] == 'T' || t[0] == 't');
        }
        if (kiwi_set_source_params_mask(S.ctx, mask.data(), (int)mask.size())) return cfail();
        return true;
    }
    if (cmd == "set_source_subparams") {   // minimizer.f90:738-770
        if (!to_floats(w, 1, &v)) return fail("failed to parse subparams");
        if (kiwi_set_source_subparams(S.ctx, v.data(), (int)v.size())) return cfail();
        return true;
    }
    if (cmd == "set_source_subparams_limits") {   // minimizer.f90:772-812: all minima, then all maxima
        if (!to_floats(w, 1, &v) || v.size() % 2) return fail("failed to parse subparam limits");
        const int n = (int)v.size() / 2;
        if (kiwi_set_source_subparams_limits(S.ctx, v.data(), v.data() + n, n)) return cfail();
        return true;
    }
    if (cmd == "get_source_subparams") {   // minimizer.f90:1199-1224
        float sub[64]; int n = 0;
        if (kiwi_get_source_subparams(S.ctx, sub, 64, &n)) return cfail();
        *answer = fmt_floats(sub, (size_t)n);
        return true;
    }
    if (cmd == "get_peak_amplitudes" || cmd == "get_arias_intensities") {   // minimizer.f90 do_get_peak_amplitudes / do_get_arias_intensities
        std::vector<float> val(S.comps.size() + 1); int n = 0;
        if (cmd == "get_peak_amplitudes") {
            if (w.size() != 2) return fail("usage: get_peak_amplitudes differentiate");
            if (kiwi_get_peak_amplitudes(S.ctx, atoi(w[1].c_str()), val.data(), (int)val.size(), &n)) return cfail();
        } else if (kiwi_get_arias_intensities(S.ctx, val.data(), (int)val.size(), &n)) return cfail();
        *answer = fmt_floats(val.data(), (size_t)n);
        return true;
    }
    if (cmd == "minimize_lm") {   // minimizer.f90:1048-1081: answer = info iterations misfit
        int info = 0, iterations = 0; float misfit = 0.f;
        if (kiwi_minimize_lm(S.ctx, &info, &iterations, &misfit)) return cfail();
        *answer = std::to_string(info) + " " + std::to_string(iterations) + " " + fmt_floats(&misfit, 1);
        return true;
    }
    if (cmd == "shift_ref_seismogram") {   // minimizer.f90:356-386
        if (!to_floats(w, 1, &v) || v.size() != 2) return fail("usage: shift_ref_seismogram ireceiver shift");
        return kiwi_shift_ref_seismogram(S.ctx, (int)v[0], v[1]) ? cfail() : true;
    }
    if (cmd == "autoshift_ref_seismogram") {   // minimizer.f90:447-483: the applied shifts in seconds
        if (!to_floats(w, 1, &v) || v.size() != 3) return fail("usage: autoshift_ref_seismogram ireceiver min-shift max-shift");
        std::vector<float> f(S.comps.size() + 1); int n = 0;
        if (kiwi_autoshift_ref_seismogram(S.ctx, (int)v[0], v[1], v[2], f.data(), (int)f.size(), &n)) return cfail();
        *answer = fmt_floats(f.data(), (size_t)n);
        return true;
    }
    if (cmd == "get_floating_shifts") {   // minimizer_engine.f90:1095-1128: seconds
        std::vector<int> s(S.comps.size() + 1); int n = 0;
        if (kiwi_get_floating_shifts(S.ctx, s.data(), (int)s.size(), &n)) return cfail();
        std::vector<float> f(n);
        for (int i = 0; i < n; i++) f[i] = (float)s[i] * S.dt;
        *answer = fmt_floats(f.data(), f.size());
        return true;
    }
    if (cmd == "output_seismograms" || cmd == "output_seismogram_spectra") {   // minimizer_engine.f90:947-1067, `table` format
        const bool spec = cmd == "output_seismogram_spectra";
        if (w.size() < (spec ? 2u : 3u)) return fail(spec ? "usage: output_seismogram_spectra filenamebase (synthetics|references) (plain|tapered|filtered)"
                                                          : "usage: output_seismograms filenamebase fileformat (synthetics|references) (plain|tapered|filtered)");
        const size_t o = spec ? 2 : 3;   // index of the probe word
        if (!spec && w[2] != "table") return fail("file format not available: " + w[2]);
        const std::string probe = w.size() > o ? w[o] : "synthetics", proc = w.size() > o + 1 ? w[o + 1] : "plain";
        const int wp = probe == "synthetics" ? 0 : (probe == "references" ? 1 : -1);
        const int pr = proc == "plain" ? 0 : (proc == "tapered" ? 1 : (proc == "filtered" ? 2 : -1));
        if (wp < 0) return fail("unknown probe name: " + probe);
        if (pr < 0) return fail("unknown processing name: " + proc);
        std::vector<float> buf(1 << 16);
        for (size_t ir = 0; ir < S.comps.size(); ir++)
            for (size_t ic = 0; ic < S.comps[ir].size(); ic++) {
                int first = 0, n = 0; float df = 0.f;
                if (spec ? kiwi_get_probe_spectrum(S.ctx, (int)ir + 1, (int)ic + 1, wp, pr, &df, &n, buf.data(), (int)buf.size())
                         : kiwi_get_probe(S.ctx, (int)ir + 1, (int)ic + 1, wp, pr, &first, &n, buf.data(), (int)buf.size())) return cfail();
                const std::string fn = w[1] + "-" + std::to_string(ir + 1) + "-" + S.comps[ir][ic] + "." + (spec ? "table" : w[2]);
                FILE* f = fopen(fn.c_str(), "w");
                if (!f) return fail("failed to write output file: " + fn);
                for (int i = 0; i < n; i++)   // receiver.f90:649 (time of sample `first`: reftime + (first-1) dt) and :688 (frequency k df)
                    fprintf(f, "%.9g %.9g\n", spec ? (double)i * (double)df : S.ref_time + (double)(first - 1 + i) * (double)S.dt, buf[i]);
                fclose(f);
            }
        return true;
    }
    if (cmd == "eval_sources") {   // batched evaluation: one candidate per line of the file, answer = global misfits
        if (w.size() != 3) return fail("usage: eval_sources source-type filename");
        const int st = source_id(w[1]);
        if (!st) return fail("unknown source type name: " + w[1]);
        const int np = kiwi_get_n_source_params(st);
        std::ifstream f(w[2]);
        if (!f) return fail("can't open file: " + w[2]);
        std::vector<float> params; std::string line;
        while (std::getline(f, line)) {
            std::vector<std::string> t = words(reduce_whitespace(line));
            if (t.empty()) continue;
            std::vector<float> p;
            if (!to_floats(t, 0, &p) || (int)p.size() != np) return fail("failed to parse source params");
            params.insert(params.end(), p.begin(), p.end());
        }
        const int ns = (int)(params.size() / (size_t)np), nm = kiwi_get_nmisfits(S.ctx);
        std::vector<float> mis((size_t)ns * nm * 2 + 2), g(ns + 1);
        std::vector<int> status(ns + 1);
        if (kiwi_eval_sources(S.ctx, st, ns, np, params.data(), mis.data(), status.data())) return cfail();
        kiwi_global_misfits(ns, nm, mis.data(), g.data());
        *answer = fmt_floats(g.data(), ns);
        return true;
    }
    return fail("unknown command: " + cmd);
}

}  // namespace

int main() {
    State S;
    std::string line;
    while (std::getline(std::cin, line)) {
        const std::string args = reduce_whitespace(line);
        if (args.empty()) continue;
        const std::vector<std::string> w = words(args);
        std::string answer, err;
        const bool ok = do_command(S, w[0], w, &answer, &err);
        if (ok) {
            if (answer.empty()) printf("%s: ok\n", w[0].c_str());
            else printf("%s: ok >\n%s\n", w[0].c_str(), answer.c_str());
        } else {
            if (err.empty()) printf("%s: nok\n", w[0].c_str());
            else printf("%s: nok >\n%s\n", w[0].c_str(), err.c_str());
        }
        fflush(stdout);
    }
    if (S.ctx) kiwi_destroy(S.ctx);
    if (S.db) kiwi_gfdb_destroy(S.db);
    return 0;
}
