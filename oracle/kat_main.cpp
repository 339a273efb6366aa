// TEST INFRASTRUCTURE ONLY (see ko_base.hpp).  The reference's own known-answer tests, restated
// against the oracle.  Every check cites the reference test it mirrors.  Prints one line per
// failure ("fail: <test>: <label>", the reference's util.f90:186-198 format) and a final
// "kat: N checks, M failures"; exit status = (M != 0).
#include "ko_engine.hpp"

using namespace ko;

static int nchecks = 0, nfail = 0;
static const char* cur = "";
static void begin(const char* name) { cur = name; }
static void check(bool ok, const char* label) {
    nchecks++;
    if (!ok) { nfail++; printf("fail: %s: %s\n", cur, label); }
}
static Strip mk(int s0, int s1, std::vector<float> v) { Strip s; strip_init(s0, s1, v.data(), (int)v.size(), s); return s; }
static bool eq(const Strip& s, std::vector<float> v) { return s.d == v; }
static bool nearf(float a, float b, float eps) { return fabsf(a - b) < eps; }
static bool neard(double a, double b, double eps) { return fabs(a - b) < eps; }

// test_sparse_trace.f90:30-123
static void test_sparse_trace() {
    begin("test_sparse_trace");
    Strip cont1 = mk(21, 40, {0, 0, 0, 1, 1, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0});
    Strip cont2 = mk(51, 51, {6});
    Trace sparse1, sparse2, joined, empty;
    trace_pack(cont1, sparse1);
    trace_pack(cont2, sparse2);
    check(sparse1.nstrips == 2, "nstrips");                                                        // :37-38
    check(sparse1.strips[0].lo == 24 && sparse1.strips[0].hi() == 27, "span1");                    // :40-41
    check(sparse1.strips[1].lo == 33 && sparse1.strips[1].hi() == 40, "span2");                    // :42-43
    trace_join(sparse1, sparse2, joined);
    check(joined.span[0] == 24 && joined.span[1] == 51, "span3");                                  // :47-48
    trace_unpack(joined, cont1);
    check(cont1.lo == 24 && cont1.hi() == 51, "span4");                                            // :52-53
    check(cont1.at(cont1.hi()) == 6.f, "last value");                                              // :55-56
    Strip accu1, accu2;
    trace_multiply_add(sparse1, accu1); trace_multiply_add(sparse2, accu1);
    trace_multiply_add(sparse2, accu2); trace_multiply_add(sparse1, accu2);
    check(cont1.d == accu1.d, "muliply-add 1");                                                    // :63-64
    check(cont1.d == accu2.d, "muliply-add 2");                                                    // :66-67
    cont2 = mk(1, 4, {3, 1, 1, 99});
    trace_pack(cont2, sparse2);
    check(sparse2.strips[0].lo == 1 && sparse2.strips[0].hi() == 4, "pack (span)");                // :73-74
    check(eq(sparse2.strips[0], {3, 1, 1, 99}), "pack (data)");                                    // :75-76
    cont1 = mk(1, 2, {1, 1}); cont2 = mk(2, 3, {1, 1});
    trace_pack(cont2, sparse2);
    trace_multiply_add(sparse2, cont1, 1.f, SHIFT_INT, -1);
    check(cont1.lo == 1 && cont1.hi() == 2 && eq(cont1, {2, 2}), "multiply-add 4");                // :82-86
    cont1 = mk(-2, 2, {1, 0, 0, 0, 1});
    trace_pack(cont1, sparse1);
    trace_join(empty, sparse1, sparse2);
    trace_unpack(sparse2, cont2);
    check(cont2.d == cont1.d, "join w empty");                                                     // :88-93
    cont1 = mk(1, 1, {0}); cont2 = mk(2, 4, {1, 1, 0});
    trace_pack(cont2, sparse2);
    trace_multiply_add(sparse2, cont1, 1.f, SHIFT_REAL, 0, -0.25f);
    check(eq(cont1, {0.25f, 1.f, 0.75f, 0.f}), "multiply-add 6");                                  // :105-110
    int ds[2];
    cont1 = mk(-2, 5, {0, 0, 1, 2, 2, 2, 2, 2}); strip_dataspan(cont1, ds);
    check(ds[0] == 0 && ds[1] == 1, "strip dataspan 1");                                           // :113-115
    cont1 = mk(-2, 5, {1, 1, 1, 2, 2, 2, 2, 3}); strip_dataspan(cont1, ds);
    check(ds[0] == -2 && ds[1] == 5, "strip dataspan 2");                                          // :117-119
    cont1 = mk(-2, 0, {0, 0, 0}); strip_dataspan(cont1, ds);
    check(ds[0] == 0 && ds[1] == -2, "strip dataspan 3");                                          // :121-123
}

// test_comparator.f90:33-113
static void test_comparator() {
    begin("test_comparator");
    float dt = 1.f;
    Probe a, b; probe_init(a, dt); probe_init(b, dt);
    Strip s1 = mk(-1, 2, {0, 0, 5, 1}), s2 = mk(1, 4, {5, 1, 1, 1});
    probe_set_array(a, s1); probe_set_array(b, s2);
    check(probes_norm(a, b) == 0.f, "1");                                                          // :43-46
    check(probes_norm(a, b, L1NORM) == 0.f, "1b");
    s1 = mk(-1, 2, {0, 0, 0, 1}); s2 = mk(1, 4, {0, 1, 1, 1});
    probe_set_array(a, s1); probe_set_array(b, s2);
    check(probes_norm(a, b) == 0.f, "2");                                                          // :48-55
    check(probes_norm(a, b, L1NORM) == 0.f, "2b");
    float eps = 0.000001f;
    s1 = mk(-4, -1, {1, 0, 0, 0}); s2 = mk(1, 4, {1, 0, 0, 0});
    probe_set_array(a, s1); probe_set_array(b, s2);
    check(nearf(probes_norm(a, b), sqrtf(2.f), eps), "3");                                         // :59-66
    check(nearf(probes_norm(a, b, L1NORM), 2.f, eps), "3b");
    s1 = mk(0, 3, {1, 2, 1, 0}); s2 = mk(1, 4, {1, 1, 0, 1});
    probe_set_array(a, s1); probe_set_array(b, s2);
    check(nearf(probes_norm(a, b), sqrtf(3.f), eps), "4");                                         // :68-75
    check(nearf(probes_norm(a, b, L1NORM), 3.f, eps), "4b");
    s1 = mk(0, 3, {0, 1, 2, 1}); s2 = mk(1, 2, {1, 2});
    probe_set_array(a, s1); probe_set_array(b, s2);
    check(nearf(probes_norm(a, b), 1.f, eps), "5");                                                // :77-84
    check(nearf(probes_norm(a, b, L1NORM), 1.f, eps), "5b");
    s1 = mk(0, 4, {0, 1, 2, 1, 0}); s2 = mk(10, 14, {0, 1, 2, 1, 0});
    probe_set_array(a, s1); probe_set_array(b, s2);
    check(nearf(probes_norm(a, b, AMPSPEC_L2NORM), 0.f, eps), "6");                                // :87-92
    s1 = mk(1, 5, {0, 1, 2, 1, 0}); s2 = mk(2, 6, {0, 1, 2, 1, 0});
    probe_set_array(a, s1); probe_set_array(b, s2);
    int shiftrange[2] = {-5, 5};
    float cc[11];
    probes_windowed_cross_corr(a, b, shiftrange, cc);
    float want[11] = {0, 0, 1, 4, 6, 4, 1, 0, 0, 0, 0};
    bool ok = true; for (int i = 0; i < 11; i++) ok = ok && cc[i] == want[i];
    check(ok, "cross correlation");                                                                // :94-102
    Plf taper; plf_make(taper, {2.5f, 3.5f}, {1.f, 1.f});
    probe_set_taper(a, taper); probe_set_taper(b, taper);
    probes_windowed_cross_corr(a, b, shiftrange, cc);
    float want2[11] = {0, 0, 0, 2, 4, 2, 0, 0, 0, 0, 0};
    ok = true; for (int i = 0; i < 11; i++) ok = ok && cc[i] == want2[i];
    check(ok, "tapered cross correlation");                                                        // :104-111
}

// test_piecewise_linear_function.f90:28-75
static void test_plf() {
    begin("test_piecewise_linear_function");
    Plf f; plf_make(f, {0.f, 1.f, 2.f, 3.f}, {0.f, 1.f, 1.f, 0.f});
    check(2.f == plf_integrate(f, -1.f, 3.f), "1");
    check(1.f == plf_integrate(f, -1.f, 1.5f), "2");
    check(6.f / 8.f + 1.f == plf_integrate(f, 0.5f, 2.5f), "3");
    check(3.f / 8.f == plf_integrate(f, 2.f, 2.5f), "4");
    check(0.f == plf_integrate(f, 2.f, 2.f), "5");
    check(0.f == plf_integrate(f, 2.5f, 2.5f), "6");
    check(1.f / 8.f - 1.f / 32.f == plf_integrate(f, 2.5f, 2.75f), "7");
    check(1.f == plf_integrate(f, 1.f, 2.f), "8");
    float a, c;
    // the reference's checks 9-11 use `.and.` (nearly vacuous, :62-75); the intended values are
    // checked here with a tolerance instead
    plf_integrate_and_centroid(f, -1.f, 6.f, a, c); check(a == 2.f && nearf(c, 1.5f, 1e-6f), "9");
    plf_integrate_and_centroid(f, 0.f, 0.5f, a, c); check(a == 1.f / 8.f && nearf(c, 1.f / 3.f, 1e-6f), "10");
    plf_integrate_and_centroid(f, 0.f, 2.f, a, c); check(a == 3.f / 2.f && nearf(c, 1.f + 2.f / 9.f, 1e-6f), "11");
}

// test_source_bilat.f90:36-145
static void test_source_bilat() {
    begin("test_source_bilat");
    for (int icase = 0; icase < 2; icase++) {
        Psm psm; Tdsm tdsm; bool omc, ok;
        float p1[14] = {0, 0, 0, 1000, 1, 90, 45, 90, 0, 2000, 0, 1000, 2000, 1};
        float p2[14] = {0, 0, 0, 1000, 1, 45, 90, 0, 0, 2000, 0, 1000, 2000, 1};
        psm_set(psm, PSM_BILAT, icase == 0 ? p1 : p2, 14, omc);
        psm_to_tdsm(psm, tdsm, 0.5f, ok);
        float epsm = 1.f / (float)tdsm.centroids.size() / 100.f;
        float msum = 0.f; bool bad = false;
        for (auto& c : tdsm.centroids) {
            float mxx = c.m[0], myy = c.m[1], mzz = c.m[2], mxy = c.m[3], mxz = c.m[4], myz = c.m[5];
            msum += mxx;
            if (icase == 0) {  // 45deg dip: -mxx == mzz, all others zero (:66-75)
                if (nearf(mxx, 0, epsm) || nearf(mzz, 0, epsm) || !nearf(mxx, -mzz, epsm) || mxx > 0 || !nearf(myy, 0, epsm) ||
                    !nearf(mxy, 0, epsm) || !nearf(mxz, 0, epsm) || !nearf(myz, 0, epsm)) bad = true;
            } else {           // 90deg dip, strike 45: -mxx == myy (:118-127)
                if (nearf(mxx, 0, epsm) || nearf(myy, 0, epsm) || !nearf(mxx, -myy, epsm) || mxx > 0 || !nearf(mzz, 0, epsm) ||
                    !nearf(mxy, 0, epsm) || !nearf(mxz, 0, epsm) || !nearf(myz, 0, epsm)) bad = true;
            }
        }
        check(!bad, icase == 0 ? "thrust1" : "thrust2");
        check(nearf(-1.f, msum, 0.01f), icase == 0 ? "thrustsum1" : "thrustsum2");
    }
}

// test_orthodrome.f90:28-134
static void test_orthodrome() {
    begin("test_orthodrome");
    auto gc = [](double la, double lo) { GeoCoords g; g.lat = la; g.lon = lo; return g; };
    GeoCoords a = gc(0, 0), b = gc(90, 0);
    check(neard(0., r2d_d(azimuth(d2r_tgc(a), d2r_tgc(b))), 0.001), "azimuth 1");
    check(neard(90., r2d_d(arcdistance(d2r_tgc(a), d2r_tgc(b))), 0.001), "arcdistance 1");
    b = gc(0, 100);
    check(neard(90., r2d_d(azimuth(d2r_tgc(a), d2r_tgc(b))), 0.001), "azimuth 2");
    check(neard(100., r2d_d(arcdistance(d2r_tgc(a), d2r_tgc(b))), 0.001), "arcdistance 2");
    a = gc(0, 10); b = gc(0, -10);
    check(neard(-90., r2d_d(azimuth(d2r_tgc(a), d2r_tgc(b))), 0.001), "azimuth 3");
    check(neard(20., r2d_d(arcdistance(d2r_tgc(a), d2r_tgc(b))), 0.001), "arcdistance 3");
    a = gc(10, 0); b = gc(-10, -0.001);
    check(neard(-180., r2d_d(azimuth(d2r_tgc(a), d2r_tgc(b))), 0.1), "azimuth 4");
    check(neard(20., r2d_d(arcdistance(d2r_tgc(a), d2r_tgc(b))), 0.001), "arcdistance 4");
    a = gc(10, 0); b = gc(-10, 0);
    check(neard(180., r2d_d(azimuth(d2r_tgc(a), d2r_tgc(b))), 0.1), "azimuth 5");
    check(neard(20., r2d_d(arcdistance(d2r_tgc(a), d2r_tgc(b))), 0.001), "arcdistance 5");
    double km = 1000.;
    // literals in the reference are default real (:88-91, :99-102, :108-111)
    a = gc((double)53.556867f, (double)9.994622f); b = gc((double)48.139743f, (double)11.560050f);
    check(neard(distance_accurate50m(d2r_tgc(a), d2r_tgc(b)) / km, 612.59, 0.05), "hamburg-munich");
    a = gc((double)53.568391f, (double)9.973672f); b = gc((double)53.580049f, (double)9.956507f);
    check(neard(distance_accurate50m(d2r_tgc(a), d2r_tgc(b)) / km, 1.73, 0.05), "geomatikum-home");
    a = gc((double)52.5167f, (double)13.4000f); b = gc((double)35.7000f, (double)139.7667f);
    check(neard(distance_accurate50m(d2r_tgc(a), d2r_tgc(b)) / km, 8941.20671, 0.05), "berlin-tokio");
    double azi, bazi, dist;
    approx_differential_azidist((float)(10. * km), (float)(10. * km), d2r_d(90.), d2r_d(-90.), 20. * km, azi, bazi, dist);
    check(neard(r2d_d(azi), 135., 1.) && neard(dist, 10. * km * (double)sqrtf(2.f), 1.), "differential azidist 1");
    approx_differential_azidist((float)(10. * km), (float)(10. * km), d2r_d(90.), d2r_d(-90.), 1111.95 * km, azi, bazi, dist);
    check(neard(r2d_d(azi), 90., 1.) && neard(dist, 1111.95 * km - 10. * km, 50.), "differential azidist 2");
}

// test_euler.f90:25-60
static void test_euler() {
    begin("test_euler");
    float rot[3][3];
    auto matches = [&](const float want_cols[9]) {  // want given column-major like reshape()
        bool ok = true;
        for (int col = 0; col < 3; col++) {
            float e[3] = {0, 0, 0}; e[col] = 1.f; float out[3];
            matvec3(rot, e, out);
            for (int r = 0; r < 3; r++) ok = ok && nearf(out[r], want_cols[col * 3 + r], 0.001f);
        }
        return ok;
    };
    float rotbeta[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1}, rotalpha[9] = {1, 0, 0, 0, 0, 1, 0, -1, 0}, rotgamma[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1};
    init_euler(pi / 2.f, 0.f, 0.f, rot); check(matches(rotalpha), "alpha");
    init_euler(0.f, pi / 2.f, 0.f, rot); check(matches(rotbeta), "beta");
    init_euler(0.f, 0.f, pi / 2.f, rot); check(matches(rotgamma), "gamma");
}

// comparator.f90:1111-1118: the fp32-log next_power_of_two equals the exact integer one
static void test_next_pow2() {
    begin("next_power_of_two");
    bool ok = true;
    for (int n = 1; n <= (1 << 21); n++) {
        int want = 1; while (want < n) want <<= 1;
        if (next_power_of_two(n) != want) { ok = false; printf("  n=%d got %d want %d\n", n, next_power_of_two(n), want); break; }
    }
    check(ok, "fp32-log vs integer for n<=2^21");
}

// test_eikonal.f90:26-56: uniform 500x1000 grid, corner times within one cell
static void test_eikonal() {
    begin("test_eikonal");
    const int nx = 500, ny = 1000;
    Field speed, times; speed.alloc(nx, ny, 2.f);
    float delta[2] = {50.f / nx, 50.f / ny}, initialpoint[2] = {0.f, 25.f}, origin[2] = {0.f, 0.f};
    float eps = std::max(delta[0], delta[1]) / 2.f;
    eikonal_solver_fmm(speed, origin, delta, initialpoint, times);
    check(nearf(times(1, 1), 12.5f, eps), "1,1");          // :45-47
    check(nearf(times(1, ny), 12.5f, eps), "1,ny");        // :48-50
    check(nearf(times(nx, 1), 27.95f, eps), "nx,1");       // :51-53
    check(nearf(times(nx, ny), 27.95f, eps), "nx,ny");     // :54-56
}
// test_heap.f90:27-60
static void test_heap() {
    begin("test_heap");
    const int nn = 1000000;
    std::vector<float> keys(nn + 1); std::vector<int> back(nn + 1, 0);
    IndexHeapO h; initheap(h, nn);
    for (int i = 1; i <= nn; i++) keys[i] = (float)(nn - (i - 1));
    for (int i = 1; i <= nn; i++) pushheap(h, i, keys.data(), back.data());
    bool ok = true;
    for (int i = 1; i <= nn / 1000; i++) if (back[h.iheap[i]] != i) { ok = false; break; }
    check(ok, "backpointer");                              // :43-48
    ok = true;
    for (int i = 1; i <= nn / 1000; i++) { int j; popheap(h, j, keys.data(), back.data()); if (fabsf((float)(nn - nn + i) - keys[j]) > 0.00001f) { ok = false; break; } }
    check(ok, "popheap");                                  // :50-56
}
// test_geometry.f90:44-114
static void test_geometry() {
    begin("test_geometry");
    HalfSpace hs; hs.point = Vec3{{0.f, 0.f, -1.f}}; hs.normal = Vec3{{0.f, -1.f, -1.f}};
    Vec3 pts[4] = {{{0.f, 2.f, -1.f}}, {{0.f, -2.f, -1.f}}, {{0.f, 0.f, -1.f}}, {{0.f, 0.f, 0.f}}};
    bool expected[4] = {true, false, true, true};
    bool ok = true;
    for (int i = 0; i < 4; i++) ok = ok && (expected[i] == point_in_halfspace(pts[i], hs));
    check(ok, "point_in_halfspace");                       // :47-51
    Vec3 pp; bool between, par;
    get_piercingpoint(pts[0], pts[1], hs, pp, between, par);
    check(pp[0] == 0.f && pp[1] == 0.f && pp[2] == -1.f && between && !par, "piercingpoint 1");     // :53-57
    get_piercingpoint(Vec3{{0.f, 2.f, -1.f}}, Vec3{{0.f, 1.f, -2.f}}, hs, pp, between, par);
    check(pp[0] == 0.f && pp[1] == 1.f && pp[2] == -2.f && !between && !par, "piercingpoint 2");    // :60-64
    get_piercingpoint(Vec3{{0.f, 2.f, 5.f}}, Vec3{{0.f, 1.f, 1.f}}, hs, pp, between, par);
    check(nearf(pp[0], 0.f, 1e-4f) && nearf(pp[1], 0.4f, 1e-4f) && nearf(pp[2], -1.4f, 1e-4f) && !between && !par, "piercingpoint 3");   // :66-70
    get_piercingpoint(Vec3{{0.f, 1.f, 0.f}}, Vec3{{0.f, 2.f, -1.0001f}}, hs, pp, between, par);
    check(pp[0] == 0.f && pp[1] == 0.f && pp[2] == 0.f && !between && par, "piercingpoint 4");      // :72-76
    Circle circle;
    init_euler(d2r_r(45.f), d2r_r(45.f), 0.f, circle.transform);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) circle.transform[i][j] = circle.transform[i][j] * 3.f;
    circle.center = Vec3{{0.f, 0.f, 1.f}};
    PolygonPts poly, trimmed;
    circle_to_polygon(circle, 7, poly);
    hs.point = Vec3{{0.f, 0.f, 0.f}}; hs.normal = Vec3{{0.f, 0.f, -1.f}};
    trim_polygon(poly, hs, trimmed);
    const float expected_circ[21] = {0.14987442f, 2.4953687f, 2.6585152f, -1.9344299f, 0.9903534f, 3.0681345f, -2.5620692f, -1.2604182f,
                                     1.9204066f, -1.2604178f, -2.5620692f, 0.07959348f, -1.1043297f, -2.5185432f, 0.f,
                                     2.3468528f, 0.9326396f, 0.f, 2.12132f, 2.1213207f, 1.0000004f};
    ok = trimmed.size() == 7;
    for (size_t i = 0; ok && i < 7; i++) for (int k = 0; k < 3; k++) ok = ok && nearf(trimmed[i][k], expected_circ[3 * i + k], 0.00001f);
    check(ok, "trimmed circle");                           // :78-89
    init_euler(d2r_r(13.f), d2r_r(100.f), 0.f, circle.transform);
    circle.center = Vec3{{-5.f, -2.f, 1.f}};
    circle_to_polygon(circle, 3600, poly);
    check(nearf(polygon_area(poly), pi, 0.0001f), "pi estimation");   // :92-100
    PolygonPts sq1 = {{{0, 0, 1}}, {{0, 2, 1}}, {{2, 2, 1}}, {{2, 0, 1}}}, sq2 = {{{1, 0, 0}}, {{1, 0, 2}}, {{1, 2, 2}}, {{1, 2, 0}}},
               sq3 = {{{0, 1, 0}}, {{2, 1, 0}}, {{2, 1, 2}}, {{0, 1, 2}}};
    check(nearf(polygon_area(sq1), 4.f, 0.00001f), "square area 1");
    check(nearf(polygon_area(sq2), 4.f, 0.00001f), "square area 2");
    check(nearf(polygon_area(sq3), 4.f, 0.00001f), "square area 3");
}

int main() {
    test_sparse_trace();
    test_comparator();
    test_plf();
    test_source_bilat();
    test_orthodrome();
    test_euler();
    test_next_pow2();
    test_eikonal();
    test_heap();
    test_geometry();
    printf("kat: %d checks, %d failures\n", nchecks, nfail);
    return nfail != 0;
}
