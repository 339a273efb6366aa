// Host-side arithmetic of the engine (see host_math.cpp).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace kh {

float d2r_r(float deg);
double d2r_d(double deg);
void azibazi(double alat, double alon, double blat, double blon, double* azi, double* bazi);
double distance_accurate50m(double alat, double alon, double blat, double blon);
void final_rotation(double bazi0, float* cl0, float* sl0);
void init_euler(float alpha, float beta, float gamma, float* mat9);
// p- and t-axis (azimuth, polar angle in degrees) of a shear source (source_bilat.f90:232-237 and the same lines of the circular and
// eikonal sources; polar / domeshot / wrap :565-594)
void principal_axes(float strike_deg, float dip_deg, float rake_deg, float pax[2], float tax[2]);
void plf_integrate_and_centroid(const float* px, const float* py, int n, float a, float b, float* area, float* centroid);
void taper_table(const std::vector<float>& x, const std::vector<float>& y, float dt, int* tp0, int* tp1, std::vector<float>* tab);
void discrete_plf_span(const std::vector<float>& x, float dt, int* s0, int* s1);
void allowed_span(int s0, int s1, int minlength, int* n0, int* n1);
void initial_probe_span(int ds0, int ds1, int* s0, int* s1);

// per-candidate preparation: grid sizes, rotation matrices, STF taps, unit moment tensor
struct SourcePrep {
    int nx = 0, ny = 0, nt = 0, ngroups = 0;
    float moment = 1.f, risetime = 0.f;
    float rot_rup[9] = {0};
    float mhat[6] = {0};
    float p[16] = {0};
    float point[3] = {0}, time = 0.f;
    std::vector<float> toff, wt;
    // eikonal sources: explicit groups with their own taps (toff = absolute centroid time)
    std::vector<float> g_north, g_east, g_depth, g_gw, g_tbase;
    std::vector<int> g_tap_begin, g_tap_count;
    bool explicit_groups = false;
};
// ---- eikonal / mt_eikonal sources (source_eikonal_host.cpp) -----------------------------------------
struct CrustProfile { float vp[8], vs[8], rho[8], thickness[7], elevation; };   // crust2x2.f90:39-44
struct Crust2x2 {                                                                // crust2x2.f90:66 `model`
    bool loaded = false;
    int ntypes = 0, nlo = 0, nla = 0;
    std::vector<CrustProfile> types;
    std::vector<int16_t> map;
    std::vector<float> elev;
};
struct Halfspace { float point[3], normal[3]; };                                 // geometry.f90:25-28
struct EikonalGroup { float north, east, depth, gw; int tap_begin, tap_count; };
struct EikonalPrep {
    int nx = 0, ny = 0;
    float moment = 1.f, risetime = 0.f;
    float mhat[6] = {0};
    std::vector<EikonalGroup> groups;          // sub-faults with a valid rupture time, iy outer / ix inner
    std::vector<float> tap_time, tap_wt;       // centroid time and time weight per (group, tap)
    std::string err;
};
bool crust2x2_load(const char* path, Crust2x2* c, std::string* err);
CrustProfile crust2x2_get_profile(const Crust2x2& c, float lat, float lon);
void crust2x2_get_profile_averages(const CrustProfile& p, float* vvp, float* vvs, float* vrho, float* vthi);
void default_constraints(const Crust2x2& c, double olat_rad, double olon_rad, float thickness_limit, std::vector<Halfspace>* out);
void eikonal_solver_fmm(const float* speed, int nx, int ny, const float origin[2], const float delta[2], const float initialpoint[2],
                        float* times);
bool prep_eikonal(const float* params, bool mt_variant, float shortest_doi, double olat_rad, double olon_rad, const Crust2x2& crust,
                  const std::vector<Halfspace>& constraints, EikonalPrep* out);
// the same in three steps (begin: geometry and speed field of the fine grid; the fast-marching solve, on the host or -- for large batches --
// on the device between the two; finish: down-sampling and the sub-source table).  `params` must stay valid until finish.
struct EikonalWork {
    const float* p = nullptr; float rot_rup[9] = {0}, rot_slip[9] = {0}; int idx[6] = {0};
    bool mt_variant = false; float shortest_doi = 0.f;
    float first[2] = {0, 0}, last[2] = {0, 0}, delta[2] = {0, 0}, initialpoint[2] = {0, 0};
    int fnx = 0, fny = 0;
    float minspeed = 0.f, invalid_speed = 0.f;
    std::vector<float> speed, times, points;     // fine grid (fnx, fny), ix fastest; points: north, east, depth per node
    // what the speed field of the fine grid is made from (psm_make_eikonal_grid)
    float center[3] = {0, 0, 0}, bord_radius = 0.f, relv = 0.f;
    CrustProfile profile;
    const std::vector<Halfspace>* constraints = nullptr;
};
// the down-sampled grid (psm_downsample_grid): per sub-fault the number of fine points, mean rupture time, mean position, duration
struct EikonalCoarse {
    int nxc = 0, nyc = 0;
    float cdelta[2] = {0, 0};
    std::vector<float> ntimes, ctimes, cpoints, cdur;
};
bool prep_eikonal_setup(const float* params, bool mt_variant, float shortest_doi, double olat_rad, double olon_rad, const Crust2x2& crust,
                        const std::vector<Halfspace>& constraints, EikonalWork* work, EikonalPrep* out);   // geometry only
bool prep_eikonal_speed_host(EikonalWork* work, EikonalPrep* out);                                        // speed field of the fine grid
void eikonal_layer_table(const CrustProfile& p, float thr[5], float vs[6]);
bool prep_eikonal_coarse_dims(const EikonalWork& w, EikonalCoarse* cg, std::string* err);
void prep_eikonal_downsample_host(EikonalWork* work, EikonalCoarse* cg);
bool prep_eikonal_table(const EikonalWork& w, const EikonalCoarse& cg, EikonalPrep* out);
bool prep_eikonal_begin(const float* params, bool mt_variant, float shortest_doi, double olat_rad, double olon_rad, const Crust2x2& crust,
                        const std::vector<Halfspace>& constraints, EikonalWork* work, EikonalPrep* out);
void prep_eikonal_solve_host(EikonalWork* work);
bool prep_eikonal_finish(EikonalWork* work, EikonalPrep* out);

bool prep_bilateral(const float* params14, float shortest_doi, SourcePrep* out);
bool prep_moment_tensor(const float* params11, float shortest_doi, SourcePrep* out);
bool prep_circular(const float* params11, float shortest_doi, SourcePrep* out);
bool prep_point_lp(const float* params13, float shortest_doi, SourcePrep* out);

}  // namespace kh
