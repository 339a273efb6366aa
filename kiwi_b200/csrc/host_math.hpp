// Host-side arithmetic of the engine (see host_math.cpp).
#pragma once
#include <vector>

namespace kh {

float d2r_r(float deg);
double d2r_d(double deg);
void azibazi(double alat, double alon, double blat, double blon, double* azi, double* bazi);
double distance_accurate50m(double alat, double alon, double blat, double blon);
void final_rotation(double bazi0, float* cl0, float* sl0);
void init_euler(float alpha, float beta, float gamma, float* mat9);
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
};
bool prep_bilateral(const float* params14, float shortest_doi, SourcePrep* out);
bool prep_moment_tensor(const float* params11, float shortest_doi, SourcePrep* out);

}  // namespace kh
