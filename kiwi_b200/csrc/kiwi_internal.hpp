// Internal declarations shared by the host translation units of libkiwi_b200.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/kiwi_b200.h"

// util.f90:106-116 error(): record a message, the caller returns non-zero
int kiwi_set_error(const char* fmt, ...);

// Host-side Green's function database (replaces t_gfdb/t_chunk storage, gfdb.f90:49-146).
// Traces are kept per (ix,iz,ig) while the database is being filled and are flattened on demand.
struct kiwi_gfdb {
    int nx = 0, nz = 0, ng = 0;
    float dt = 0.f, dx = 0.f, dz = 0.f, firstx = 0.f, firstz = 0.f;
    // flat form
    std::vector<int> span0, len;
    std::vector<long long> offset;
    std::vector<float> data;
    // fill form (moved into the flat form by flatten())
    std::vector<std::vector<float>> pending;
    bool flat = true;
    size_t ntr() const { return (size_t)nx * nz * ng; }
    size_t idx(int ix, int iz, int ig) const { return ((size_t)(ix - 1) * nz + (iz - 1)) * ng + (ig - 1); }
    void flatten();
};

// trace_pack (sparse_trace.f90:443-555) restricted to what the dense slab needs: returns the
// [first,last] sample window to keep of data[0..n) (first nonzero .. last nonzero plus one
// trailing zero if there is one); an all-zero trace keeps its first sample.
void kiwi_pack_window(const float* data, int n, int* first, int* last);
