// microbenchmark: gather of 512-byte row pieces (the access pattern of k_synth) out of a region of W bytes, as a function of W
// (L2-resident vs HBM-resident), per-lane LDG.128 into registers vs cp.async (LDGSTS) + LDS.  B200, natural clocks.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2bw l2bw.cu && ./l2bw
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int PIECE = 512, WARPS = 8;

__device__ __forceinline__ unsigned su32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ size_t hash_piece(size_t i, size_t n) {   // cheap: a handful of 32-bit operations
    unsigned h = (unsigned)i * 2654435761u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    return (size_t)__umulhi(h, (unsigned)n);
}
__device__ __forceinline__ float4 ldg_nc(const void* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// each warp reads `n` pieces; DEPTH pieces in flight per lane in registers
template <int DEPTH>
__global__ void __launch_bounds__(256, 2) k_ldg(const char* __restrict__ src, size_t n, size_t region_pieces, float* out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t gw = (size_t)blockIdx.x * WARPS + warp;
    float4 acc = make_float4(0, 0, 0, 0);
    float4 v[DEPTH];
#pragma unroll
    for (int s = 0; s < DEPTH; s++) v[s] = ldg_nc(src + hash_piece(gw * n + s, region_pieces) * PIECE + lane * 16);
    for (size_t i = 0; i < n; i += DEPTH) {
#pragma unroll
        for (int s = 0; s < DEPTH; s++) {
            acc.x += v[s].x; acc.y += v[s].y; acc.z += v[s].z; acc.w += v[s].w;
            if (i + s + DEPTH < n) v[s] = ldg_nc(src + hash_piece(gw * n + i + s + DEPTH, region_pieces) * PIECE + lane * 16);
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.f) out[0] = 1.f;
}

template <int STAGES>
__global__ void __launch_bounds__(256, 2) k_ldgsts(const char* __restrict__ src, size_t n, size_t region_pieces, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t gw = (size_t)blockIdx.x * WARPS + warp;
    unsigned ring = su32(smem) + warp * STAGES * PIECE + lane * 16;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int s = 0; s < STAGES; s++) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + s * PIECE), "l"(src + hash_piece(gw * n + s, region_pieces) * PIECE + lane * 16) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    int st = 0;
    for (size_t i = 0; i < n; i++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ring + st * PIECE) : "memory");
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        if (i + STAGES < n)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + st * PIECE), "l"(src + hash_piece(gw * n + i + STAGES, region_pieces) * PIECE + lane * 16) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        st = st == STAGES - 1 ? 0 : st + 1;
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.f) out[0] = 1.f;
}

template <class F>
float time_ms(F f, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); CHECK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int i = 0; i < reps; i++) f();
    cudaEventRecord(b); CHECK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main() {
    const size_t total = (size_t)4 << 30;
    char* d; float* out;
    CHECK(cudaMalloc(&d, total)); CHECK(cudaMemset(d, 0, total)); CHECK(cudaMalloc(&out, 4));
    cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
    printf("%s  SMs %d  L2 %d MB\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20);
    const int ctas = p.multiProcessorCount * 2 * 8;
    const size_t n = 4096;   // pieces per warp
    const double bytes = (double)ctas * WARPS * n * PIECE;
    CHECK(cudaFuncSetAttribute(k_ldgsts<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, WARPS * 12 * PIECE));
    CHECK(cudaFuncSetAttribute(k_ldgsts<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, WARPS * 24 * PIECE));
    for (size_t mb : {8, 16, 32, 48, 64, 96, 128, 256, 1024, 4096}) {
        const size_t rp = (mb << 20) / PIECE;
        float t1 = time_ms([&] { k_ldg<4><<<ctas, 256>>>(d, n, rp, out); }, 3);
        float t2 = time_ms([&] { k_ldg<8><<<ctas, 256>>>(d, n, rp, out); }, 3);
        float t3 = time_ms([&] { k_ldg<16><<<ctas, 256>>>(d, n, rp, out); }, 3);
        float t4 = time_ms([&] { k_ldgsts<12><<<ctas, 256, WARPS * 12 * PIECE>>>(d, n, rp, out); }, 3);
        float t5 = time_ms([&] { k_ldgsts<24><<<ctas, 256, WARPS * 24 * PIECE>>>(d, n, rp, out); }, 3);
        CHECK(cudaGetLastError());
        printf("region %5zu MB : LDG depth4 %7.0f  depth8 %7.0f  depth16 %7.0f   LDGSTS+LDS 12 stages %7.0f  24 stages %7.0f  GB/s\n", mb,
               bytes / t1 / 1e6, bytes / t2 / 1e6, bytes / t3 / 1e6, bytes / t4 / 1e6, bytes / t5 / 1e6);
    }
    return 0;
}
