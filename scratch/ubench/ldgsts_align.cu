// microbenchmark: shared-memory write wavefronts of LDGSTS.128 (cp.async.cg 16 B per lane) as a function of the alignment of the
// 512-byte source piece and of the number of active lanes.  Run under ncu:
//   ncu --metrics smsp__sass_l1tex_data_pipe_lsu_wavefronts_mem_shared_op_ldgsts.sum,smsp__inst_executed_op_ldgsts.sum,gpu__time_duration.sum ./ldgsts_align
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int WARPS = 8, STAGES = 12;
__device__ __forceinline__ unsigned su32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ size_t hash_piece(size_t i, size_t n) {
    unsigned h = (unsigned)i * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    return (size_t)__umulhi(h, (unsigned)n);
}
// pieces are 1024 bytes apart; the 512 bytes read start `off` bytes into the piece; lanes >= nlanes are idle
__global__ void __launch_bounds__(256, 2) k(const char* __restrict__ src, size_t n, size_t region_pieces, int off, int nlanes, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t gw = (size_t)blockIdx.x * WARPS + warp;
    unsigned ring = su32(smem) + warp * STAGES * 512 + lane * 16;
    float4 acc = make_float4(0, 0, 0, 0);
    const bool on = lane < nlanes;
    for (int s = 0; s < STAGES; s++) {
        if (on) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + s * 512), "l"(src + hash_piece(gw * n + s, region_pieces) * 1024 + off + lane * 16) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    int st = 0;
    for (size_t i = 0; i < n; i++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ring + st * 512) : "memory");
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        if (i + STAGES < n && on)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + st * 512), "l"(src + hash_piece(gw * n + i + STAGES, region_pieces) * 1024 + off + lane * 16) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        st = st == STAGES - 1 ? 0 : st + 1;
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.f) out[0] = 1.f;
}
int main() {
    const size_t total = (size_t)64 << 20;
    char* d; float* out;
    CHECK(cudaMalloc(&d, total + 4096)); CHECK(cudaMemset(d, 0, total + 4096)); CHECK(cudaMalloc(&out, 4));
    CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, WARPS * STAGES * 512));
    const int ctas = 148 * 2 * 4; const size_t n = 2048;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int nl : {32, 22}) for (int off : {0, 16, 32, 48, 64, 96}) {
        k<<<ctas, 256, WARPS * STAGES * 512>>>(d, n, total / 1024, off, nl, out);
        CHECK(cudaDeviceSynchronize());
        cudaEventRecord(a);
        k<<<ctas, 256, WARPS * STAGES * 512>>>(d, n, total / 1024, off, nl, out);
        cudaEventRecord(b); CHECK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("lanes %2d offset %3d : %.3f ms  %.0f GB/s\n", nl, off, ms, (double)ctas * WARPS * n * nl * 16 / ms / 1e6);
    }
    return 0;
}
