// Micro-benchmark: how fast can plain kernels READ two 235 MB arrays on this GPU?  (roofline sanity for mm_seg_moments)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ float4 ld4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// A: one array, grid-stride float4
__global__ void read_one(const float4* a, long long n4, float* out) {
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = ld4(a + i); s += v.x + v.y + v.z + v.w;
    }
    if (s == 12345.678f) out[0] = s;
}
// B: two arrays, same index
__global__ void read_two(const float4* a, const float4* b, long long n4, float* out) {
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = ld4(a + i), w = ld4(b + i); s += v.x + v.y + v.z + v.w + w.x + w.y + w.z + w.w;
    }
    if (s == 12345.678f) out[0] = s;
}
// C: two arrays, a warp owns 512-element spans (4 float4 per lane and array, all issued up front)
__global__ void read_spans(const float4* a, const float4* b, long long n4, float* out) {
    float s = 0.f;
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x / 32);
    const long long n_spans = n4 / 128;
    for (long long sp = (long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); sp < n_spans; sp += warps) {
        const float4* pa = a + sp * 128 + lane; const float4* pb = b + sp * 128 + lane;
        float4 v0 = ld4(pa), v1 = ld4(pa + 32), v2 = ld4(pa + 64), v3 = ld4(pa + 96);
        float4 w0 = ld4(pb), w1 = ld4(pb + 32), w2 = ld4(pb + 64), w3 = ld4(pb + 96);
        s += v0.x + v1.y + v2.z + v3.w + w0.x + w1.y + w2.z + w3.w + v0.y + w0.y;
    }
    if (s == 12345.678f) out[0] = s;
}
template <class F> float timeit(F f, int reps = 20) {
    for (int i = 0; i < 3; ++i) f();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); for (int i = 0; i < reps; ++i) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / reps;
}
int main() {
    const long long n = 58873218LL / 4 * 4, n4 = n / 4;
    float *a, *b, *c, *out;
    CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMalloc(&c, 2 * n * 4)); CK(cudaMalloc(&out, 4));
    CK(cudaMemset(a, 0, n * 4)); CK(cudaMemset(b, 0, n * 4)); CK(cudaMemset(c, 0, 2 * n * 4));
    const double gb = 2.0 * n * 4 / 1e9;
    for (int bps : {2, 4, 8, 16}) for (int th : {256, 512, 1024}) {
        if (bps * th > 2048) continue;
        int g = 148 * bps;
        float t1 = timeit([&] { read_one<<<g, th>>>((const float4*)c, 2 * n4, out); });
        float t2 = timeit([&] { read_two<<<g, th>>>((const float4*)a, (const float4*)b, n4, out); });
        float t3 = timeit([&] { read_spans<<<g, th>>>((const float4*)a, (const float4*)b, n4, out); });
        printf("blocks/SM %2d threads %4d : one array %.1f GB/s | two arrays %.1f GB/s | spans %.1f GB/s\n", bps, th, gb / t1 * 1e3, gb / t2 * 1e3, gb / t3 * 1e3);
    }
    float tc = timeit([&] { cudaMemcpyAsync(c, a, n * 4, cudaMemcpyDeviceToDevice); });
    printf("memcpy D2D 235 MB: %.1f GB/s (read+write)\n", 2.0 * n * 4 / 1e9 / tc * 1e3);
    CK(cudaDeviceSynchronize());
    return 0;
}
