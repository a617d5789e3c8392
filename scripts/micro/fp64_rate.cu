// FP64 pipe throughput on this GPU: DFMA, DADD, DMUL, F2F.F64.F32, I2F.F64 (independent chains, all SMs busy).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_rate scripts/micro/fp64_rate.cu && ./fp64_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(double* out, int iters, double seed, float fseed, int iseed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    float f = fseed + threadIdx.x;
    int n = iseed + threadIdx.x;
    for (int i = 0; i < iters; ++i) {
        if (OP == 0) { a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c); a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c); }
        if (OP == 1) { a0 += c; a1 += c; a2 += c; a3 += c; a4 += c; a5 += c; a6 += c; a7 += c; }
        if (OP == 2) { a0 *= m; a1 *= m; a2 *= m; a3 *= m; a4 *= m; a5 *= m; a6 *= m; a7 *= m; }
        if (OP == 3) {   // float -> double conversions (8 per iteration), consumed by cheap integer ops
            long long acc = 0;
#pragma unroll
            for (int u = 0; u < 8; ++u) { double d = (double)(f + (float)u); acc ^= __double_as_longlong(d); }
            f = __int_as_float((__float_as_int(f) ^ (int)acc) & 0x3fffffff | 0x3f800000);
            n ^= (int)(acc >> 32);
        }
        if (OP == 4) {   // int -> double conversions
            long long acc = 0;
#pragma unroll
            for (int u = 0; u < 8; ++u) { double d = (double)(n + u); acc ^= __double_as_longlong(d); }
            n = (n ^ (int)(acc >> 20)) & 0xffff;
        }
        if (OP == 5) {   // FP32 FMA for comparison
            float b0 = f, b1 = f + 1, b2 = f + 2, b3 = f + 3, b4 = f + 4, b5 = f + 5, b6 = f + 6, b7 = f + 7;
            for (int u = 0; u < 1; ++u) { b0 = fmaf(b0, 1.0001f, 1e-6f); b1 = fmaf(b1, 1.0001f, 1e-6f); b2 = fmaf(b2, 1.0001f, 1e-6f); b3 = fmaf(b3, 1.0001f, 1e-6f);
                b4 = fmaf(b4, 1.0001f, 1e-6f); b5 = fmaf(b5, 1.0001f, 1e-6f); b6 = fmaf(b6, 1.0001f, 1e-6f); b7 = fmaf(b7, 1.0001f, 1e-6f); }
            f = b0 + b1 + b2 + b3 + b4 + b5 + b6 + b7;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + f + n;
}

template <int OP>
void run(const char* name, double* out) {
    const int blocks = 148 * 8, threads = 256, iters = 4096;
    k<OP><<<blocks, threads>>>(out, 64, 1.0, 1.0f, 1);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, iters, 1.0, 1.0f, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * threads * iters * 8;
    printf("%-22s %8.3f ms  %8.3f Tops/s  (%.1f lanes/clk/SM at 1.965 GHz)\n", name, ms, ops / ms / 1e9,
           ops / (ms * 1e-3) / 148 / 1.965e9);
}

int main() {
    double* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(double));
    run<0>("DFMA", out);
    run<1>("DADD", out);
    run<2>("DMUL", out);
    run<3>("F2F.F64.F32 (+ints)", out);
    run<4>("I2F.F64 (+ints)", out);
    run<5>("FFMA (+adds)", out);
    return 0;
}
