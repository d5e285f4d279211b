// Throughput of the conversion and FP64 pipes on sm_100a (development probe).
// Each kernel runs ITERS x UNROLL independent ops per thread; prints ops/clk/SM.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, UNROLL = 8;

template <int MODE>
__global__ void k(float* out, float seed, double dseed)
{
    float f[UNROLL];
    double d[UNROLL];
#pragma unroll
    for (int i = 0; i < UNROLL; i++)
    {
        f[i] = seed + threadIdx.x + i;
        d[i] = dseed + threadIdx.x + i;
    }
    for (int it = 0; it < ITERS; it++)
    {
#pragma unroll
        for (int i = 0; i < UNROLL; i++)
        {
            if (MODE == 0) // F2F f32->f64 then tiny int op to keep a chain
            {
                double x = (double)f[i];
                f[i]     = __int_as_float(__double2hiint(x) ^ it);
            }
            else if (MODE == 1) // F2F f64->f32
            {
                float x = (float)d[i];
                d[i]    = __hiloint2double(__float_as_int(x) ^ it, 12345);
            }
            else if (MODE == 2) // DFMA
                d[i] = fma(d[i], 1.0000001, dseed);
            else if (MODE == 3) // DADD
                d[i] = d[i] + dseed;
            else if (MODE == 4) // FFMA
                f[i] = fmaf(f[i], 1.0000001f, seed);
            else if (MODE == 5) // DMUL
                d[i] = d[i] * 1.0000001;
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < UNROLL; i++)
        s += f[i] + (float)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name)
{
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k<MODE><<<148 * 8, 256>>>(out, 1.f, 1.0);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<MODE><<<148 * 8, 256>>>(out, 1.f, 1.0);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double ops = 148.0 * 8 * 256 * ITERS * UNROLL;
    printf("%-12s %8.3f ms  %7.2f ops/clk/SM (at %d MHz nominal)\n", name, ms,
        ops / (ms * 1e-3) / (clk * 1e3) / 148.0, clk / 1000);
    cudaFree(out);
}

int main()
{
    run<0>("F2F.64.32");
    run<1>("F2F.32.64");
    run<2>("DFMA");
    run<3>("DADD");
    run<5>("DMUL");
    run<4>("FFMA");
    return 0;
}
