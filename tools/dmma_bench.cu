// Development probe: FP64 throughput of one B200 for DFMA and for the DMMA (mma.sync f64)
// shapes ptxas accepts for sm_100a, with operands in registers (peak) and with operands
// re-read from shared memory every k-step (what a tiled kernel sees).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/dmma_bench tools/dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], const double (&a)[2], double b)
{
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4])
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int MODE> __global__ void k_peak(double *out, int iters, double seed)
{
    const int NI = 8; // independent accumulator chains
    double a[8], b[4];
    for (int i = 0; i < 8; ++i)
        a[i] = seed + threadIdx.x * 1e-9 + i;
    for (int i = 0; i < 4; ++i)
        b[i] = seed * 0.5 + i;
    double acc = 0.;
    if (MODE == 0)
    {
        double c[16];
        for (int i = 0; i < 16; ++i)
            c[i] = i;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 16; ++i)
                c[i] = fma(a[i & 7], b[i & 3], c[i]);
        for (int i = 0; i < 16; ++i)
            acc += c[i];
    }
    else if (MODE == 1)
    {
        double c[NI][2];
        for (int i = 0; i < NI; ++i)
            c[i][0] = c[i][1] = i;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < NI; ++i)
                dmma884(c[i][0], c[i][1], a[i], b[i & 3]);
        for (int i = 0; i < NI; ++i)
            acc += c[i][0] + c[i][1];
    }
    else
    {
        double c[NI][4];
        for (int i = 0; i < NI; ++i)
            for (int j = 0; j < 4; ++j)
                c[i][j] = i + j;
        double a2[2] = {a[0], a[1]}, a4[4] = {a[0], a[1], a[2], a[3]}, b2[2] = {b[0], b[1]};
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < NI; ++i)
            {
                if (MODE == 2)
                    dmma1684(c[i], a2, b[0]);
                else if (MODE == 3)
                    dmma1688(c[i], a4, b2);
                else
                    dmma16816(c[i], a, b);
            }
        for (int i = 0; i < NI; ++i)
            for (int j = 0; j < 4; ++j)
                acc += c[i][j];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// GEMM-like inner loop from shared memory: each warp owns a 32 x 32 tile of C (WM x WN
// fragments of m16n8), k-loop over a KT-deep slab of A (col-major, ld 32+pad) and B in smem.
template <int MODE> __global__ void k_smem(double *out, int iters)
{
    __shared__ double As[16][36]; // [k][m]
    __shared__ double Bs[16][36]; // [k][n]
    for (int i = threadIdx.x; i < 16 * 36; i += blockDim.x)
    {
        (&As[0][0])[i] = i * 1e-3;
        (&Bs[0][0])[i] = i * 2e-3;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    double acc = 0.;
    if (MODE == 1)
    {
        // m8n8k4: 4 x 4 fragments of 8x8 = 32 x 32 tile; per k4 step: 4 A loads, 4 B loads, 16 mma
        double c[4][4][2];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                c[i][j][0] = c[i][j][1] = 0.;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
            {
                double a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                {
                    a[i] = As[k4 * 4 + t][i * 8 + g];
                    b[i] = Bs[k4 * 4 + t][i * 8 + g];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
            }
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                acc += c[i][j][0] + c[i][j][1];
    }
    else if (MODE == 0)
    {
        // DFMA register tile 4 x 8 per thread (32 lanes: 8 x 4 threads -> 32 x 32), k = 16
        double c[4][8];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 8; ++j)
                c[i][j] = 0.;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int k = 0; k < 16; ++k)
            {
                double a[4], b[8];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    a[i] = As[k][g * 4 + i];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    b[j] = Bs[k][t * 8 + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        c[i][j] = fma(a[i], b[j], c[i][j]);
            }
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 8; ++j)
                acc += c[i][j];
    }
    else
    {
        // m16n8k8 / m16n8k16: 2 x 4 fragments of 16x8 = 32 x 32 tile
        double c[2][4][4];
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 4; ++j)
                for (int q = 0; q < 4; ++q)
                    c[i][j][q] = 0.;
        for (int it = 0; it < iters; ++it)
        {
            if (MODE == 3)
            {
#pragma unroll
                for (int k8 = 0; k8 < 2; ++k8)
                {
                    double a[2][4], b[4][2];
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                    {
                        a[i][0] = As[k8 * 8 + t][i * 16 + g];
                        a[i][1] = As[k8 * 8 + t][i * 16 + g + 8];
                        a[i][2] = As[k8 * 8 + t + 4][i * 16 + g];
                        a[i][3] = As[k8 * 8 + t + 4][i * 16 + g + 8];
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                    {
                        b[j][0] = Bs[k8 * 8 + t][j * 8 + g];
                        b[j][1] = Bs[k8 * 8 + t + 4][j * 8 + g];
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            dmma1688(c[i][j], a[i], b[j]);
                }
            }
            else
            {
                double a[2][8], b[4][4];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                    {
                        a[i][2 * q] = As[q * 4 + t][i * 16 + g];
                        a[i][2 * q + 1] = As[q * 4 + t][i * 16 + g + 8];
                    }
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        b[j][q] = Bs[q * 4 + t][j * 8 + g];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        dmma16816(c[i][j], a[i], b[j]);
            }
        }
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 4; ++j)
                for (int q = 0; q < 4; ++q)
                    acc += c[i][j][q];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class F> static double time_ms(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount;
    double *out;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    const int iters = 4096;
    const char *names[5] = {"DFMA", "DMMA m8n8k4", "DMMA m16n8k4", "DMMA m16n8k8", "DMMA m16n8k16"};
    const double fl_per_thread[5] = {16 * 2., 8 * 256 * 2. / 32, 8 * 512 * 2. / 32, 8 * 1024 * 2. / 32, 8 * 2048 * 2. / 32};
    for (int warps = 4; warps <= 32; warps *= 2)
    {
        const int threads = warps * 32;
        const int blocks = sms * (threads >= 1024 ? 2 : (2048 / threads > 4 ? 4 : 2048 / threads));
        double ms[5];
        ms[0] = time_ms([&] { k_peak<0><<<blocks, threads>>>(out, iters, 1.0); });
        ms[1] = time_ms([&] { k_peak<1><<<blocks, threads>>>(out, iters, 1.0); });
        ms[2] = time_ms([&] { k_peak<2><<<blocks, threads>>>(out, iters, 1.0); });
        ms[3] = time_ms([&] { k_peak<3><<<blocks, threads>>>(out, iters, 1.0); });
        ms[4] = time_ms([&] { k_peak<4><<<blocks, threads>>>(out, iters, 1.0); });
        for (int m = 0; m < 5; ++m)
            printf("regs  %-14s blocks %4d x %4d thr: %8.3f ms  %7.2f TFLOP/s\n", names[m], blocks, threads, ms[m],
                   fl_per_thread[m] * iters * (double)blocks * threads / (ms[m] * 1e-3) / 1e12);
    }
    for (int warps = 4; warps <= 16; warps *= 2)
    {
        const int threads = warps * 32;
        const int blocks = sms * (2048 / threads > 4 ? 4 : 2048 / threads);
        const int it2 = 2048;
        // flops per warp per iteration: 32 x 32 x 16 x 2
        const double fl = 32. * 32 * 16 * 2 * it2 * (double)blocks * warps;
        double m0 = time_ms([&] { k_smem<0><<<blocks, threads>>>(out, it2); });
        double m1 = time_ms([&] { k_smem<1><<<blocks, threads>>>(out, it2); });
        double m3 = time_ms([&] { k_smem<3><<<blocks, threads>>>(out, it2); });
        double m4 = time_ms([&] { k_smem<4><<<blocks, threads>>>(out, it2); });
        printf("smem  warps/blk %2d blocks %4d: DFMA4x8 %6.2f  m8n8k4 %6.2f  m16n8k8 %6.2f  m16n8k16 %6.2f TFLOP/s\n", warps,
               blocks, fl / (m0 * 1e-3) / 1e12, fl / (m1 * 1e-3) / 1e12, fl / (m3 * 1e-3) / 1e12, fl / (m4 * 1e-3) / 1e12);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
