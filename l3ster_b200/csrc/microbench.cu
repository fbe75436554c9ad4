// fp64 peak microbenchmarks (mode 3: DFMA and DMMA interleaved, sum of both flop counts): MEASURED_PEAKS.json carries no fp64 figure, so bench.py measures the denominators of the
// fp64 roofline itself, in the same run: (mode 0) dependent-free DFMA chains on the CUDA cores, (mode 1) DMMA
// mma.sync.m8n8k4.f64 on the tensor cores, (mode 2) an HBM copy (read + write bytes, like the driver's figure).
#include "../../include/l3ster_b200.h"

#include <cuda_runtime.h>

namespace
{
__global__ void dfmaKernel(double* out, int iters, double a, double b)
{
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
        acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 16; ++i)
            acc[i] = fma(acc[i], a, b);
    double s = 0.;
#pragma unroll
    for (int i = 0; i < 16; ++i)
        s += acc[i];
    if (s == 123.456)
        out[0] = s;
}
__global__ void dmmaKernel(double* out, int iters, double a, double b)
{
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    double s = 0.;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        s += c[i][0] + c[i][1];
    if (s == 123.456)
        out[0] = s;
}
// both instruction streams interleaved in every warp: do the fp64 CUDA-core pipe and the fp64 tensor path overlap?
__global__ void mixedKernel(double* out, int iters, double a, double b)
{
    double acc[16];
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 16; ++i)
        acc[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i)
        {
            acc[2 * i]     = fma(acc[2 * i], a, b);
            acc[2 * i + 1] = fma(acc[2 * i + 1], a, b);
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
        }
    double s = 0.;
#pragma unroll
    for (int i = 0; i < 16; ++i)
        s += acc[i];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        s += c[i][0] + c[i][1];
    if (s == 123.456)
        out[0] = s;
}
__global__ void copyKernel(const double2* __restrict__ in, double2* __restrict__ out, long long n)
{
    for (long long i = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; i < n; i += static_cast< long long >(gridDim.x) * blockDim.x)
        out[i] = in[i];
}
} // namespace

extern "C" int l3b_microbench(l3b_context* ctx, int mode, double* result)
{
    cudaStream_t s  = static_cast< cudaStream_t >(l3b_context_stream(ctx));
    cudaEvent_t  e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double* buf = nullptr;
    float   ms  = 0.f;
    int     rc  = 0;
    if (mode == 0 or mode == 1 or mode == 3)
    {
        cudaMalloc(&buf, 64);
        const int iters = 4096, blocks = 148 * 8, threads = 256;
        double    best  = 0.;
        for (int rep = 0; rep < 4; ++rep)
        {
            cudaEventRecord(e0, s);
            if (mode == 0)
                dfmaKernel<<< blocks, threads, 0, s >>>(buf, iters, 1.0000001, 1e-9);
            else if (mode == 1)
                dmmaKernel<<< blocks, threads, 0, s >>>(buf, iters, 1.0000001, 1e-9);
            else
                mixedKernel<<< blocks, threads, 0, s >>>(buf, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1, s);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            // DFMA: 16 FMAs/thread/iter; DMMA m8n8k4: 8*8*4 MACs per warp-instruction, 8 per iteration
            const double f_fma = 2. * 16 * iters * static_cast< double >(blocks) * threads;
            const double f_mma = 2. * 256 * 8 * iters * static_cast< double >(blocks) * (threads / 32);
            const double flops = mode == 0 ? f_fma : mode == 1 ? f_mma : f_fma + f_mma;
            best               = flops / (ms * 1e-3) / 1e12 > best ? flops / (ms * 1e-3) / 1e12 : best;
        }
        *result = best;
    }
    else if (mode == 2)
    {
        const long long n = 1ll << 27; // 2 GiB per buffer
        double2 *       a = nullptr, *b = nullptr;
        if (cudaMalloc(&a, n * sizeof(double2)) != cudaSuccess or cudaMalloc(&b, n * sizeof(double2)) != cudaSuccess)
            rc = L3B_ERR_CUDA;
        else
        {
            cudaMemsetAsync(a, 0, n * sizeof(double2), s);
            double best = 0.;
            for (int rep = 0; rep < 5; ++rep)
            {
                cudaEventRecord(e0, s);
                copyKernel<<< 148 * 16, 512, 0, s >>>(a, b, n);
                cudaEventRecord(e1, s);
                cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1);
                const double gbs = 2. * n * sizeof(double2) / (ms * 1e-3) / 1e9;
                best             = gbs > best ? gbs : best;
            }
            *result = best;
        }
        cudaFree(a);
        cudaFree(b);
    }
    else
        rc = L3B_ERR_INVALID_ARG;
    if (cudaGetLastError() != cudaSuccess)
        rc = L3B_ERR_CUDA;
    cudaFree(buf);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return rc;
}
