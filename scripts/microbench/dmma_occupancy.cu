#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void dmmaK(double* out, int iters, double a, double b)
{
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    double s = 0.;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}
// with fragment loads from shared memory between the DMMAs (like the assembly kernel's k4 step)
__global__ void dmmaLds(double* out, int iters)
{
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 64 * 132; i += blockDim.x) sm[i] = 1e-3 * i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tq = lane & 3;
    double c[4][4][2];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j][0] = c[i][j][1] = 0.;
    const double* pa = sm + (warp % 4) * 32 + g;
    const double* pb = sm + 32 * 132 + (warp / 4 % 4) * 32 + g;
    for (int it = 0; it < iters; ++it)
#pragma unroll 2
        for (int k4 = 0; k4 < 8; ++k4)
        {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = pa[(k4 * 4 + tq) * 132 + i * 8]; b[i] = pb[(k4 * 4 + tq) * 132 + i * 8]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][j][0]), "+d"(c[i][j][1]) : "d"(a[i]), "d"(b[j]));
        }
    double s = 0.;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j][0] + c[i][j][1];
    if (s == 123.456) out[0] = s;
}
int main()
{
    double* buf; cudaMalloc(&buf, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int threads : {128, 256, 512, 1024})
    {
        const int iters = 2048; float ms;
        dmmaK<16><<<148, threads>>>(buf, 64, 1.0000001, 1e-9);
        cudaEventRecord(e0); dmmaK<16><<<148, threads>>>(buf, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("regs-only 16 acc: %4d threads/SM: %.2f TFLOP/s\n", threads, 2. * 256 * 16 * iters * 148. * (threads / 32) / (ms * 1e-3) / 1e12);
    }
    cudaFuncSetAttribute(dmmaLds, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 132 * 8);
    for (int threads : {128, 256, 512})
    {
        const int iters = 512; float ms;
        dmmaLds<<<148, threads, 64 * 132 * 8>>>(buf, 8);
        cudaEventRecord(e0); dmmaLds<<<148, threads, 64 * 132 * 8>>>(buf, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("with LDS fragments: %4d threads/SM: %.2f TFLOP/s\n", threads, 2. * 256 * 16 * 8 * iters * 148. * (threads / 32) / (ms * 1e-3) / 1e12);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
