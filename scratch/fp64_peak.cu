#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int MODE>
__global__ void k(double *out, int iters, double a, double b)
{
    double x[ILP];
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (MODE == 0) x[i] = __fma_rn(x[i], a, b);
            else if (MODE == 1) x[i] = __dmul_rn(x[i], a);
            else if (MODE == 2) x[i] = __dadd_rn(x[i], b);
            else { x[i] = __dmul_rn(x[i], a); x[i] = __dadd_rn(x[i], b); }
        }
    }
    double s = 0;
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP, int MODE>
void run(const char *name, int blocks, int threads)
{
    double *d; cudaMalloc(&d, sizeof(double) * blocks * threads);
    int iters = 20000;
    k<ILP, MODE><<<blocks, threads>>>(d, 100, 1.0000001, 1e-9);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<ILP, MODE><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * iters * ILP * (MODE == 3 ? 2 : 1);
    printf("%-28s blocks %4d x %4d thr  ILP %d : %.2f T inst/s\n", name, blocks, threads, ILP, ops / ms / 1e9);
    cudaFree(d);
}
int main()
{
    run<8, 0>("DFMA", 148 * 4, 256);
    run<8, 1>("DMUL", 148 * 4, 256);
    run<8, 2>("DADD", 148 * 4, 256);
    run<8, 3>("DMUL+DADD", 148 * 4, 256);
    run<4, 0>("DFMA 8 warps/SM", 148 * 2, 128);
    run<2, 0>("DFMA 8 warps/SM ILP2", 148 * 2, 128);
    run<1, 0>("DFMA 8 warps/SM ILP1", 148 * 2, 128);
    run<1, 0>("DFMA 16 warps/SM ILP1", 148 * 4, 128);
    run<1, 0>("DFMA 32 warps/SM ILP1", 148 * 8, 128);
    return 0;
}
