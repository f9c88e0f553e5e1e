// Micro-benchmark: FP64 dependent-issue latency and the (chains x warps) needed to saturate the pipe.
// One CTA per SM with W warps per SMSP (block = 128*W threads), CH independent DFMA chains per thread.
#include <cstdio>
#include <cuda_runtime.h>

template <int CH, int MODE>
__global__ void k(double* out, long long* cyc, int iters, double x, double y) {
    double a[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) a[i] = threadIdx.x + i;
    double xr = x + threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (MODE == 0) a[i] = fma(a[i], x, y);       // r,c,c
                if (MODE == 1) a[i] = fma(a[i], xr, y);      // r,r,c
                if (MODE == 2) a[i] = __dmul_rn(a[i], xr);
                if (MODE == 3) a[i] = __dadd_rn(a[i], xr);
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int CH, int MODE>
void run(int W) {
    int iters = 2000;
    double* d; long long* c; long long h;
    cudaMalloc(&d, 148 * 1024 * 8); cudaMalloc(&c, 8);
    k<CH, MODE><<<148, 128 * W>>>(d, c, iters, 0.999999, 1e-9);
    k<CH, MODE><<<148, 128 * W>>>(d, c, iters, 0.999999, 1e-9);
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    double per_round = (double)h / (iters * 8.0);          // cycles per "one op on every chain"
    printf("mode %d chains %2d warps/SMSP %d : %7.2f cycles per round  -> %5.2f cycles per warp-instr per SMSP (pipe floor 2.0)\n",
           MODE, CH, W, per_round, per_round / (CH * W));
    cudaFree(d); cudaFree(c);
}

int main() {
    run<1, 0>(1); run<2, 0>(1); run<4, 0>(1); run<8, 0>(1); run<16, 0>(1);
    run<1, 0>(2); run<2, 0>(2); run<4, 0>(2); run<8, 0>(2);
    run<1, 0>(4); run<2, 0>(4); run<4, 0>(4); run<8, 0>(4);
    run<1, 0>(8); run<2, 0>(8); run<4, 0>(8);
    run<1, 1>(1); run<4, 1>(1); run<4, 1>(4); run<8, 1>(2);
    run<1, 2>(1); run<4, 2>(4); run<1, 3>(1); run<4, 3>(4);
    return 0;
}
