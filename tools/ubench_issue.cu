// Micro-benchmark: does non-FP64 issue overlap with FP64 pipe occupancy on sm_100a?
// Each loop iteration issues 16 independent FP64 ops (variant-dependent) and NI independent ALU ops.
// Prints cycles per iteration per SMSP-resident warp set.   nvcc -arch=sm_100a -O3 tools/ubench_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NI, int MODE>
__global__ void __launch_bounds__(256) k(double* out, unsigned* iout, int iters, double x, double y, unsigned m) {
    double a[16];
    unsigned u[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = threadIdx.x + i; u[i] = threadIdx.x * 7 + i; }
    double xr = x + threadIdx.x * 1e-9, yr = y + threadIdx.x * 1e-9;  // register (non-uniform) operands
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) a[i] = fma(a[i], x, y);            // 1 register + 2 uniform/const operands
            if (MODE == 1) a[i] = fma(a[i], xr, yr);          // 3 distinct register operands
            if (MODE == 2) a[i] = __dmul_rn(a[i], xr);        // DMUL 2 regs
            if (MODE == 3) a[i] = __dadd_rn(a[i], xr);        // DADD 2 regs
            if (MODE == 4) a[i] = fma(a[i], xr, y);           // 2 registers + const
            if (i < NI) u[i] = (u[i] ^ m) + 0x9e3779b9u;       // LOP3 + IADD-ish: ALU pipe
        }
        if (NI > 16) {
#pragma unroll
            for (int i = 0; i < NI - 16; ++i) u[i & 15] = (u[i & 15] ^ m) + 0x9e3779b9u;
        }
    }
    double s = 0; unsigned t = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) { s += a[i]; t += u[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int NI, int MODE>
void run(const char* name, int blocks_per_sm) {
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    int blocks = p.multiProcessorCount * blocks_per_sm, threads = 256, iters = 20000;
    double* d; unsigned* ui;
    cudaMalloc(&d, blocks * threads * 8); cudaMalloc(&ui, blocks * threads * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NI, MODE><<<blocks, threads>>>(d, ui, 100, 0.999999, 1e-9, 0x5bd1e995u);
    cudaEventRecord(e0);
    k<NI, MODE><<<blocks, threads>>>(d, ui, iters, 0.999999, 1e-9, 0x5bd1e995u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    // warps per SMSP = blocks_per_sm*8/4 ; cycles per iteration per warp-set on one SMSP:
    double cyc = ms * 1e-3 * clk * 1e3 / iters / (blocks_per_sm * 2.0);
    printf("%-34s NI=%2d blocks/SM=%d  %8.3f ms  cycles/iter/warp = %6.2f  (16 FP64 ops: ideal 32)\n", name, NI, blocks_per_sm, ms, cyc);
    cudaFree(d); cudaFree(ui);
}

int main() {
    for (int bps : {1, 2}) {
        run<0, 0>("DFMA r,c,c", bps); run<4, 0>("DFMA r,c,c", bps); run<8, 0>("DFMA r,c,c", bps);
        run<16, 0>("DFMA r,c,c", bps); run<32, 0>("DFMA r,c,c", bps);
        run<0, 1>("DFMA r,r,r", bps); run<8, 1>("DFMA r,r,r", bps); run<16, 1>("DFMA r,r,r", bps);
        run<0, 4>("DFMA r,r,c", bps); run<16, 4>("DFMA r,r,c", bps);
        run<0, 2>("DMUL r,r", bps); run<16, 2>("DMUL r,r", bps);
        run<0, 3>("DADD r,r", bps); run<16, 3>("DADD r,r", bps);
    }
    return 0;
}
