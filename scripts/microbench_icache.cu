// Instruction-fetch microbenchmark: a loop whose straight-line body holds BODY DFMAs (4 independent chains).
#include <cstdio>
#include <cuda_runtime.h>

template <int BODY>
__global__ void k(double* out, int iters, double b, double c) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < BODY / 4; ++u) {
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = (a0 + a1) + (a2 + a3);
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}

template <int BODY>
void run(int warps_per_sched, double* d) {
    int iters = 400000 / BODY;
    int threads = 128 * warps_per_sched;
    k<BODY><<<148, threads>>>(d, iters, 0.999, 1e-9);
    cudaDeviceSynchronize();
    double cyc;
    cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    double per_inst = cyc / ((double)iters * BODY);
    printf("body %5d DFMA (%6d B) warps/sched %d: %.3f warp-DFMA/cycle/scheduler\n", BODY, BODY * 16, warps_per_sched,
           warps_per_sched / per_inst);
}

int main() {
    double* d;
    cudaMalloc(&d, 8 * 148 * 1024);
    for (int w : {1, 4}) {
        run<64>(w, d); run<128>(w, d); run<256>(w, d); run<512>(w, d); run<768>(w, d); run<1024>(w, d); run<2048>(w, d);
        run<4096>(w, d); run<8192>(w, d);
    }
    return 0;
}
