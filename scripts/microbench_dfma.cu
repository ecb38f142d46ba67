// DFMA latency / throughput microbenchmark for B200 (development aid, not part of the library).
// Each thread runs ILP independent chains of dependent DFMAs; W warps per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double* out, int iters, double b, double c) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b, c);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}

template <int ILP>
void run(int warps_per_sched, double* d) {
    int iters = 2000;
    int threads = 128 * warps_per_sched;
    if (threads > 1024) return;
    k<ILP><<<148, threads>>>(d, iters, 0.999, 1e-9);
    cudaDeviceSynchronize();
    double cyc;
    cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    double per_inst = cyc / (iters * 8.0 * ILP);          // cycles per DFMA per warp
    double sched_rate = warps_per_sched / per_inst;      // warp-DFMA per cycle per scheduler
    printf("ILP %d warps/sched %d: %.2f cycles per DFMA per warp -> %.3f warp-DFMA/cycle/scheduler\n", ILP, warps_per_sched,
           per_inst, sched_rate);
}

int main() {
    double* d;
    cudaMalloc(&d, 8 * 148 * 1024);
    for (int w : {1, 2, 4, 8}) {
        run<1>(w, d); run<2>(w, d); run<3>(w, d); run<4>(w, d); run<6>(w, d); run<8>(w, d);
    }
    return 0;
}
