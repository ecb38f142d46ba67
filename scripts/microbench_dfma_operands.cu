// FP64 pipe throughput with distinct register operands (development aid, not part of the library).
// The plain DFMA microbenchmark (a = fma(a, b, c), b and c shared by all chains) reaches 0.5 warp-instructions per cycle
// per scheduler.  Real code has three different register operands per DFMA and a mix of DFMA / DMUL / DADD; this checks
// whether operand delivery (register banks, reuse cache) lowers the sustainable rate.
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: a_i = fma(a_i, b, c)            (shared b, c)
// MODE 1: a_i = fma(a_i, b_i, c_i)        (three distinct registers per instruction)
// MODE 2: a_i = fma(b_i, c_i, a_i)
// MODE 3: mix: a_i = fma(a_i, b_i, c_i); b_i = b_i * d; c_i = c_i + a_(i+1)
// MODE 4: a_i = fma(a_i, b_i, K[i]) with K in constant memory
__constant__ double K[8] = {1e-9, 2e-9, 3e-9, 4e-9, 5e-9, 6e-9, 7e-9, 8e-9};

template <int MODE, int ILP>
__global__ void k(double* out, int iters, double b0, double c0) {
    double a[ILP], b[ILP], c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = threadIdx.x + i; b[i] = b0 + 1e-12 * (threadIdx.x + 3 * i); c[i] = c0 * (1 + i + threadIdx.x); }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (MODE == 0) a[i] = fma(a[i], b0, c0);
                if (MODE == 1) a[i] = fma(a[i], b[i], c[i]);
                if (MODE == 2) a[i] = fma(b[i], c[i], a[i]);
                if (MODE == 3) { a[i] = fma(a[i], b[i], c[i]); b[i] = b[i] * b0; c[i] = c[i] + a[(i + 1) % ILP]; }
                if (MODE == 4) a[i] = fma(a[i], b[i], K[i]);
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i] + b[i] + c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}

template <int MODE, int ILP>
void run(int warps_per_sched, double* d) {
    int iters = 2000;
    int threads = 128 * warps_per_sched;
    k<MODE, ILP><<<148, threads>>>(d, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    double cyc;
    cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    const double n_inst = iters * 8.0 * ILP * (MODE == 3 ? 3 : 1);
    printf("mode %d ILP %d warps/sched %d: %.3f warp-FP64-instr/cycle/scheduler (peak 0.5)\n", MODE, ILP, warps_per_sched,
           warps_per_sched * n_inst / cyc);
}

int main() {
    double* d;
    cudaMalloc(&d, 8 * 148 * 1024);
    for (int w : {1, 2, 4}) {
        run<0, 8>(w, d); run<1, 8>(w, d); run<2, 8>(w, d); run<3, 8>(w, d); run<4, 8>(w, d);
        run<1, 4>(w, d); run<3, 4>(w, d); run<3, 2>(w, d);
    }
    return 0;
}
