// Follow-up to microbench_dfma_operands.cu: which operand patterns cost a third register-read cycle?
#include <cstdio>
#include <cuda_runtime.h>
// MODE 0: a = fma(a, b_i, c_i)  3 distinct            MODE 1: a = fma(b_i, b_i, a)   (square + add: 2 distinct)
// MODE 2: a = fma(a, a, b_i)    2 distinct            MODE 3: a = fma(a, b_i, a)     2 distinct
// MODE 4: a = a * b_i + imm                           MODE 5: a_i = fma(a_i, b_i, U) with U warp-uniform from shfl
// MODE 6: DMUL a_i = a_i * b_i                        MODE 7: DADD a_i = a_i + b_i
// MODE 8: a_i = fma(a_i, b_(i&1), c_i): consecutive instructions share operand B in pairs (reuse cache)
template <int MODE, int ILP>
__global__ void k(double* out, int iters, double b0, double c0) {
    double a[ILP], b[ILP], c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = 1.0 + 1e-3 * (threadIdx.x + i); b[i] = b0 + 1e-12 * (threadIdx.x + 3 * i); c[i] = c0 * (1 + i + threadIdx.x); }
    const double U = __shfl_sync(0xffffffffu, b[0], 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (MODE == 0) a[i] = fma(a[i], b[i], c[i]);
                if (MODE == 1) a[i] = fma(b[i], b[i], a[i]);
                if (MODE == 2) a[i] = fma(a[i], a[i], b[i]);
                if (MODE == 3) a[i] = fma(a[i], b[i], a[i]);
                if (MODE == 4) a[i] = fma(a[i], b[i], 1.0);
                if (MODE == 5) a[i] = fma(a[i], b[i], U);
                if (MODE == 6) a[i] = a[i] * b[i];
                if (MODE == 7) a[i] = a[i] + b[i];
                if (MODE == 8) a[i] = fma(a[i], b[i & 1], c[i]);
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
void run(int w, double* d) {
    int iters = 2000;
    k<MODE, ILP><<<148, 128 * w>>>(d, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    double cyc;
    cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    printf("mode %d ILP %d warps/sched %d: %.3f instr/cycle/scheduler\n", MODE, ILP, w, w * iters * 8.0 * ILP / cyc);
}
int main() {
    double* d;
    cudaMalloc(&d, 8 * 148 * 1024);
    const int w = 4;
    run<0, 8>(w, d); run<1, 8>(w, d); run<2, 8>(w, d); run<3, 8>(w, d); run<4, 8>(w, d); run<5, 8>(w, d); run<6, 8>(w, d);
    run<7, 8>(w, d); run<8, 8>(w, d);
    return 0;
}
