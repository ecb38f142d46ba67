// Is a named barrier shared by a SUBSET of the warps of a CTA (bar.sync id, count with count < blockDim) accepted by
// compute-sanitizer --tool synccheck?  The line-march kernel synchronises a leader warp with its followers this way.
//   nvcc -arch=sm_100a -o named_barrier scripts/microbench_named_barrier.cu && compute-sanitizer --tool synccheck ./named_barrier
#include <cstdio>
// variant 2: the two warps of a team reach the barrier from DIFFERENT instructions (leader code / follower code), as in k_march
__device__ __noinline__ void follower(int team, volatile int* buf, int* out) {
    for (int it = 0; it < 4; ++it) {
        asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(64) : "memory");
        const int v = buf[team];
        asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(64) : "memory");
        if ((threadIdx.x & 31) == 0) out[team] = v;
    }
}
__global__ void k2(int* out) {
    const int warp = threadIdx.x >> 5, team = warp >> 1;
    __shared__ int buf[4];
    if (warp & 1) { follower(team, buf, out); return; }
    for (int it = 0; it < 4; ++it) {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) buf[team] = it + team;
        __syncwarp();
        asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(64) : "memory");
        asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(64) : "memory");
    }
}
__global__ void k(int* out) {
    const int warp = threadIdx.x >> 5, team = warp >> 1;
    __shared__ int buf[4];
    for (int it = 0; it < 4; ++it) {
        if ((warp & 1) == 0 && (threadIdx.x & 31) == 0) buf[team] = it + team;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(64) : "memory");
        const int v = buf[team];
        asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(64) : "memory");
        if (threadIdx.x == 32 * (2 * team + 1)) out[team] = v;
    }
}
int main() {
    int* d;
    cudaMalloc(&d, 16);
    k<<<1, 128>>>(d);
    int h[2] = {-1, -1};
    cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("out %d %d (expected 3 4): %s\n", h[0], h[1], cudaGetErrorString(cudaGetLastError()));
    k2<<<1, 128>>>(d);
    cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("different instructions: out %d %d (expected 3 4): %s\n", h[0], h[1], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
