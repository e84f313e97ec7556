// Microbenchmark: issue cost of MATCH.ANY (warp match) on sm_100a, next to a SHFL and an LDS/FADD/STS read-modify-write.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/match_bench tools/micro/match_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_match(int iters, int spread, unsigned* out, long long* cyc) {
    unsigned v = (threadIdx.x & 31) % spread + blockIdx.x, acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        unsigned m = __match_any_sync(0xffffffffu, v);
        acc += m;
        v += (m & 1);  // dependent chain
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_match_ind(int iters, int spread, unsigned* out, long long* cyc) {
    unsigned v0 = (threadIdx.x & 31) % spread, v1 = v0 + 3, v2 = v0 * 7, v3 = v0 ^ 5, acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        acc += __match_any_sync(0xffffffffu, v0 + i);
        acc += __match_any_sync(0xffffffffu, v1 + i);
        acc += __match_any_sync(0xffffffffu, v2 + i);
        acc += __match_any_sync(0xffffffffu, v3 + i);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_shfl(int iters, unsigned* out, long long* cyc) {
    unsigned v = threadIdx.x, acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        acc += __shfl_sync(0xffffffffu, v + i, (threadIdx.x + 1) & 31);
        acc += __shfl_sync(0xffffffffu, v ^ i, (threadIdx.x + 2) & 31);
        acc += __shfl_sync(0xffffffffu, v * 3 + i, (threadIdx.x + 3) & 31);
        acc += __shfl_sync(0xffffffffu, v - i, (threadIdx.x + 4) & 31);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_rmw(int iters, float* outf, long long* cyc) {
    extern __shared__ float pl[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) pl[i] = 0.f;
    __syncthreads();
    volatile float* p = pl + (threadIdx.x >> 5) * 128;
    int a = threadIdx.x & 31;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        p[a] = p[a] + 1.0f;
        a = (a + 33) & 127;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    outf[blockIdx.x * blockDim.x + threadIdx.x] = pl[threadIdx.x];
}

int main() {
    unsigned* out; long long* cyc; float* outf;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&outf, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    for (int warps : {1, 4, 8, 16, 32}) {
        long long c;
        for (int spread : {32, 8, 1}) {
            k_match<<<148, warps * 32>>>(iters, spread, out, cyc);
            cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("match.any dependent  warps/SM=%2d spread=%2d: %.1f cyc per match per warp, %.2f cyc/match/SM\n", warps, spread,
                   (double)c / iters, (double)c / iters / warps);
            k_match_ind<<<148, warps * 32>>>(iters, spread, out, cyc);
            cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("match.any independent warps/SM=%2d spread=%2d: %.1f cyc per match per warp, %.2f cyc/match/SM\n", warps, spread,
                   (double)c / iters / 4, (double)c / iters / 4 / warps);
        }
        k_shfl<<<148, warps * 32>>>(iters, out, cyc);
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("shfl independent     warps/SM=%2d: %.1f cyc per shfl per warp, %.2f cyc/shfl/SM\n", warps, (double)c / iters / 4,
               (double)c / iters / 4 / warps);
        k_rmw<<<148, warps * 32, 16384>>>(iters, outf, cyc);
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("LDS/FADD/STS chain   warps/SM=%2d: %.1f cyc per rmw per warp, %.2f cyc/rmw/SM\n", warps, (double)c / iters,
               (double)c / iters / warps);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
