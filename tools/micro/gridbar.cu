// Grid-barrier and launch floors for the merge loop: the software barrier of merge.cuh
// (arrival counter + generation word) with one 1024-thread CTA per SM, cooperative launch; and the cost of a
// kernel launch + synchronise from the host, i.e. what a host round trip per merge would cost.
#include <cstdio>
#include <cstdint>
#include <chrono>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef long long i64;
__device__ __forceinline__ void grid_barrier(i64* state) {
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile i64* gen = &state[1];
        const i64 g = *gen;
        __threadfence();
        if (atomicAdd((u64*)&state[0], 1ULL) == (u64)gridDim.x - 1) {
            *(volatile i64*)&state[0] = 0;
            __threadfence();
            atomicAdd((u64*)&state[1], 1ULL);
        } else {
            while (*gen == g) { }
        }
        __threadfence();
    }
    __syncthreads();
}
__global__ void __launch_bounds__(1024) k_bar(i64* state, i64* out, int iters) {
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) grid_barrier(state);
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = (t1 - t0) / iters;
}
__global__ void k_empty(i64* out) { if (threadIdx.x == 0) out[1] = 1; }
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    i64 *state, *out; cudaMalloc(&state, 64); cudaMalloc(&out, 64); cudaMemset(state, 0, 64);
    int iters = 2000;
    void* args[] = {&state, &out, &iters};
    for (int rep = 0; rep < 3; rep++) {
        cudaLaunchCooperativeKernel((void*)k_bar, dim3(sms), dim3(1024), args, 0, 0);
        cudaDeviceSynchronize();
    }
    i64 r[2]; cudaMemcpy(r, out, 16, cudaMemcpyDeviceToHost);
    // launch + sync floor
    for (int i = 0; i < 100; i++) { k_empty<<<1, 32>>>(out); cudaDeviceSynchronize(); }
    auto a = std::chrono::steady_clock::now();
    for (int i = 0; i < 2000; i++) { k_empty<<<1, 32>>>(out); cudaDeviceSynchronize(); }
    auto b = std::chrono::steady_clock::now();
    double us_launch = std::chrono::duration<double, std::micro>(b - a).count() / 2000;
    printf("grid barrier (%d CTAs x 1024 threads): %lld cycles = %.2f us at %.0f MHz | kernel launch + synchronize: %.2f us | %s\n",
           sms, r[0], r[0] / (khz / 1000.0), khz / 1000.0, us_launch, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
