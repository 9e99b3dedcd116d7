// Latency micro-benchmarks that shape the merge loop design: dependent L2 loads, global atomics with a
// returned value, shared-memory atomics, block barriers.  One CTA, like the leader.  nvcc -arch=sm_100a lat.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__global__ void k_lat(u64* buf, int n, u64* out, int iters) {
    __shared__ u64 sh[1024];
    const int tid = threadIdx.x;
    sh[tid] = tid;
    __syncthreads();
    long long t0, t1;
    // (a) dependent ld.cg chain (pointer chase through buf: buf[i] holds the next index)
    if (tid == 0) {
        u64 p = 0;
        t0 = clock64();
        for (int i = 0; i < iters; i++) p = __ldcg(&buf[p]);
        t1 = clock64();
        out[0] = (t1 - t0) / iters; out[20] = p;
    }
    __syncthreads();
    // (b) dependent atomicAdd-with-return chain on distinct addresses
    if (tid == 0) {
        u64 p = 1;
        t0 = clock64();
        for (int i = 0; i < iters; i++) p = (atomicAdd(&buf[n + (p & 4095) * 16], 1ULL) & 1) + i * 7 + 1;
        t1 = clock64();
        out[1] = (t1 - t0) / iters; out[21] = p;
    }
    __syncthreads();
    // (c) dependent atomicCAS chain
    if (tid == 0) {
        u64 p = 1;
        t0 = clock64();
        for (int i = 0; i < iters; i++) p = (atomicCAS(&buf[n + 65536 + (p & 4095) * 16], 0ULL, (u64)i + 1) & 1) + i * 5 + 1;
        t1 = clock64();
        out[2] = (t1 - t0) / iters; out[22] = p;
    }
    __syncthreads();
    // (d) __syncthreads with all threads of the CTA
    t0 = clock64();
    for (int i = 0; i < iters; i++) __syncthreads();
    t1 = clock64();
    if (tid == 0) out[3] = (t1 - t0) / iters;
    // (e) dependent shared-memory atomic chain
    if (tid == 0) {
        u64 p = 1;
        t0 = clock64();
        for (int i = 0; i < iters; i++) p = atomicAdd(&sh[p & 1023], 1ULL) + i;
        t1 = clock64();
        out[4] = (t1 - t0) / iters; out[24] = p;
    }
    __syncthreads();
    // (f) RED (no return) followed by a barrier and a ld.cg of the same address by another thread: is it visible? cost?
    {
        int bad = 0;
        t0 = clock64();
        for (int i = 0; i < iters; i++) {
            if (tid == 1) atomicAdd(&buf[n + 200000], 1ULL);
            __syncthreads();
            if (tid == 33) { u64 v = __ldcg(&buf[n + 200000]); if (v != (u64)i + 1) bad++; }
            __syncthreads();
        }
        t1 = clock64();
        if (tid == 33) { out[5] = (t1 - t0) / iters; out[6] = bad; }
    }
    // (g) dependent ld (default caching, L1 allowed) chain
    if (tid == 0) {
        u64 p = 0;
        t0 = clock64();
        for (int i = 0; i < iters; i++) p = buf[p];
        t1 = clock64();
        out[7] = (t1 - t0) / iters; out[27] = p;
    }
    // (h) 8 independent atomics with return issued back to back, then all consumed
    if (tid == 0) {
        u64 acc = 0;
        t0 = clock64();
        for (int i = 0; i < iters; i++) {
            u64 r[8];
#pragma unroll
            for (int u = 0; u < 8; u++) r[u] = atomicAdd(&buf[n + 300000 + ((i * 8 + u) & 4095) * 16], 1ULL);
#pragma unroll
            for (int u = 0; u < 8; u++) acc += r[u];
        }
        t1 = clock64();
        out[8] = (t1 - t0) / iters; out[28] = acc;
    }
}
int main() {
    const int n = 1 << 20;            // 8 MB chase region (L2 resident, far beyond L1)
    u64* h = (u64*)malloc((n + 400000 + 70000) * 8);
    // random cycle
    for (int i = 0; i < n; i++) h[i] = i;
    uint64_t s = 88172645463325252ULL;
    for (int i = n - 1; i > 0; i--) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; int j = s % i; u64 t = h[i]; h[i] = h[j]; h[j] = t; }
    u64 *d, *o;
    cudaMalloc(&d, (n + 400000 + 70000) * 8); cudaMalloc(&o, 64 * 8);
    cudaMemset(d, 0, (n + 400000 + 70000) * 8);
    cudaMemcpy(d, h, n * 8, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; rep++) {
        cudaMemset(d + n, 0, (400000 + 70000) * 8);
        k_lat<<<1, 1024>>>(d, n, o, 2000);
        cudaDeviceSynchronize();
    }
    u64 r[64]; cudaMemcpy(r, o, 64 * 8, cudaMemcpyDeviceToHost);
    printf("cycles: ldcg chain %llu | atomicAdd-return chain %llu | atomicCAS chain %llu | syncthreads(1024) %llu | smem atomic chain %llu | RED+bar+ldcg+bar %llu (bad=%llu) | ld (L1) chain %llu | 8 indep atomics %llu\n",
           r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8]);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
