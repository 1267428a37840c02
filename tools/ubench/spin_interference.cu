// Does a grid-barrier spin (one thread per waiting CTA polling one word with ld.acquire.gpu) slow down the memory
// accesses of the CTAs that are still working? CTA 0 pointer-chases through an L2-resident buffer and reports cycles per
// dependent load while the other 147 CTAs (a) have exited, (b) spin with ld.acquire.gpu, (c) spin with ld.relaxed + backoff.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <numeric>
#include <random>
#include <algorithm>
__device__ __forceinline__ unsigned ld_acq(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_rlx(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__global__ void k(const int* chain, int steps, unsigned* flag, int mode, long long* out, int warps_chasing) {
    if (blockIdx.x == 0) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        long long t0 = clock64();
        int p = (warp * 977 + lane * 31) & ((1 << 20) - 1);
        if (warp < warps_chasing) for (int i = 0; i < steps; i++) p = __ldcg(chain + p);
        long long t1 = clock64();
        __syncthreads();
        if (threadIdx.x == 0) { out[0] = (t1 - t0) / steps; out[1] = p; __threadfence(); atomicExch(flag, 1u); }
    } else if (mode == 1) {
        if (threadIdx.x == 0) while (ld_acq(flag) == 0u) {}
        __syncthreads();
    } else if (mode == 2) {
        if (threadIdx.x == 0) while (ld_rlx(flag) == 0u) { __nanosleep(200); }
        __syncthreads();
    } else if (mode == 3) {  // every thread of the waiting CTAs parked in __syncthreads except the poller (same as 1), pollers of all warps
        if ((threadIdx.x & 31) == 0) while (ld_acq(flag) == 0u) {}
        __syncthreads();
    }
}
int main() {
    const int n = 1 << 20;
    std::vector<int> h(n); std::iota(h.begin(), h.end(), 0); std::mt19937 g(1); std::shuffle(h.begin(), h.end(), g);
    std::vector<int> c(n); for (int i = 0; i < n; i++) c[h[i]] = h[(i + 1) % n];
    int* d; unsigned* flag; long long* out;
    cudaMalloc(&d, n * 4); cudaMalloc(&flag, 4); cudaMalloc(&out, 16);
    cudaMemcpy(d, c.data(), n * 4, cudaMemcpyHostToDevice);
    const char* names[] = {"others exited", "others spin ld.acquire.gpu", "others spin ld.relaxed + nanosleep(200)", "others: 32 pollers per CTA ld.acquire"};
    for (int wc : {1, 32}) for (int mode = 0; mode < 4; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaMemset(flag, 0, 4);
            void* args[] = {&d, (void*)&(const int&)2000, &flag, &mode, &out, &wc};
            int steps = 2000; args[1] = &steps;
            cudaLaunchCooperativeKernel((void*)k, dim3(148), dim3(1024), args, 0, 0);
            cudaDeviceSynchronize();
            long long r[2]; cudaMemcpy(r, out, 16, cudaMemcpyDeviceToHost);
            if (rep) printf("%2d warps chasing, %-45s: %lld cycles per dependent L2 load\n", wc, names[mode], r[0]);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
