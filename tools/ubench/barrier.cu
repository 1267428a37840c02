// Microbenchmarks that size the frame kernel's design (B200, sm_100a): cost of a grid-wide barrier between the phases of
// a persistent kernel (one atomic counter, release/acquire at gpu scope), of a hardware thread-block-cluster barrier, and
// of a dependent kernel boundary with programmatic dependent launch - the three ways of ordering two phases of a frame.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o barrier barrier.cu && ./barrier
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned ld_acq(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_rel(unsigned* p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }

__global__ void k_grid_barrier(unsigned* ctr, int rounds, unsigned long long* sink) {
    unsigned target = 0;
    unsigned long long acc = 0;
    for (int r = 0; r < rounds; r++) {
        acc += r * threadIdx.x;  // a token amount of work
        __syncthreads();
        if (threadIdx.x == 0) {
            target += gridDim.x;
            red_rel(ctr);
            while (ld_acq(ctr) < target) {}
        }
        __syncthreads();
    }
    if (acc == 0xdeadbeefull) *sink = acc;
}

__global__ void k_cluster_barrier(int rounds, unsigned long long* sink) {
    unsigned long long acc = 0;
    for (int r = 0; r < rounds; r++) {
        acc += r * threadIdx.x;
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (acc == 0xdeadbeefull) *sink = acc;
}

__global__ void k_empty_pdl(unsigned long long* sink) {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 9999) *sink = 1;
}

// a phase with one dependent L2 round trip per thread (what the short phases of a frame look like)
__global__ void k_grid_barrier_work(unsigned* ctr, int rounds, int* buf, int n) {
    unsigned target = 0;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    for (int r = 0; r < rounds; r++) {
        int v = buf[idx % n];
        buf[(idx + v + 1) % n] = v + 1;
        __syncthreads();
        if (threadIdx.x == 0) {
            target += gridDim.x;
            __threadfence();
            red_rel(ctr);
            while (ld_acq(ctr) < target) {}
        }
        __syncthreads();
    }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main() {
    unsigned* ctr; unsigned long long* sink; int* buf;
    CK(cudaMalloc(&ctr, 4)); CK(cudaMalloc(&sink, 8)); CK(cudaMalloc(&buf, 1 << 22));
    CK(cudaMemset(buf, 0, 1 << 22));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int rounds = 2000;
    const int cfgs[][2] = {{1, 1024}, {1, 512}, {2, 512}, {2, 256}, {4, 256}, {1, 256}};
    for (auto& c : cfgs) {
        int blocks = sms * c[0], threads = c[1];
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaMemset(ctr, 0, 4));
            int r = rounds; void* args[] = {&ctr, &r, &sink};
            cudaEventRecord(a);
            CK(cudaLaunchCooperativeKernel((void*)k_grid_barrier, dim3(blocks), dim3(threads), args, 0, 0));
            cudaEventRecord(b); CK(cudaEventSynchronize(b));
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (rep) printf("grid barrier  %4d blocks x %4d threads: %.3f us per barrier\n", blocks, threads, 1e3 * ms / rounds);
        }
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaMemset(ctr, 0, 4));
            int r = rounds, n = 1 << 20; void* args[] = {&ctr, &r, &buf, &n};
            cudaEventRecord(a);
            CK(cudaLaunchCooperativeKernel((void*)k_grid_barrier_work, dim3(blocks), dim3(threads), args, 0, 0));
            cudaEventRecord(b); CK(cudaEventSynchronize(b));
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (rep) printf("  + one dependent load/store per thread:   %.3f us per phase\n", 1e3 * ms / rounds);
        }
    }
    for (int cs : {2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        int nclusters = sms / cs;
        cfg.gridDim = dim3(nclusters * cs); cfg.blockDim = dim3(512);
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (cs > 8) CK(cudaFuncSetAttribute(k_cluster_barrier, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        int maxc = 0; cudaOccupancyMaxActiveClusters(&maxc, k_cluster_barrier, &cfg);
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(a);
            cudaError_t e = cudaLaunchKernelEx(&cfg, k_cluster_barrier, rounds, sink);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (e != cudaSuccess) { printf("cluster %d: %s\n", cs, cudaGetErrorString(e)); cudaGetLastError(); break; }
            if (rep) printf("cluster barrier, cluster of %2d x 512 threads (%d clusters launched, %d co-resident max): %.3f us per barrier\n", cs, nclusters, maxc, 1e3 * ms / rounds);
        }
    }
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(sms); cfg.blockDim = dim3(256);
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(a);
            for (int i = 0; i < 1000; i++) cudaLaunchKernelEx(&cfg, k_empty_pdl, sink);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (rep) printf("empty kernel boundary with PDL (148 x 256): %.3f us per launch\n", ms);
        }
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(a);
            for (int i = 0; i < 1000; i++) k_empty_pdl<<<sms, 256>>>(sink);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (rep) printf("empty kernel boundary, plain launch:        %.3f us per launch\n", ms);
        }
    }
    printf("done\n");
    return 0;
}
