// mor_debug.cuh — the reference's VISUALIZE outputs (IncludeAll.h:32), produced on request from the frame state
// that is resident on the device anyway. Not part of the per-frame chain: nothing here runs unless asked for.
//
//   cluster_collection (cpp:226-229, published at cpp:553-558): the points of all size-valid clusters, cluster
//   after cluster in cluster order, ascending cloud index inside a cluster.
#pragma once
#include "mor_kernels.cuh"

namespace mor {

struct CollectionPtrs {
    const int* cid;          // cluster of every cloud point, -1 = none
    const float4* pts;       // cloud (x, y, z, intensity)
    const int* cl_size;
    const int* counts;
    int* cursor;             // [kmax] next free slot of every cluster's segment
    int* turn;               // [2]: ticket counter, tile whose turn it is
    float4* out;             // 2 x float4 per point: pcl::PointXYZI records
};

constexpr int kCollBlock = 1024;
constexpr int kCollSlots = 2048;  // >= 2 x distinct clusters a tile can hold

// Segment offsets of the clusters (exclusive scan of the sizes, one block) and the reset of the turn counters.
__global__ void __launch_bounds__(kCollBlock) k_collection_offsets(CollectionPtrs a) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int K = a.counts[MOR_CNT_K];
    if (threadIdx.x == 0) { s_carry = 0; a.turn[0] = 0; a.turn[1] = 0; }
    __syncthreads();
    for (int base = 0; base < K; base += kCollBlock) {
        const int k = base + threadIdx.x;
        const int v = k < K ? a.cl_size[k] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(kFull, inc, o); if ((threadIdx.x & 31) >= o) inc += t; }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = s_warp[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(kFull, w, o); if ((int)threadIdx.x >= o) w += t; }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const int before = s_carry + (threadIdx.x >= 32 ? s_warp[(threadIdx.x >> 5) - 1] : 0) + inc - v;
        if (k < K) a.cursor[k] = before;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += s_warp[31];
        __syncthreads();
    }
}

// Stable multi-way partition of the cloud by cluster. A tile (1024 consecutive cloud points) meets few distinct
// clusters, so it first ranks its points per cluster in shared memory (hash of the clusters present, warps taken in
// order), then - when the tiles before it have done so - moves the cursors of those clusters forward by its counts
// in one step, and finally writes its points. Only the cursor update is serial over the tiles.
__global__ void __launch_bounds__(kCollBlock) k_cluster_collection(CollectionPtrs a) {
    __shared__ int s_key[kCollSlots], s_cnt[kCollSlots], s_base[kCollSlots];
    __shared__ int s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(&a.turn[0], 1);
    for (int t = threadIdx.x; t < kCollSlots; t += kCollBlock) { s_key[t] = -1; s_cnt[t] = 0; }
    __syncthreads();
    const int tile = s_tile;
    const int nc = a.counts[MOR_CNT_NC];
    const int c = tile * kCollBlock + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = c < nc ? a.cid[c] : -1;
    const unsigned grp = __match_any_sync(kFull, k);
    const int leader = __ffs(grp) - 1;
    const int rank_in_warp = __popc(grp & ((1u << lane) - 1u));
    int slot = -1;
    if (k >= 0 && lane == leader) {  // find or claim the slot of cluster k
        unsigned hsh = ((unsigned)k * 0x9E3779B1u) >> 21;  // 11 bits
        while (true) {
            const int old = atomicCAS(&s_key[hsh], -1, k);
            if (old == -1 || old == k) break;
            hsh = (hsh + 1) & (kCollSlots - 1);
        }
        slot = (int)hsh;
    }
    __syncthreads();
    int before_in_tile = 0;
    for (int w = 0; w < kCollBlock / 32; w++) {  // warps in order: the running count of a cluster inside the tile
        if (warp == w && slot >= 0) { before_in_tile = s_cnt[slot]; s_cnt[slot] = before_in_tile + __popc(grp); }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        while (ld_parent(a.turn + 1) != tile) {}
        __threadfence();
    }
    __syncthreads();
    for (int t = threadIdx.x; t < kCollSlots; t += kCollBlock) {
        const int key = s_key[t];
        if (key >= 0) {
            const int b = ld_parent(a.cursor + key);
            s_base[t] = b;
            st_parent(a.cursor + key, b + s_cnt[t]);
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); st_parent(a.turn + 1, tile + 1); }  // release: cursors before the turn
    const int base = __shfl_sync(grp, slot >= 0 ? s_base[slot] + before_in_tile : 0, leader);
    if (k >= 0) {
        const float4 p = a.pts[c];
        const int o = base + rank_in_warp;
        a.out[2 * o] = make_float4(p.x, p.y, p.z, 1.0f);
        a.out[2 * o + 1] = make_float4(p.w, 0.f, 0.f, 0.f);
    }
}

// ---------------------------------------------------------------------------------------------- radius ties
// north_star: "any tie-breaking divergence at the exact clustering radius counted and reported". The predicate of the
// hot path is evaluated bit for bit like FLANN's L2_Simple, so nothing diverges against the oracle; what CAN differ
// against a real PCL build are point pairs whose squared distance sits within a few units in the last place of r^2
// (another rounding of the distance, a pruned tree search). This diagnostic counts them: all pairs (i < j) of the
// current frame's `cloud` with d2 in [lo, hi], brute force over tiles of 1024 points in shared memory.
__global__ void __launch_bounds__(1024) k_radius_ties(const float4* __restrict__ pts, const int* __restrict__ counts, float lo, float hi,
                                                      unsigned long long* __restrict__ out) {
    __shared__ float sx[1024], sy[1024], sz[1024];
    const int nc = counts[MOR_CNT_NC];
    const int tiles = (nc + 1023) / 1024;
    unsigned long long mine = 0;
    // block b takes the tile pairs (ti <= tj) with index b, b + gridDim.x, ...
    const long long npairs = (long long)tiles * (tiles + 1) / 2;
    for (long long pidx = blockIdx.x; pidx < npairs; pidx += gridDim.x) {
        int ti = 0;
        long long rem = pidx;
        while (rem >= tiles - ti) { rem -= tiles - ti; ti++; }
        const int tj = ti + (int)rem;
        __syncthreads();
        const int j = tj * 1024 + threadIdx.x;
        if (j < nc) { const float4 p = pts[j]; sx[threadIdx.x] = p.x; sy[threadIdx.x] = p.y; sz[threadIdx.x] = p.z; }
        __syncthreads();
        const int i = ti * 1024 + threadIdx.x;
        if (i < nc) {
            const float4 p = pts[i];
            const int jn = min(1024, nc - tj * 1024);
            for (int t = 0; t < jn; t++) {
                if (ti == tj && t <= (int)threadIdx.x) continue;  // i < j
                const float d = sqdist3(p.x, p.y, p.z, sx[t], sy[t], sz[t]);
                mine += (d >= lo && d <= hi) ? 1ull : 0ull;
            }
        }
    }
    mine = warp_sum_u64(mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(out, mine);
}

}  // namespace mor
