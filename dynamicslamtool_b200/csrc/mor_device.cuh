// mor_device.cuh — device-side building blocks shared by the MOR kernels (sm_100a).
//
//  * decoupled look-back tile prefix (single-pass stable scans / partitions)
//  * lock-free union-find (roots ordered by a hashed priority => shallow forests)
//  * order-preserving float<->uint keys for atomic min/max
//  * exact, order-independent fixed-point accumulation of float coordinates (centroids)
//  * the bit-exact distance / transform arithmetic shared with the CPU oracle
//
// The translation unit is compiled with -fmad=false: every float/double op below rounds exactly
// like the reference's default x86-64 build (no FMA contraction), see SURVEY §7 "Hard parts".
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mor {

constexpr int kBlock = 256;          // threads per block for the streaming kernels
constexpr int kItems = 4;            // items per thread in the scan kernels
constexpr int kTile = kBlock * kItems;
constexpr unsigned kFull = 0xFFFFFFFFu;

// ----------------------------------------------------------------------------- programmatic dependent launch
// Every kernel of the frame chain starts with this: it lets the NEXT kernel of the stream be scheduled right away
// (its blocks become resident and park in griddepcontrol.wait) and then waits until the PREVIOUS kernel has
// completed and its writes are visible. With kernels that last 7-50 us and never fill the GPU this hides the launch
// and block-dispatch latency of every boundary. A kernel launched without the PDL attribute passes straight through.
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ----------------------------------------------------------------------------- arithmetic (A6, A12)
// FLANN L2_Simple<float>: ((dx*dx) + dy*dy) + dz*dz, every op rounded to float, no FMA.
__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
    float d = __fsub_rn(ax, bx);
    float r = __fmul_rn(d, d);
    d = __fsub_rn(ay, by);
    r = __fadd_rn(r, __fmul_rn(d, d));
    d = __fsub_rn(az, bz);
    r = __fadd_rn(r, __fmul_rn(d, d));
    return r;
}

struct Affine12 {  // row-major 3x4 float (Eigen::Affine3f of pcl_ros::transformPointCloud)
    float m[12];
};

// pcl::transformPointCloud dense branch (PCL 1.8): m0*x + m1*y + m2*z + m3, left to right in float.
__device__ __forceinline__ float3 xform(const Affine12& M, float x, float y, float z) {
    float3 o;
    o.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M.m[0], x), __fmul_rn(M.m[1], y)), __fmul_rn(M.m[2], z)), M.m[3]);
    o.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M.m[4], x), __fmul_rn(M.m[5], y)), __fmul_rn(M.m[6], z)), M.m[7]);
    o.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M.m[8], x), __fmul_rn(M.m[9], y)), __fmul_rn(M.m[10], z)), M.m[11]);
    return o;
}

// ----------------------------------------------------------------------------- ordered float keys
__device__ __forceinline__ unsigned fkey(float f) {
    unsigned b = __float_as_uint(f);
    return b ^ ((unsigned)((int)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(unsigned k) {
    unsigned b = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
    return __uint_as_float(b);
}

// ----------------------------------------------------------------------------- exact coordinate sums
// x (|x| < 2^15 m) is split into hi = floor(x * 2^16) and lo = (x*2^16 - hi) * 2^30; both are
// integers for every float with |x| >= 2^-23 (smaller magnitudes lose < 2^-46 m). Integer sums are
// exact and order-independent, so the centroid is deterministic under any atomic ordering.
__device__ __forceinline__ void split_fixed(float x, long long& hi, long long& lo) {
    double s = (double)x * 65536.0;
    double f = floor(s);
    hi = (long long)f;
    lo = (long long)((s - f) * 1073741824.0);
}
__device__ __forceinline__ double join_fixed_mean(long long hi, long long lo, double n) {
    return ((double)hi * (1.0 / 65536.0) + (double)lo * (1.0 / 70368744177664.0)) / n;
}

// ----------------------------------------------------------------------------- union-find
// parent[] over cell indices, read and written concurrently by the whole group: all accesses are relaxed gpu-scope
// (served by L2, never by a stale L1 line), hooks are atomicCAS on roots only. Every pointer leads to a SMALLER index,
// so the forest is acyclic whatever the interleaving, and the root of a set is its smallest index.
__device__ __forceinline__ int ld_parent(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_parent(int* p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v));
}
__device__ __forceinline__ int uf_find(int* parent, int x) {
    int p = ld_parent(parent + x);
    while (p != x) {  // path halving; only non-roots are rewritten, always to an ancestor
        int gp = ld_parent(parent + p);
        if (gp != p) st_parent(parent + x, gp);
        x = p;
        p = gp;
    }
    return x;
}
// Read-only find: for a pass that stores each node's root itself (a concurrent path-halving store of another thread could
// land after that store and replace the root by an intermediate ancestor).
__device__ __forceinline__ int uf_find_ro(const int* parent, int x) {
    int p = ld_parent(parent + x);
    while (p != x) { x = p; p = ld_parent(parent + x); }
    return x;
}
// Returns the root of the merged set.
__device__ __forceinline__ int uf_union(int* parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return a;
        if (a < b) { int t = a; a = b; b = t; }  // a is the larger root: it goes under b
        const int old = atomicCAS(&parent[a], a, b);
        if (old == a) return b;
    }
}

// The same union with both climbs in flight together: one round trip per level instead of two, and - the usual case
// when the forest has just been flattened - one round trip to see that both nodes are roots, one for the hook.
__device__ __forceinline__ void uf_union_pair(int* parent, int a, int b) {
    while (true) {
        const int pa = ld_parent(parent + a), pb = ld_parent(parent + b);
        if (pa != a || pb != b) { a = pa; b = pb; continue; }
        if (a == b) return;
        const int hi = max(a, b), lo = min(a, b);
        if (atomicCAS(&parent[hi], hi, lo) == hi) return;
    }
}

// ----------------------------------------------------------------------------- look-back tile prefix
// status word: bits 63..62 = 0 invalid / 1 aggregate / 2 inclusive prefix; low 62 bits = value.
constexpr unsigned long long kStAgg = 1ull << 62, kStPre = 2ull << 62, kStMask = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v));
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// Called by every thread of the block (blockDim.x >= 32). `aggregate` is the tile's total, valid in
// thread 0. Returns the exclusive prefix of the tile (sum of the aggregates of tiles < tile).
// Tiles must be numbered in the order the blocks started (dynamic ticket), which makes the wait
// deadlock-free.
__device__ __forceinline__ unsigned long long tile_exclusive_prefix(unsigned long long* status, int tile,
                                                                    unsigned long long aggregate) {
    __shared__ unsigned long long s_prefix;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        aggregate = __shfl_sync(kFull, aggregate, 0);
        if (tile == 0) {
            if (lane == 0) { st_status(&status[0], kStPre | aggregate); s_prefix = 0; }
        } else {
            if (lane == 0) st_status(&status[tile], kStAgg | aggregate);
            unsigned long long excl = 0;
            int look = tile - 1;
            while (true) {
                const int idx = look - lane;
                unsigned long long v = idx >= 0 ? ld_status(&status[idx]) : kStPre;
                while (__any_sync(kFull, (v >> 62) == 0)) v = idx >= 0 ? ld_status(&status[idx]) : kStPre;
                const unsigned pmask = __ballot_sync(kFull, (v >> 62) == 2);
                const int firstp = pmask ? (__ffs(pmask) - 1) : 31;
                excl += warp_sum_u64(lane <= firstp ? (v & kStMask) : 0ull);
                if (pmask) break;
                look -= 32;
            }
            if (lane == 0) { st_status(&status[tile], kStPre | (excl + aggregate)); s_prefix = excl; }
        }
    }
    __syncthreads();
    return s_prefix;
}

// The same prefix for the frame kernel, where all tiles of a scan are in flight at once (a tile per CTA and round, every
// CTA resident): instead of a chained look-back - up to tile/32 dependent round trips - the whole CTA reads the
// aggregates of ALL predecessor tiles at the same time and sums them: one round trip. A tile posts its aggregate before
// it waits, so the wait cannot deadlock. `aggregate` must be valid in thread 0. BLOCK = blockDim.x.
template <int BLOCK>
__device__ __forceinline__ unsigned long long tile_prefix_wide(unsigned long long* status, int tile, unsigned long long aggregate) {
    __shared__ unsigned long long s_part[BLOCK / 32];
    __shared__ unsigned long long s_prefix;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) st_status(&status[tile], kStAgg | aggregate);
    unsigned long long sum = 0;
    for (int idx = threadIdx.x; idx < tile; idx += BLOCK) {
        unsigned long long v;
        do { v = ld_status(&status[idx]); } while ((v >> 62) == 0);
        sum += v & kStMask;
    }
    sum = warp_sum_u64(sum);
    if (lane == 0) s_part[warp] = sum;
    __syncthreads();
    if (warp == 0) {
        unsigned long long v = lane < BLOCK / 32 ? s_part[lane] : 0ull;
        v = warp_sum_u64(v);
        if (lane == 0) s_prefix = v;
    }
    __syncthreads();
    const unsigned long long r = s_prefix;
    __syncthreads();  // s_part / s_prefix may be reused by a following call
    return r;
}

// Block-wide exclusive scan of one value per thread (kBlock threads). Returns exclusive prefix within
// the block; *total (valid in all threads) = block sum.
template <typename T, int BLOCK = kBlock>
__device__ __forceinline__ T block_exclusive_scan(T v, T* total) {
    __shared__ T s_warp[BLOCK / 32];
    __shared__ T s_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : T(0);
        T winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T t = __shfl_up_sync(kFull, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < (int)(blockDim.x >> 5)) s_warp[lane] = winc - w;
        if (lane == 31) s_total = winc;
    }
    __syncthreads();
    *total = s_total;
    T res = inc - v + s_warp[warp];
    __syncthreads();  // s_warp / s_total may be reused by a following call
    return res;
}

// ----------------------------------------------------------------------------- 64-bit hash set
constexpr unsigned long long kEmptyKey = ~0ull;
__device__ __forceinline__ unsigned hash64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned)k;
}
__device__ __forceinline__ void hset_insert(unsigned long long* tab, unsigned mask, unsigned long long key) {
    unsigned h = hash64(key) & mask;
    while (true) {
        unsigned long long cur = tab[h];
        if (cur == key) return;
        if (cur == kEmptyKey) {
            unsigned long long old = atomicCAS(&tab[h], kEmptyKey, key);
            if (old == kEmptyKey || old == key) return;
        }
        h = (h + 1) & mask;
    }
}
__device__ __forceinline__ bool hset_contains(const unsigned long long* tab, unsigned mask, unsigned long long key) {
    unsigned h = hash64(key) & mask;
    while (true) {
        unsigned long long cur = tab[h];
        if (cur == key) return true;
        if (cur == kEmptyKey) return false;
        h = (h + 1) & mask;
    }
}

}  // namespace mor
