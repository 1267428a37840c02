// mor_kernels.cuh — the per-frame MOR kernels (sm_100a), crop ground mode + clustering + matching +
// moving tests + tracking + output. Launched by mor_b200.cu. Every kernel cites the reference lines
// (src/MovingObjectRemoval.cpp unless noted) whose behaviour it reproduces; DESIGN.md has the data
// layout and the roofline of each.
// Every kernel exists twice: NAME(FramePtrs) for one sequence (arguments in the constant bank) and
// NAME_batch(const FramePtrs*) for S sequences in one launch (blockIdx.z selects the sequence; the per-sequence
// state - scratch, tickets, tables - is disjoint, so the bodies are identical).
#pragma once
#ifndef LINK_BLOCK_SINGLE
#define LINK_BLOCK_SINGLE 128
#endif
#include "mor_device.cuh"
#include "../../include/mor_b200.h"

namespace mor {

constexpr int kSingle = 1024;  // threads of the single-block bookkeeping kernels
constexpr int kBoxMinCount = 8;  // cells with more points than this carry a tight bounding box
constexpr int kGridPad = 2;  // empty cells on the low side of every axis: backward neighbour look-ups need no bounds checks
constexpr int kRootCheckCount = 48;  // far pass: cells above this size are first checked for a common root

enum ErrBits { ERR_CLUSTER_CAP = 1, ERR_MOVING_CAP = 2, ERR_LATTICE_RANGE = 4, ERR_GROUND_CAP = 8, ERR_GRID_CAP = 16 };

struct GridDesc {  // uniform grid over `cloud`; cell edge h = r/sqrt(3)*(1-2^-10): two points of one cell are always
                   // within r (one union-find node per cell) and d<r => |dcell| <= 2 per axis. The origin lies kGridPad
                   // cells below the data on every axis, so every real cell has coordinates >= kGridPad
    double ox, oy, oz, inv_h;
    int nx, ny, nz, ncells;
};

struct Scratch {  // zeroed at the start of every frame (one memset, together with cell_count)
    int ticket_ingest, ticket_cells, ticket_out, n_roots;
    int moving_total, stats_blocks_done, out_blocks_done, flatten_blocks_done;
    int moving_blocks_done, pad5, pad6, pad7;
    unsigned box_inv_min[3], box_max[3];  // dynamic grid: bbox of `cloud` (ordered keys; mins stored inverted so 0 is neutral)
    int pad3, pad4;
};

struct TrackState {  // persists across frames (MovingObjectRemoval members, .h:109-128)
    int n_mo[2];      // mo_vec.size(), double-buffered like mo_centroid / mo_conf (see k_filter_output)
    int res_count, res_head;    // res_vec deque
    int corr_count, corr_head;  // corrs_vec deque
    int frames;
    int extract_overflow;
    int n_markers;    // mo_vec entries the last filterCloud looked up (one bounding-box marker each, cpp:640-642)
};

// Everything a kernel needs, passed by value (fits the 4 KB parameter space comfortably).
struct FramePtrs {
    // ---- input
    const uint8_t* in; uint32_t n, step, off_x, off_y, off_z, off_i; int vec16;
    // ---- config
    float trim_x, trim_y, trim_z, gp_limit, r2, volume_constraint, pde_lb, pde_ub, pde_thr, leave_off, catch_up;
    long long min_cluster, max_cluster;
    int method, opc_factor, moving_confidence, static_confidence;
    int kmax, momax, ring_depth;
    GridDesc grid;            // static mode: the config crop box (known at create)
    GridDesc* dgrid;          // the grid every kernel after k_ingest/k_keys reads (device copy; rewritten per frame in dynamic mode)
    int dynamic_grid;         // 1: the config box would need too many cells -> per-frame grid over the bounding box of `cloud`
    int max_cells; double cell_h;
    // ---- per-frame scratch
    Scratch* scratch; unsigned long long* st_ingest; unsigned long long* st_cells; unsigned long long* st_out;
    int* cell_count; int* cell_start;
    uint8_t* point_class; uint8_t* removed_mask;
    int* cloud_src; float4* gpts; int* gsrc;
    int* cell_key; int* skey;
    int* parent; int* label; int* comp_size; int* root_list; int* cid_of_root;
    int* comp; int* minidx; unsigned long long* done;  // indexed by sorted position (cell leaders)
    int* scid;        // cluster id per sorted position (coalesced companion of spts for the method-1 search)
    uint4* cell_box;  // [2*N] per leader position: {min x,y,z keys, count}, {max x,y,z keys, 0}
    unsigned long long* acc_sum;  // [kmax*6] hi/lo per axis
    unsigned* acc_box;            // [kmax*6] min xyz, max xyz keys
    unsigned* pacc_box;           // [kmax*6] transformed prev clusters
    float4* tpts;                 // transformed prev cloud points (w = prev cluster id bits, -1 if none)
    float* pct;                   // [kmax*3] transformed prev centroids
    float* pbbox;                 // [kmax*6] decoded
    int* recip_q; int* recip_m; int* match_q; int* match_m; float* match_dist; double* match_score;
    int* match_of_prev; int* mid_of_prev; int* mid_of_cur; double* anchor; int* newcount;
    unsigned long long* lattice; unsigned lattice_mask;
    uint8_t* cluster_removed; int* found;
    int* marker_cluster;          // [momax] cluster each mo_vec entry was matched to by the last filterCloud
    float4* out;
    // ---- ping-pong frame state: cur / prev
    float4* pts; float4* spts; int* cid; int* cl_root; int* cl_size; float* cl_centroid; uint8_t* cl_flags; float* cl_bbox; int* counts;
    const float4* p_pts; const float4* p_spts; const int* p_cid; const int* p_cl_root; const int* p_cl_size; const float* p_cl_centroid;
    const uint8_t* p_cl_flags; const int* p_counts;
    // ---- persistent tracking state
    TrackState* track; float* mo_centroid; int* mo_conf;
    uint8_t* res_ring; int* res_len; int* corr_ring; int* corr_len;
    Affine12 M; int two_frames;
    int mo_parity;  // which half of the mo_vec double buffer is current
    int pde_ring;   // method 1: search reach in cells, ceil(sqrt(pde_ub)/h)
    int tiles_pts, tiles_cells;  // sizes of the scan status arrays
    unsigned lattice_words16;    // lattice size in 16-byte units (cleared by k_ingest)
};

// ===================================================================================== K1
// pcl::fromPCLPointCloud2 (cpp:523) + PassThrough x, y (cpp:66-74, A1) + CropBox with removed indices
// (cpp:78-86, A2), fused with the stable two-way partition into `cloud` / gp_indices order, the grid
// cell key of every cloud point and the per-cell histogram. kIngestItems consecutive points per thread
// (tile = 1024 points): the decoupled look-back chain has n/1024 links instead of n/256.
#ifndef INGEST_BLOCK
#define INGEST_BLOCK 256
#endif
constexpr int kIngestBlock = INGEST_BLOCK;
constexpr int kIngestTile = 1024;
constexpr int kIngestItems = kIngestTile / kIngestBlock;

__device__ __forceinline__ void k_ingest_body(const FramePtrs& a) {
    pdl_prologue();
    __shared__ int s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(&a.scratch->ticket_ingest, 1);
    if (a.two_frames) {
        // housekeeping for the two-frame stages, spread over the grid and overlapped with the ticket's round
        // trip: empty octree-lattice hash set, neutral bounding boxes for the transformed previous clusters
        const uint32_t gtid = blockIdx.x * kIngestBlock + threadIdx.x, stride = gridDim.x * kIngestBlock;
        if (a.method == 2) {
            uint4* lat = reinterpret_cast<uint4*>(a.lattice);
            for (uint32_t t = gtid; t < a.lattice_words16; t += stride) lat[t] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        }
        const uint32_t kp6 = (uint32_t)a.p_counts[MOR_CNT_K] * 6u;
        for (uint32_t t = gtid; t < kp6; t += stride) a.pacc_box[t] = (t % 6u) < 3u ? 0xFFFFFFFFu : 0u;
    }
    __syncthreads();
    const int tile = s_tile;
    const uint32_t i0 = (uint32_t)tile * kIngestTile + threadIdx.x * kIngestItems;
    float4 v[kIngestItems];
    int cls[kIngestItems];
    unsigned long long packed = 0ull;
#pragma unroll
    for (int k = 0; k < kIngestItems; k++) {
        const uint32_t i = i0 + k;
        v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        cls[k] = 0;
        if (i < a.n) {
            const uint8_t* p = a.in + (size_t)i * a.step;
            if (a.vec16) {
                v[k] = __ldg(reinterpret_cast<const float4*>(p));
            } else {
                v[k].x = __ldg(reinterpret_cast<const float*>(p + a.off_x));
                v[k].y = __ldg(reinterpret_cast<const float*>(p + a.off_y));
                v[k].z = __ldg(reinterpret_cast<const float*>(p + a.off_z));
                v[k].w = a.off_i != 0xFFFFFFFFu ? __ldg(reinterpret_cast<const float*>(p + a.off_i)) : 0.f;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kIngestItems; k++) {
        const uint32_t i = i0 + k;
        if (i < a.n) {
            const float x = v[k].x, y = v[k].y, z = v[k].z;
            const bool fin = isfinite(x) && isfinite(y) && isfinite(z);
            const bool in_xy = fin && !(x < -a.trim_x || x > a.trim_x) && !(y < -a.trim_y || y > a.trim_y);
            if (in_xy) cls[k] = (z < a.gp_limit || z > a.trim_z) ? 2 : 1;  // x,y box tests of CropBox are implied by the trim
            a.point_class[i] = (uint8_t)cls[k];
            a.removed_mask[i] = cls[k] ? 1 : 0;
            packed += (cls[k] == 1 ? 1ull : 0ull) | (cls[k] == 2 ? (1ull << 31) : 0ull);
        }
    }
    unsigned long long total;
    const unsigned long long in_block = block_exclusive_scan<unsigned long long, kIngestBlock>(packed, &total);
    const unsigned long long before = tile_exclusive_prefix(a.st_ingest, tile, total);
    unsigned long long mine = before + in_block;
#pragma unroll
    for (int k = 0; k < kIngestItems; k++) {
        int key = -1;
        if (cls[k] == 1 && !a.dynamic_grid) {
            const GridDesc& g = a.grid;
            int cx = (int)floor(((double)v[k].x - g.ox) * g.inv_h);
            int cy = (int)floor(((double)v[k].y - g.oy) * g.inv_h);
            int cz = (int)floor(((double)v[k].z - g.oz) * g.inv_h);
            cx = min(max(cx, kGridPad), g.nx - 1); cy = min(max(cy, kGridPad), g.ny - 1); cz = min(max(cz, kGridPad), g.nz - 1);
            key = (cz * g.ny + cy) * g.nx + cx;
            atomicAdd(&a.cell_count[key], 1);  // result unused: a fire-and-forget RED, no round trip on the critical path
        }
        if (cls[k] == 1) {
            const int c = (int)(mine & 0x7FFFFFFFull);
            mine += 1ull;
            a.pts[c] = v[k];
            a.cloud_src[c] = (int)(i0 + k);
            a.cell_box[2 * c] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u);  // slot c doubles as a sorted position
            a.cell_box[2 * c + 1] = make_uint4(0u, 0u, 0u, 0u);
            if (!a.dynamic_grid) a.cell_key[c] = key;
        } else if (cls[k] == 2) {
            const int gi = (int)((mine >> 31) & 0x7FFFFFFFull);
            mine += 1ull << 31;
            a.gpts[gi] = v[k];
            a.gsrc[gi] = (int)(i0 + k);
        }
    }
    if (a.dynamic_grid) {  // bounding box of `cloud`: warp redux, then one set of atomics per warp
        unsigned ix = 0u, iy = 0u, iz = 0u, mx = 0u, my = 0u, mz = 0u;
#pragma unroll
        for (int k = 0; k < kIngestItems; k++)
            if (cls[k] == 1) {
                const unsigned kx = fkey(v[k].x), ky = fkey(v[k].y), kz = fkey(v[k].z);
                ix = max(ix, ~kx); iy = max(iy, ~ky); iz = max(iz, ~kz); mx = max(mx, kx); my = max(my, ky); mz = max(mz, kz);
            }
        ix = __reduce_max_sync(kFull, ix); iy = __reduce_max_sync(kFull, iy); iz = __reduce_max_sync(kFull, iz);
        mx = __reduce_max_sync(kFull, mx); my = __reduce_max_sync(kFull, my); mz = __reduce_max_sync(kFull, mz);
        if ((threadIdx.x & 31) == 0 && (ix | mx)) {
            atomicMax(&a.scratch->box_inv_min[0], ix); atomicMax(&a.scratch->box_inv_min[1], iy); atomicMax(&a.scratch->box_inv_min[2], iz);
            atomicMax(&a.scratch->box_max[0], mx); atomicMax(&a.scratch->box_max[1], my); atomicMax(&a.scratch->box_max[2], mz);
        }
    }
    const int last_tile = a.n ? (int)((a.n - 1) / kIngestTile) : 0;
    if (tile == last_tile && threadIdx.x == 0) {
        const unsigned long long all = before + total;
        const int nc = (int)(all & 0x7FFFFFFFull), ng = (int)((all >> 31) & 0x7FFFFFFFull);
        int* c = a.counts;
        for (int k = 0; k < MOR_NCOUNTS; k++) c[k] = 0;
        c[MOR_CNT_N] = (int)a.n; c[MOR_CNT_NT] = nc + ng; c[MOR_CNT_NC] = nc; c[MOR_CNT_NG] = ng;
        c[MOR_CNT_TWO_FRAMES] = a.two_frames;
        if (a.two_frames) { c[MOR_CNT_KPREV] = a.p_counts[MOR_CNT_K]; c[MOR_CNT_NCPREV] = a.p_counts[MOR_CNT_NC]; }
        c[MOR_CNT_FRAME] = a.track->frames + 1;
        a.track->frames += 1;
    }
}
__global__ void __launch_bounds__(kIngestBlock) k_ingest(FramePtrs a) { k_ingest_body(a); }
__global__ void __launch_bounds__(kIngestBlock) k_ingest_batch(const FramePtrs* __restrict__ P) { k_ingest_body(P[blockIdx.z]); }


// ===================================================================================== K1b (dynamic grid only)
// When the config crop box would need more cells than the table holds (e.g. trim "disabled" with huge
// values), the grid is laid over the bounding box of this frame's `cloud` instead. Every block derives
// the same GridDesc from the reduced box; block 0 publishes it for the later kernels.
__device__ __forceinline__ GridDesc grid_from_box(const FramePtrs& a, bool* too_big) {
    GridDesc g;
    const Scratch* sc = a.scratch;
    const double inv_h = 1.0 / a.cell_h;
    double lo[3], hi[3];
#pragma unroll
    for (int q = 0; q < 3; q++) { lo[q] = (double)fkey_inv(~sc->box_inv_min[q]); hi[q] = (double)fkey_inv(sc->box_max[q]); }
    if (a.counts[MOR_CNT_NC] == 0) { lo[0] = lo[1] = lo[2] = 0; hi[0] = hi[1] = hi[2] = 0; }
    const double pad = (double)kGridPad * a.cell_h;
    g.ox = lo[0] - pad; g.oy = lo[1] - pad; g.oz = lo[2] - pad; g.inv_h = inv_h;
    const double fx = floor((hi[0] - g.ox) * inv_h) + 1.0, fy = floor((hi[1] - g.oy) * inv_h) + 1.0, fz = floor((hi[2] - g.oz) * inv_h) + 1.0;
    *too_big = fx * fy * fz > (double)a.max_cells;
    if (*too_big) { g.nx = g.ny = g.nz = kGridPad + 1; }  // memory-safe degenerate grid (one real cell); the frame is flagged
    else { g.nx = (int)fx; g.ny = (int)fy; g.nz = (int)fz; }
    g.ncells = g.nx * g.ny * g.nz;
    return g;
}

__device__ __forceinline__ void k_keys_body(const FramePtrs& a) {
    pdl_prologue();
    __shared__ GridDesc s_g;
    if (threadIdx.x == 0) {
        bool too_big;
        s_g = grid_from_box(a, &too_big);
        if (blockIdx.x == 0) {
            *a.dgrid = s_g;
            if (too_big) atomicOr(&a.counts[MOR_CNT_ERRFLAGS], ERR_GRID_CAP);
        }
    }
    __syncthreads();
    const int c = blockIdx.x * kBlock + threadIdx.x;
    if (c >= a.counts[MOR_CNT_NC]) return;
    const GridDesc& g = s_g;
    const float4 p = a.pts[c];
    int cx = (int)floor(((double)p.x - g.ox) * g.inv_h);
    int cy = (int)floor(((double)p.y - g.oy) * g.inv_h);
    int cz = (int)floor(((double)p.z - g.oz) * g.inv_h);
    cx = min(max(cx, kGridPad), g.nx - 1); cy = min(max(cy, kGridPad), g.ny - 1); cz = min(max(cz, kGridPad), g.nz - 1);
    const int key = (cz * g.ny + cy) * g.nx + cx;
    a.cell_key[c] = key;
    atomicAdd(&a.cell_count[key], 1);
}
__global__ void __launch_bounds__(kBlock) k_keys(FramePtrs a) { k_keys_body(a); }
__global__ void __launch_bounds__(kBlock) k_keys_batch(const FramePtrs* __restrict__ P) { k_keys_body(P[blockIdx.z]); }


// ===================================================================================== K2
// Exclusive scan of the per-cell histogram (counting sort of the cell keys) -> cell_start[0..ncells]. The histogram
// itself is left in place: k_scatter counts it back down to zero (rank = atomicSub - 1), which both hands out the
// slots of a cell and leaves the table clean for the next frame - the dense table is read once and written once.
// Persistent blocks pull tiles by ticket, so the launch does not depend on the (possibly device-side)
// cell count.
// Tile size: 2048 cells for one sequence (more, shorter tiles overlap better with the neighbouring kernels of the
// chain: +5 % frame rate), 4096 in batches (fewer look-back hops per byte: +3 %). Arrays are sized for the smaller.
constexpr int kScanItems = 8, kScanItemsBatch = 16;
constexpr int kScanTile = kBlock * kScanItems, kScanTileBatch = kBlock * kScanItemsBatch;

template <int ITEMS>
__device__ __forceinline__ void k_scan_cells_body(const FramePtrs& a) {
    pdl_prologue();
    __shared__ int s_tile;
    const int ncells = a.dgrid->ncells;
    const int ntiles = (ncells + (kBlock * ITEMS) - 1) / (kBlock * ITEMS);
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(&a.scratch->ticket_cells, 1);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= ntiles) return;
        const int base = tile * (kBlock * ITEMS) + threadIdx.x * ITEMS;
        int v[ITEMS];
        if (base + ITEMS <= ncells) {
            const int4* src = reinterpret_cast<const int4*>(a.cell_count + base);
#pragma unroll
            for (int k = 0; k < ITEMS / 4; k++) {
                const int4 t = src[k];
                v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < ITEMS; k++) v[k] = (base + k < ncells) ? a.cell_count[base + k] : 0;
        }
        int sum = 0;
#pragma unroll
        for (int k = 0; k < ITEMS; k++) sum += v[k];
        int total;
        const int in_block = block_exclusive_scan<int>(sum, &total);
        const int before = (int)tile_exclusive_prefix(a.st_cells, tile, (unsigned long long)total);
        int run = before + in_block;
        if (base + ITEMS <= ncells) {
            int4* d0 = reinterpret_cast<int4*>(a.cell_start + base);
#pragma unroll
            for (int k = 0; k < ITEMS / 4; k++) {
                int4 t;
                t.x = run; run += v[4 * k];
                t.y = run; run += v[4 * k + 1];
                t.z = run; run += v[4 * k + 2];
                t.w = run; run += v[4 * k + 3];
                d0[k] = t;
            }
        } else {
#pragma unroll
            for (int k = 0; k < ITEMS; k++) {
                if (base + k < ncells) a.cell_start[base + k] = run;
                run += v[k];
            }
        }
        if (tile == ntiles - 1 && threadIdx.x == 0) a.cell_start[ncells] = before + total;
    }
}
__global__ void __launch_bounds__(kBlock) k_scan_cells(FramePtrs a) { k_scan_cells_body<kScanItems>(a); }
__global__ void __launch_bounds__(kBlock, 8) k_scan_cells_batch(const FramePtrs* __restrict__ P) { k_scan_cells_body<kScanItemsBatch>(P[blockIdx.z]); }


// ===================================================================================== K3
// Scatter cloud points into cell-sorted order (float4 xyz + cloud index) for the neighbour search and
// reset the per-position union-find state. The first point of a cell (its "leader" position
// cell_start[key]) is the union-find node of the whole cell.
__device__ __forceinline__ void k_scatter_body(const FramePtrs& a) {
    pdl_prologue();
    const int c = blockIdx.x * kBlock + threadIdx.x;
    if (c >= a.counts[MOR_CNT_NC]) return;
    const int key = a.cell_key[c];
    const int pos = a.cell_start[key] + atomicSub(&a.cell_count[key], 1) - 1;  // slots of a cell are handed out last to first
    float4 p = a.pts[c];
    p.w = __int_as_float(c);
    a.spts[pos] = p;
    a.skey[pos] = key;
    a.parent[c] = c;  // c doubles as a sorted position here: both index spaces are [0, N_c)
    a.comp_size[c] = 0;
    a.minidx[c] = 0x7FFFFFFF;
    a.done[c] = 0ull;
    // tight bounding box of every crowded cell (prunes the point-vs-cell scans of k_link_cells);
    // lanes of a warp that fall into the same cell are combined with redux before the atomics
    const int cnt = a.cell_start[key + 1] - a.cell_start[key];
    if (cnt > kBoxMinCount) {
        const unsigned grp = __match_any_sync(__activemask(), key);
        const unsigned kx = fkey(p.x), ky = fkey(p.y), kz = fkey(p.z);
        const unsigned mnx = __reduce_min_sync(grp, kx), mny = __reduce_min_sync(grp, ky), mnz = __reduce_min_sync(grp, kz);
        const unsigned mxx = __reduce_max_sync(grp, kx), mxy = __reduce_max_sync(grp, ky), mxz = __reduce_max_sync(grp, kz);
        if ((int)(__ffs(grp) - 1) == (int)(threadIdx.x & 31)) {
            unsigned* b = reinterpret_cast<unsigned*>(a.cell_box + 2 * a.cell_start[key]);
            atomicMin(b + 0, mnx); atomicMin(b + 1, mny); atomicMin(b + 2, mnz);
            atomicMax(b + 4, mxx); atomicMax(b + 5, mxy); atomicMax(b + 6, mxz);
        }
    }
}
__global__ void __launch_bounds__(kBlock) k_scatter(FramePtrs a) { k_scatter_body(a); }
__global__ void __launch_bounds__(kBlock) k_scatter_batch(const FramePtrs* __restrict__ P) { k_scatter_body(P[blockIdx.z]); }


// ===================================================================================== K4
// pcl::EuclideanClusterExtraction radius graph (cpp:213-218; A5-A7) on cell granularity. Every point q
// looks at the 62 "backward" cells of its 5x5x5 neighbourhood (13 x-rows, each a contiguous run of the
// sorted array); the first point of such a cell with L2_Simple distance < r2 (strict) connects the two
// cells. A 62-bit mask per cell records the pairs already united, so each connected cell pair costs
// one atomicOr + one union, however many point pairs realise it.
__device__ __forceinline__ unsigned long long ld_done(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// Examines the sorted positions j..e-1 (the cells of one x-run of a neighbour row) for point q; see k_link_cells.
// `rowkey` is the key of the cell straight "above" q's cell in that row: the pair bit of cell kj is row*5 + (kj - rowkey + 2).
template <int PHASE>
__device__ __forceinline__ void link_scan_range(const FramePtrs& a, const float4 q, int lead, int row, int rowkey, int j, const int e,
                                                unsigned long long& dmask) {
    const float r2 = a.r2, r2_prune = a.r2 * 1.00001f;
    while (j < e) {
        const int kj = a.skey[j];
        const int cell_end = a.cell_start[kj + 1];
        const int bit = row * 5 + (kj - rowkey + 2);
        bool skip = (dmask >> bit) & 1ull;
        if (!skip && cell_end - j > kBoxMinCount) {
            // conservative point-to-box distance: no point of the cell can be closer than this
            const uint4 lo = a.cell_box[2 * j], hi = a.cell_box[2 * j + 1];
            const float ex = fmaxf(fmaxf(fkey_inv(lo.x) - q.x, q.x - fkey_inv(hi.x)), 0.f);
            const float ey = fmaxf(fmaxf(fkey_inv(lo.y) - q.y, q.y - fkey_inv(hi.y)), 0.f);
            const float ez = fmaxf(fmaxf(fkey_inv(lo.z) - q.z, q.z - fkey_inv(hi.z)), 0.f);
            skip = ex * ex + ey * ey + ez * ez > r2_prune;
            if (PHASE == 2 && !skip && cell_end - j > kRootCheckCount) {
                // far pass: the near pass has already merged most of a dense surface; two cells of one
                // component need no point tests (the pair is marked done, which is all the mask means)
                // one lane per (cell pair) of the converged lanes walks the two paths and shares the verdict
                const unsigned grp = __match_any_sync(__activemask(), (lead << 6) | bit);
                const int leader_lane = __ffs(grp) - 1;
                int same = 0;
                if ((int)(threadIdx.x & 31) == leader_lane) {
                    same = uf_find(a.parent, lead) == uf_find(a.parent, j) ? 1 : 0;
                    if (same) atomicOr(a.done + lead, 1ull << bit);
                }
                same = __shfl_sync(grp, same, leader_lane);
                if (same) { dmask |= 1ull << bit; skip = true; }
            }
        }
        if (!skip) {
            bool hit = false;
            const int other = j;  // leader position of the neighbour cell
            const int last = cell_end - 1;
            for (int it = 0; j < cell_end && !hit; j += 4) {  // always 4 independent loads in flight (indices clamped)
                const float4 p0 = a.spts[j], p1 = a.spts[min(j + 1, last)], p2 = a.spts[min(j + 2, last)], p3 = a.spts[min(j + 3, last)];
                const float d0 = sqdist3(q.x, q.y, q.z, p0.x, p0.y, p0.z), d1 = sqdist3(q.x, q.y, q.z, p1.x, p1.y, p1.z);
                const float d2 = sqdist3(q.x, q.y, q.z, p2.x, p2.y, p2.z), d3 = sqdist3(q.x, q.y, q.z, p3.x, p3.y, p3.z);
                hit = fminf(fminf(d0, d1), fminf(d2, d3)) < r2;
                if (((++it) & 63) == 63 && !hit) {  // somebody else of my cell may have connected this pair meanwhile
                    dmask |= ld_done(a.done + lead);
                    if ((dmask >> bit) & 1ull) break;
                }
            }
            if (hit) {
                // lanes of the warp that found the same cell pair at the same time elect one publisher: the
                // done word of a crowded cell would otherwise take thousands of same-address atomics
                const unsigned grp = __match_any_sync(__activemask(), (lead << 6) | bit);
                if ((int)(__ffs(grp) - 1) == (int)(threadIdx.x & 31)) {
                    dmask |= ld_done(a.done + lead);
                    if (!((dmask >> bit) & 1ull)) {
                        const unsigned long long old = atomicOr(a.done + lead, 1ull << bit);
                        if (!((old >> bit) & 1ull)) uf_union(a.parent, lead, other);
                    }
                }
                dmask |= 1ull << bit;
            }
        }
        j = cell_end;
    }
}

// grid = (point tiles, rows): one thread per (point q, x-row of the backward neighbourhood), so the serial
// chain of a thread is a handful of cells and a warp walks the same cells for neighbouring q.
// PHASE 1 = the 13 backward cells of the 3x3x3 block (5 rows); PHASE 2 = the 49 cells at offset 2 (13 rows).
// Neighbour rows are addressed linearly from the point's own key (key + dy*nx + dz*nx*ny +- 2): the grid carries
// kGridPad empty cells on the low side of every axis, so a backward offset never leaves the table and an offset
// that runs over the high end of a row / layer lands in the next row's / layer's padding, which is always empty.
// Block size: at 32 registers an SM holds 64 warps either way; one sequence alone is a latency chain in which a
// block should retire as soon as its slowest warp does (small blocks), a batch of sequences keeps the SMs full and
// pays per block (large blocks).
constexpr int kLinkBlock = LINK_BLOCK_SINGLE, kLinkBlockBatch = 256;
template <int PHASE, int BLOCK>
__device__ __forceinline__ void k_link_cells_body(const FramePtrs& a, int row_index) {
    pdl_prologue();
    const int s = blockIdx.x * BLOCK + threadIdx.x;
    const int nc = a.counts[MOR_CNT_NC];
    if (s >= nc) return;
    // canonical row ids (they fix the bit layout): 0-4: dz=-2, 5-9: dz=-1, 10-12: dz=0 with dy=-2,-1,0
    int row = row_index;
    if (PHASE == 1) row = row_index == 0 ? 12 : (row_index == 1 ? 11 : 4 + row_index);  // near rows: 12, 11, 6, 7, 8
    const int dz = row < 5 ? -2 : (row < 10 ? -1 : 0);
    const int dy = row < 10 ? (row % 5) - 2 : row - 12;
    const bool near_row = dz >= -1 && dy >= -1 && dy <= 1;
    const int key = a.skey[s];
    const int nx = a.dgrid->nx, ny = a.dgrid->ny;
    const int rowkey = key + dy * nx + dz * nx * ny;
    // the x-runs of this (point, row): most of them are empty, so their bounds are fetched before anything else
    int ka, kb, kc = 0, kd = -1;
    if (PHASE == 1) { ka = rowkey - 1; kb = row == 12 ? rowkey - 1 : rowkey + 1; }
    else if (!near_row) { ka = rowkey - 2; kb = rowkey + 2; }
    else { ka = kb = rowkey - 2; if (row != 12) { kc = kd = rowkey + 2; } }
    const int j0 = a.cell_start[ka], e0 = a.cell_start[kb + 1];
    int j1 = 0, e1 = 0;
    if (kd >= kc) { j1 = a.cell_start[kc]; e1 = a.cell_start[kd + 1]; }
    if (j0 >= e0 && j1 >= e1) return;
    const float4 q = a.spts[s];
    const int lead = a.cell_start[key];
    unsigned long long dmask = ld_done(a.done + lead);
    if (j0 < e0) link_scan_range<PHASE>(a, q, lead, row, rowkey, j0, e0, dmask);
    if (j1 < e1) link_scan_range<PHASE>(a, q, lead, row, rowkey, j1, e1, dmask);
}
// One launch for both passes: blockIdx.y 0..4 = near pass, 5..17 = far pass. Blocks are dispatched y-major, so the
// near rows start first and most of their unions are in place when the far rows run their root checks; the two
// passes' tails overlap instead of adding up (the far pass is correct with any amount of near-pass progress: its
// root check is only a shortcut).
__global__ void __launch_bounds__(kLinkBlock, 2048 / kLinkBlock) k_link_cells(FramePtrs a) {
    if (blockIdx.y < 5) k_link_cells_body<1, kLinkBlock>(a, blockIdx.y); else k_link_cells_body<2, kLinkBlock>(a, blockIdx.y - 5);
}
__global__ void __launch_bounds__(kLinkBlockBatch, 2048 / kLinkBlockBatch) k_link_cells_batch(const FramePtrs* __restrict__ P) {
    if (blockIdx.y < 5) k_link_cells_body<1, kLinkBlockBatch>(P[blockIdx.z], blockIdx.y); else k_link_cells_body<2, kLinkBlockBatch>(P[blockIdx.z], blockIdx.y - 5);
}

// ===================================================================================== K5
// Pointer-jump every cell leader to its root, count component sizes and reduce the minimum cloud
// index of every component (the canonical label) with warp-aggregated atomics, collect the roots.
__device__ void select_block(const FramePtrs& a);

__device__ __forceinline__ void k_flatten_body(const FramePtrs& a) {
    pdl_prologue();
    const int s = blockIdx.x * kSingle + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int nc = a.counts[MOR_CNT_NC];
    const bool act = s < nc;
    int c = 0, r = -1 - lane;
    if (act) {
        c = __float_as_int(a.spts[s].w);
        const int lead = a.cell_start[a.skey[s]];
        r = uf_find(a.parent, lead);
        a.comp[s] = r;
        if (r == s) a.root_list[atomicAdd(&a.scratch->n_roots, 1)] = s;
    }
    const unsigned same = __match_any_sync(kFull, r);
    const int mn = __reduce_min_sync(same, c);
    if (act && (int)(__ffs(same) - 1) == lane) {
        atomicAdd(&a.comp_size[r], __popc(same));
        atomicMin(&a.minidx[r], mn);
    }
    // the last block to finish selects and orders the clusters (K6) - no separate single-block launch
    __shared__ int s_last;
    __threadfence();  // root_list entries are plain stores of arbitrary threads
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(&a.scratch->flatten_blocks_done, 1) == (int)gridDim.x - 1;
        __threadfence();
    }
    __syncthreads();
    if (s_last) select_block(a);
}
__global__ void __launch_bounds__(kSingle) k_flatten(FramePtrs a) { k_flatten_body(a); }
__global__ void __launch_bounds__(kSingle) k_flatten_batch(const FramePtrs* __restrict__ P) { k_flatten_body(P[blockIdx.z]); }


// ===================================================================================== K6
// Size filter min <= size <= max (cpp:215-216), cluster order = size descending then min index
// ascending (A9 canonical rule) by a shared-memory bitonic sort of (~size, root) keys.
__device__ void select_block(const FramePtrs& a) {
    extern __shared__ unsigned long long keys[];
    __shared__ int s_k;
    if (threadIdx.x == 0) s_k = 0;
    __syncthreads();
    const int n_roots = __ldcg(&a.scratch->n_roots);
    for (int t = threadIdx.x; t < n_roots; t += kSingle) {
        const int rpos = a.root_list[t];
        const int sz = __ldcg(&a.comp_size[rpos]);
        const int root = __ldcg(&a.minidx[rpos]);  // min cloud index of the component = canonical label
        if ((long long)sz >= a.min_cluster && (long long)sz <= a.max_cluster) {
            const int slot = atomicAdd(&s_k, 1);
            if (slot < a.kmax) keys[slot] = ((unsigned long long)(0xFFFFFFFFu - (unsigned)sz) << 32) | (unsigned)root;
        } else {
            a.cid_of_root[root] = -1;
        }
    }
    __syncthreads();
    int K = s_k;
    if (K > a.kmax) {  // capacity exceeded: keep the first kmax found, flag the frame
        if (threadIdx.x == 0) atomicOr(&a.counts[MOR_CNT_ERRFLAGS], ERR_CLUSTER_CAP);
        // the dropped roots must not keep a stale cluster id
        for (int t = threadIdx.x; t < n_roots; t += kSingle) a.cid_of_root[__ldcg(&a.minidx[a.root_list[t]])] = -1;
        K = a.kmax;
    }
    int P = 1;
    while (P < K) P <<= 1;
    for (int t = K + threadIdx.x; t < P; t += kSingle) keys[t] = ~0ull;
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += kSingle) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const unsigned long long A = keys[lo], B = keys[hi];
                if ((A > B) == up) { keys[lo] = B; keys[hi] = A; }
            }
            __syncthreads();
        }
    }
    int nk = 0;
    for (int k = threadIdx.x; k < K; k += kSingle) {
        const unsigned long long kk = keys[k];
        const int root = (int)(unsigned)(kk & 0xFFFFFFFFull);
        const int sz = (int)(0xFFFFFFFFu - (unsigned)(kk >> 32));
        a.cl_root[k] = root; a.cl_size[k] = sz; a.cid_of_root[root] = k; a.cl_flags[k] = 0;
        nk += sz;
#pragma unroll
        for (int q = 0; q < 6; q++) a.acc_sum[k * 6 + q] = 0ull;
#pragma unroll
        for (int q = 0; q < 3; q++) { a.acc_box[k * 6 + q] = 0xFFFFFFFFu; a.acc_box[k * 6 + 3 + q] = 0u; }
    }
    atomicAdd(&a.counts[MOR_CNT_NK], nk);
    if (threadIdx.x == 0) a.counts[MOR_CNT_K] = K;
    // the ingest / cell scans of this frame are complete: reset their look-back state for the next frame
    for (int t = threadIdx.x; t < a.tiles_pts; t += kSingle) a.st_ingest[t] = 0ull;
    for (int t = threadIdx.x; t < a.tiles_cells; t += kSingle) a.st_cells[t] = 0ull;
    if (threadIdx.x == 0) {
        a.scratch->ticket_ingest = 0; a.scratch->ticket_cells = 0; a.scratch->n_roots = 0; a.scratch->stats_blocks_done = 0;
        a.scratch->flatten_blocks_done = 0; a.scratch->moving_blocks_done = 0;
    }
    if (threadIdx.x < 3) { a.scratch->box_inv_min[threadIdx.x] = 0u; a.scratch->box_max[threadIdx.x] = 0u; }
}

// ===================================================================================== K7
// Per-cluster statistics (cpp:221-244): cluster id of every point, exact coordinate sums for
// compute3DCentroid<double> (A10) and getMinMax3D bounding boxes (cpp:272-275), warp-aggregated.
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ unsigned warp_min_u(unsigned v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ unsigned warp_max_u(unsigned v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// Per-cluster accumulation of coordinate sums (optional) and bounding boxes with two levels of
// aggregation before the global atomics: warp (shuffles) and block (shared memory). In sorted order a
// block of kStatBlock consecutive points lies inside one cluster most of the time, so a 30k-point
// cluster costs ~30 sets of atomics instead of 30k. Every thread of the block must call.
constexpr int kStatBlock = 1024;

template <bool WITH_SUMS>
__device__ __forceinline__ void block_cluster_accumulate(unsigned long long* acc_sum, unsigned* acc_box, int k, bool valid, float x, float y, float z) {
    __shared__ int s_k[kStatBlock / 32];                       // cluster of the warp, -1 = no valid lane, -2 = mixed
    __shared__ unsigned long long s_sum[kStatBlock / 32][6];
    __shared__ unsigned s_box[kStatBlock / 32][6];
    __shared__ int s_mode;                                     // >= 0: the whole block is cluster s_mode; -1: per-warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // warp level: lanes are grouped by cluster (match.any) and every group is reduced with redux
    const unsigned vmask = __ballot_sync(kFull, valid);
    const unsigned grp = __match_any_sync(kFull, valid ? k : -1);
    const bool leader = valid && (int)(__ffs(grp) - 1) == lane;
    const bool uniform = vmask != 0 && (grp & vmask) == vmask && valid;  // true in the valid lanes of a one-cluster warp
    const bool warp_uniform = __any_sync(kFull, uniform);
    unsigned bx[6];
    {
        const unsigned kx = fkey(x), ky = fkey(y), kz = fkey(z);
        bx[0] = __reduce_min_sync(grp, kx); bx[1] = __reduce_min_sync(grp, ky); bx[2] = __reduce_min_sync(grp, kz);
        bx[3] = __reduce_max_sync(grp, kx); bx[4] = __reduce_max_sync(grp, ky); bx[5] = __reduce_max_sync(grp, kz);
    }
    unsigned long long sm[6] = {0, 0, 0, 0, 0, 0};
    if (WITH_SUMS) {
        const float v[3] = {x, y, z};
#pragma unroll
        for (int q = 0; q < 3; q++) {
            long long h, l;
            split_fixed(v[q], h, l);  // h in [-2^31, 2^31), l in [0, 2^30): summed in 16/15-bit pieces so redux.add (32-bit) cannot overflow
            const long long sh = ((long long)__reduce_add_sync(grp, (int)(h >> 16)) << 16) + (long long)__reduce_add_sync(grp, (int)(h & 0xFFFF));
            const long long sl = ((long long)__reduce_add_sync(grp, (int)(l >> 15)) << 15) + (long long)__reduce_add_sync(grp, (int)(l & 0x7FFF));
            sm[q * 2] = (unsigned long long)sh; sm[q * 2 + 1] = (unsigned long long)sl;
        }
    }
    if (warp_uniform && leader) {
#pragma unroll
        for (int q = 0; q < 6; q++) { s_box[warp][q] = bx[q]; if (WITH_SUMS) s_sum[warp][q] = sm[q]; }
    }
    const int k_first = __shfl_sync(kFull, k, vmask ? __ffs(vmask) - 1 : 0);
    if (lane == 0) s_k[warp] = !vmask ? -1 : (warp_uniform ? k_first : -2);
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        const int wk = lane < nw ? s_k[lane] : -1;
        const unsigned has = __ballot_sync(kFull, wk != -1);
        const int first = has ? __shfl_sync(kFull, wk, __ffs(has) - 1) : -1;
        const bool same = __all_sync(kFull, wk == -1 || (wk == first && wk >= 0));
        if (lane == 0) s_mode = (has && same) ? first : -1;
    }
    __syncthreads();
    const int mode = s_mode;
    if (mode >= 0) {  // the whole block is one cluster: one set of atomics
        const int nw = blockDim.x >> 5;
        if (threadIdx.x < 6) {
            unsigned v = threadIdx.x < 3 ? 0xFFFFFFFFu : 0u;
            for (int w = 0; w < nw; w++)
                if (s_k[w] >= 0) v = threadIdx.x < 3 ? min(v, s_box[w][threadIdx.x]) : max(v, s_box[w][threadIdx.x]);
            if (threadIdx.x < 3) atomicMin(acc_box + mode * 6 + threadIdx.x, v); else atomicMax(acc_box + mode * 6 + threadIdx.x, v);
        } else if (WITH_SUMS && threadIdx.x >= 32 && threadIdx.x < 38) {
            const int q = threadIdx.x - 32;
            unsigned long long v = 0;
            for (int w = 0; w < nw; w++)
                if (s_k[w] >= 0) v += s_sum[w][q];
            atomicAdd(acc_sum + mode * 6 + q, v);
        }
    } else if (leader) {  // one set of atomics per (warp, cluster) group
        unsigned* b = acc_box + k * 6;
        atomicMin(b + 0, bx[0]); atomicMin(b + 1, bx[1]); atomicMin(b + 2, bx[2]);
        atomicMax(b + 3, bx[3]); atomicMax(b + 4, bx[4]); atomicMax(b + 5, bx[5]);
        if (WITH_SUMS) {
#pragma unroll
            for (int q = 0; q < 6; q++) atomicAdd(acc_sum + k * 6 + q, sm[q]);
        }
    }
}

__device__ void match_block(const FramePtrs& a);

__device__ __forceinline__ void k_cluster_stats_body(const FramePtrs& a) {
    pdl_prologue();
    const int s = blockIdx.x * kStatBlock + threadIdx.x;
    const int nc = a.counts[MOR_CNT_NC];
    if (blockIdx.x * kStatBlock >= nc && !(nc == 0 && blockIdx.x == 0)) return;  // whole blocks stay alive for the barriers
    float4 p = make_float4(0, 0, 0, 0);
    int k = -1;
    if (s < nc) {
        p = a.spts[s];
        const int c = __float_as_int(p.w);
        const int lab = a.minidx[a.comp[s]];
        a.label[c] = lab;
        k = a.cid_of_root[lab];
        a.cid[c] = k;
        a.scid[s] = k;
    }
    block_cluster_accumulate<true>(a.acc_sum, a.acc_box, k, k >= 0, p.x, p.y, p.z);
    // the last block to finish turns the accumulators into centroids (compute3DCentroid<double>, A10)
    // and bounding boxes
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();  // the block's accumulator atomics are ordered before the ticket
        s_last = atomicAdd(&a.scratch->stats_blocks_done, 1) == max((nc + kStatBlock - 1) / kStatBlock, 1) - 1;
        __threadfence();
    }
    __syncthreads();
    if (!s_last) return;
    const int K = a.counts[MOR_CNT_K];
    for (int c = threadIdx.x; c < K; c += kStatBlock) {
        const double n = (double)a.cl_size[c];
#pragma unroll
        for (int q = 0; q < 3; q++)
            a.cl_centroid[c * 3 + q] = (float)join_fixed_mean((long long)__ldcg(&a.acc_sum[c * 6 + q * 2]), (long long)__ldcg(&a.acc_sum[c * 6 + q * 2 + 1]), n);
#pragma unroll
        for (int q = 0; q < 6; q++) a.cl_bbox[c * 6 + q] = fkey_inv(__ldcg(&a.acc_box[c * 6 + q]));
    }
    // ... and, with two frames, goes straight on to the cluster correspondences (K9)
    if (a.two_frames) {
        __syncthreads();
        match_block(a);
    }
}
__global__ void __launch_bounds__(kStatBlock) k_cluster_stats(FramePtrs a) { k_cluster_stats_body(a); }
__global__ void __launch_bounds__(kStatBlock, 2) k_cluster_stats_batch(const FramePtrs* __restrict__ P) { k_cluster_stats_body(P[blockIdx.z]); }


// ===================================================================================== K8
// pcl_ros::transformPointCloud of every previous-frame cluster (cpp:544-551, A12) and the bounding box
// of the transformed points (getMinMax3D runs after the transform, cpp:272).
__device__ __forceinline__ void k_transform_prev_body(const FramePtrs& a) {
    pdl_prologue();
    const int s = blockIdx.x * kStatBlock + threadIdx.x;
    const int ncp = a.p_counts[MOR_CNT_NC];
    if (blockIdx.x * kStatBlock >= ncp) return;
    int k = -1;
    float3 t = make_float3(0, 0, 0);
    if (s < ncp) {
        const float4 p = a.p_spts[s];
        const int c = __float_as_int(p.w);
        k = a.p_cid[c];
        if (k >= 0) {
            t = xform(a.M, p.x, p.y, p.z);
            a.tpts[c] = make_float4(t.x, t.y, t.z, __int_as_float(k));
        } else {
            a.tpts[c] = make_float4(0, 0, 0, __int_as_float(-1));
        }
    }
    block_cluster_accumulate<false>(nullptr, a.pacc_box, k, k >= 0, t.x, t.y, t.z);
}
__global__ void __launch_bounds__(kStatBlock) k_transform_prev(FramePtrs a) { k_transform_prev_body(a); }
__global__ void __launch_bounds__(kStatBlock) k_transform_prev_batch(const FramePtrs* __restrict__ P) { k_transform_prev_body(P[blockIdx.z]); }


// Block-wide ordered compaction helper for the single-block kernels: returns the exclusive rank of
// `flag` among all threads, *total = number of set flags. kSingle threads.
__device__ __forceinline__ int single_block_rank(bool flag, int* total) {
    __shared__ int s_w[kSingle / 32];
    __shared__ int s_tot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(kFull, flag);
    if (lane == 0) s_w[warp] = __popc(m);
    __syncthreads();
    if (warp == 0) {
        int v = s_w[lane], inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += t;
        }
        s_w[lane] = inc - v;
        if (lane == 31) s_tot = inc;
    }
    __syncthreads();
    const int r = s_w[warp] + __popc(m & ((1u << lane) - 1u));
    *total = s_tot;
    __syncthreads();
    return r;
}

// ===================================================================================== K9
// Finalise centroids/boxes, transform the previous centroids (cpp:540-541), reciprocal 1-NN between
// centroid sets (cpp:291-294, A16), volume constraint (cpp:264-283, A17), per-match octree anchors.
__device__ __forceinline__ int nn_brute(const float* pts, int n, float qx, float qy, float qz, float* out_d) {
    int best = -1;
    float bd = 3.402823466e+38f;
    for (int i = 0; i < n; i++) {
        const float d = sqdist3(qx, qy, qz, pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]);
        if (d < bd) { bd = d; best = i; }  // ties -> lowest index
    }
    *out_d = bd;
    return best;
}

__device__ void match_block(const FramePtrs& a) {
    const int K = a.counts[MOR_CNT_K], Kp = a.p_counts[MOR_CNT_K];
    // previous centroids and boxes into the current frame
    for (int i = threadIdx.x; i < Kp; i += kSingle) {
        const float3 t = xform(a.M, a.p_cl_centroid[i * 3], a.p_cl_centroid[i * 3 + 1], a.p_cl_centroid[i * 3 + 2]);
        a.pct[i * 3] = t.x; a.pct[i * 3 + 1] = t.y; a.pct[i * 3 + 2] = t.z;
#pragma unroll
        for (int q = 0; q < 6; q++) a.pbbox[i * 6 + q] = fkey_inv(__ldcg(&a.pacc_box[i * 6 + q]));
        a.match_of_prev[i] = -1; a.mid_of_prev[i] = -1;
    }
    for (int j = threadIdx.x; j < K; j += kSingle) a.mid_of_cur[j] = -1;
    __syncthreads();
    // reciprocal correspondences, ascending query index
    int n_recip = 0;
    for (int base = 0; base < Kp; base += kSingle) {
        const int i = base + threadIdx.x;
        bool ok = false; int j = -1; float d = 0.f;
        if (i < Kp && K > 0) {
            j = nn_brute(a.cl_centroid, K, a.pct[i * 3], a.pct[i * 3 + 1], a.pct[i * 3 + 2], &d);
            float dr;
            const int ir = nn_brute(a.pct, Kp, a.cl_centroid[j * 3], a.cl_centroid[j * 3 + 1], a.cl_centroid[j * 3 + 2], &dr);
            ok = (ir == i);
        }
        int tot;
        const int r = single_block_rank(ok, &tot);
        if (ok) { a.recip_q[n_recip + r] = i; a.recip_m[n_recip + r] = j; a.match_dist[n_recip + r] = d; }
        n_recip += tot;
    }
    __syncthreads();
    // volume constraint; match_dist is rewritten in place (rank <= index, chunked with barriers)
    int n_match = 0, p1 = 0, p2 = 0;
    for (int base = 0; base < n_recip; base += kSingle) {
        const int u = base + threadIdx.x;
        bool ok = false; int i = -1, j = -1; float d = 0.f;
        if (u < n_recip) {
            i = a.recip_q[u]; j = a.recip_m[u]; d = a.match_dist[u];
            const float* bp = a.pbbox + i * 6; const float* bc = a.cl_bbox + j * 6;
            const double volp = (double)__fmul_rn(__fmul_rn(__fsub_rn(bp[3], bp[0]), __fsub_rn(bp[4], bp[1])), __fsub_rn(bp[5], bp[2]));
            const double volc = (double)__fmul_rn(__fmul_rn(__fsub_rn(bc[3], bc[0]), __fsub_rn(bc[4], bc[1])), __fsub_rn(bc[5], bc[2]));
            ok = (fabs(volp - volc) / (volp + volc)) < (double)a.volume_constraint;  // NaN -> false
        }
        int tot;
        const int r = single_block_rank(ok, &tot);
        if (ok) {
            const int m = n_match + r;
            a.match_q[m] = i; a.match_m[m] = j; a.match_dist[m] = d;
            a.match_of_prev[i] = j; a.mid_of_prev[i] = m; a.mid_of_cur[j] = m;
            a.newcount[m] = 0; a.match_score[m] = 0.0;
            // OctreePointCloudChangeDetector lattice anchor from the first point of the previous cluster
            // (its min-index point) - PCL 1.8 adoptBoundingBoxToPoint + getKeyBitSize (A13)
            const float4 f = a.tpts[a.p_cl_root[i]];
            const double res = (double)0.1f, eps = (double)1.1920928955078125e-07f;
            const double fv[3] = {(double)f.x, (double)f.y, (double)f.z};
#pragma unroll
            for (int q = 0; q < 3; q++) {
                const double lo = fv[q] - res / 2, hi = fv[q] + res / 2;
                const double side = 2.0 * res - eps;
                const double over = (side - (hi - lo)) / 2.0;
                a.anchor[m * 3 + q] = lo - over;
            }
            atomicAdd(&a.counts[MOR_CNT_P1], a.p_cl_size[i]);
            atomicAdd(&a.counts[MOR_CNT_P2], a.cl_size[j]);
        }
        n_match += tot;
    }
    (void)p1; (void)p2;
    if (threadIdx.x == 0) {
        a.counts[MOR_CNT_MU] = n_recip; a.counts[MOR_CNT_M] = n_match;
        a.counts[MOR_CNT_NKPREV] = a.p_counts[MOR_CNT_NK];
    }
}

// ===================================================================================== K10 / K11 (method 2)
// pcl::octree::OctreePointCloudChangeDetector (cpp:319-330, A13): the leaf lattice of every matched
// pair is floor((p - anchor)/res) in double; the occupied leaves of the transformed previous cluster
// go into one global hash set keyed (match, ix, iy, iz); the score is the number of points of the
// current cluster whose leaf is absent.
__device__ __forceinline__ bool lattice_key(const FramePtrs& a, int m, float x, float y, float z, unsigned long long* key) {
    const double res = (double)0.1f;
    const long long ix = (long long)floor(((double)x - a.anchor[m * 3]) / res);
    const long long iy = (long long)floor(((double)y - a.anchor[m * 3 + 1]) / res);
    const long long iz = (long long)floor(((double)z - a.anchor[m * 3 + 2]) / res);
    const bool ok = ix >= -32768 && ix < 32768 && iy >= -32768 && iy < 32768 && iz >= -32768 && iz < 32768;
    *key = ((unsigned long long)(unsigned)m << 48) | ((unsigned long long)(ix + 32768) << 32) | ((unsigned long long)(iy + 32768) << 16) |
           (unsigned long long)(iz + 32768);
    return ok;
}

__device__ __forceinline__ void k_lattice_insert_body(const FramePtrs& a) {
    pdl_prologue();
    const int c = blockIdx.x * kBlock + threadIdx.x;
    if (c >= a.p_counts[MOR_CNT_NC]) return;
    const float4 t = a.tpts[c];
    const int k = __float_as_int(t.w);
    if (k < 0) return;
    const int m = a.mid_of_prev[k];
    if (m < 0) return;
    unsigned long long key;
    if (!lattice_key(a, m, t.x, t.y, t.z, &key)) { atomicOr(&a.counts[MOR_CNT_ERRFLAGS], ERR_LATTICE_RANGE); return; }
    hset_insert(a.lattice, a.lattice_mask, key);
}
__global__ void __launch_bounds__(kBlock) k_lattice_insert(FramePtrs a) { k_lattice_insert_body(a); }
__global__ void __launch_bounds__(kBlock) k_lattice_insert_batch(const FramePtrs* __restrict__ P) { k_lattice_insert_body(P[blockIdx.z]); }


__device__ void chain_block(const FramePtrs& a);

// Shared tail of the two moving-test kernels: the last block to finish turns the scores into flags and runs the
// consistency chain (K12).
__device__ __forceinline__ void moving_test_epilogue(const FramePtrs& a) {
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(&a.scratch->moving_blocks_done, 1) == (int)gridDim.x - 1;
        __threadfence();
    }
    __syncthreads();
    if (s_last) chain_block(a);
}

__device__ __forceinline__ void k_lattice_count_body(const FramePtrs& a) {
    pdl_prologue();
    const int s = blockIdx.x * kSingle + threadIdx.x;
    int m = -1;
    bool is_new = false;
    if (s < a.counts[MOR_CNT_NC]) {
        const float4 p = a.spts[s];
        const int c = __float_as_int(p.w);
        const int k = a.cid[c];
        m = k >= 0 ? a.mid_of_cur[k] : -1;
        if (m >= 0) {
            unsigned long long key;
            if (lattice_key(a, m, p.x, p.y, p.z, &key)) is_new = !hset_contains(a.lattice, a.lattice_mask, key);
            else { atomicOr(&a.counts[MOR_CNT_ERRFLAGS], ERR_LATTICE_RANGE); is_new = true; }
        }
    }
    const unsigned same = __match_any_sync(kFull, is_new ? m : -1 - (int)(threadIdx.x & 31));
    if (is_new && (int)(__ffs(same) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&a.newcount[m], __popc(same));
    moving_test_epilogue(a);
}
__global__ void __launch_bounds__(kSingle) k_lattice_count(FramePtrs a) { k_lattice_count_body(a); }
__global__ void __launch_bounds__(kSingle, 2) k_lattice_count_batch(const FramePtrs* __restrict__ P) { k_lattice_count_body(P[blockIdx.z]); }


// ===================================================================================== K10' (method 1)
// CorrespondenceEstimation::determineCorrespondences (cpp:343-361): for every point of the transformed
// previous cluster the nearest point of the matched current cluster; only squared distances inside
// (pde_lb, pde_ub) count, so the search is bounded by sqrt(pde_ub) on the clustering grid.
__device__ __forceinline__ void k_pde_count_body(const FramePtrs& a) {
    pdl_prologue();
    const int ring = a.pde_ring;
    const int c = blockIdx.x * kSingle + threadIdx.x;
    if (c < a.p_counts[MOR_CNT_NC]) {
        const float4 t = a.tpts[c];
        const int kp = __float_as_int(t.w);
        const int m = kp >= 0 ? a.mid_of_prev[kp] : -1;
        if (m >= 0) {
            const int target = a.match_m[m];
            const GridDesc g = *a.dgrid;
            const int cx = (int)floor(((double)t.x - g.ox) * g.inv_h), cy = (int)floor(((double)t.y - g.oy) * g.inv_h), cz = (int)floor(((double)t.z - g.oz) * g.inv_h);
            const float h = (float)(1.0 / g.inv_h);
            float best = 3.402823466e+38f;
            // Shells of growing Chebyshev distance r around the query's cell. A point in shell r is at least (r-1)*h away,
            // so the search stops as soon as that bound exceeds the best distance (or pde_ub: farther neighbours never
            // count), and a neighbour at d2 <= pde_lb settles the answer (the nearest one is then <= pde_lb: not counted).
            bool settled = false;
            for (int r = 0; r <= ring && !settled; r++) {
                if (r > 1) {
                    const float lb = (float)(r - 1) * h * 0.99999f;
                    if (lb * lb >= fminf(best, a.pde_ub)) break;
                }
                for (int dz = -r; dz <= r && !settled; dz++) {
                    const int zz = cz + dz;
                    if (zz < 0 || zz >= g.nz) continue;
                    for (int dy = -r; dy <= r && !settled; dy++) {
                        const int yy = cy + dy;
                        if (yy < 0 || yy >= g.ny) continue;
                        const int base = (zz * g.ny + yy) * g.nx;
                        const bool full = (dz == -r || dz == r || dy == -r || dy == r);  // rows on the shell's faces: whole x-run
                        // otherwise only the two end cells x = cx -+ r belong to the shell
                        for (int part = 0; part < (full || r == 0 ? 1 : 2); part++) {
                            int xa, xb;
                            if (full || r == 0) { xa = cx - r; xb = cx + r; } else if (part == 0) { xa = xb = cx - r; } else { xa = xb = cx + r; }
                            xa = max(xa, 0); xb = min(xb, g.nx - 1);
                            if (xa > xb) continue;
                            const int b = a.cell_start[base + xa], e = a.cell_start[base + xb + 1];
                            for (int j = b; j < e; j++) {
                                if (a.scid[j] != target) continue;
                                const float4 q = a.spts[j];
                                best = fminf(best, sqdist3(t.x, t.y, t.z, q.x, q.y, q.z));
                            }
                            if (best <= a.pde_lb) { settled = true; break; }
                        }
                    }
                }
            }
            if (best > a.pde_lb && best < a.pde_ub) atomicAdd(&a.newcount[m], 1);
        }
    }
    moving_test_epilogue(a);
}
__global__ void __launch_bounds__(kSingle) k_pde_count(FramePtrs a) { k_pde_count_body(a); }
__global__ void __launch_bounds__(kSingle) k_pde_count_batch(const FramePtrs* __restrict__ P) { k_pde_count_body(P[blockIdx.z]); }


// ===================================================================================== K12
// Detection flags (cpp:580-606) and the N-frame consistency chain: checkMovingClusterChain
// (cpp:478-514), recurseFindClusterChain (cpp:415-453), pushCentroid (cpp:455-476). corrs_vec /
// res_vec are device-resident ring buffers; a correspondence map is stored as match_of_prev[].
__device__ void chain_block(const FramePtrs& a) {
    const int K = a.counts[MOR_CNT_K], Kp = a.p_counts[MOR_CNT_K], M = a.counts[MOR_CNT_M];
    const int D = a.ring_depth, kmax = a.kmax;
    TrackState* ts = a.track;
    __shared__ int s_flag_any;
    for (int m = threadIdx.x; m < M; m += kSingle) {
        const unsigned long long n1 = (unsigned long long)a.p_cl_size[a.match_q[m]], n2 = (unsigned long long)a.cl_size[a.match_m[m]];
        double score, thr;
        if (a.method == 1) {
            score = (double)__ldcg(&a.newcount[m]) / (double)((n1 + n2) / 2ull);  // cpp:361
            thr = (double)a.pde_thr;                                       // cpp:586
        } else {
            score = (double)__ldcg(&a.newcount[m]);                                        // cpp:330
            thr = (double)((n1 + n2) / (unsigned long long)(long long)a.opc_factor);      // cpp:590, unsigned division
        }
        a.match_score[m] = score;
        a.cl_flags[a.match_m[m]] = score > thr ? 1 : 0;
    }
    __syncthreads();
    // ---- checkMovingClusterChain: push buffers
    const int corr_slot = (ts->corr_head + ts->corr_count) % D;
    for (int i = threadIdx.x; i < Kp; i += kSingle) a.corr_ring[corr_slot * kmax + i] = a.match_of_prev[i];
    int res_count = ts->res_count;
    const int res_head = ts->res_head;
    if (res_count == 0) {
        const int slot = res_head % D;
        for (int i = threadIdx.x; i < Kp; i += kSingle) a.res_ring[slot * kmax + i] = a.p_cl_flags[i];
        if (threadIdx.x == 0) a.res_len[slot] = Kp;
        res_count = 1;
    }
    {
        const int slot = (res_head + res_count) % D;
        for (int j = threadIdx.x; j < K; j += kSingle) a.res_ring[slot * kmax + j] = a.cl_flags[j];
        if (threadIdx.x == 0) a.res_len[slot] = K;
        res_count += 1;
    }
    if (threadIdx.x == 0) a.corr_len[corr_slot] = Kp;
    const int corr_count = ts->corr_count + 1;
    const int corr_head = ts->corr_head;
    __syncthreads();
    float* const mo_centroid = a.mo_centroid + (size_t)a.mo_parity * a.momax * 3;
    int* const mo_conf = a.mo_conf + (size_t)a.mo_parity * a.momax;
    int n_mo = ts->n_mo[a.mo_parity];
    if (res_count >= a.moving_confidence) {
        const int r0 = res_head % D;
        const int len0 = a.res_len[r0];
        // every flagged cluster of the oldest frame is followed through all buffered maps
        for (int i = threadIdx.x; i < len0; i += kSingle) {
            int track = -1;
            if (a.res_ring[r0 * kmax + i]) {
                track = i;
                for (int col = 0; col < corr_count && track >= 0; col++) {
                    const int cs = (corr_head + col) % D;
                    const int j = track < a.corr_len[cs] ? a.corr_ring[cs * kmax + track] : -1;
                    const int rs = (res_head + col + 1) % D;
                    track = (j >= 0 && a.res_ring[rs * kmax + j]) ? j : -1;
                }
            }
            a.found[i] = track;
        }
        __syncthreads();
        // keep the surviving chain ends, in ascending i (in place: write index <= read index)
        int n_found = 0;
        for (int base = 0; base < len0; base += kSingle) {
            const int i = base + threadIdx.x;
            const int f = i < len0 ? a.found[i] : -1;
            int tot;
            const int r = single_block_rank(f >= 0, &tot);
            if (f >= 0) a.found[n_found + r] = f;
            n_found += tot;
            __syncthreads();
        }
        // pushCentroid in that order; the scan over mo_vec is parallel, the append is serial
        for (int i = 0; i < n_found; i++) {
            const int f = a.found[i];
            const float px = a.cl_centroid[f * 3], py = a.cl_centroid[f * 3 + 1], pz = a.cl_centroid[f * 3 + 2];
            if (threadIdx.x == 0) s_flag_any = 0;
            __syncthreads();
            bool close = false;
            for (int t = threadIdx.x; t < n_mo; t += kSingle) {
                const double dx = (double)__fsub_rn(px, mo_centroid[t * 3]), dy = (double)__fsub_rn(py, mo_centroid[t * 3 + 1]),
                             dz = (double)__fsub_rn(pz, mo_centroid[t * 3 + 2]);
                const double dist = sqrt(dx * dx + dy * dy + dz * dz);
                close |= dist < (double)a.catch_up;
            }
            if (close) s_flag_any = 1;
            __syncthreads();
            const bool any = s_flag_any != 0;
            if (!any) {
                if (n_mo < a.momax) {
                    if (threadIdx.x == 0) {
                        mo_centroid[n_mo * 3] = px; mo_centroid[n_mo * 3 + 1] = py; mo_centroid[n_mo * 3 + 2] = pz;
                        mo_conf[n_mo] = a.static_confidence + 1;  // MovingObjectCentroid ctor, .h:91
                    }
                    n_mo++;
                } else if (threadIdx.x == 0) {
                    atomicOr(&a.counts[MOR_CNT_ERRFLAGS], ERR_MOVING_CAP);
                }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        if (res_count >= a.moving_confidence) {  // pop_front both deques, cpp:511-512
            ts->corr_head = (corr_head + 1) % D; ts->corr_count = corr_count - 1;
            ts->res_head = (res_head + 1) % D; ts->res_count = res_count - 1;
        } else {
            ts->corr_count = corr_count; ts->res_count = res_count; ts->res_head = res_head % D;
        }
        ts->n_mo[a.mo_parity] = n_mo;
        a.counts[MOR_CNT_NMO] = n_mo;
    }
}

// ===================================================================================== K13 + K14
// filterCloud (cpp:613-696) in one kernel.
//  Tracking part (cpp:630-671): 1-NN of every confirmed mover among the current centroids, unconditional
//  selection of that cluster (cpp:644-648), confidence update and erase. It is tiny (|mo_vec| x K distance
//  evaluations), so EVERY block recomputes the selection into shared memory instead of waiting for a separate
//  single-block kernel; block 0 alone writes the updated mo_vec into the other half of a double buffer.
//  Output part: ExtractIndices(negative) of the moving points + append of the ground points (cpp:673-684),
//  written as pcl::PointXYZI wire records (cpp:690); stable single-pass compaction over [cloud | ground],
//  kOutItems consecutive points per thread.
#ifndef OUT_BLOCK
#define OUT_BLOCK 256
#endif
constexpr int kOutBlock = OUT_BLOCK;
constexpr int kOutTile = 1024;
constexpr int kOutItems = kOutTile / kOutBlock;
constexpr int kRemovedBits = 16384;  // = max kmax

__device__ __forceinline__ void k_filter_output_body(const FramePtrs& a) {
    pdl_prologue();
    const int mo_parity = a.mo_parity;
    __shared__ unsigned s_removed[kRemovedBits / 32];
    __shared__ int s_tile, s_total, s_keep_base;
    const int K = a.counts[MOR_CNT_K];
    const int nc = a.counts[MOR_CNT_NC], ng = a.counts[MOR_CNT_NG];
    TrackState* ts = a.track;
    const int n_mo = ts->n_mo[mo_parity];
    const float* mo_c_in = a.mo_centroid + (size_t)mo_parity * a.momax * 3;
    const int* mo_f_in = a.mo_conf + (size_t)mo_parity * a.momax;
    float* mo_c_out = a.mo_centroid + (size_t)(mo_parity ^ 1) * a.momax * 3;
    int* mo_f_out = a.mo_conf + (size_t)(mo_parity ^ 1) * a.momax;
    if (threadIdx.x == 0) { s_tile = atomicAdd(&a.scratch->ticket_out, 1); s_total = 0; s_keep_base = 0; }
    for (int t = threadIdx.x; t < (K + 31) / 32; t += kOutBlock) s_removed[t] = 0u;
    __syncthreads();
    const int tile = s_tile;
    const bool writer = tile == 0;  // the first block to start also owns the mo_vec update
    // ---- tracking (redundant in every block; K == 0: un-built kd-tree in the reference (UB) -> entries untouched)
    for (int base = 0; base < n_mo && K > 0; base += kOutBlock) {
        const int t = base + threadIdx.x;
        bool keep = false;
        float cx = 0, cy = 0, cz = 0; int conf = 0;
        if (t < n_mo) {
            cx = mo_c_in[t * 3]; cy = mo_c_in[t * 3 + 1]; cz = mo_c_in[t * 3 + 2];
            conf = mo_f_in[t];
            float d;
            const int k = nn_brute(a.cl_centroid, K, cx, cy, cz, &d);
            if (writer) a.marker_cluster[t] = k;
            atomicOr(&s_removed[k >> 5], 1u << (k & 31));
            atomicAdd(&s_total, a.cl_size[k]);
            if (!a.cl_flags[k] || d > a.leave_off) {  // cpp:650
                conf--;
                keep = conf != 0;
            } else {
                cx = a.cl_centroid[k * 3]; cy = a.cl_centroid[k * 3 + 1]; cz = a.cl_centroid[k * 3 + 2];
                if (conf < a.static_confidence + 1) conf++;
                keep = true;
            }
        }
        if (writer) {  // ordered erase (cpp:655-660): survivors keep their relative order
            int tot;
            const int r = block_exclusive_scan<int, kOutBlock>(keep ? 1 : 0, &tot);
            if (keep) {
                const int o = s_keep_base + r;
                mo_c_out[o * 3] = cx; mo_c_out[o * 3 + 1] = cy; mo_c_out[o * 3 + 2] = cz;
                mo_f_out[o] = conf;
            }
            __syncthreads();
            if (threadIdx.x == 0) s_keep_base += tot;
        }
        __syncthreads();
    }
    __syncthreads();
    const int overflow = s_total > nc ? 1 : 0;  // ExtractIndices: more indices than points => error, empty output (A18)
    if (writer) {
        if (K == 0) {  // entries untouched: copy through
            for (int t = threadIdx.x; t < n_mo; t += kOutBlock) {
                mo_c_out[t * 3] = mo_c_in[t * 3]; mo_c_out[t * 3 + 1] = mo_c_in[t * 3 + 1]; mo_c_out[t * 3 + 2] = mo_c_in[t * 3 + 2];
                mo_f_out[t] = mo_f_in[t];
            }
        }
        for (int k = threadIdx.x; k < K; k += kOutBlock) a.cluster_removed[k] = (s_removed[k >> 5] >> (k & 31)) & 1u;
        if (threadIdx.x == 0) {
            const int kept = K > 0 ? s_keep_base : n_mo;
            ts->n_mo[mo_parity ^ 1] = kept;  // the host flips the parity after this launch
            a.counts[MOR_CNT_NMO] = kept;
            ts->extract_overflow = overflow;
            ts->n_markers = K > 0 ? n_mo : 0;
            a.counts[MOR_CNT_EXTRACT_OVERFLOW] = overflow;
        }
    }
    // ---- output compaction
    const int total_items = nc + ng;
    const int last_tile = total_items ? (total_items - 1) / kOutTile : 0;
    if (tile <= last_tile) {
    const int t0 = tile * kOutTile + threadIdx.x * kOutItems;
    float4 p[kOutItems];
    bool keep[kOutItems];
    int nkeep = 0;
#pragma unroll
    for (int k = 0; k < kOutItems; k++) {
        const int t = t0 + k;
        keep[k] = false;
        p[k] = make_float4(0, 0, 0, 0);
        if (t < nc) {
            p[k] = a.pts[t];
            const int c = a.cid[t];
            const bool removed = overflow || (c >= 0 && ((s_removed[c >> 5] >> (c & 31)) & 1u));
            keep[k] = !removed;
            if (removed) a.removed_mask[a.cloud_src[t]] = 2;
        } else if (t < total_items) {
            p[k] = a.gpts[t - nc];
            keep[k] = true;
        }
        nkeep += keep[k] ? 1 : 0;
    }
    int tot;
    const int in_block = block_exclusive_scan<int, kOutBlock>(nkeep, &tot);
    const int before = (int)tile_exclusive_prefix(a.st_out, tile, (unsigned long long)tot);
    int o = before + in_block;
#pragma unroll
    for (int k = 0; k < kOutItems; k++) {
        if (keep[k]) {
            a.out[2 * o] = make_float4(p[k].x, p[k].y, p[k].z, 1.0f);
            a.out[2 * o + 1] = make_float4(p[k].w, 0.f, 0.f, 0.f);
            o++;
        }
    }
    if (tile == last_tile && threadIdx.x == 0) a.counts[MOR_CNT_NOUT] = before + tot;
    }
    // the last block to finish resets the look-back state, so filterCloud may be called again at any time
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&a.scratch->out_blocks_done, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (s_last) {
        for (int t = threadIdx.x; t < a.tiles_pts; t += kOutBlock) a.st_out[t] = 0ull;
        if (threadIdx.x == 0) { a.scratch->ticket_out = 0; a.scratch->out_blocks_done = 0; }
    }
}
__global__ void __launch_bounds__(kOutBlock) k_filter_output(FramePtrs a) { k_filter_output_body(a); }
__global__ void __launch_bounds__(kOutBlock, 2048 / kOutBlock) k_filter_output_batch(const FramePtrs* __restrict__ P) { k_filter_output_body(P[blockIdx.z]); }


}  // namespace mor
