// mor_b200.cu — handle, launch sequence and C ABI (include/mor_b200.h) of the B200-native MOR hot path.
//
// One handle = one CUDA device + one stream + device-resident SoA frame state that persists across
// frames (previous frame's clusters, mo_vec, the corrs_vec / res_vec ring buffers). A frame is
// pushRawCloudAndPose (reference cpp:516-611) = one H2D copy + ~14 kernel launches with no host
// synchronisation, then filterCloud (cpp:613-696) = 2 launches + the D2H copy of the output cloud.
// There is no CPU fallback: every entry point that computes needs the device.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mor_kernels.cuh"
#include "mor_ground.cuh"
#include "mor_debug.cuh"

using namespace mor;

#define MOR_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            h->last_error = std::string(#call) + ": " + cudaGetErrorString(e__);                    \
            return MOR_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

namespace {

// ---- host-side pose delta (double, tf semantics A11) and the Eigen float affine (A12) -------------
struct Tf { double m[3][3]; double o[3]; };

// tf::Transform(tf::Quaternion(x,y,z,w), tf::Vector3) via Matrix3x3::setRotation; q is NOT normalised
// (tf::poseMsgToTF, reference cpp:524).
Tf tf_from_pose(const double p[7]) {
    Tf t;
    const double qx = p[3], qy = p[4], qz = p[5], qw = p[6];
    const double s = 2.0 / (qx * qx + qy * qy + qz * qz + qw * qw);
    const double xs = qx * s, ys = qy * s, zs = qz * s;
    const double wx = qw * xs, wy = qw * ys, wz = qw * zs, xx = qx * xs, xy = qx * ys, xz = qx * zs, yy = qy * ys, yz = qy * zs, zz = qz * zs;
    t.m[0][0] = 1.0 - (yy + zz); t.m[0][1] = xy - wz; t.m[0][2] = xz + wy;
    t.m[1][0] = xy + wz; t.m[1][1] = 1.0 - (xx + zz); t.m[1][2] = yz - wx;
    t.m[2][0] = xz - wy; t.m[2][1] = yz + wx; t.m[2][2] = 1.0 - (xx + yy);
    t.o[0] = p[0]; t.o[1] = p[1]; t.o[2] = p[2];
    return t;
}

// cb.ps.inverseTimes(ca.ps) (cpp:536): basis cur^T * prev, origin cur^T * (prev.o - cur.o).
// Then tf getRotation (double) -> Eigen::Quaternionf -> Translation3f * q (pcl_ros::transformPointCloud).
void pose_delta_affine(const double prev7[7], const double cur7[7], float M[12]) {
    const Tf a = tf_from_pose(cur7), b = tf_from_pose(prev7);
    double R[3][3], o[3];
    const double v[3] = {b.o[0] - a.o[0], b.o[1] - a.o[1], b.o[2] - a.o[2]};
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) R[i][j] = a.m[0][i] * b.m[0][j] + a.m[1][i] * b.m[1][j] + a.m[2][i] * b.m[2][j];
        o[i] = a.m[0][i] * v[0] + a.m[1][i] * v[1] + a.m[2][i] * v[2];
    }
    double q[4];
    const double trace = R[0][0] + R[1][1] + R[2][2];
    if (trace > 0.0) {
        double s = std::sqrt(trace + 1.0);
        q[3] = s * 0.5;
        s = 0.5 / s;
        q[0] = (R[2][1] - R[1][2]) * s; q[1] = (R[0][2] - R[2][0]) * s; q[2] = (R[1][0] - R[0][1]) * s;
    } else {
        const int i = R[0][0] < R[1][1] ? (R[1][1] < R[2][2] ? 2 : 1) : (R[0][0] < R[2][2] ? 2 : 0);
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        double s = std::sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0);
        q[i] = s * 0.5;
        s = 0.5 / s;
        q[3] = (R[k][j] - R[j][k]) * s; q[j] = (R[j][i] + R[i][j]) * s; q[k] = (R[k][i] + R[i][k]) * s;
    }
    const float x = (float)q[0], y = (float)q[1], z = (float)q[2], w = (float)q[3];
    const float tx = 2.0f * x, ty = 2.0f * y, tz = 2.0f * z;
    const float twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    M[0] = 1.0f - (tyy + tzz); M[1] = txy - twz; M[2] = txz + twy; M[3] = (float)o[0];
    M[4] = txy + twz; M[5] = 1.0f - (txx + tzz); M[6] = tyz - twx; M[7] = (float)o[1];
    M[8] = txz - twy; M[9] = tyz + twx; M[10] = 1.0f - (txx + tyy); M[11] = (float)o[2];
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

struct mor_handle {
    mor_config cfg;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t side = nullptr;  // runs k_transform_prev (depends only on the previous frame + pose) beside the clustering chain
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    uint32_t spec_out = 0;        // speculative size of the output D2H copy (points), from the previous frame
    cudaStream_t last_stream = nullptr;  // stream that carries this handle's latest work (differs after a batched step)
    // batched stepping (this handle as the leader of a batch): device array of per-sequence arguments + pinned ring
    FramePtrs* d_batch = nullptr; FramePtrs* h_batch = nullptr; uint32_t batch_cap = 0; int batch_slot = 0;
    cudaEvent_t batch_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t nmax = 0, kmax = 0, momax = 0;
    int ring_depth = 0, pde_ring = 0;
    bool dynamic_grid = false; int max_cells = 0; double cell_h = 0;
    uint32_t static_cell_cap = 0;
    int num_sms = 148;
    size_t select_smem = 0;
    GridDesc grid;
    std::string last_error;

    // one big device arena + carved pointers
    uint8_t* arena = nullptr;
    size_t arena_bytes = 0;
    uint8_t* d_in = nullptr;
    size_t d_in_bytes = 0;
    uint8_t* zero_region = nullptr;  // Scratch + scan status + cell_count: one memset per frame
    size_t zero_bytes = 0;
    size_t lattice_cap = 0;
    FramePtrs base;  // pointers that do not change from frame to frame
    int* coll_cursor = nullptr; int* coll_turn = nullptr;  // mor_get_cluster_collection scratch
    GroundPtrs ground;  // voxel-covariance ground removal state (ground_mode 1/2 only)
    // ping-pong
    float4* pts[2]; float4* spts[2]; int* cid[2]; int* cl_root[2]; int* cl_size[2]; float* cl_centroid[2]; uint8_t* cl_flags[2]; float* cl_bbox[2]; int* counts[2];
    int cur = 0;
    bool have_cur = false, have_prev = false, filtered = false;
    int mo_parity = 0;
    double cur_pose[7], prev_pose[7];
    float M[12];
    bool two_frames = false;
    uint32_t n_input = 0, n_prev_input = 0;
    uint64_t launches = 0;
    FramePtrs frame;  // arguments of the current frame
    // timing
    bool timing = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int32_t* h_counts = nullptr;  // pinned
    // generic event slots (bench.py brackets its timed regions with these, on the handle's stream)
    cudaEvent_t slot_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // per-kernel profiling (off by default: two extra event records per launch)
    bool profiling = false;
    std::vector<cudaEvent_t> prof_pool;
    std::vector<int> prof_ids;  // kernel id of every recorded pair of the current frame
    double prof_ms[40] = {0};
    uint64_t prof_n[40] = {0};
};

namespace {

enum KernelId { KID_INGEST = 0, KID_KEYS, KID_SCAN_CELLS, KID_SCATTER, KID_NEIGHBORS, KID_FLATTEN, KID_STATS,
                KID_TRANSFORM_PREV, KID_LATTICE_INSERT, KID_LATTICE_COUNT, KID_PDE,
                KID_OUTPUT, KID_G_INGEST, KID_G_KEYS, KID_G_SCAN_VOX, KID_G_SCATTER, KID_G_EVAL, KID_G_MODE, KID_G_MARK, KID_G_PARTITION, KID__COUNT };
const char* const kKernelNames[KID__COUNT] = {"k_ingest", "k_keys", "k_scan_cells", "k_scatter", "k_link_cells", "k_flatten+select",
                                              "k_cluster_stats+match", "k_transform_prev", "k_lattice_insert", "k_lattice_count+chain", "k_pde_count+chain",
                                              "k_filter_output", "k_ingest_raw", "k_ground_keys", "k_scan_voxels", "k_ground_scatter",
                                              "k_voxel_eval", "k_ground_mode", "k_ground_mark", "k_ground_partition"};

inline void prof_begin(mor_handle* h, int id) {
    if (!h->profiling) return;
    const size_t i = h->prof_ids.size() * 2;
    while (h->prof_pool.size() < i + 2) { cudaEvent_t e; cudaEventCreate(&e); h->prof_pool.push_back(e); }
    cudaEventRecord(h->prof_pool[i], h->stream);
    h->prof_ids.push_back(id);
}
inline void prof_end(mor_handle* h) {
    if (!h->profiling) return;
    cudaEventRecord(h->prof_pool[h->prof_ids.size() * 2 - 1], h->stream);
}
void prof_harvest(mor_handle* h) {  // stream must be idle
    for (size_t i = 0; i < h->prof_ids.size(); i++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->prof_pool[2 * i], h->prof_pool[2 * i + 1]) == cudaSuccess) { h->prof_ms[h->prof_ids[i]] += ms; h->prof_n[h->prof_ids[i]]++; }
    }
    h->prof_ids.clear();
}
// Launch with the programmatic-stream-serialisation attribute (see pdl_prologue in mor_device.cuh).
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}
// Kernel launch of the frame chain: PDL normally, plain (and bracketed by events) under per-kernel profiling.
#define MOR_KLAUNCH(id, kernel, grid, block, smem, ...)                                           \
    do {                                                                                           \
        if (h->profiling) { prof_begin(h, id); kernel<<<grid, block, smem, st>>>(__VA_ARGS__); prof_end(h); } \
        else launch_pdl(kernel, dim3(grid), dim3(block), smem, st, __VA_ARGS__);                  \
        h->launches++;                                                                             \
    } while (0)

// Every device operation of the hot path goes through this: counted (mor_get_launch_count) and, when
// profiling is on, bracketed by events on the handle's stream.
#define MOR_LAUNCH(id, ...) do { prof_begin(h, id); __VA_ARGS__; prof_end(h); h->launches++; } while (0)

template <typename T>
T* carve(uint8_t*& p, size_t count) {
    T* r = reinterpret_cast<T*>(p);
    p += align_up(count * sizeof(T), 256);
    return r;
}

int build_grid(mor_handle* h) {
    const mor_config& c = h->cfg;
    // A7: r2 = (float)((double)tol * (double)tol); effective radius = sqrt(r2). Cell edge h = r/sqrt(3) shrunk by
    // 2^-10: the diagonal of a cell stays below r even after float rounding of the distance (every two points of
    // a cell are neighbours), and d < r implies a cell offset of at most 2 per axis (r/h = 1.734).
    const float r2 = (float)((double)c.ec_distance_threshold * (double)c.ec_distance_threshold);
    if (!(r2 > 0.f) || !(c.trim_x > 0.f) || !(c.trim_y > 0.f)) return MOR_ERR_CONFIG_VALUE;
    const double hcell = std::sqrt((double)r2) / std::sqrt(3.0) * (1.0 - 1.0 / 1024.0);
    GridDesc g;
    const double pad = (double)kGridPad * hcell;  // empty low-side cells: see k_link_cells
    g.ox = -(double)c.trim_x - pad; g.oy = -(double)c.trim_y - pad;
    double zlo, zhi;
    if (c.ground_mode == MOR_GROUND_CROP) { zlo = (double)c.gp_limit; zhi = (double)c.trim_z; }
    else { zlo = -64.0; zhi = 64.0; }  // voxel modes: z is not cropped; generous fixed slab
    if (!(zhi >= zlo)) zhi = zlo;
    g.oz = zlo - pad; g.inv_h = 1.0 / hcell;
    const double fx = std::floor(((double)c.trim_x - g.ox) / hcell) + 1, fy = std::floor(((double)c.trim_y - g.oy) / hcell) + 1, fz = std::floor((zhi - g.oz) / hcell) + 1;
    if (fx < 1 || fy < 1 || fz < 1) return MOR_ERR_CONFIG_VALUE;
    h->cell_h = hcell;
    // Dense cell table, three regimes by the number of cells the config crop box needs:
    //  * up to 2^22: the grid covers the crop box and is fixed at create time (scanning it costs ~10 us per frame);
    //  * up to static_cap (mor_limits.max_cells, default 2^27 = 2 x 512 MB of tables): tables for the whole box are
    //    allocated, but each frame's grid is laid over the bounding box of its cloud (k_keys), so the scan pays for
    //    the occupied extent only - a small radius over a large box, or the voxel modes' uncropped z slab;
    //  * beyond (e.g. trimming "disabled" with huge values): per-frame bounding-box grid in tables of 2^24 cells (or
    //    mor_limits.max_cells); a frame whose box needs more is rejected with MOR_ERR_CAPACITY.
    const double need = fx * fy * fz;
    const double static_cap = h->static_cell_cap ? (double)h->static_cell_cap : 134217728.0;
    if (need <= 4194304.0) {
        h->dynamic_grid = false;
        g.nx = (int)fx; g.ny = (int)fy; g.nz = (int)fz; g.ncells = g.nx * g.ny * g.nz;
        h->max_cells = 1 << 24;  // the voxel modes' ball-query grid shares the tables
    } else {
        h->dynamic_grid = true;
        if (need <= static_cap) h->max_cells = need > 16777216.0 ? (int)need : (1 << 24);
        else h->max_cells = h->static_cell_cap ? (int)h->static_cell_cap : (1 << 24);
        g.nx = g.ny = g.nz = kGridPad + 1; g.ncells = h->max_cells;
    }
    h->grid = g;
    h->pde_ring = (int)std::ceil(std::sqrt((double)c.pde_ub) / hcell);
    if (h->pde_ring < 1) h->pde_ring = 1;
    return MOR_OK;
}

int allocate(mor_handle* h) {
    const size_t N = h->nmax, K = h->kmax, MO = h->momax, D = (size_t)h->ring_depth;
    const size_t ncells = h->cfg.ground_mode != MOR_GROUND_CROP ? (size_t)h->max_cells : (size_t)h->grid.ncells;
    const size_t tiles_pts = N / kBlock + 2, tiles_cells = ncells / kScanTile + 2;
    size_t lat = 1;
    while (lat < 2 * N) lat <<= 1;
    h->lattice_cap = lat;
    h->d_in_bytes = N * 32;
    // ---- size pass (mirror of the carve pass below)
    auto plan = [&](uint8_t* p0) -> uint8_t* {
        uint8_t* p = p0;
        FramePtrs& b = h->base;
        h->d_in = carve<uint8_t>(p, h->d_in_bytes);
        h->zero_region = p;
        b.scratch = carve<Scratch>(p, 1);
        b.st_ingest = carve<unsigned long long>(p, tiles_pts);
        b.st_cells = carve<unsigned long long>(p, tiles_cells);
        b.st_out = carve<unsigned long long>(p, tiles_pts);
        b.cell_count = carve<int>(p, ncells + 16);
        h->zero_bytes = (size_t)(p - h->zero_region);
        b.cell_start = carve<int>(p, ncells + 16);
        b.dgrid = carve<GridDesc>(p, 1);
        b.point_class = carve<uint8_t>(p, N); b.removed_mask = carve<uint8_t>(p, N);
        b.cloud_src = carve<int>(p, N); b.gpts = carve<float4>(p, N); b.gsrc = carve<int>(p, N);
        b.cell_key = carve<int>(p, N); b.skey = carve<int>(p, N);
        b.parent = carve<int>(p, N); b.label = carve<int>(p, N); b.comp_size = carve<int>(p, N); b.root_list = carve<int>(p, N); b.cid_of_root = carve<int>(p, N);
        b.comp = carve<int>(p, N); b.scid = carve<int>(p, N); b.minidx = carve<int>(p, N); b.done = carve<unsigned long long>(p, N); b.cell_box = carve<uint4>(p, 2 * N);
        b.acc_sum = carve<unsigned long long>(p, K * 6); b.acc_box = carve<unsigned>(p, K * 6); b.pacc_box = carve<unsigned>(p, K * 6);
        b.tpts = carve<float4>(p, N); b.pct = carve<float>(p, K * 3); b.pbbox = carve<float>(p, K * 6);
        b.recip_q = carve<int>(p, K); b.recip_m = carve<int>(p, K); b.match_q = carve<int>(p, K); b.match_m = carve<int>(p, K);
        b.match_dist = carve<float>(p, K); b.match_score = carve<double>(p, K);
        b.match_of_prev = carve<int>(p, K); b.mid_of_prev = carve<int>(p, K); b.mid_of_cur = carve<int>(p, K);
        b.anchor = carve<double>(p, K * 3); b.newcount = carve<int>(p, K);
        b.lattice = carve<unsigned long long>(p, lat);
        b.cluster_removed = carve<uint8_t>(p, K); b.found = carve<int>(p, K);
        b.marker_cluster = carve<int>(p, MO); h->coll_cursor = carve<int>(p, K); h->coll_turn = carve<int>(p, 2);
        b.out = carve<float4>(p, N * 2);
        for (int f = 0; f < 2; f++) {
            h->pts[f] = carve<float4>(p, N); h->spts[f] = carve<float4>(p, N); h->cid[f] = carve<int>(p, N);
            h->cl_root[f] = carve<int>(p, K); h->cl_size[f] = carve<int>(p, K); h->cl_centroid[f] = carve<float>(p, K * 3);
            h->cl_flags[f] = carve<uint8_t>(p, K); h->cl_bbox[f] = carve<float>(p, K * 6); h->counts[f] = carve<int>(p, MOR_NCOUNTS);
        }
        if (h->cfg.ground_mode != MOR_GROUND_CROP) {
            GroundPtrs& g = h->ground;
            const size_t vc = (size_t)h->max_cells;
            g.rpts = carve<float4>(p, N); g.rsrc = carve<int>(p, N); g.is_ground = carve<uint8_t>(p, N); g.vkey = carve<int>(p, N);
            g.vox_count = carve<int>(p, vc + 1); g.vox_ord = carve<int>(p, vc + 1); g.tiles_vox = (int)(vc / kTile + 2);  // k_scan_voxels pulls kTile-voxel tiles
            g.st_vox = carve<unsigned long long>(p, g.tiles_vox);
            g.vox_n = carve<int>(p, N); g.vacc = carve<unsigned long long>(p, N * 6); g.vox_info = carve<float>(p, N * 8);
            g.bin_hist = carve<int>(p, 65536); g.ggrid = carve<GridDesc>(p, 1); g.vdesc = carve<VoxDesc>(p, 1); g.gstate = carve<int>(p, 8);
        }
        b.track = carve<TrackState>(p, 1); b.mo_centroid = carve<float>(p, 2 * MO * 3); b.mo_conf = carve<int>(p, 2 * MO);
        b.res_ring = carve<uint8_t>(p, D * K); b.res_len = carve<int>(p, D); b.corr_ring = carve<int>(p, D * K); b.corr_len = carve<int>(p, D);
        return p;
    };
    h->arena_bytes = (size_t)(plan(nullptr) - (uint8_t*)nullptr) + 256;
    MOR_CUDA(cudaMalloc(&h->arena, h->arena_bytes));
    plan(h->arena);
    MOR_CUDA(cudaMallocHost(&h->h_counts, sizeof(int32_t) * MOR_NCOUNTS));
    for (int i = 0; i < 4; i++) MOR_CUDA(cudaEventCreate(&h->ev[i]));
    int P = 1;
    while (P < (int)K) P <<= 1;
    h->select_smem = (size_t)P * sizeof(unsigned long long);
    MOR_CUDA(cudaFuncSetAttribute(k_flatten, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->select_smem));
    MOR_CUDA(cudaFuncSetAttribute(k_flatten_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->select_smem));
    MOR_CUDA(cudaStreamSynchronize(h->stream));
    return MOR_OK;
}

void fill_static(mor_handle* h) {
    FramePtrs& b = h->base;
    const mor_config& c = h->cfg;
    b.trim_x = c.trim_x; b.trim_y = c.trim_y; b.trim_z = c.trim_z; b.gp_limit = c.gp_limit;
    b.r2 = (float)((double)c.ec_distance_threshold * (double)c.ec_distance_threshold);
    b.volume_constraint = c.volume_constraint; b.pde_lb = c.pde_lb; b.pde_ub = c.pde_ub; b.pde_thr = c.pde_distance_threshold;
    b.leave_off = c.leave_off_distance; b.catch_up = c.catch_up_distance;
    b.min_cluster = c.min_cluster_size; b.max_cluster = c.max_cluster_size;
    b.method = c.method_choice; b.opc_factor = c.opc_normalization_factor;
    b.moving_confidence = c.n_bad; b.static_confidence = c.n_good;
    b.kmax = (int)h->kmax; b.momax = (int)h->momax; b.ring_depth = h->ring_depth;
    b.grid = h->grid;
    b.dynamic_grid = h->dynamic_grid ? 1 : 0; b.max_cells = h->max_cells; b.cell_h = h->cell_h;
    b.lattice_mask = (unsigned)(h->lattice_cap - 1);
    b.pde_ring = h->pde_ring;
    b.lattice_words16 = (unsigned)(h->lattice_cap / 2);
    b.tiles_pts = (int)(h->nmax / kBlock + 2); b.tiles_cells = (int)((h->cfg.ground_mode != MOR_GROUND_CROP ? (size_t)h->max_cells : (size_t)h->grid.ncells) / kScanTile + 2);
    if (c.ground_mode != MOR_GROUND_CROP) {
        GroundPtrs& g = h->ground;
        g.mode = c.ground_mode; g.leaf = c.gp_leaf; g.inv_leaf = 1.0f / c.gp_leaf; g.r2 = (float)((double)c.gp_leaf * (double)c.gp_leaf);
        g.bin_gap = c.bin_gap; g.planarity = c.gp_planarity; g.bin_width = c.gp_bin_width;
        g.ball_cell_h = (double)c.gp_leaf * (1.0 + 1.0 / 1024.0);
    }
}

inline unsigned blocks_for(uint32_t n) { return n ? (n + kBlock - 1) / kBlock : 1; }

// Arguments of the current frame (everything the kernels read) from the handle's host-side state.
void fill_frame(mor_handle* h, const uint8_t* d_points, uint32_t n, uint32_t step, uint32_t ox, uint32_t oy, uint32_t oz, uint32_t oi) {
    FramePtrs& a = h->frame;
    a = h->base;
    const int cur = h->cur, prev = cur ^ 1;
    a.in = d_points; a.n = n; a.step = step; a.off_x = ox; a.off_y = oy; a.off_z = oz; a.off_i = oi;
    a.vec16 = (step == 16 && ox == 0 && oy == 4 && oz == 8 && oi == 12 && ((uintptr_t)d_points % 16) == 0) ? 1 : 0;
    a.pts = h->pts[cur]; a.spts = h->spts[cur]; a.cid = h->cid[cur]; a.cl_root = h->cl_root[cur]; a.cl_size = h->cl_size[cur];
    a.cl_centroid = h->cl_centroid[cur]; a.cl_flags = h->cl_flags[cur]; a.cl_bbox = h->cl_bbox[cur]; a.counts = h->counts[cur];
    a.p_pts = h->pts[prev]; a.p_spts = h->spts[prev]; a.p_cid = h->cid[prev]; a.p_cl_root = h->cl_root[prev]; a.p_cl_size = h->cl_size[prev];
    a.p_cl_centroid = h->cl_centroid[prev]; a.p_cl_flags = h->cl_flags[prev]; a.p_counts = h->counts[prev];
    a.two_frames = h->two_frames ? 1 : 0;
    a.mo_parity = h->mo_parity;
    std::memcpy(a.M.m, h->M, sizeof(h->M));
}

int enqueue_push(mor_handle* h, const uint8_t* d_points, uint32_t n, uint32_t step, uint32_t ox, uint32_t oy, uint32_t oz, uint32_t oi) {
    cudaStream_t st = h->stream;
    fill_frame(h, d_points, n, step, ox, oy, oz, oi);
    FramePtrs& a = h->frame;
    const unsigned gb = blocks_for(n);
    if (h->cfg.ground_mode != MOR_GROUND_CROP) {
        // voxel-covariance ground removal (reference cpp:90-200, repaired): 8 launches, then the common pipeline
        const GroundPtrs& g = h->ground;
        FramePtrs ag = a;
        ag.dgrid = g.ggrid;  // the cell scan of this stage runs over the ball-query grid
        MOR_KLAUNCH(KID_G_INGEST, k_ingest_raw, gb, kBlock, 0, a, g);
        MOR_KLAUNCH(KID_G_KEYS, k_ground_keys, gb, kBlock, 0, a, g);
        MOR_KLAUNCH(KID_SCAN_CELLS, k_scan_cells, h->num_sms * 8, kBlock, 0, ag);
        MOR_KLAUNCH(KID_G_SCAN_VOX, k_scan_voxels, h->num_sms * 8, kBlock, 0, a, g);
        MOR_KLAUNCH(KID_G_SCATTER, k_ground_scatter, gb, kBlock, 0, a, g);
        MOR_KLAUNCH(KID_G_EVAL, k_voxel_eval, h->num_sms * 8, kBlock, 0, a, g);  // warps stride over the voxels
        MOR_KLAUNCH(KID_G_MODE, k_ground_mode, 1, kSingle, 0, a, g);
        MOR_KLAUNCH(KID_G_MARK, k_ground_mark, h->num_sms * 8, kBlock, 0, a, g);
        MOR_KLAUNCH(KID_G_PARTITION, k_ground_partition, gb, kBlock, 0, a, g);
    } else {
        MOR_KLAUNCH(KID_INGEST, k_ingest, n ? (n + kIngestTile - 1) / kIngestTile : 1, kIngestBlock, 0, a);
    }
    // The transform of the previous frame's clusters needs only the previous frame, the pose delta and the neutral
    // boxes written by the ingest kernel: it runs on a side stream beside the clustering chain and is joined before
    // k_match. (With per-kernel profiling on it stays in line so that its events bracket it alone.)
    const bool fork = h->two_frames && !h->profiling;
    if (fork) {
        MOR_CUDA(cudaEventRecord(h->ev_fork, st));
        MOR_CUDA(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
        k_transform_prev<<<(h->n_prev_input + kStatBlock - 1) / kStatBlock + (h->n_prev_input ? 0 : 1), kStatBlock, 0, h->side>>>(a);
        h->launches++;
        MOR_CUDA(cudaEventRecord(h->ev_join, h->side));
    }
    if (h->dynamic_grid) MOR_KLAUNCH(KID_KEYS, k_keys, gb, kBlock, 0, a);
    {
        const int tiles = (h->grid.ncells + kScanTile - 1) / kScanTile;
        const int scan_blocks = h->dynamic_grid ? h->num_sms * 8 : (tiles < h->num_sms * 8 ? tiles : h->num_sms * 8);
        MOR_KLAUNCH(KID_SCAN_CELLS, k_scan_cells, scan_blocks, kBlock, 0, a);
    }
    MOR_KLAUNCH(KID_SCATTER, k_scatter, gb, kBlock, 0, a);
    MOR_KLAUNCH(KID_NEIGHBORS, k_link_cells, dim3((n + kLinkBlock - 1) / kLinkBlock + (n ? 0 : 1), 18), kLinkBlock, 0, a);  // near pass (5 rows) + far pass (13 rows)
    const unsigned g1k = n ? (n + kSingle - 1) / kSingle : 1;
    MOR_KLAUNCH(KID_FLATTEN, k_flatten, g1k, kSingle, h->select_smem, a);  // + cluster selection in its last block
    if (fork) MOR_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0));  // k_cluster_stats' last block runs the correspondences
    else if (h->two_frames) MOR_LAUNCH(KID_TRANSFORM_PREV, (k_transform_prev<<<(h->n_prev_input + kStatBlock - 1) / kStatBlock + (h->n_prev_input ? 0 : 1), kStatBlock, 0, st>>>(a)));
    MOR_KLAUNCH(KID_STATS, k_cluster_stats, (n + kStatBlock - 1) / kStatBlock + (n ? 0 : 1), kStatBlock, 0, a);
    if (h->two_frames) {
        const unsigned gp = blocks_for(h->n_prev_input);
        if (h->cfg.method_choice == 2) {
            MOR_KLAUNCH(KID_LATTICE_INSERT, k_lattice_insert, gp, kBlock, 0, a);
            MOR_KLAUNCH(KID_LATTICE_COUNT, k_lattice_count, g1k, kSingle, 0, a);  // + flags and consistency chain in its last block
        } else {
            const unsigned gp1k = h->n_prev_input ? (h->n_prev_input + kSingle - 1) / kSingle : 1;
            MOR_KLAUNCH(KID_PDE, k_pde_count, gp1k, kSingle, 0, a);
        }
    }
    MOR_CUDA(cudaGetLastError());
    return MOR_OK;
}

// ca = cb; cb = new frame (cpp:520-521), pose delta cb.ps^-1 * ca.ps (cpp:536)
void advance_frame(mor_handle* h, uint32_t n, const double pose7[7]) {
    if (h->have_cur) { h->cur ^= 1; std::memcpy(h->prev_pose, h->cur_pose, sizeof(h->cur_pose)); h->have_prev = true; h->n_prev_input = h->n_input; }
    std::memcpy(h->cur_pose, pose7, sizeof(h->cur_pose));
    h->n_input = n;
    h->two_frames = h->have_prev;  // ca->init && cb->init (cpp:534)
    std::memset(h->M, 0, sizeof(h->M));
    if (h->two_frames) pose_delta_affine(h->prev_pose, h->cur_pose, h->M);
    h->have_cur = true;
    h->filtered = false;
}

// Work of a handle may have been enqueued on another handle's stream by mor_batch_step_device.
int join_foreign_stream(mor_handle* h) {
    if (h->last_stream && h->last_stream != h->stream) {
        MOR_CUDA(cudaStreamSynchronize(h->last_stream));
        h->last_stream = h->stream;
    }
    return MOR_OK;
}

int do_push(mor_handle* h, const void* data, bool on_device, uint32_t n, uint32_t step, uint32_t ox, uint32_t oy, uint32_t oz, uint32_t oi, const double pose7[7]) {
    if (!h || (!data && n) || !pose7) return MOR_ERR_ARG;
    if (step < 12 || step % 4 || ox % 4 || oy % 4 || oz % 4 || (oi != 0xFFFFFFFFu && oi % 4)) return MOR_ERR_ARG;
    if (ox + 4 > step || oy + 4 > step || oz + 4 > step || (oi != 0xFFFFFFFFu && oi + 4 > step)) return MOR_ERR_ARG;
    if (n > h->nmax) { h->last_error = "frame larger than mor_limits.max_points"; return MOR_ERR_CAPACITY; }
    MOR_CUDA(cudaSetDevice(h->device));
    { int js = join_foreign_stream(h); if (js != MOR_OK) return js; }
    if (h->timing) MOR_CUDA(cudaEventRecord(h->ev[0], h->stream));
    if (h->profiling && !h->prof_ids.empty()) { MOR_CUDA(cudaStreamSynchronize(h->stream)); prof_harvest(h); }
    const uint8_t* d_points = (const uint8_t*)data;
    if (!on_device) {
        const size_t bytes = (size_t)n * step;
        if (bytes > h->d_in_bytes) { h->last_error = "n * point_step exceeds the staging buffer (32 B/point)"; return MOR_ERR_CAPACITY; }
        if (bytes) MOR_CUDA(cudaMemcpyAsync(h->d_in, data, bytes, cudaMemcpyHostToDevice, h->stream));
        d_points = h->d_in;
    }
    advance_frame(h, n, pose7);
    int st = enqueue_push(h, d_points, n, step, ox, oy, oz, oi);
    if (st != MOR_OK) return st;
    if (h->timing) MOR_CUDA(cudaEventRecord(h->ev[1], h->stream));
    return MOR_OK;
}

int do_filter(mor_handle* h, void* out, bool on_device, uint32_t cap_points, uint32_t* n_out) {
    if (!h) return MOR_ERR_ARG;
    if (!h->have_cur) return MOR_ERR_STATE;
    MOR_CUDA(cudaSetDevice(h->device));
    { int js = join_foreign_stream(h); if (js != MOR_OK) return js; }
    cudaStream_t st = h->stream;
    if (h->timing) MOR_CUDA(cudaEventRecord(h->ev[2], st));
    FramePtrs& a = h->frame;
    if (on_device && out) a.out = (float4*)out;  // write the records straight into the caller's device buffer
    else a.out = h->base.out;
    if (on_device && out && cap_points < h->n_input) { h->last_error = "device output buffer must hold n_input points"; return MOR_ERR_CAPACITY; }
    a.mo_parity = h->mo_parity;
    MOR_KLAUNCH(KID_OUTPUT, k_filter_output, h->n_input ? (h->n_input + kOutTile - 1) / kOutTile : 1, kOutBlock, 0, a);
    h->mo_parity ^= 1;  // the kernel wrote the updated mo_vec into the other half
    a.mo_parity = h->mo_parity;
    MOR_CUDA(cudaGetLastError());
    h->filtered = true;
    if (h->timing) MOR_CUDA(cudaEventRecord(h->ev[3], st));
    if (on_device && !n_out && !h->profiling) return MOR_OK;  // fully asynchronous device-resident mode
    MOR_CUDA(cudaMemcpyAsync(h->h_counts, a.counts, sizeof(int32_t) * MOR_NCOUNTS, cudaMemcpyDeviceToHost, st));
    // The size of the output is only known on the device. Instead of a second round trip (sync on the count, then
    // copy), the cloud copy is issued speculatively with the previous frame's size plus a margin and topped up in
    // the rare case the frame turned out larger.
    uint32_t spec = 0;
    if (!on_device && out) {
        spec = h->spec_out ? h->spec_out + h->spec_out / 32 + 1024 : h->n_input;
        if (spec > h->n_input) spec = h->n_input;
        if (spec > cap_points) spec = cap_points;
        if (spec) MOR_CUDA(cudaMemcpyAsync(out, a.out, (size_t)spec * 32, cudaMemcpyDeviceToHost, st));
    }
    MOR_CUDA(cudaStreamSynchronize(st));
    if (h->profiling) prof_harvest(h);
    const uint32_t no = (uint32_t)h->h_counts[MOR_CNT_NOUT];
    if (n_out) *n_out = no;
    h->spec_out = no;
    if (h->h_counts[MOR_CNT_ERRFLAGS]) {  // a device-side capacity was exceeded: the frame's results are not reference-exact
        char msg[160];
        std::snprintf(msg, sizeof msg, "device capacity exceeded (error bits 0x%x: 1=clusters 2=moving 4=lattice 8=ground grid 16=grid cells)", h->h_counts[MOR_CNT_ERRFLAGS]);
        h->last_error = msg;
        return MOR_ERR_CAPACITY;
    }
    if (!on_device) {
        if (no > cap_points) return MOR_ERR_CAPACITY;
        if (no > spec) {
            MOR_CUDA(cudaMemcpyAsync((uint8_t*)out + (size_t)spec * 32, (const uint8_t*)a.out + (size_t)spec * 32, (size_t)(no - spec) * 32, cudaMemcpyDeviceToHost, st));
            MOR_CUDA(cudaStreamSynchronize(st));
        }
    }
    return MOR_OK;
}

// The state of a handle that has seen no frame: all device tables zero, the grid descriptor in place, no tracked
// objects, empty buffers (the reference's freshly constructed object, cpp:368-391). Ordered on the handle's stream.
int reset_state(mor_handle* h) {
    MOR_CUDA(cudaMemsetAsync(h->arena, 0, h->arena_bytes, h->stream));
    GridDesc g0 = h->grid;
    if (h->dynamic_grid) { g0.nx = g0.ny = g0.nz = kGridPad + 1; g0.ncells = g0.nx * g0.ny * g0.nz; }
    MOR_CUDA(cudaMemcpyAsync(h->base.dgrid, &g0, sizeof(g0), cudaMemcpyHostToDevice, h->stream));
    MOR_CUDA(cudaStreamSynchronize(h->stream));  // g0 is on this stack frame
    h->cur = 0; h->have_cur = h->have_prev = h->filtered = h->two_frames = false;
    h->mo_parity = 0; h->n_input = h->n_prev_input = 0; h->spec_out = 0;
    h->last_stream = h->stream;
    return MOR_OK;
}

}  // namespace

// ============================================================================================ C ABI
extern "C" {

int mor_create_ex(const char* config_path, int n_bad, int n_good, int device, const mor_limits* limits, mor_handle** out) {
    if (!out || !config_path) return MOR_ERR_ARG;
    *out = nullptr;
    mor_config cfg;
    int st = mor_parse_config(config_path, &cfg);
    if (st != MOR_OK) return st;
    cfg.n_bad = n_bad; cfg.n_good = n_good;
    if (cfg.ground_mode != MOR_GROUND_CROP && !(cfg.gp_leaf > 0.f)) return MOR_ERR_CONFIG_VALUE;
    mor_handle* h = new mor_handle();
    h->cfg = cfg; h->device = device;
    h->nmax = limits && limits->max_points ? limits->max_points : 300000u;
    h->kmax = limits && limits->max_clusters ? limits->max_clusters : 8192u;
    h->momax = limits && limits->max_moving ? limits->max_moving : 1024u;
    h->static_cell_cap = limits ? limits->max_cells : 0u;
    if (h->kmax > 16384u) h->kmax = 16384u;  // k_select_clusters sorts the clusters of a frame in shared memory (128 KB of keys)
    h->ring_depth = (n_bad > 1 ? n_bad : 1) + 2;
    st = build_grid(h);
    if (st != MOR_OK) { delete h; return st; }
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
    if (e != cudaSuccess) { cudaGetLastError(); mor_destroy(h); return MOR_ERR_CUDA; }  // mor_destroy releases whatever exists
    st = allocate(h);
    if (st != MOR_OK) { cudaGetLastError(); mor_destroy(h); return st; }
    fill_static(h);
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) h->num_sms = sms;
    }
    st = reset_state(h);
    if (st != MOR_OK) { mor_destroy(h); return st; }
    *out = h;
    return MOR_OK;
}

int mor_reset(mor_handle* h) {
    if (!h) return MOR_ERR_ARG;
    MOR_CUDA(cudaSetDevice(h->device));
    { int js = join_foreign_stream(h); if (js != MOR_OK) return js; }
    return reset_state(h);
}

int mor_create(const char* config_path, int n_bad, int n_good, int device, mor_handle** out) {
    return mor_create_ex(config_path, n_bad, n_good, device, nullptr, out);
}

int mor_destroy(mor_handle* h) {
    if (!h) return MOR_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->side) { cudaStreamSynchronize(h->side); cudaStreamDestroy(h->side); }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    for (auto& e : h->slot_ev) if (e) cudaEventDestroy(e);
    for (auto& e : h->prof_pool) cudaEventDestroy(e);
    if (h->h_counts) cudaFreeHost(h->h_counts);
    if (h->d_batch) cudaFree(h->d_batch);
    if (h->h_batch) cudaFreeHost(h->h_batch);
    for (auto& e : h->batch_ev) if (e) cudaEventDestroy(e);
    if (h->arena) cudaFree(h->arena);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return MOR_OK;
}

int mor_get_config(const mor_handle* h, mor_config* out) {
    if (!h || !out) return MOR_ERR_ARG;
    *out = h->cfg;
    return MOR_OK;
}

int mor_get_limits(const mor_handle* h, mor_limits* out) {
    if (!h || !out) return MOR_ERR_ARG;
    *out = mor_limits{};
    out->max_points = h->nmax; out->max_clusters = h->kmax; out->max_moving = h->momax; out->max_cells = (uint32_t)h->max_cells;
    return MOR_OK;
}

const char* mor_last_error(const mor_handle* h) { return h ? h->last_error.c_str() : ""; }

int mor_push_raw_cloud_and_pose(mor_handle* h, const void* data, uint32_t n, uint32_t point_step, uint32_t off_x, uint32_t off_y, uint32_t off_z,
                                uint32_t off_i, const double pose7[7]) {
    return do_push(h, data, false, n, point_step, off_x, off_y, off_z, off_i, pose7);
}
int mor_push_raw_cloud_and_pose_device(mor_handle* h, const void* d_data, uint32_t n, uint32_t point_step, uint32_t off_x, uint32_t off_y,
                                       uint32_t off_z, uint32_t off_i, const double pose7[7]) {
    return do_push(h, d_data, true, n, point_step, off_x, off_y, off_z, off_i, pose7);
}
int mor_filter_cloud(mor_handle* h, void* out, uint32_t cap_points, uint32_t* n_out) {
    if (!out && cap_points) return MOR_ERR_ARG;
    return do_filter(h, out, false, cap_points, n_out);
}
int mor_filter_cloud_device(mor_handle* h, void* d_out, uint32_t cap_points, uint32_t* n_out) { return do_filter(h, d_out, true, cap_points, n_out); }

// One pushRawCloudAndPose + filterCloud for S independent sequences in one set of launches (BASELINE config 5).
int mor_batch_step_device(mor_handle* const* hs, uint32_t S, const void* const* d_data, const uint32_t* n, uint32_t point_step, uint32_t off_x,
                          uint32_t off_y, uint32_t off_z, uint32_t off_i, const double* poses7, void* const* d_out) {
    if (!hs || !S || !d_data || !n || !poses7 || !d_out || !hs[0]) return MOR_ERR_ARG;
    mor_handle* h = hs[0];  // leader: owns the stream and the argument array of the batch
    if (point_step < 12 || point_step % 4 || off_x % 4 || off_y % 4 || off_z % 4 || (off_i != 0xFFFFFFFFu && off_i % 4)) return MOR_ERR_ARG;
    uint32_t n_max = 0, np_max = 0;
    for (uint32_t s = 0; s < S; s++) {
        mor_handle* g = hs[s];
        if (!g || g->device != h->device || g->nmax != h->nmax || g->kmax != h->kmax || g->cfg.method_choice != h->cfg.method_choice ||
            g->dynamic_grid != h->dynamic_grid || g->grid.ncells != h->grid.ncells || g->have_prev != h->have_prev || g->have_cur != h->have_cur ||
            g->profiling) { h->last_error = "batched handles must share device, limits, config and frame count"; return MOR_ERR_ARG; }
        if (g->cfg.ground_mode != MOR_GROUND_CROP) { h->last_error = "batched stepping supports ground_mode 0 only"; return MOR_ERR_ARG; }
        if (n[s] > g->nmax || (!d_data[s] && n[s]) || !d_out[s]) return n[s] > g->nmax ? MOR_ERR_CAPACITY : MOR_ERR_ARG;
        for (uint32_t t = 0; t < s; t++) if (hs[t] == g) return MOR_ERR_ARG;
    }
    MOR_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    if (S > h->batch_cap) {
        MOR_CUDA(cudaStreamSynchronize(st));
        if (h->d_batch) cudaFree(h->d_batch);
        if (h->h_batch) cudaFreeHost(h->h_batch);
        h->d_batch = nullptr; h->h_batch = nullptr; h->batch_cap = 0;
        MOR_CUDA(cudaMalloc(&h->d_batch, sizeof(FramePtrs) * S * 4));
        MOR_CUDA(cudaMallocHost(&h->h_batch, sizeof(FramePtrs) * S * 4));
        for (auto& e : h->batch_ev) if (!e) MOR_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->batch_cap = S;
    }
    // ring of 4 argument slots: the H2D copy of a slot runs later on the stream, so a slot is only rewritten once the
    // step that used it four steps ago has consumed it
    const int slot = h->batch_slot;
    h->batch_slot = (slot + 1) & 3;
    MOR_CUDA(cudaEventSynchronize(h->batch_ev[slot]));
    FramePtrs* hp = h->h_batch + (size_t)slot * h->batch_cap;
    FramePtrs* dp = h->d_batch + (size_t)slot * h->batch_cap;
    for (uint32_t s = 0; s < S; s++) {
        mor_handle* g = hs[s];
        if (g->last_stream != st) {  // earlier work of this handle ran elsewhere (its own stream or another batch): wait for it once
            if (g->last_stream) MOR_CUDA(cudaStreamSynchronize(g->last_stream));
            if (g->stream != st) MOR_CUDA(cudaStreamSynchronize(g->stream));
        }
        advance_frame(g, n[s], poses7 + 7 * s);
        fill_frame(g, (const uint8_t*)d_data[s], n[s], point_step, off_x, off_y, off_z, off_i);
        g->frame.out = (float4*)d_out[s];
        hp[s] = g->frame;
        n_max = n[s] > n_max ? n[s] : n_max;
        np_max = g->n_prev_input > np_max ? g->n_prev_input : np_max;
        g->last_stream = st;
    }
    MOR_CUDA(cudaMemcpyAsync(dp, hp, sizeof(FramePtrs) * S, cudaMemcpyHostToDevice, st));
    MOR_CUDA(cudaEventRecord(h->batch_ev[slot], st));
    const bool two = h->two_frames;
    const unsigned gb = blocks_for(n_max), g1k = n_max ? (n_max + kSingle - 1) / kSingle : 1;
    launch_pdl(k_ingest_batch, dim3(n_max ? (n_max + kIngestTile - 1) / kIngestTile : 1, 1, S), dim3(kIngestBlock), 0, st, (const FramePtrs*)dp);
    if (two) {  // the transform of the previous clusters runs beside the clustering chain (see enqueue_push)
        MOR_CUDA(cudaEventRecord(h->ev_fork, st));
        MOR_CUDA(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
        k_transform_prev_batch<<<dim3(np_max ? (np_max + kStatBlock - 1) / kStatBlock : 1, 1, S), kStatBlock, 0, h->side>>>(dp);
        MOR_CUDA(cudaEventRecord(h->ev_join, h->side));
    }
    if (h->dynamic_grid) launch_pdl(k_keys_batch, dim3(gb, 1, S), dim3(kBlock), 0, st, (const FramePtrs*)dp);
    {
        const int tiles = (h->grid.ncells + kScanTileBatch - 1) / kScanTileBatch;
        const int per_seq = h->num_sms * 8 / (int)S > 8 ? h->num_sms * 8 / (int)S : 8;
        const int scan_blocks = h->dynamic_grid ? per_seq : (tiles < per_seq ? tiles : per_seq);
        launch_pdl(k_scan_cells_batch, dim3(scan_blocks, 1, S), dim3(kBlock), 0, st, (const FramePtrs*)dp);
    }
    launch_pdl(k_scatter_batch, dim3(gb, 1, S), dim3(kBlock), 0, st, (const FramePtrs*)dp);
    launch_pdl(k_link_cells_batch, dim3((n_max + kLinkBlockBatch - 1) / kLinkBlockBatch + (n_max ? 0 : 1), 18, S), dim3(kLinkBlockBatch), 0, st, (const FramePtrs*)dp);
    launch_pdl(k_flatten_batch, dim3(g1k, 1, S), dim3(kSingle), h->select_smem, st, (const FramePtrs*)dp);
    if (two) MOR_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0));
    launch_pdl(k_cluster_stats_batch, dim3(g1k, 1, S), dim3(kStatBlock), 0, st, (const FramePtrs*)dp);
    h->launches += 6 + (h->dynamic_grid ? 1 : 0);
    if (two) {
        if (h->cfg.method_choice == 2) {
            launch_pdl(k_lattice_insert_batch, dim3(blocks_for(np_max), 1, S), dim3(kBlock), 0, st, (const FramePtrs*)dp);
            launch_pdl(k_lattice_count_batch, dim3(g1k, 1, S), dim3(kSingle), 0, st, (const FramePtrs*)dp);
            h->launches += 3;
        } else {
            launch_pdl(k_pde_count_batch, dim3(np_max ? (np_max + kSingle - 1) / kSingle : 1, 1, S), dim3(kSingle), 0, st, (const FramePtrs*)dp);
            h->launches += 2;
        }
    }
    launch_pdl(k_filter_output_batch, dim3(n_max ? (n_max + kOutTile - 1) / kOutTile : 1, 1, S), dim3(kOutBlock), 0, st, (const FramePtrs*)dp);
    h->launches += 1;
    MOR_CUDA(cudaGetLastError());
    for (uint32_t s = 0; s < S; s++) {
        hs[s]->mo_parity ^= 1;
        hs[s]->frame.mo_parity = hs[s]->mo_parity;
        hs[s]->filtered = true;
    }
    return MOR_OK;
}

int mor_sync(mor_handle* h) {
    if (!h) return MOR_ERR_ARG;
    MOR_CUDA(cudaSetDevice(h->device));
    { int js = join_foreign_stream(h); if (js != MOR_OK) return js; }
    MOR_CUDA(cudaStreamSynchronize(h->stream));
    return MOR_OK;
}

int mor_alloc_pinned(size_t bytes, void** out) { return cudaMallocHost(out, bytes) == cudaSuccess ? MOR_OK : MOR_ERR_CUDA; }
int mor_free_pinned(void* p) { return cudaFreeHost(p) == cudaSuccess ? MOR_OK : MOR_ERR_CUDA; }
int mor_device_alloc(int device, size_t bytes, void** out) {
    if (cudaSetDevice(device) != cudaSuccess) return MOR_ERR_CUDA;
    return cudaMalloc(out, bytes) == cudaSuccess ? MOR_OK : MOR_ERR_CUDA;
}
int mor_device_free(int device, void* p) {
    if (cudaSetDevice(device) != cudaSuccess) return MOR_ERR_CUDA;
    return cudaFree(p) == cudaSuccess ? MOR_OK : MOR_ERR_CUDA;
}
int mor_device_upload(int device, void* d_dst, const void* src, size_t bytes) {
    if (cudaSetDevice(device) != cudaSuccess) return MOR_ERR_CUDA;
    return cudaMemcpy(d_dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess ? MOR_OK : MOR_ERR_CUDA;
}
int mor_device_download(int device, void* dst, const void* d_src, size_t bytes) {
    if (cudaSetDevice(device) != cudaSuccess) return MOR_ERR_CUDA;
    return cudaMemcpy(dst, d_src, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? MOR_OK : MOR_ERR_CUDA;
}

int mor_get_launch_count(const mor_handle* h, uint64_t* out) {
    if (!h || !out) return MOR_ERR_ARG;
    *out = h->launches;
    return MOR_OK;
}
int mor_set_timing(mor_handle* h, int enabled) {
    if (!h) return MOR_ERR_ARG;
    h->timing = enabled != 0;
    return MOR_OK;
}
int mor_get_last_device_ms(mor_handle* h, float* push_ms, float* filter_ms) {
    if (!h || !h->timing) return MOR_ERR_STATE;
    MOR_CUDA(cudaStreamSynchronize(h->stream));
    if (push_ms) MOR_CUDA(cudaEventElapsedTime(push_ms, h->ev[0], h->ev[1]));
    if (filter_ms) MOR_CUDA(cudaEventElapsedTime(filter_ms, h->ev[2], h->ev[3]));
    return MOR_OK;
}

int mor_event_record(mor_handle* h, int slot) {
    if (!h || slot < 0 || slot >= 8) return MOR_ERR_ARG;
    MOR_CUDA(cudaSetDevice(h->device));
    if (!h->slot_ev[slot]) MOR_CUDA(cudaEventCreate(&h->slot_ev[slot]));
    MOR_CUDA(cudaEventRecord(h->slot_ev[slot], h->stream));
    return MOR_OK;
}
int mor_event_elapsed_ms(mor_handle* h, int slot_a, int slot_b, float* ms) {
    if (!h || !ms || slot_a < 0 || slot_a >= 8 || slot_b < 0 || slot_b >= 8 || !h->slot_ev[slot_a] || !h->slot_ev[slot_b]) return MOR_ERR_ARG;
    MOR_CUDA(cudaSetDevice(h->device));
    MOR_CUDA(cudaEventSynchronize(h->slot_ev[slot_b]));
    MOR_CUDA(cudaEventElapsedTime(ms, h->slot_ev[slot_a], h->slot_ev[slot_b]));
    return MOR_OK;
}
int mor_set_kernel_profiling(mor_handle* h, int enabled) {
    if (!h) return MOR_ERR_ARG;
    MOR_CUDA(cudaSetDevice(h->device));
    MOR_CUDA(cudaStreamSynchronize(h->stream));
    if (!h->prof_ids.empty()) prof_harvest(h);
    h->profiling = enabled != 0;
    if (enabled) { std::memset(h->prof_ms, 0, sizeof(h->prof_ms)); std::memset(h->prof_n, 0, sizeof(h->prof_n)); }
    return MOR_OK;
}
int mor_get_kernel_profile(mor_handle* h, int index, char name[32], double* total_ms, uint64_t* launches) {
    if (!h || index < 0 || index >= KID__COUNT || !name || !total_ms || !launches) return MOR_ERR_ARG;
    std::snprintf(name, 32, "%s", kKernelNames[index]);
    *total_ms = h->prof_ms[index];
    *launches = h->prof_n[index];
    return MOR_OK;
}

// ---- parity taps ----------------------------------------------------------------------------------
int mor_tap(mor_handle* h, int tap, void* dst, size_t cap_bytes, size_t* n_bytes) {
    if (!h) return MOR_ERR_ARG;
    if (!h->have_cur) return MOR_ERR_STATE;
    MOR_CUDA(cudaSetDevice(h->device));
    { int js = join_foreign_stream(h); if (js != MOR_OK) return js; }
    MOR_CUDA(cudaStreamSynchronize(h->stream));
    const FramePtrs& a = h->frame;
    int32_t c[MOR_NCOUNTS];
    MOR_CUDA(cudaMemcpy(c, a.counts, sizeof(c), cudaMemcpyDeviceToHost));
    const void* src = nullptr;
    size_t bytes = 0;
    std::vector<uint8_t> host;  // for taps assembled on the host
    const size_t N = c[MOR_CNT_N], NC = c[MOR_CNT_NC], K = c[MOR_CNT_K], KP = c[MOR_CNT_KPREV], M = c[MOR_CNT_M], MU = c[MOR_CNT_MU], NMO = c[MOR_CNT_NMO],
                 NCP = c[MOR_CNT_NCPREV];
    switch (tap) {
        case MOR_TAP_COUNTS: host.assign((uint8_t*)c, (uint8_t*)c + sizeof(c)); break;
        case MOR_TAP_POINT_CLASS: src = a.point_class; bytes = N; break;
        case MOR_TAP_LABELS: src = a.label; bytes = NC * 4; break;
        case MOR_TAP_CLUSTER_ID: src = a.cid; bytes = NC * 4; break;
        case MOR_TAP_CLUSTER_ROOT: src = a.cl_root; bytes = K * 4; break;
        case MOR_TAP_CLUSTER_SIZE: src = a.cl_size; bytes = K * 4; break;
        case MOR_TAP_CENTROIDS: src = a.cl_centroid; bytes = K * 12; break;
        case MOR_TAP_TRANSFORM: host.assign((uint8_t*)h->M, (uint8_t*)h->M + sizeof(h->M)); break;
        case MOR_TAP_PREV_CENTROIDS_T: src = a.pct; bytes = KP * 12; break;
        case MOR_TAP_PREV_POINTS_T: {
            std::vector<float4> t(NCP);
            if (NCP) MOR_CUDA(cudaMemcpy(t.data(), a.tpts, NCP * sizeof(float4), cudaMemcpyDeviceToHost));
            host.resize(NCP * 12);
            float* o = (float*)host.data();
            for (size_t i = 0; i < NCP; i++) {
                int k; std::memcpy(&k, &t[i].w, 4);
                const float nan = std::nanf("");
                o[i * 3] = k >= 0 ? t[i].x : nan; o[i * 3 + 1] = k >= 0 ? t[i].y : nan; o[i * 3 + 2] = k >= 0 ? t[i].z : nan;
            }
        } break;
        case MOR_TAP_MATCH_QUERY: src = a.match_q; bytes = M * 4; break;
        case MOR_TAP_MATCH_MATCH: src = a.match_m; bytes = M * 4; break;
        case MOR_TAP_MATCH_DIST: src = a.match_dist; bytes = M * 4; break;
        case MOR_TAP_MATCH_SCORE: src = a.match_score; bytes = M * 8; break;
        case MOR_TAP_FLAGS: src = a.cl_flags; bytes = K; break;
        case MOR_TAP_MO_CENTROIDS: src = a.mo_centroid + (size_t)h->mo_parity * h->momax * 3; bytes = NMO * 12; break;
        case MOR_TAP_MO_CONF: src = a.mo_conf + (size_t)h->mo_parity * h->momax; bytes = NMO * 4; break;
        case MOR_TAP_REMOVED_MASK: if (!h->filtered) return MOR_ERR_STATE; src = a.removed_mask; bytes = N; break;
        case MOR_TAP_CLUSTER_REMOVED: if (!h->filtered) return MOR_ERR_STATE; src = a.cluster_removed; bytes = K; break;
        case MOR_TAP_RECIP_QUERY: src = a.recip_q; bytes = MU * 4; break;
        case MOR_TAP_RECIP_MATCH: src = a.recip_m; bytes = MU * 4; break;
        case MOR_TAP_GROUND_VOXELS: if (h->cfg.ground_mode != MOR_GROUND_CROP) { src = h->ground.vox_info; bytes = (size_t)c[MOR_CNT_NVOX] * 32; } break;
        case MOR_TAP_CLUSTER_BBOX: src = a.cl_bbox; bytes = K * 24; break;
        case MOR_TAP_PREV_BBOX_T: src = a.pbbox; bytes = KP * 24; break;
        default: return MOR_ERR_ARG;
    }
    if (!host.empty() || tap == MOR_TAP_PREV_POINTS_T) bytes = host.size();
    if (n_bytes) *n_bytes = bytes;
    if (bytes > cap_bytes) return MOR_ERR_CAPACITY;
    if (!bytes) return MOR_OK;
    if (!host.empty()) std::memcpy(dst, host.data(), bytes);
    else MOR_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return MOR_OK;
}

#define MOR_NAMED_TAP(name, tapid, type)                                           \
    int name(mor_handle* h, type* dst, size_t n) { size_t nb = 0; return mor_tap(h, tapid, dst, n * sizeof(type), &nb); }
MOR_NAMED_TAP(mor_get_counts, MOR_TAP_COUNTS, int32_t)
MOR_NAMED_TAP(mor_get_trim_mask, MOR_TAP_POINT_CLASS, uint8_t)
MOR_NAMED_TAP(mor_get_ground_mask, MOR_TAP_POINT_CLASS, uint8_t)
MOR_NAMED_TAP(mor_get_labels, MOR_TAP_LABELS, int32_t)
MOR_NAMED_TAP(mor_get_cluster_order, MOR_TAP_CLUSTER_ROOT, int32_t)
MOR_NAMED_TAP(mor_get_centroids, MOR_TAP_CENTROIDS, float)
MOR_NAMED_TAP(mor_get_transform, MOR_TAP_TRANSFORM, float)
MOR_NAMED_TAP(mor_get_transformed_xyz, MOR_TAP_PREV_POINTS_T, float)
MOR_NAMED_TAP(mor_get_scores, MOR_TAP_MATCH_SCORE, double)
MOR_NAMED_TAP(mor_get_flags, MOR_TAP_FLAGS, uint8_t)
MOR_NAMED_TAP(mor_get_removed_mask, MOR_TAP_REMOVED_MASK, uint8_t)
int mor_get_matches(mor_handle* h, int32_t* query, int32_t* match, size_t n) {
    size_t nb = 0;
    int st = mor_tap(h, MOR_TAP_MATCH_QUERY, query, n * 4, &nb);
    return st != MOR_OK ? st : mor_tap(h, MOR_TAP_MATCH_MATCH, match, n * 4, &nb);
}
int mor_get_mo_vec(mor_handle* h, float* xyz, int32_t* conf, size_t n) {
    size_t nb = 0;
    int st = mor_tap(h, MOR_TAP_MO_CENTROIDS, xyz, n * 12, &nb);
    return st != MOR_OK ? st : mor_tap(h, MOR_TAP_MO_CONF, conf, n * 4, &nb);
}

// ---- VISUALIZE outputs (IncludeAll.h:32), on request
int mor_get_cluster_collection(mor_handle* h, void* out, uint32_t cap_points, uint32_t* n_out) {
    if (!h || !n_out) return MOR_ERR_ARG;
    if (!h->have_cur) return MOR_ERR_STATE;
    MOR_CUDA(cudaSetDevice(h->device));
    { int js = join_foreign_stream(h); if (js != MOR_OK) return js; }
    const FramePtrs& a = h->frame;
    CollectionPtrs c;
    c.cid = a.cid; c.pts = a.pts; c.cl_size = a.cl_size; c.counts = a.counts; c.cursor = h->coll_cursor; c.turn = h->coll_turn;
    c.out = h->base.out;  // free between calls: mor_filter_cloud has copied its result out before it returns
    k_collection_offsets<<<1, kCollBlock, 0, h->stream>>>(c);
    k_cluster_collection<<<h->n_input ? (h->n_input + kCollBlock - 1) / kCollBlock : 1, kCollBlock, 0, h->stream>>>(c);
    h->launches += 2;
    MOR_CUDA(cudaGetLastError());
    MOR_CUDA(cudaMemcpyAsync(h->h_counts, a.counts, sizeof(int32_t) * MOR_NCOUNTS, cudaMemcpyDeviceToHost, h->stream));
    MOR_CUDA(cudaStreamSynchronize(h->stream));
    const uint32_t nk = (uint32_t)h->h_counts[MOR_CNT_NK];
    *n_out = nk;
    if (!out) return MOR_OK;
    if (nk > cap_points) return MOR_ERR_CAPACITY;
    if (nk) MOR_CUDA(cudaMemcpy(out, c.out, (size_t)nk * 32, cudaMemcpyDeviceToHost));
    return MOR_OK;
}

int mor_get_moving_markers(mor_handle* h, mor_marker* out, uint32_t cap, uint32_t* n_out) {
    if (!h || !n_out) return MOR_ERR_ARG;
    if (!h->have_cur || !h->filtered) return MOR_ERR_STATE;
    MOR_CUDA(cudaSetDevice(h->device));
    { int js = join_foreign_stream(h); if (js != MOR_OK) return js; }
    MOR_CUDA(cudaStreamSynchronize(h->stream));
    const FramePtrs& a = h->frame;
    TrackState ts;
    MOR_CUDA(cudaMemcpy(&ts, a.track, sizeof ts, cudaMemcpyDeviceToHost));
    const uint32_t n = (uint32_t)ts.n_markers;
    *n_out = n;
    if (!out || !n) return MOR_OK;
    if (n > cap) return MOR_ERR_CAPACITY;
    MOR_CUDA(cudaMemcpy(h->h_counts, a.counts, sizeof(int32_t) * MOR_NCOUNTS, cudaMemcpyDeviceToHost));
    const size_t K = (size_t)h->h_counts[MOR_CNT_K];
    std::vector<int> which(n);
    std::vector<float> cen(K * 3), box(K * 6);
    MOR_CUDA(cudaMemcpy(which.data(), a.marker_cluster, n * sizeof(int), cudaMemcpyDeviceToHost));
    MOR_CUDA(cudaMemcpy(cen.data(), a.cl_centroid, K * 12, cudaMemcpyDeviceToHost));
    MOR_CUDA(cudaMemcpy(box.data(), a.cl_bbox, K * 24, cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < n; i++) {
        const int k = which[i];
        mor_marker& m = out[i];
        for (int q = 0; q < 3; q++) {
            m.position[q] = cen[(size_t)k * 3 + q];
            const float ext = box[(size_t)k * 6 + 3 + q] - box[(size_t)k * 6 + q];
            m.scale[q] = ext == 0.f ? 0.1f : ext;  // cpp:40-47
        }
        m.color[0] = 0.8f; m.color[1] = 0.1f; m.color[2] = 0.4f; m.color[3] = 0.5f;  // cpp:622, :53
        m.id = (int32_t)i + 1;  // cpp:622, :669: filterCloud's counter starts at 1 and advances once per looked-up entry
        m.cluster = k;
    }
    return MOR_OK;
}

}  // extern "C"
