// =====================================================================================
// mor_oracle.cpp — CPU ORACLE for the MOR per-frame filtering hot path.
//
// TEST INFRASTRUCTURE ONLY. Nothing in the product path (dynamicslamtool_b200/) includes,
// links, loads or calls this file. Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it, and there only as the checker / the
// CPU baseline.
//
// PARITY UNPINNED: the reference (prabinrath/dynamicslamtool) ships no golden vectors, no
// known-answer tests and no fixtures, and it cannot be compiled offline (needs ROS, PCL 1.8,
// FLANN, Eigen, Boost - none present, no network). This file is a dependency-free C++17
// restatement of src/MovingObjectRemoval.cpp that encodes the third-party semantics listed
// as A1..A18 in SURVEY.md §8c (PCL 1.8 / FLANN / tf source knowledge). Each function cites
// the reference file:line it follows. Cross-checks that ARE available offline
// (scipy cKDTree + connected_components, brute force) live in tests/test_oracle_*.py; the check
// against a real PCL installation is oracle/pcl_probe/ (one probe per assumption; needs PCL, so it
// has never been run from this repository).
//
// Build: g++ -O2 -ffp-contract=off (no -march=native: FMA contraction would break the float
// bit-parity with a default x86-64 PCL/FLANN build). Single-threaded like the reference.
// =====================================================================================
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <limits>
#include <memory>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../include/mor_b200.h"

namespace {

// ------------------------------------------------------------------ basic types
struct PointXYZI {  // pcl::PointXYZI payload (A3): x,y,z,intensity
    float x, y, z, intensity;
};
struct PointXYZ {
    float x, y, z;
};
struct Correspondence {  // pcl::Correspondence
    int index_query, index_match;
    float distance;
};

// ------------------------------------------------------------------ config (cpp:698-864)
struct Config {
    mor_config c;
    bool seen[32];
};

enum KeyId {
    K_gp_limit, K_gp_leaf, K_bin_gap, K_min_cluster_size, K_max_cluster_size, K_volume_constraint,
    K_pde_lb, K_pde_ub, K_output_topic, K_debug_topic, K_marker_topic, K_input_pointcloud_topic,
    K_input_odometry_topic, K_output_fid, K_debug_fid, K_leave_off_distance, K_catch_up_distance,
    K_trim_x, K_trim_y, K_trim_z, K_ec_distance_threshold, K_opc_normalization_factor,
    K_pde_distance_threshold, K_method_choice, K_NREF,
    K_ground_mode = K_NREF, K_gp_planarity, K_gp_bin_width, K_NALL
};
const char* const kKeyNames[K_NALL] = {
    "gp_limit", "gp_leaf", "bin_gap", "min_cluster_size", "max_cluster_size", "volume_constraint",
    "pde_lb", "pde_ub", "output_topic", "debug_topic", "marker_topic", "input_pointcloud_topic",
    "input_odometry_topic", "output_fid", "debug_fid", "leave_off_distance", "catch_up_distance",
    "trim_x", "trim_y", "trim_z", "ec_distance_threshold", "opc_normalization_factor",
    "pde_distance_threshold", "method_choice", "ground_mode", "gp_planarity", "gp_bin_width"};

void copy_str(char* dst, const std::string& s) {
    std::snprintf(dst, 64, "%s", s.c_str());
}

// setVariables, src/MovingObjectRemoval.cpp:698-864. Same grammar: '#' or <3 chars => skipped;
// everything before the first ':' is the key, every later non-':' character is the value.
int parse_config_file(const char* path, int n_bad, int n_good, mor_config* out) {
    std::ifstream f(path);
    if (!f.is_open()) return MOR_ERR_CONFIG_OPEN;  // cpp:703-707 (exit(0) in the reference)
    mor_config c;
    std::memset(&c, 0, sizeof(c));
    bool seen[K_NALL] = {false};
    std::string line;
    while (std::getline(f, line)) {
        if (line.empty() || line[0] == '#' || line.length() < 3) continue;  // cpp:712
        std::string p1, p2;
        bool flag = true;
        for (char ch : line) {  // cpp:718-733
            if (ch == ':') { flag = false; continue; }
            if (flag) p1.push_back(ch); else p2.push_back(ch);
        }
        int key = -1;
        for (int k = 0; k < K_NALL; k++)
            if (p1 == kKeyNames[k]) key = k;
        if (key < 0) return MOR_ERR_CONFIG_KEY;  // cpp:856-860
        try {
            switch (key) {
                case K_gp_limit: c.gp_limit = std::stof(p2); break;
                case K_gp_leaf: c.gp_leaf = std::stof(p2); break;
                case K_bin_gap: c.bin_gap = std::stof(p2); break;
                case K_min_cluster_size: c.min_cluster_size = std::stol(p2); break;
                case K_max_cluster_size: c.max_cluster_size = std::stol(p2); break;
                case K_volume_constraint: c.volume_constraint = std::stof(p2); break;
                case K_pde_lb: c.pde_lb = std::stof(p2); break;
                case K_pde_ub: c.pde_ub = std::stof(p2); break;
                case K_output_topic: copy_str(c.output_topic, p2); break;
                case K_debug_topic: copy_str(c.debug_topic, p2); break;
                case K_marker_topic: copy_str(c.marker_topic, p2); break;
                case K_input_pointcloud_topic: copy_str(c.input_pointcloud_topic, p2); break;
                case K_input_odometry_topic: copy_str(c.input_odometry_topic, p2); break;
                case K_output_fid: copy_str(c.output_fid, p2); break;
                case K_debug_fid: copy_str(c.debug_fid, p2); break;
                case K_leave_off_distance: c.leave_off_distance = std::stof(p2); break;
                case K_catch_up_distance: c.catch_up_distance = std::stof(p2); break;
                case K_trim_x: c.trim_x = std::stof(p2); break;
                case K_trim_y: c.trim_y = std::stof(p2); break;
                case K_trim_z: c.trim_z = std::stof(p2); break;
                case K_ec_distance_threshold: c.ec_distance_threshold = std::stof(p2); break;
                // cpp:843: int member assigned from std::stof => truncation toward zero
                case K_opc_normalization_factor: c.opc_normalization_factor = (int)std::stof(p2); break;
                case K_pde_distance_threshold: c.pde_distance_threshold = std::stof(p2); break;
                case K_method_choice: c.method_choice = std::stoi(p2); break;
                case K_ground_mode: c.ground_mode = std::stoi(p2); break;
                case K_gp_planarity: c.gp_planarity = std::stof(p2); break;
                case K_gp_bin_width: c.gp_bin_width = std::stof(p2); break;
            }
        } catch (...) {
            return MOR_ERR_CONFIG_VALUE;
        }
        seen[key] = true;
    }
    // the reference leaves unseen members uninitialised; defined behaviour: error
    for (int k = 0; k < K_NREF; k++)
        if (!seen[k]) return MOR_ERR_CONFIG_MISSING;
    if (!seen[K_ground_mode]) c.ground_mode = MOR_GROUND_CROP;
    if (!seen[K_gp_planarity]) c.gp_planarity = 0.01f;
    if (!seen[K_gp_bin_width]) c.gp_bin_width = c.gp_leaf;
    if (c.method_choice != 1 && c.method_choice != 2) return MOR_ERR_CONFIG_VALUE;  // cpp:568-593 UB otherwise
    if (c.ground_mode < 0 || c.ground_mode > 2) return MOR_ERR_CONFIG_VALUE;
    if (c.opc_normalization_factor == 0 && c.method_choice == 2) return MOR_ERR_CONFIG_VALUE;  // cpp:590 div by 0
    c.n_bad = n_bad;
    c.n_good = n_good;
    *out = c;
    return MOR_OK;
}

// ------------------------------------------------------------------ FLANN L2_Simple<float> (A6)
inline float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
    float result = 0.0f;
    float d;
    d = ax - bx; result += d * d;
    d = ay - by; result += d * d;
    d = az - bz; result += d * d;
    return result;
}

// ------------------------------------------------------------------ exact kd-tree (stands in for
// pcl::search::KdTree -> FLANN KDTreeSingleIndex, leaf size 15). Pruning keeps a safety margin so
// the result set equals the brute-force set { j : L2_Simple(q, p_j) < r2 } exactly (A6, A8).
struct KdTree {
    struct Node {
        int lo, hi;        // range in idx
        int left, right;   // children or -1
        int dim;
        float split_lo, split_hi;  // max of left / min of right on dim
    };
    const float* px = nullptr; const float* py = nullptr; const float* pz = nullptr;
    std::vector<float> sx, sy, sz;  // leaf-ordered copies for locality
    std::vector<int> idx;
    std::vector<Node> nodes;
    int n = 0;

    void build(const std::vector<float>& x, const std::vector<float>& y, const std::vector<float>& z) {
        n = (int)x.size();
        idx.resize(n);
        for (int i = 0; i < n; i++) idx[i] = i;
        nodes.clear();
        nodes.reserve(n / 4 + 8);
        px = x.data(); py = y.data(); pz = z.data();
        if (n > 0) build_rec(0, n);
        sx.resize(n); sy.resize(n); sz.resize(n);
        for (int i = 0; i < n; i++) { sx[i] = x[idx[i]]; sy[i] = y[idx[i]]; sz[i] = z[idx[i]]; }
    }
    float coord(int i, int d) const { return d == 0 ? px[i] : (d == 1 ? py[i] : pz[i]); }
    int build_rec(int lo, int hi) {
        int id = (int)nodes.size();
        nodes.push_back(Node{lo, hi, -1, -1, 0, 0.f, 0.f});
        if (hi - lo <= 15) return id;
        float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int i = lo; i < hi; i++)
            for (int d = 0; d < 3; d++) {
                float v = coord(idx[i], d);
                mn[d] = std::min(mn[d], v); mx[d] = std::max(mx[d], v);
            }
        int dim = 0;
        for (int d = 1; d < 3; d++) if (mx[d] - mn[d] > mx[dim] - mn[dim]) dim = d;
        if (!(mx[dim] > mn[dim])) return id;  // all points identical: keep as a (big) leaf
        int mid = (lo + hi) / 2;
        std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi,
                         [&](int a, int b) { return coord(a, dim) < coord(b, dim); });
        float slo = -FLT_MAX, shi = FLT_MAX;
        for (int i = lo; i < mid; i++) slo = std::max(slo, coord(idx[i], dim));
        shi = coord(idx[mid], dim);
        for (int i = mid; i < hi; i++) shi = std::min(shi, coord(idx[i], dim));
        int l = build_rec(lo, mid);
        int r = build_rec(mid, hi);
        nodes[id].left = l; nodes[id].right = r; nodes[id].dim = dim;
        nodes[id].split_lo = slo; nodes[id].split_hi = shi;
        return id;
    }
    // all j with sqdist < r2 (strict). Appends to out (unsorted, like FLANN with sorted=false).
    void radius(float qx, float qy, float qz, float r2, std::vector<int>& out) const {
        out.clear();
        if (n == 0) return;
        const double margin = (double)r2 * (1.0 + 1e-5) + 1e-30;
        int stack[128]; int sp = 0; stack[sp++] = 0;
        while (sp) {
            const Node& nd = nodes[stack[--sp]];
            if (nd.left < 0) {
                for (int i = nd.lo; i < nd.hi; i++)
                    if (sqdist3(qx, qy, qz, sx[i], sy[i], sz[i]) < r2) out.push_back(idx[i]);
                continue;
            }
            float q = nd.dim == 0 ? qx : (nd.dim == 1 ? qy : qz);
            double dl = (double)q - (double)nd.split_lo;  // >0 => query right of everything in left
            double dr = (double)nd.split_hi - (double)q;  // >0 => query left of everything in right
            if (!(dl > 0 && dl * dl > margin)) stack[sp++] = nd.left;
            if (!(dr > 0 && dr * dr > margin)) stack[sp++] = nd.right;
        }
    }
    // exact 1-NN; ties resolved to the lowest original index (canonical rule, A16).
    int nearest(float qx, float qy, float qz, float* out_d) const {
        int best = -1; float bd = FLT_MAX;
        if (n == 0) { *out_d = bd; return -1; }
        int stack[128]; int sp = 0; stack[sp++] = 0;
        while (sp) {
            const Node& nd = nodes[stack[--sp]];
            if (nd.left < 0) {
                for (int i = nd.lo; i < nd.hi; i++) {
                    float d = sqdist3(qx, qy, qz, sx[i], sy[i], sz[i]);
                    if (d < bd || (d == bd && idx[i] < best)) { bd = d; best = idx[i]; }
                }
                continue;
            }
            float q = nd.dim == 0 ? qx : (nd.dim == 1 ? qy : qz);
            double dl = (double)q - (double)nd.split_lo;
            double dr = (double)nd.split_hi - (double)q;
            double margin = (double)bd * (1.0 + 1e-5) + 1e-30;
            bool go_l = !(dl > 0 && dl * dl > margin);
            bool go_r = !(dr > 0 && dr * dr > margin);
            // visit the nearer child first (push it last)
            if (dl <= 0) { if (go_r) stack[sp++] = nd.right; if (go_l) stack[sp++] = nd.left; }
            else { if (go_l) stack[sp++] = nd.left; if (go_r) stack[sp++] = nd.right; }
        }
        *out_d = bd;
        return best;
    }
};

// ------------------------------------------------------------------ per-frame container (T1, .h:7-56)
struct Cluster {
    std::vector<int> indices;        // into cloud, ascending (A5)
    std::vector<PointXYZI> points;   // deep copy (cpp:223-237); transformed in place when it is `ca` (cpp:550)
};

struct FrameCloud {  // MovingObjectDetectionCloud
    std::vector<PointXYZI> raw_cloud, cloud;
    std::vector<int> raw_src;    // raw_cloud index -> input index (bookkeeping for the mask taps)
    std::vector<int> cloud_src;  // cloud index -> raw_cloud index
    std::vector<int> gp_indices; // into raw_cloud, ascending
    std::vector<Cluster> clusters;
    std::vector<PointXYZ> centroid_collection;
    std::vector<uint8_t> detection_results;
    std::vector<int> labels;      // tap: min cloud index of each point's component
    std::vector<int> cluster_id;  // tap
    std::vector<float> ground_voxels;  // tap (modes 1/2)
    double ps[7] = {0, 0, 0, 0, 0, 0, 1};  // position xyz, orientation xyzw (tf::Pose source)
    uint32_t n_input = 0;
    std::vector<uint8_t> point_class;  // tap, over input indices
    bool init = false;
    int size_tie_groups = 0;
};

struct MovingObjectCentroid {  // .h:83-94
    PointXYZ centroid;
    int confidence, max_confidence;
    MovingObjectCentroid(PointXYZ c, int n_good) : centroid(c), confidence(n_good + 1), max_confidence(n_good + 1) {}
    bool decreaseConfidence() { confidence--; return confidence == 0; }
    void increaseConfidence() { if (confidence < max_confidence) confidence++; }
};

// ------------------------------------------------------------------ tf (A11): double precision
struct TfTransform {
    double m[3][3];
    double o[3];
};

// tf::Transform(Quaternion, Vector3) -> Matrix3x3::setRotation (no normalisation of q)
TfTransform tf_from_pose(const double p[7]) {
    TfTransform t;
    const double x = p[3], y = p[4], z = p[5], w = p[6];
    double d = x * x + y * y + z * z + w * w;
    double s = 2.0 / d;
    double xs = x * s, ys = y * s, zs = z * s;
    double wx = w * xs, wy = w * ys, wz = w * zs;
    double xx = x * xs, xy = x * ys, xz = x * zs;
    double yy = y * ys, yz = y * zs, zz = z * zs;
    t.m[0][0] = 1.0 - (yy + zz); t.m[0][1] = xy - wz;         t.m[0][2] = xz + wy;
    t.m[1][0] = xy + wz;         t.m[1][1] = 1.0 - (xx + zz); t.m[1][2] = yz - wx;
    t.m[2][0] = xz - wy;         t.m[2][1] = yz + wx;         t.m[2][2] = 1.0 - (xx + yy);
    t.o[0] = p[0]; t.o[1] = p[1]; t.o[2] = p[2];
    return t;
}

// tf::Transform::inverseTimes: (a^-1 * b): basis a.R^T * b.R, origin a.R^T * (b.o - a.o)
TfTransform tf_inverse_times(const TfTransform& a, const TfTransform& b) {
    TfTransform r;
    double v[3] = {b.o[0] - a.o[0], b.o[1] - a.o[1], b.o[2] - a.o[2]};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            r.m[i][j] = a.m[0][i] * b.m[0][j] + a.m[1][i] * b.m[1][j] + a.m[2][i] * b.m[2][j];
    for (int i = 0; i < 3; i++) r.o[i] = a.m[0][i] * v[0] + a.m[1][i] * v[1] + a.m[2][i] * v[2];
    return r;
}

// tf::Matrix3x3::getRotation (double) -> Eigen::Quaternionf -> toRotationMatrix (float) (A12)
void tf_to_affine3f(const TfTransform& t, float M[12]) {
    double q[4];
    double trace = t.m[0][0] + t.m[1][1] + t.m[2][2];
    if (trace > 0.0) {
        double s = std::sqrt(trace + 1.0);
        q[3] = s * 0.5;
        s = 0.5 / s;
        q[0] = (t.m[2][1] - t.m[1][2]) * s;
        q[1] = (t.m[0][2] - t.m[2][0]) * s;
        q[2] = (t.m[1][0] - t.m[0][1]) * s;
    } else {
        int i = t.m[0][0] < t.m[1][1] ? (t.m[1][1] < t.m[2][2] ? 2 : 1) : (t.m[0][0] < t.m[2][2] ? 2 : 0);
        int j = (i + 1) % 3, k = (i + 2) % 3;
        double s = std::sqrt(t.m[i][i] - t.m[j][j] - t.m[k][k] + 1.0);
        q[i] = s * 0.5;
        s = 0.5 / s;
        q[3] = (t.m[k][j] - t.m[j][k]) * s;
        q[j] = (t.m[j][i] + t.m[i][j]) * s;
        q[k] = (t.m[k][i] + t.m[i][k]) * s;
    }
    const float x = (float)q[0], y = (float)q[1], z = (float)q[2], w = (float)q[3];
    const float tx = 2.0f * x, ty = 2.0f * y, tz = 2.0f * z;
    const float twx = tx * w, twy = ty * w, twz = tz * w;
    const float txx = tx * x, txy = ty * x, txz = tz * x;
    const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
    M[0] = 1.0f - (tyy + tzz); M[1] = txy - twz;          M[2] = txz + twy;           M[3] = (float)t.o[0];
    M[4] = txy + twz;          M[5] = 1.0f - (txx + tzz); M[6] = tyz - twx;           M[7] = (float)t.o[1];
    M[8] = txz - twy;          M[9] = tyz + twx;          M[10] = 1.0f - (txx + tyy); M[11] = (float)t.o[2];
}

// pcl::transformPointCloud dense branch (PCL 1.8): left-to-right float, no FMA (A12)
inline void xform(const float M[12], float x, float y, float z, float& ox, float& oy, float& oz) {
    ox = M[0] * x + M[1] * y + M[2] * z + M[3];
    oy = M[4] * x + M[5] * y + M[6] * z + M[7];
    oz = M[8] * x + M[9] * y + M[10] * z + M[11];
}

inline bool finite3(float x, float y, float z) { return std::isfinite(x) && std::isfinite(y) && std::isfinite(z); }

// ------------------------------------------------------------------ closed-form symmetric 3x3 eigen
// (ground mode 2 only; no reference behaviour). Smallest-eigenvalue eigenvector of a PSD matrix,
// trigonometric method in double. Shared formula with the CUDA kernel (same operation order).
void smallest_eigvec_sym3(const double a[6] /*xx,xy,xz,yy,yz,zz*/, double& lmin, double n[3], double& tr) {
    const double xx = a[0], xy = a[1], xz = a[2], yy = a[3], yz = a[4], zz = a[5];
    tr = xx + yy + zz;
    const double p1 = xy * xy + xz * xz + yz * yz;
    const double q = tr / 3.0;
    const double p2 = (xx - q) * (xx - q) + (yy - q) * (yy - q) + (zz - q) * (zz - q) + 2.0 * p1;
    const double p = std::sqrt(p2 / 6.0);
    if (!(p > 1e-300)) { lmin = q; n[0] = 0; n[1] = 0; n[2] = 1; return; }
    const double b00 = (xx - q) / p, b11 = (yy - q) / p, b22 = (zz - q) / p;
    const double b01 = xy / p, b02 = xz / p, b12 = yz / p;
    double r = (b00 * (b11 * b22 - b12 * b12) - b01 * (b01 * b22 - b12 * b02) + b02 * (b01 * b12 - b11 * b02)) / 2.0;
    r = std::min(1.0, std::max(-1.0, r));
    const double phi = std::acos(r) / 3.0;
    // eigenvalues: q + 2p cos(phi + 2k pi/3); smallest is k=1
    lmin = q + 2.0 * p * std::cos(phi + 2.0943951023931954923);
    // eigenvector: cross products of rows of (A - lmin I); take the largest
    const double r0[3] = {xx - lmin, xy, xz}, r1[3] = {xy, yy - lmin, yz}, r2[3] = {xz, yz, zz - lmin};
    double c0[3] = {r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0]};
    double c1[3] = {r0[1] * r2[2] - r0[2] * r2[1], r0[2] * r2[0] - r0[0] * r2[2], r0[0] * r2[1] - r0[1] * r2[0]};
    double c2[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
    double d0 = c0[0] * c0[0] + c0[1] * c0[1] + c0[2] * c0[2];
    double d1 = c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2];
    double d2 = c2[0] * c2[0] + c2[1] * c2[1] + c2[2] * c2[2];
    const double* c = c0; double d = d0;
    if (d1 > d) { c = c1; d = d1; }
    if (d2 > d) { c = c2; d = d2; }
    if (!(d > 1e-300)) { n[0] = 0; n[1] = 0; n[2] = 1; return; }
    const double inv = 1.0 / std::sqrt(d);
    n[0] = c[0] * inv; n[1] = c[1] * inv; n[2] = c[2] * inv;
    if (n[2] < 0 || (n[2] == 0 && (n[1] < 0 || (n[1] == 0 && n[0] < 0)))) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
}

}  // namespace

// ------------------------------------------------------------------ the MOR facade (.h:96-168)
struct mor_handle {
    mor_config cfg;
    std::vector<MovingObjectCentroid> mo_vec;                         // .h:109
    std::deque<std::shared_ptr<std::vector<Correspondence>>> corrs_vec;  // .h:112
    std::deque<std::vector<uint8_t>> res_vec;                         // .h:115
    std::shared_ptr<FrameCloud> ca, cb;                               // .h:121
    int moving_confidence, static_confidence;                         // .h:127

    // taps of the last push/filter
    float M[12];
    bool two_frames = false;
    std::vector<float> prev_centroids_t, prev_points_t, prev_bbox_t, cluster_bbox;
    std::vector<Correspondence> recip, matches;
    std::vector<double> scores;
    std::vector<uint8_t> removed_mask, cluster_removed;
    int n_out = 0, extract_overflow = 0, P1 = 0, P2 = 0, n_kprev = 0, frames = 0;
    bool filtered = false;
    std::vector<PointXYZI> f_cloud;
    std::vector<mor_marker> markers;  // VISUALIZE: marker_pub.publish(mark_cluster(...)) calls of the last filterCloud
    std::string last_error;
    // F3 literal probe (oracle_set_literal_ground_probe): voxels whose accept test (cpp:145) comes out differently under
    // the reference's literal float arithmetic (A14/A15) than under the order-independent definition above
    bool literal_probe = false;
    long long literal_tested = 0, literal_differ = 0;

    // ---------------- mark_cluster, cpp:7-58: pcl::compute3DCentroid into a Vector4f (float sums in index order,
    // divided by n), getMinMax3D extents, zero extents widened to 0.1; `id` is filterCloud's running counter (cpp:622, :669)
    static mor_marker mark_cluster(const std::vector<PointXYZI>& pts, int k, int id) {
        mor_marker m;
        float sx = 0.f, sy = 0.f, sz = 0.f;
        float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (const PointXYZI& p : pts) {
            sx += p.x; sy += p.y; sz += p.z;
            const float v[3] = {p.x, p.y, p.z};
            for (int q = 0; q < 3; q++) { mn[q] = std::min(mn[q], v[q]); mx[q] = std::max(mx[q], v[q]); }
        }
        const float n = (float)pts.size();
        m.position[0] = sx / n; m.position[1] = sy / n; m.position[2] = sz / n;
        for (int q = 0; q < 3; q++) { const float e = mx[q] - mn[q]; m.scale[q] = e == 0.f ? 0.1f : e; }
        m.color[0] = 0.8f; m.color[1] = 0.1f; m.color[2] = 0.4f; m.color[3] = 0.5f;
        m.id = id; m.cluster = k;
        return m;
    }

    // ---------------- groundPlaneRemoval(x,y,z), cpp:62-88 (ACTIVE path, cpp:526)
    void trim_xy(FrameCloud& fc, const std::vector<PointXYZI>& in, float x, float y) {
        // PassThrough "x" then "y" (A1): non-finite xyz dropped, inclusive limits, order kept
        fc.raw_cloud.clear(); fc.raw_src.clear();
        const float nx = -x, ny = -y;
        for (size_t i = 0; i < in.size(); i++) {
            const PointXYZI& p = in[i];
            if (!finite3(p.x, p.y, p.z)) continue;
            if (p.x < nx || p.x > x) continue;  // cpp:66-70
            if (p.y < ny || p.y > y) continue;  // cpp:71-74
            fc.raw_cloud.push_back(p);
            fc.raw_src.push_back((int)i);
        }
    }
    void ground_crop(FrameCloud& fc, float x, float y, float z) {
        // CropBox(extract_removed=true), min(-x,-y,gp_limit) max(x,y,z) (A2), cpp:78-86
        const float mnx = -x, mny = -y, mnz = cfg.gp_limit;
        fc.cloud.clear(); fc.cloud_src.clear(); fc.gp_indices.clear();
        for (size_t i = 0; i < fc.raw_cloud.size(); i++) {
            const PointXYZI& p = fc.raw_cloud[i];
            bool outside = (p.x < mnx || p.y < mny || p.z < mnz) || (p.x > x || p.y > y || p.z > z);
            if (outside) fc.gp_indices.push_back((int)i);
            else { fc.cloud.push_back(p); fc.cloud_src.push_back((int)i); }
        }
    }

    // ---------------- groundPlaneRemoval(x,y), cpp:90-200 (DEAD + crashing in the reference: the call is commented
    // out at cpp:527 and gp_i is a null shared_ptr at cpp:182-188). Repaired semantics:
    //   * gp_i allocated; ground indices deduplicated and ascending; no accepted voxel => nothing removed;
    //   * mode bin tie => smallest key;
    //   * order-independent arithmetic so a parallel implementation can reproduce it: voxel centroids and ball
    //     statistics are accumulated in double (relative to the voxel centroid) instead of float in index order.
    // mode 1 = literal test (|S_xz|, |S_yz|, |S_zz| < 0.001 on the un-normalised scatter, Z bins of cpp:166);
    // mode 2 = eigen-normal generalisation (north_star item 2; no reference behaviour): planarity + normal test,
    //          bins along each voxel's own normal, every bin holding >= 25 % of the mode bin is ground.
    void ground_voxel(FrameCloud& fc, int mode) {
        const std::vector<PointXYZI>& raw = fc.raw_cloud;
        const int n = (int)raw.size();
        fc.cloud.clear(); fc.cloud_src.clear(); fc.gp_indices.clear(); fc.ground_voxels.clear();
        std::vector<uint8_t> is_ground(n, 0);
        const float leaf = cfg.gp_leaf;
        if (n > 0 && leaf > 0) {
            // --- VoxelGrid (A14), cpp:110-113: float index arithmetic, output in ascending voxel index
            const float inv_leaf = 1.0f / leaf;
            float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            for (const auto& p : raw) {
                mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
                mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
            }
            int64_t minb[3], div[3];
            for (int d = 0; d < 3; d++) {
                minb[d] = (int64_t)std::floor(mn[d] * inv_leaf);
                div[d] = (int64_t)std::floor(mx[d] * inv_leaf) - minb[d] + 1;
            }
            struct Vox { int64_t idx; double sx, sy, sz; int n; };
            std::vector<Vox> vox;
            std::vector<std::vector<int>> vox_members;  // (literal probe only) the points of every voxel, ascending index
            {
                std::vector<std::pair<int64_t, int>> keyed(n);
                for (int i = 0; i < n; i++) {
                    int64_t i0 = (int64_t)std::floor(raw[i].x * inv_leaf) - minb[0];
                    int64_t i1 = (int64_t)std::floor(raw[i].y * inv_leaf) - minb[1];
                    int64_t i2 = (int64_t)std::floor(raw[i].z * inv_leaf) - minb[2];
                    keyed[i] = {i0 + i1 * div[0] + i2 * div[0] * div[1], i};
                }
                std::stable_sort(keyed.begin(), keyed.end(), [](const std::pair<int64_t, int>& a, const std::pair<int64_t, int>& b) { return a.first < b.first; });
                for (int s = 0; s < n;) {
                    int e = s;
                    Vox v{keyed[s].first, 0, 0, 0, 0};
                    while (e < n && keyed[e].first == keyed[s].first) {
                        const PointXYZI& p = raw[keyed[e].second];
                        v.sx += p.x; v.sy += p.y; v.sz += p.z; v.n++; e++;
                    }
                    if (literal_probe) {
                        vox_members.emplace_back();
                        for (int t = s; t < e; t++) vox_members.back().push_back(keyed[t].second);
                    }
                    vox.push_back(v);
                    s = e;
                }
            }
            // --- ball query r = gp_leaf around every voxel centroid (cpp:115-125), strict < (A6, A7)
            std::vector<float> rx(n), ry(n), rz(n);
            for (int i = 0; i < n; i++) { rx[i] = raw[i].x; ry[i] = raw[i].y; rz[i] = raw[i].z; }
            KdTree tree; tree.build(rx, ry, rz);
            const float r2 = (float)((double)leaf * (double)leaf);
            std::vector<int> ind;
            std::vector<long long> acc_key;
            std::vector<std::vector<int>> index_bank;
            fc.ground_voxels.assign(vox.size() * 8, 0.f);
            for (size_t v = 0; v < vox.size(); v++) {
                const float qx = (float)(vox[v].sx / vox[v].n), qy = (float)(vox[v].sy / vox[v].n), qz = (float)(vox[v].sz / vox[v].n);
                float* gv = &fc.ground_voxels[v * 8];
                gv[0] = qx; gv[1] = qy; gv[2] = qz;
                tree.radius(qx, qy, qz, r2, ind);
                if (ind.size() <= 3) continue;  // cpp:131
                double m[3] = {0, 0, 0}, a[6] = {0, 0, 0, 0, 0, 0};  // moments of d = p - q
                for (int j : ind) {
                    const double dx = (double)raw[j].x - (double)qx, dy = (double)raw[j].y - (double)qy, dz = (double)raw[j].z - (double)qz;
                    m[0] += dx; m[1] += dy; m[2] += dz;
                    a[0] += dx * dx; a[1] += dx * dy; a[2] += dx * dz; a[3] += dy * dy; a[4] += dy * dz; a[5] += dz * dz;
                }
                const double nn = (double)ind.size();
                // un-normalised scatter about the mean (computeCovarianceMatrix, A15): S_ab = sum(d_a d_b) - sum(d_a) sum(d_b) / n
                const double S[6] = {a[0] - m[0] * m[0] / nn, a[1] - m[0] * m[1] / nn, a[2] - m[0] * m[2] / nn,
                                     a[3] - m[1] * m[1] / nn, a[4] - m[1] * m[2] / nn, a[5] - m[2] * m[2] / nn};
                bool ok; long long key;
                if (mode == MOR_GROUND_VOXEL_COV) {
                    ok = std::fabs(S[2]) < 0.001 && std::fabs(S[4]) < 0.001 && std::fabs(S[5]) < 0.001;  // cpp:145
                    if (literal_probe) {
                        // The same voxel the way the reference's own arithmetic would run: VoxelGrid centroid as
                        // float sums in index order, radiusSearch results sorted by distance, pcl::compute3DCentroid into a
                        // Vector4f (float sums / n) and pcl::computeCovarianceMatrix<float> (float products about that centroid,
                        // un-normalised), both over the neighbours in that order (cpp:137-144).
                        float fx = 0.f, fy = 0.f, fz = 0.f;
                        for (int j : vox_members[v]) { fx += raw[j].x; fy += raw[j].y; fz += raw[j].z; }
                        const float inv = (float)vox[v].n;
                        const float lqx = fx / inv, lqy = fy / inv, lqz = fz / inv;
                        std::vector<int> li;
                        tree.radius(lqx, lqy, lqz, r2, li);
                        bool lok = false;
                        if (li.size() > 3) {
                            std::vector<std::pair<float, int>> byd(li.size());
                            for (size_t t = 0; t < li.size(); t++) byd[t] = {sqdist3(lqx, lqy, lqz, raw[li[t]].x, raw[li[t]].y, raw[li[t]].z), li[t]};
                            std::sort(byd.begin(), byd.end());
                            float cx = 0.f, cy = 0.f, cz = 0.f;
                            for (const auto& e : byd) { cx += raw[e.second].x; cy += raw[e.second].y; cz += raw[e.second].z; }
                            const float fn = (float)byd.size();
                            cx /= fn; cy /= fn; cz /= fn;
                            float c02 = 0.f, c12 = 0.f, c22 = 0.f;
                            for (const auto& e : byd) {
                                const float dx = raw[e.second].x - cx, dy = raw[e.second].y - cy, dz = raw[e.second].z - cz;
                                c12 += dy * dz; c22 += dz * dz; c02 += dx * dz;
                            }
                            lok = std::fabs(c02) < 0.001 && std::fabs(c12) < 0.001 && std::fabs(c22) < 0.001;
                        }
                        literal_tested++;
                        if (lok != ok) literal_differ++;
                    }
                    key = (long long)(int)(qz * 10);  // cpp:166: (float)((int)(z*10))/bin_gap is monotone in this integer
                    gv[5] = 0; gv[6] = 0; gv[7] = 1;
                } else {
                    double lmin, nrm[3], tr;
                    smallest_eigvec_sym3(S, lmin, nrm, tr);
                    ok = tr > 0 && (lmin / tr) < (double)cfg.gp_planarity && nrm[2] > 0.7;
                    gv[5] = (float)nrm[0]; gv[6] = (float)nrm[1]; gv[7] = (float)nrm[2];
                    const double off = nrm[0] * (double)qx + nrm[1] * (double)qy + nrm[2] * (double)qz;  // plane offset along the voxel's own normal
                    key = (long long)std::floor(off / (double)cfg.gp_bin_width);
                }
                if (key < -32768) key = -32768;
                if (key > 32767) key = 32767;
                gv[4] = (float)key;
                if (ok) {
                    gv[3] = 1;
                    acc_key.push_back(key);
                    index_bank.push_back(ind);
                }
            }
            // --- bin histogram and mode (cpp:161-178)
            if (!acc_key.empty()) {
                std::vector<int> hist(65536, 0);
                for (long long k : acc_key) hist[k + 32768]++;
                // mode 1: the reference compares keys (float)kz/bin_gap; "smallest key" is the smallest kz for bin_gap > 0
                // and the largest kz for bin_gap < 0
                const bool ascending = !(mode == MOR_GROUND_VOXEL_COV && cfg.bin_gap < 0);
                int best = -1, best_cnt = 0;
                for (int t = 0; t < 65536; t++) {
                    const int k = ascending ? t : 65535 - t;
                    if (hist[k] > best_cnt) { best_cnt = hist[k]; best = k; }
                }
                const int thr = mode == MOR_GROUND_VOXEL_EIGEN ? std::max(1, (best_cnt + 3) / 4) : best_cnt;
                for (size_t a2 = 0; a2 < acc_key.size(); a2++) {
                    const int k = (int)(acc_key[a2] + 32768);
                    const bool take = mode == MOR_GROUND_VOXEL_EIGEN ? hist[k] >= thr : k == best;
                    if (take) for (int j : index_bank[a2]) is_ground[j] = 1;  // cpp:184-191
                }
            }
        }
        // ExtractIndices(negative) (cpp:194-198) + repaired gp_indices: deduplicated, ascending
        for (int i = 0; i < n; i++) {
            if (is_ground[i]) fc.gp_indices.push_back(i);
            else { fc.cloud.push_back(raw[i]); fc.cloud_src.push_back(i); }
        }
    }

    // ---------------- computeClusters, cpp:202-262
    void compute_clusters(FrameCloud& fc, float distance_threshold) {
        fc.clusters.clear(); fc.detection_results.clear(); fc.centroid_collection.clear();
        const int n = (int)fc.cloud.size();
        fc.labels.assign(n, -1); fc.cluster_id.assign(n, -1);
        std::vector<float> x(n), y(n), z(n);
        for (int i = 0; i < n; i++) { x[i] = fc.cloud[i].x; y[i] = fc.cloud[i].y; z[i] = fc.cloud[i].z; }
        KdTree tree; tree.build(x, y, z);
        // A7: radius*radius in double, cast to float
        const float r2 = (float)((double)distance_threshold * (double)distance_threshold);
        // pcl::extractEuclideanClusters (A5): BFS in ascending seed order
        std::vector<uint8_t> processed(n, 0);
        std::vector<int> nn;
        std::vector<std::vector<int>> found;
        for (int i = 0; i < n; i++) {
            if (processed[i]) continue;
            std::vector<int> seed_queue;
            size_t sq_idx = 0;
            seed_queue.push_back(i);
            processed[i] = 1;
            while (sq_idx < seed_queue.size()) {
                int s = seed_queue[sq_idx];
                tree.radius(x[s], y[s], z[s], r2, nn);
                for (int j : nn) {
                    if (processed[j]) continue;
                    seed_queue.push_back(j);
                    processed[j] = 1;
                }
                sq_idx++;
            }
            for (int j : seed_queue) fc.labels[j] = i;  // i is the min index of the component
            if ((int64_t)seed_queue.size() >= cfg.min_cluster_size && (int64_t)seed_queue.size() <= cfg.max_cluster_size) {
                std::sort(seed_queue.begin(), seed_queue.end());
                found.push_back(std::move(seed_queue));
            }
        }
        // A9: std::sort(rbegin, rend, size-less) => size descending; canonical tie rule: discovery
        // order (= min index ascending), which is what libstdc++ yields for <= 16 clusters.
        std::stable_sort(found.begin(), found.end(),
                         [](const std::vector<int>& a, const std::vector<int>& b) { return a.size() > b.size(); });
        fc.size_tie_groups = 0;
        for (size_t k = 1; k < found.size(); k++)
            if (found[k].size() == found[k - 1].size()) fc.size_tie_groups++;
        for (size_t k = 0; k < found.size(); k++) {
            Cluster c;
            c.indices = std::move(found[k]);
            c.points.reserve(c.indices.size());
            for (int j : c.indices) { c.points.push_back(fc.cloud[j]); fc.cluster_id[j] = (int)k; }  // cpp:224-230
            // compute3DCentroid<PointXYZI,double> (A10), cpp:239-243
            double sx = 0, sy = 0, sz = 0;
            for (const auto& p : c.points) { sx += p.x; sy += p.y; sz += p.z; }
            const double dn = (double)c.points.size();
            sx /= dn; sy /= dn; sz /= dn;
            fc.centroid_collection.push_back(PointXYZ{(float)sx, (float)sy, (float)sz});
            fc.clusters.push_back(std::move(c));
        }
        fc.detection_results.assign(fc.clusters.size(), 0);  // cpp:250-254
    }

    // ---------------- volumeConstraint, cpp:264-283 (A17)
    static void minmax3(const std::vector<PointXYZI>& pts, float mn[3], float mx[3]) {
        mn[0] = mn[1] = mn[2] = FLT_MAX; mx[0] = mx[1] = mx[2] = -FLT_MAX;
        for (const auto& p : pts) {
            mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
            mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
        }
    }
    static bool volume_constraint(const std::vector<PointXYZI>& fp, const std::vector<PointXYZI>& fcur, double threshold) {
        float mn[3], mx[3];
        minmax3(fp, mn, mx);
        double volp = (mx[0] - mn[0]) * (mx[1] - mn[1]) * (mx[2] - mn[2]);  // float product, widened
        minmax3(fcur, mn, mx);
        double volc = (mx[0] - mn[0]) * (mx[1] - mn[1]) * (mx[2] - mn[2]);
        return (std::fabs(volp - volc) / (volp + volc)) < threshold;  // NaN => false
    }

    // ---------------- calculateCorrespondenceCentroid, cpp:285-307 (A16)
    static int nn_centroid(const std::vector<PointXYZ>& pts, const PointXYZ& q, float* d) {
        int best = -1; float bd = FLT_MAX;
        for (size_t i = 0; i < pts.size(); i++) {
            float dd = sqdist3(q.x, q.y, q.z, pts[i].x, pts[i].y, pts[i].z);
            if (dd < bd) { bd = dd; best = (int)i; }  // ties => lowest index
        }
        *d = bd;
        return best;
    }
    void correspondence_centroid(FrameCloud& a, FrameCloud& b, std::vector<Correspondence>& ufmp, std::vector<Correspondence>& fmp) {
        ufmp.clear(); fmp.clear();
        if (a.centroid_collection.empty() || b.centroid_collection.empty()) return;
        for (size_t i = 0; i < a.centroid_collection.size(); i++) {
            float d, dr;
            int j = nn_centroid(b.centroid_collection, a.centroid_collection[i], &d);
            int ir = nn_centroid(a.centroid_collection, b.centroid_collection[j], &dr);
            if (ir != (int)i) continue;
            ufmp.push_back(Correspondence{(int)i, j, d});
        }
        for (const auto& c : ufmp)  // cpp:297-306
            if (volume_constraint(a.clusters[c.index_query].points, b.clusters[c.index_match].points, (double)cfg.volume_constraint))
                fmp.push_back(c);
    }

    // ---------------- getClusterPointcloudChangeVector, cpp:309-334 (method 2; A13 as refined in DESIGN.md)
    // OctreePointCloudChangeDetector(resolution): leaf lattice = floor((p - min)/res) with
    // min = first_point - res/2 - oversize, oversize = ((2*res - eps_f) - res)/2   (PCL 1.8
    // OctreePointCloud::adoptBoundingBoxToPoint + getKeyBitSize on the first point; later bounding-box
    // growth shifts min by multiples of res). Score = #points of c2 in leaves with no c1 point.
    static void octree_anchor(const PointXYZI& first, double res, double mn[3]) {
        const double minValue = (double)std::numeric_limits<float>::epsilon();
        const double f[3] = {(double)first.x, (double)first.y, (double)first.z};
        for (int d = 0; d < 3; d++) {
            double lo = f[d] - res / 2, hi = f[d] + res / 2;
            double side = 2.0 * res - minValue;
            double over = (side - (hi - lo)) / 2.0;
            mn[d] = lo - over;
        }
    }
    struct Key3 { int64_t a, b, c; bool operator==(const Key3& o) const { return a == o.a && b == o.b && c == o.c; } };
    struct Key3Hash { size_t operator()(const Key3& k) const { return (size_t)(k.a * 73856093LL ^ k.b * 19349663LL ^ k.c * 83492791LL); } };
    std::vector<double> cluster_change_vector(FrameCloud& a, FrameCloud& b, const std::vector<Correspondence>& mp, float resolution) {
        std::vector<double> changed;
        const double res = (double)resolution;
        for (const auto& m : mp) {
            const auto& c1 = a.clusters[m.index_query].points;
            const auto& c2 = b.clusters[m.index_match].points;
            double mn[3];
            octree_anchor(c1[0], res, mn);
            std::unordered_set<Key3, Key3Hash> occ;
            occ.reserve(c1.size() * 2);
            for (const auto& p : c1)
                occ.insert(Key3{(int64_t)std::floor(((double)p.x - mn[0]) / res), (int64_t)std::floor(((double)p.y - mn[1]) / res), (int64_t)std::floor(((double)p.z - mn[2]) / res)});
            size_t cnt = 0;
            for (const auto& p : c2) {
                Key3 k{(int64_t)std::floor(((double)p.x - mn[0]) / res), (int64_t)std::floor(((double)p.y - mn[1]) / res), (int64_t)std::floor(((double)p.z - mn[2]) / res)};
                if (!occ.count(k)) cnt++;
            }
            changed.push_back((double)cnt);
        }
        return changed;
    }

    // ---------------- getPointDistanceEstimateVector, cpp:336-366 (method 1)
    std::vector<double> point_distance_estimate_vector(FrameCloud& a, FrameCloud& b, const std::vector<Correspondence>& mp) {
        std::vector<double> estimates;
        for (const auto& m : mp) {
            const auto& c1 = a.clusters[m.index_query].points;
            const auto& c2 = b.clusters[m.index_match].points;
            std::vector<float> x(c2.size()), y(c2.size()), z(c2.size());
            for (size_t i = 0; i < c2.size(); i++) { x[i] = c2[i].x; y[i] = c2[i].y; z[i] = c2[i].z; }
            KdTree tree; tree.build(x, y, z);
            double count = 0;
            for (const auto& p : c1) {
                float d;
                tree.nearest(p.x, p.y, p.z, &d);
                if (d > cfg.pde_lb && d < cfg.pde_ub) count++;  // cpp:356 (squared distance vs bounds)
            }
            estimates.push_back(count / (double)((c1.size() + c2.size()) / 2));  // cpp:361, integer /2
        }
        return estimates;
    }

    // ---------------- recurseFindClusterChain, cpp:415-453
    int recurse_find_cluster_chain(int col, int track) {
        if (col == (int)corrs_vec.size()) return track;
        for (size_t j = 0; j < corrs_vec[col]->size(); j++) {
            if ((*corrs_vec[col])[j].index_query == track) {
                if (res_vec[col + 1][(*corrs_vec[col])[j].index_match]) return recurse_find_cluster_chain(col + 1, (*corrs_vec[col])[j].index_match);
                return -1;
            }
        }
        return -1;
    }
    // ---------------- pushCentroid, cpp:455-476
    void push_centroid(PointXYZ pt) {
        for (size_t i = 0; i < mo_vec.size(); i++) {
            // float differences, squared/summed/sqrt in double (std::pow(float,int) promotes)
            double dx = (double)(pt.x - mo_vec[i].centroid.x), dy = (double)(pt.y - mo_vec[i].centroid.y), dz = (double)(pt.z - mo_vec[i].centroid.z);
            double dist = std::sqrt(dx * dx + dy * dy + dz * dz);
            if (dist < (double)cfg.catch_up_distance) return;
        }
        mo_vec.push_back(MovingObjectCentroid(pt, static_confidence));
    }
    // ---------------- checkMovingClusterChain, cpp:478-514
    void check_moving_cluster_chain(std::shared_ptr<std::vector<Correspondence>> mp, const std::vector<uint8_t>& res_ca, const std::vector<uint8_t>& res_cb) {
        corrs_vec.push_back(mp);
        if (res_vec.size() == 0) res_vec.push_back(res_ca);
        res_vec.push_back(res_cb);
        if ((int)res_vec.size() >= moving_confidence) {
            for (size_t i = 0; i < res_vec[0].size(); i++) {
                if (res_vec[0][i]) {
                    int found = recurse_find_cluster_chain(0, (int)i);
                    if (found != -1) push_centroid(cb->centroid_collection[found]);
                }
            }
            corrs_vec.pop_front();
            res_vec.pop_front();
        }
    }

    // ---------------- pushRawCloudAndPose, cpp:516-611
    int push(const void* data, uint32_t n, uint32_t step, uint32_t ox, uint32_t oy, uint32_t oz, uint32_t oi, const double pose7[7]) {
        ca = cb;                                // cpp:520
        cb = std::make_shared<FrameCloud>();    // cpp:521
        filtered = false;
        // fromPCLPointCloud2 (A3), cpp:523
        std::vector<PointXYZI> in(n);
        const uint8_t* base = (const uint8_t*)data;
        for (uint32_t i = 0; i < n; i++) {
            const uint8_t* p = base + (size_t)i * step;
            std::memcpy(&in[i].x, p + ox, 4); std::memcpy(&in[i].y, p + oy, 4); std::memcpy(&in[i].z, p + oz, 4);
            if (oi != UINT32_MAX) std::memcpy(&in[i].intensity, p + oi, 4); else in[i].intensity = 0.f;
        }
        std::memcpy(cb->ps, pose7, sizeof(double) * 7);  // poseMsgToTF, cpp:524
        cb->n_input = n;
        trim_xy(*cb, in, cfg.trim_x, cfg.trim_y);
        if (cfg.ground_mode == MOR_GROUND_CROP) ground_crop(*cb, cfg.trim_x, cfg.trim_y, cfg.trim_z);  // cpp:526
        else ground_voxel(*cb, cfg.ground_mode);                                                          // cpp:527
        cb->point_class.assign(n, 0);
        for (size_t i = 0; i < cb->cloud_src.size(); i++) cb->point_class[cb->raw_src[cb->cloud_src[i]]] = 1;
        for (int g : cb->gp_indices) cb->point_class[cb->raw_src[g]] = 2;

        compute_clusters(*cb, cfg.ec_distance_threshold);  // cpp:529
        cb->init = true;                                    // cpp:532
        cluster_bbox.assign(cb->clusters.size() * 6, 0.f);
        for (size_t k = 0; k < cb->clusters.size(); k++) minmax3(cb->clusters[k].points, &cluster_bbox[k * 6], &cluster_bbox[k * 6 + 3]);

        two_frames = ca && ca->init && cb->init;  // cpp:534
        recip.clear(); matches.clear(); scores.clear(); prev_centroids_t.clear(); prev_points_t.clear(); prev_bbox_t.clear();
        P1 = P2 = n_kprev = 0;
        std::memset(M, 0, sizeof(M));
        if (two_frames) {
            TfTransform t = tf_inverse_times(tf_from_pose(cb->ps), tf_from_pose(ca->ps));  // cpp:536
            tf_to_affine3f(t, M);
            for (auto& c : ca->centroid_collection) xform(M, c.x, c.y, c.z, c.x, c.y, c.z);  // cpp:540-541
            prev_points_t.assign(ca->cloud.size() * 3, std::numeric_limits<float>::quiet_NaN());
            for (auto& cl : ca->clusters) {  // cpp:544-551
                for (size_t j = 0; j < cl.points.size(); j++) {
                    PointXYZI& p = cl.points[j];
                    xform(M, p.x, p.y, p.z, p.x, p.y, p.z);
                    int ci = cl.indices[j];
                    prev_points_t[ci * 3 + 0] = p.x; prev_points_t[ci * 3 + 1] = p.y; prev_points_t[ci * 3 + 2] = p.z;
                }
                n_kprev += (int)cl.points.size();
            }
            for (auto& c : ca->centroid_collection) { prev_centroids_t.push_back(c.x); prev_centroids_t.push_back(c.y); prev_centroids_t.push_back(c.z); }
            prev_bbox_t.assign(ca->clusters.size() * 6, 0.f);
            for (size_t k = 0; k < ca->clusters.size(); k++) minmax3(ca->clusters[k].points, &prev_bbox_t[k * 6], &prev_bbox_t[k * 6 + 3]);

            auto mp = std::make_shared<std::vector<Correspondence>>();
            correspondence_centroid(*ca, *cb, recip, *mp);  // cpp:564
            std::vector<double> param_vec;
            if (cfg.method_choice == 1) param_vec = point_distance_estimate_vector(*ca, *cb, *mp);       // cpp:571
            else param_vec = cluster_change_vector(*ca, *cb, *mp, 0.1f);                                   // cpp:575
            for (size_t j = 0; j < mp->size(); j++) {  // cpp:580-606
                size_t n1 = ca->clusters[(*mp)[j].index_query].points.size(), n2 = cb->clusters[(*mp)[j].index_match].points.size();
                P1 += (int)n1; P2 += (int)n2;
                double threshold;
                if (cfg.method_choice == 1) threshold = (double)cfg.pde_distance_threshold;
                else threshold = (double)((n1 + n2) / (size_t)cfg.opc_normalization_factor);  // cpp:590 unsigned integer division
                cb->detection_results[(*mp)[j].index_match] = param_vec[j] > threshold ? 1 : 0;
            }
            matches = *mp; scores = param_vec;
            check_moving_cluster_chain(mp, ca->detection_results, cb->detection_results);  // cpp:608
        }
        frames++;
        return MOR_OK;
    }

    // ---------------- filterCloud, cpp:613-696
    int filter() {
        if (!cb || !cb->init) return MOR_ERR_STATE;
        const size_t K = cb->clusters.size();
        std::vector<int> moving_points;
        cluster_removed.assign(K, 0);
        markers.clear();
        int id = 1;  // cpp:622; advanced once per looked-up entry (cpp:669), so the markers of a frame carry 1, 2, 3, ...
        for (int i = 0; i < (int)mo_vec.size(); i++) {  // cpp:630
            if (K == 0) continue;  // defined behaviour for the un-built tree: entries untouched
            float d;
            int k = nn_centroid(cb->centroid_collection, mo_vec[i].centroid, &d);  // cpp:636
            markers.push_back(mark_cluster(cb->clusters[k].points, k, id));        // cpp:640-642 (VISUALIZE)
            for (int j : cb->clusters[k].indices) moving_points.push_back(j);      // cpp:644-648
            cluster_removed[k] = 1;
            if (!cb->detection_results[k] || d > cfg.leave_off_distance) {         // cpp:650
                if (mo_vec[i].decreaseConfidence()) { mo_vec.erase(mo_vec.begin() + i); i--; }
            } else {
                mo_vec[i].centroid = cb->centroid_collection[k];                   // cpp:664
                mo_vec[i].increaseConfidence();
            }
            id++;  // cpp:669
        }
        // ExtractIndices(negative) (A18), cpp:673-678
        f_cloud.clear();
        removed_mask.assign(cb->n_input, 0);
        extract_overflow = moving_points.size() > cb->cloud.size() ? 1 : 0;
        std::vector<uint8_t> rm(cb->cloud.size(), 0);
        for (int j : moving_points) rm[j] = 1;
        if (!extract_overflow) {
            for (size_t i = 0; i < cb->cloud.size(); i++)
                if (!rm[i]) f_cloud.push_back(cb->cloud[i]);
        }
        for (size_t i = 0; i < cb->cloud.size(); i++) {
            int src = cb->raw_src[cb->cloud_src[i]];
            removed_mask[src] = (rm[i] || extract_overflow) ? 2 : 1;
        }
        for (int g : cb->gp_indices) {  // cpp:681-684
            f_cloud.push_back(cb->raw_cloud[g]);
            removed_mask[cb->raw_src[g]] = 1;
        }
        n_out = (int)f_cloud.size();
        filtered = true;
        return MOR_OK;
    }
};

// ====================================================================================== C ABI
extern "C" {

int oracle_parse_config(const char* path, mor_config* out) { return parse_config_file(path, 0, 0, out); }

int oracle_create_ex(const char* config_path, int n_bad, int n_good, int /*device*/, const mor_limits* /*limits*/, mor_handle** out) {
    if (!out || !config_path) return MOR_ERR_ARG;
    mor_config c;
    int st = parse_config_file(config_path, n_bad, n_good, &c);
    if (st != MOR_OK) return st;
    mor_handle* h = new mor_handle();
    h->cfg = c;
    h->moving_confidence = n_bad;  // cpp:368
    h->static_confidence = n_good;
    h->ca = std::make_shared<FrameCloud>();  // cpp:387-388 (init=false)
    h->cb = std::make_shared<FrameCloud>();
    *out = h;
    return MOR_OK;
}
int oracle_create(const char* config_path, int n_bad, int n_good, int device, mor_handle** out) {
    return oracle_create_ex(config_path, n_bad, n_good, device, nullptr, out);
}
int oracle_destroy(mor_handle* h) { delete h; return MOR_OK; }
// a reset handle is a newly constructed object with the same configuration (cpp:368-391)
int oracle_reset(mor_handle* h) {
    if (!h) return MOR_ERR_ARG;
    const mor_config c = h->cfg;
    const int nb = h->moving_confidence, ng = h->static_confidence;
    *h = mor_handle();
    h->cfg = c; h->moving_confidence = nb; h->static_confidence = ng;
    h->ca = std::make_shared<FrameCloud>();
    h->cb = std::make_shared<FrameCloud>();
    return MOR_OK;
}
// VISUALIZE outputs: cluster_collection (cpp:226-229, :553-558) and the markers of the last filterCloud
int oracle_get_cluster_collection(mor_handle* h, void* out, uint32_t cap_points, uint32_t* n_out) {
    if (!h || !n_out) return MOR_ERR_ARG;
    if (!h->cb || !h->cb->init) return MOR_ERR_STATE;
    size_t n = 0;
    for (const Cluster& c : h->cb->clusters) n += c.points.size();
    *n_out = (uint32_t)n;
    if (!out) return MOR_OK;
    if (n > cap_points) return MOR_ERR_CAPACITY;
    float* o = (float*)out;
    for (const Cluster& c : h->cb->clusters)
        for (const PointXYZI& p : c.points) { o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = 1.0f; o[4] = p.intensity; o[5] = o[6] = o[7] = 0.f; o += 8; }
    return MOR_OK;
}
int oracle_get_moving_markers(mor_handle* h, mor_marker* out, uint32_t cap, uint32_t* n_out) {
    if (!h || !n_out) return MOR_ERR_ARG;
    if (!h->filtered) return MOR_ERR_STATE;
    *n_out = (uint32_t)h->markers.size();
    if (!out || h->markers.empty()) return MOR_OK;
    if (h->markers.size() > cap) return MOR_ERR_CAPACITY;
    std::copy(h->markers.begin(), h->markers.end(), out);
    return MOR_OK;
}
int oracle_get_config(const mor_handle* h, mor_config* out) { if (!h || !out) return MOR_ERR_ARG; *out = h->cfg; return MOR_OK; }

int oracle_push_raw_cloud_and_pose(mor_handle* h, const void* data, uint32_t n, uint32_t point_step, uint32_t off_x, uint32_t off_y,
                                   uint32_t off_z, uint32_t off_i, const double pose7[7]) {
    if (!h || (!data && n) || !pose7) return MOR_ERR_ARG;
    return h->push(data, n, point_step, off_x, off_y, off_z, off_i, pose7);
}

int oracle_filter_cloud(mor_handle* h, void* out, uint32_t cap_points, uint32_t* n_out) {
    if (!h) return MOR_ERR_ARG;
    int st = h->filter();
    if (st != MOR_OK) return st;
    if (n_out) *n_out = (uint32_t)h->n_out;
    if ((uint32_t)h->n_out > cap_points) return MOR_ERR_CAPACITY;
    // toPCLPointCloud2 of PointXYZI: 32 B records x,y,z,1.0f,intensity,pad (cpp:690)
    uint8_t* o = (uint8_t*)out;
    for (int i = 0; i < h->n_out; i++) {
        float rec[8] = {h->f_cloud[i].x, h->f_cloud[i].y, h->f_cloud[i].z, 1.0f, h->f_cloud[i].intensity, 0.f, 0.f, 0.f};
        std::memcpy(o + (size_t)i * 32, rec, 32);
    }
    return MOR_OK;
}

int oracle_sync(mor_handle*) { return MOR_OK; }

int oracle_tap(mor_handle* h, int tap, void* dst, size_t cap_bytes, size_t* n_bytes) {
    if (!h) return MOR_ERR_ARG;
    std::vector<uint8_t> buf;
    auto put = [&](const void* p, size_t bytes) { buf.resize(bytes); if (bytes) std::memcpy(buf.data(), p, bytes); };
    FrameCloud& b = *h->cb;
    switch (tap) {
        case MOR_TAP_COUNTS: {
            int32_t c[MOR_NCOUNTS] = {0};
            c[MOR_CNT_N] = (int)b.n_input; c[MOR_CNT_NT] = (int)b.raw_cloud.size(); c[MOR_CNT_NC] = (int)b.cloud.size();
            c[MOR_CNT_NG] = (int)b.gp_indices.size(); c[MOR_CNT_K] = (int)b.clusters.size();
            c[MOR_CNT_KPREV] = h->two_frames ? (int)h->ca->clusters.size() : 0;
            c[MOR_CNT_M] = (int)h->matches.size(); c[MOR_CNT_NMO] = (int)h->mo_vec.size();
            c[MOR_CNT_NOUT] = h->filtered ? h->n_out : 0; c[MOR_CNT_NKPREV] = h->n_kprev; c[MOR_CNT_P1] = h->P1; c[MOR_CNT_P2] = h->P2;
            c[MOR_CNT_TWO_FRAMES] = h->two_frames; c[MOR_CNT_EXTRACT_OVERFLOW] = h->filtered ? h->extract_overflow : 0;
            c[MOR_CNT_MU] = (int)h->recip.size(); c[MOR_CNT_NCPREV] = h->two_frames ? (int)h->ca->cloud.size() : 0;
            int nk = 0; for (auto& cl : b.clusters) nk += (int)cl.indices.size();
            c[MOR_CNT_NK] = nk; c[MOR_CNT_NVOX] = (int)(b.ground_voxels.size() / 8); c[MOR_CNT_FRAME] = h->frames;
            c[20] = b.size_tie_groups;
            put(c, sizeof(c));
        } break;
        case MOR_TAP_POINT_CLASS: put(b.point_class.data(), b.point_class.size()); break;
        case MOR_TAP_LABELS: put(b.labels.data(), b.labels.size() * 4); break;
        case MOR_TAP_CLUSTER_ID: put(b.cluster_id.data(), b.cluster_id.size() * 4); break;
        case MOR_TAP_CLUSTER_ROOT: { std::vector<int32_t> v; for (auto& c : b.clusters) v.push_back(c.indices[0]); put(v.data(), v.size() * 4); } break;
        case MOR_TAP_CLUSTER_SIZE: { std::vector<int32_t> v; for (auto& c : b.clusters) v.push_back((int)c.indices.size()); put(v.data(), v.size() * 4); } break;
        case MOR_TAP_CENTROIDS: { std::vector<float> v; for (auto& c : b.centroid_collection) { v.push_back(c.x); v.push_back(c.y); v.push_back(c.z); } put(v.data(), v.size() * 4); } break;
        case MOR_TAP_TRANSFORM: put(h->M, sizeof(h->M)); break;
        case MOR_TAP_PREV_CENTROIDS_T: put(h->prev_centroids_t.data(), h->prev_centroids_t.size() * 4); break;
        case MOR_TAP_PREV_POINTS_T: put(h->prev_points_t.data(), h->prev_points_t.size() * 4); break;
        case MOR_TAP_MATCH_QUERY: { std::vector<int32_t> v; for (auto& m : h->matches) v.push_back(m.index_query); put(v.data(), v.size() * 4); } break;
        case MOR_TAP_MATCH_MATCH: { std::vector<int32_t> v; for (auto& m : h->matches) v.push_back(m.index_match); put(v.data(), v.size() * 4); } break;
        case MOR_TAP_MATCH_DIST: { std::vector<float> v; for (auto& m : h->matches) v.push_back(m.distance); put(v.data(), v.size() * 4); } break;
        case MOR_TAP_MATCH_SCORE: put(h->scores.data(), h->scores.size() * 8); break;
        case MOR_TAP_FLAGS: put(b.detection_results.data(), b.detection_results.size()); break;
        case MOR_TAP_MO_CENTROIDS: { std::vector<float> v; for (auto& m : h->mo_vec) { v.push_back(m.centroid.x); v.push_back(m.centroid.y); v.push_back(m.centroid.z); } put(v.data(), v.size() * 4); } break;
        case MOR_TAP_MO_CONF: { std::vector<int32_t> v; for (auto& m : h->mo_vec) v.push_back(m.confidence); put(v.data(), v.size() * 4); } break;
        case MOR_TAP_REMOVED_MASK: if (!h->filtered) return MOR_ERR_STATE; put(h->removed_mask.data(), h->removed_mask.size()); break;
        case MOR_TAP_CLUSTER_REMOVED: if (!h->filtered) return MOR_ERR_STATE; put(h->cluster_removed.data(), h->cluster_removed.size()); break;
        case MOR_TAP_RECIP_QUERY: { std::vector<int32_t> v; for (auto& m : h->recip) v.push_back(m.index_query); put(v.data(), v.size() * 4); } break;
        case MOR_TAP_RECIP_MATCH: { std::vector<int32_t> v; for (auto& m : h->recip) v.push_back(m.index_match); put(v.data(), v.size() * 4); } break;
        case MOR_TAP_GROUND_VOXELS: put(b.ground_voxels.data(), b.ground_voxels.size() * 4); break;
        case MOR_TAP_CLUSTER_BBOX: put(h->cluster_bbox.data(), h->cluster_bbox.size() * 4); break;
        case MOR_TAP_PREV_BBOX_T: put(h->prev_bbox_t.data(), h->prev_bbox_t.size() * 4); break;
        default: return MOR_ERR_ARG;
    }
    if (n_bytes) *n_bytes = buf.size();
    if (buf.size() > cap_bytes) return MOR_ERR_CAPACITY;
    if (!buf.empty()) std::memcpy(dst, buf.data(), buf.size());
    return MOR_OK;
}

// Building blocks exposed for the oracle's own cross-checks (tests/test_oracle_*.py)
// F3 literal probe (test infrastructure): count, over the frames pushed while it is on, the voxels whose accept flag
// differs between the literal float arithmetic of cpp:137-145 and the order-independent definition both sides implement.
int oracle_set_literal_ground_probe(mor_handle* h, int enabled) {
    if (!h) return MOR_ERR_ARG;
    h->literal_probe = enabled != 0; h->literal_tested = 0; h->literal_differ = 0;
    return MOR_OK;
}
int oracle_get_literal_ground_stats(mor_handle* h, long long* tested, long long* differ) {
    if (!h || !tested || !differ) return MOR_ERR_ARG;
    *tested = h->literal_tested; *differ = h->literal_differ;
    return MOR_OK;
}

// Point pairs of the current frame's `cloud` whose squared distance lies within `ulps` units in the last place of r^2
// (the pairs a differently rounded distance or a pruned tree search could classify the other way; north_star: "any
// tie-breaking divergence at the exact clustering radius counted and reported"). Brute force over all pairs.
int oracle_count_radius_ties(mor_handle* h, int ulps, uint64_t* pairs) {
    if (!h || !pairs || ulps < 0) return MOR_ERR_ARG;
    if (!h->cb) return MOR_ERR_STATE;
    const std::vector<PointXYZI>& c = h->cb->cloud;
    const float r2 = (float)((double)h->cfg.ec_distance_threshold * (double)h->cfg.ec_distance_threshold);
    float lo = r2, hi = r2;
    for (int u = 0; u < ulps; u++) { lo = std::nextafterf(lo, 0.f); hi = std::nextafterf(hi, FLT_MAX); }
    uint64_t cnt = 0;
    const size_t n = c.size();
    for (size_t i = 0; i < n; i++)
        for (size_t j = i + 1; j < n; j++) {
            const float d = sqdist3(c[i].x, c[i].y, c[i].z, c[j].x, c[j].y, c[j].z);
            if (d >= lo && d <= hi) cnt++;
        }
    *pairs = cnt;
    return MOR_OK;
}

int oracle_pose_delta(const double pose_prev7[7], const double pose_cur7[7], float m12[12]) {
    TfTransform t = tf_inverse_times(tf_from_pose(pose_cur7), tf_from_pose(pose_prev7));
    tf_to_affine3f(t, m12);
    return MOR_OK;
}
int oracle_sym3_smallest_eig(const double a6[6], double* lmin, double n3[3]) {
    double tr;
    smallest_eigvec_sym3(a6, *lmin, n3, tr);
    return MOR_OK;
}

}  // extern "C"
