// pcl_probe.cpp — checks the third-party semantics A1..A18 that oracle/mor_oracle.cpp and the CUDA
// kernels restate, against a real PCL installation. See README.md in this directory. NOT part of the repository's
// build: it needs PCL (1.8 is what the reference names), which the build environment of this repository lacks.
//
// Every probe: seeded input -> the PCL class the reference calls (file:line in the table of README.md) -> comparison
// with the prediction of the restated formula. One PASS/FAIL line each; exit code = number of failures.
#include <pcl/common/centroid.h>
#include <pcl/common/transforms.h>
#include <pcl/filters/crop_box.h>
#include <pcl/filters/extract_indices.h>
#include <pcl/filters/passthrough.h>
#include <pcl/filters/voxel_grid.h>
#include <pcl/octree/octree.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/registration/correspondence_estimation.h>
#include <pcl/search/kdtree.h>
#include <pcl/segmentation/extract_clusters.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <set>
#include <tuple>
#include <vector>

typedef pcl::PointXYZI P;
typedef pcl::PointCloud<P> Cloud;

static int g_fail = 0;
static void report(const char* name, bool ok, const char* detail = "") {
    std::printf("%-18s %s %s\n", name, ok ? "PASS" : "FAIL", detail);
    if (!ok) g_fail++;
}

// SplitMix64: the same tiny generator everywhere, so a failing probe can be reproduced from its seed
struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
    float uni(float lo, float hi) { return lo + (hi - lo) * (float)((next() >> 40) * (1.0 / 16777216.0)); }
};

static P pt(float x, float y, float z, float i = 0.f) { P p; p.x = x; p.y = y; p.z = z; p.intensity = i; return p; }
static Cloud::Ptr cloud_of(const std::vector<P>& v) {
    Cloud::Ptr c(new Cloud);
    c->points.assign(v.begin(), v.end());
    c->width = (uint32_t)v.size(); c->height = 1; c->is_dense = false;
    return c;
}
static float sqdist(const P& a, const P& b) {  // FLANN L2_Simple<float>: ((dx*dx)+(dy*dy))+(dz*dz)
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return ((dx * dx) + (dy * dy)) + (dz * dz);
}

// ------------------------------------------------------------------------------------------------ A1
static void probe_passthrough() {
    const float lim = 3.f, above = std::nextafter(lim, FLT_MAX), nan = std::numeric_limits<float>::quiet_NaN();
    std::vector<P> in = {pt(-lim, 0, 0, 1), pt(lim, 0, 0, 2), pt(above, 0, 0, 3), pt(-above, 0, 0, 4), pt(nan, 0, 0, 5), pt(0, nan, 0, 6),
                         pt(0, 0, std::numeric_limits<float>::infinity(), 7), pt(1, 2, 3, 8)};
    Cloud::Ptr c = cloud_of(in), out(new Cloud);
    pcl::PassThrough<P> pass;
    pass.setInputCloud(c);
    pass.setFilterFieldName("x");
    pass.setFilterLimits(-lim, lim);
    pass.filter(*out);
    const float want[3] = {1, 2, 8};  // inclusive limits; any non-finite coordinate drops the point; order kept
    bool ok = out->points.size() == 3;
    for (size_t i = 0; ok && i < 3; i++) ok = out->points[i].intensity == want[i];
    report("passthrough", ok, "(A1)");
}

// ------------------------------------------------------------------------------------------------ A2
static void probe_cropbox() {
    const float x = 3.f, y = 3.f, zlo = -0.5f, zhi = 5.f;
    std::vector<P> in = {pt(x, y, zhi, 0), pt(-x, -y, zlo, 1), pt(0, 0, std::nextafter(zlo, -FLT_MAX), 2), pt(0, 0, 0, 3),
                         pt(0, 0, std::nextafter(zhi, FLT_MAX), 4), pt(1, 1, 1, 5), pt(0, 0, -2, 6)};
    Cloud::Ptr c = cloud_of(in), out(new Cloud);
    pcl::CropBox<P> box(true);
    box.setMin(Eigen::Vector4f(-x, -y, zlo, 1.f));
    box.setMax(Eigen::Vector4f(x, y, zhi, 1.f));
    box.setInputCloud(c);
    box.filter(*out);
    pcl::IndicesConstPtr rem = box.getRemovedIndices();
    const float kept[4] = {0, 1, 3, 5};
    const int removed[3] = {2, 4, 6};
    bool ok = out->points.size() == 4 && rem && rem->size() == 3;
    for (size_t i = 0; ok && i < 4; i++) ok = out->points[i].intensity == kept[i];
    for (size_t i = 0; ok && i < 3; i++) ok = (*rem)[i] == removed[i];
    report("cropbox", ok, "(A2)");
}

// ------------------------------------------------------------------------------------------------ A5..A9
static std::vector<pcl::PointIndices> clusters_of(const Cloud::Ptr& c, float tol, int mn, int mx) {
    pcl::search::KdTree<P>::Ptr tree(new pcl::search::KdTree<P>);
    tree->setInputCloud(c);
    std::vector<pcl::PointIndices> out;
    pcl::EuclideanClusterExtraction<P> ec;
    ec.setClusterTolerance(tol);
    ec.setMinClusterSize(mn);
    ec.setMaxClusterSize(mx);
    ec.setSearchMethod(tree);
    ec.setInputCloud(c);
    ec.extract(out);
    return out;
}

static void probe_cluster_radius() {
    const float tol = 0.11f;
    const float r2 = (float)((double)tol * (double)tol);  // A7
    // largest float d with d*d < r2 and the next one up (d*d >= r2), both measured from the origin so that dx == d
    float d = std::sqrt(r2);
    while (d * d >= r2) d = std::nextafter(d, 0.f);
    while (std::nextafter(d, FLT_MAX) * std::nextafter(d, FLT_MAX) < r2) d = std::nextafter(d, FLT_MAX);
    const float d_in = d, d_out = std::nextafter(d, FLT_MAX);
    const bool in_joined = clusters_of(cloud_of({pt(0, 0, 0), pt(d_in, 0, 0)}), tol, 2, 10).size() == 1;
    const bool out_joined = clusters_of(cloud_of({pt(0, 0, 0), pt(d_out, 0, 0)}), tol, 2, 10).size() == 1;
    char msg[160];
    std::snprintf(msg, sizeof msg, "(A6/A7: d2<r2 strict; r2=%.9g d_in=%.9g joined=%d, d_out=%.9g joined=%d)", r2, d_in, (int)in_joined, d_out, (int)out_joined);
    report("cluster_radius", in_joined && !out_joined, msg);
}

static void probe_cluster_order() {
    // chains of points 0.05 apart (tolerance 0.11): sizes 5, 9, 3 (< min), 9, 12 (> max), far from each other
    std::vector<P> in;
    const int sizes[5] = {5, 9, 3, 9, 12};
    std::vector<std::vector<int>> members(5);
    // interleave the chains so that "indices ascending inside a cluster" is not trivially the insertion order
    for (int step = 0; step < 12; step++)
        for (int b = 0; b < 5; b++)
            if (step < sizes[b]) { members[b].push_back((int)in.size()); in.push_back(pt(10.f * b + 0.05f * step, 0, 0)); }
    std::vector<pcl::PointIndices> cl = clusters_of(cloud_of(in), 0.11f, 4, 10);
    // expected: the two 9-chains (b=1 then b=3: discovery order = ascending minimum index), then the 5-chain
    const int want[3] = {1, 3, 0};
    bool ok = cl.size() == 3;
    for (int k = 0; ok && k < 3; k++) ok = cl[k].indices == members[want[k]];
    report("cluster_order", ok, "(A5/A9: whole component size-tested, indices ascending, size desc then discovery order)");
}

// ------------------------------------------------------------------------------------------------ A10
static void probe_centroid() {
    Rng r(10);
    std::vector<P> in;
    for (int i = 0; i < 1000; i++) in.push_back(pt(r.uni(-40, 40), r.uni(-40, 40), r.uni(-2, 3)));
    Cloud::Ptr c = cloud_of(in);
    c->is_dense = true;
    Eigen::Vector4d got;
    pcl::compute3DCentroid(*c, got);
    double sx = 0, sy = 0, sz = 0;
    for (const P& p : in) { sx += p.x; sy += p.y; sz += p.z; }
    const double n = (double)in.size();
    report("centroid", got[0] == sx / n && got[1] == sy / n && got[2] == sz / n, "(A10: sequential double sums / n, bit for bit)");
}

// ------------------------------------------------------------------------------------------------ A12
static void probe_transform() {
    Rng r(12);
    Eigen::Quaterniond qd(r.uni(-1, 1), r.uni(-1, 1), r.uni(-1, 1), r.uni(-1, 1));
    qd.normalize();
    const Eigen::Quaternionf q((float)qd.w(), (float)qd.x(), (float)qd.y(), (float)qd.z());
    const Eigen::Affine3f t = Eigen::Translation3f(r.uni(-2, 2), r.uni(-2, 2), r.uni(-1, 1)) * q;
    std::vector<P> in;
    for (int i = 0; i < 5000; i++) in.push_back(pt(r.uni(-40, 40), r.uni(-40, 40), r.uni(-2, 3), (float)i));
    Cloud::Ptr c = cloud_of(in);
    c->is_dense = true;
    Cloud out;
    pcl::transformPointCloud(*c, out, t);
    const Eigen::Matrix4f& m = t.matrix();
    int diff = 0;
    for (size_t i = 0; i < in.size(); i++) {
        const P& p = in[i];
        const float x = ((m(0, 0) * p.x + m(0, 1) * p.y) + m(0, 2) * p.z) + m(0, 3);
        const float y = ((m(1, 0) * p.x + m(1, 1) * p.y) + m(1, 2) * p.z) + m(1, 3);
        const float z = ((m(2, 0) * p.x + m(2, 1) * p.y) + m(2, 2) * p.z) + m(2, 3);
        if (x != out.points[i].x || y != out.points[i].y || z != out.points[i].z || out.points[i].intensity != p.intensity) diff++;
    }
    char msg[96];
    std::snprintf(msg, sizeof msg, "(A12: %d of %zu points differ from the left-to-right float formula)", diff, in.size());
    report("transform", diff == 0, msg);
}

// ------------------------------------------------------------------------------------------------ A13'
static long predicted_new_voxels(const std::vector<P>& c1, const std::vector<P>& c2, bool refined) {
    const double res = (double)0.1f, eps = (double)FLT_EPSILON;
    double anchor[3];
    const double f[3] = {(double)c1[0].x, (double)c1[0].y, (double)c1[0].z};
    for (int q = 0; q < 3; q++) {
        const double lo = f[q] - res / 2, hi = f[q] + res / 2;
        const double over = refined ? ((2.0 * res - eps) - (hi - lo)) / 2.0 : 0.0;
        anchor[q] = lo - over;
    }
    auto cell = [&](const P& p) {
        return std::make_tuple((long long)std::floor(((double)p.x - anchor[0]) / res), (long long)std::floor(((double)p.y - anchor[1]) / res),
                               (long long)std::floor(((double)p.z - anchor[2]) / res));
    };
    std::set<std::tuple<long long, long long, long long>> occ;
    for (const P& p : c1) occ.insert(cell(p));
    long n = 0;
    for (const P& p : c2) n += occ.count(cell(p)) ? 0 : 1;
    return n;
}

static void probe_octree_anchor() {
    int agree_refined = 0, agree_naive = 0;
    const int trials = 20;
    for (int t = 0; t < trials; t++) {
        Rng r(1300 + t);
        std::vector<P> c1, c2;
        const float ox = r.uni(-20, 20), oy = r.uni(-20, 20), oz = r.uni(-1, 2);
        for (int i = 0; i < 3000; i++) c1.push_back(pt(ox + r.uni(0, 2), oy + r.uni(0, 2), oz + r.uni(0, 1.5f)));
        for (int i = 0; i < 3000; i++) c2.push_back(pt(ox + 0.15f + r.uni(0, 2), oy + r.uni(0, 2), oz + r.uni(0, 1.5f)));
        pcl::octree::OctreePointCloudChangeDetector<P> oc(0.1f);
        oc.setInputCloud(cloud_of(c1));
        oc.addPointsFromInputCloud();
        oc.switchBuffers();
        oc.setInputCloud(cloud_of(c2));
        oc.addPointsFromInputCloud();
        std::vector<int> idx;
        oc.getPointIndicesFromNewVoxels(idx);
        agree_refined += (long)idx.size() == predicted_new_voxels(c1, c2, true);
        agree_naive += (long)idx.size() == predicted_new_voxels(c1, c2, false);
    }
    char msg[160];
    std::snprintf(msg, sizeof msg, "(A13': %d/%d trials match anchor first-res+eps/2; %d/%d match the unrefined first-res/2)", agree_refined, trials, agree_naive, trials);
    report("octree_anchor", agree_refined == trials, msg);
}

// ------------------------------------------------------------------------------------------------ A14
static void probe_voxelgrid() {
    Rng r(14);
    std::vector<P> in;
    for (int i = 0; i < 4000; i++) in.push_back(pt(r.uni(-10, 10), r.uni(-10, 10), r.uni(-1, 1), r.uni(0, 1)));
    const float leaf = 0.5f;
    Cloud out;
    pcl::VoxelGrid<P> vg;
    vg.setInputCloud(cloud_of(in));
    vg.setLeafSize(leaf, leaf, leaf);
    vg.filter(out);
    // prediction: min_b = floor(min_p / leaf) with inverse_leaf = 1/leaf in float, idx ascending, float mean of x,y,z
    const float inv = 1.0f / leaf;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (const P& p : in) { const float v[3] = {p.x, p.y, p.z}; for (int q = 0; q < 3; q++) { mn[q] = std::min(mn[q], v[q]); mx[q] = std::max(mx[q], v[q]); } }
    int mnb[3], dv[3];
    for (int q = 0; q < 3; q++) { mnb[q] = (int)std::floor(mn[q] * inv); dv[q] = (int)std::floor(mx[q] * inv) - mnb[q] + 1; }
    std::vector<std::pair<int, int>> keyed;
    for (size_t i = 0; i < in.size(); i++) {
        const int a = (int)std::floor(in[i].x * inv) - mnb[0], b = (int)std::floor(in[i].y * inv) - mnb[1], c = (int)std::floor(in[i].z * inv) - mnb[2];
        keyed.push_back(std::make_pair(a + b * dv[0] + c * dv[0] * dv[1], (int)i));
    }
    std::sort(keyed.begin(), keyed.end());
    std::vector<P> want;
    for (size_t i = 0; i < keyed.size();) {
        size_t j = i;
        float sx = 0, sy = 0, sz = 0;
        while (j < keyed.size() && keyed[j].first == keyed[i].first) { const P& p = in[keyed[j].second]; sx += p.x; sy += p.y; sz += p.z; j++; }
        const float n = (float)(j - i);
        want.push_back(pt(sx / n, sy / n, sz / n));
        i = j;
    }
    bool ok = want.size() == out.points.size();
    double worst = 0;
    for (size_t i = 0; ok && i < want.size(); i++)
        worst = std::max(worst, (double)std::max(std::fabs(want[i].x - out.points[i].x), std::max(std::fabs(want[i].y - out.points[i].y), std::fabs(want[i].z - out.points[i].z))));
    char msg[128];
    std::snprintf(msg, sizeof msg, "(A14: %zu voxels vs %zu predicted, order by index, worst centroid difference %.3g)", out.points.size(), want.size(), worst);
    report("voxelgrid", ok && worst < 1e-5, msg);
}

// ------------------------------------------------------------------------------------------------ A16
static void probe_reciprocal() {
    Rng r(16);
    pcl::PointCloud<pcl::PointXYZ>::Ptr src(new pcl::PointCloud<pcl::PointXYZ>), tgt(new pcl::PointCloud<pcl::PointXYZ>);
    for (int i = 0; i < 60; i++) src->points.push_back(pcl::PointXYZ(r.uni(-20, 20), r.uni(-20, 20), r.uni(-1, 2)));
    for (int i = 0; i < 55; i++) tgt->points.push_back(pcl::PointXYZ(r.uni(-20, 20), r.uni(-20, 20), r.uni(-1, 2)));
    src->width = 60; src->height = 1; tgt->width = 55; tgt->height = 1;
    pcl::registration::CorrespondenceEstimation<pcl::PointXYZ, pcl::PointXYZ> ce;
    ce.setInputSource(src);
    ce.setInputTarget(tgt);
    pcl::Correspondences got;
    ce.determineReciprocalCorrespondences(got);
    auto d2 = [](const pcl::PointXYZ& a, const pcl::PointXYZ& b) { const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z; return ((dx * dx) + (dy * dy)) + (dz * dz); };
    auto nn = [&](const pcl::PointCloud<pcl::PointXYZ>& c, const pcl::PointXYZ& q, float* d) {
        int best = -1; float bd = FLT_MAX;
        for (size_t i = 0; i < c.points.size(); i++) { const float v = d2(q, c.points[i]); if (v < bd) { bd = v; best = (int)i; } }
        *d = bd;
        return best;
    };
    std::vector<std::tuple<int, int, float>> want;
    for (int i = 0; i < 60; i++) {
        float d, dr;
        const int j = nn(*tgt, src->points[i], &d);
        if (nn(*src, tgt->points[j], &dr) == i) want.push_back(std::make_tuple(i, j, d));
    }
    bool ok = want.size() == got.size();
    for (size_t k = 0; ok && k < want.size(); k++) ok = got[k].index_query == std::get<0>(want[k]) && got[k].index_match == std::get<1>(want[k]) && got[k].distance == std::get<2>(want[k]);
    report("reciprocal", ok, "(A16: reciprocal pairs, ascending query, distance = squared float)");
}

// ------------------------------------------------------------------------------------------------ A18
static void probe_extract_overflow() {
    std::vector<P> in = {pt(0, 0, 0), pt(1, 0, 0), pt(2, 0, 0)};
    Cloud::Ptr c = cloud_of(in);
    pcl::PointIndices::Ptr idx(new pcl::PointIndices);
    idx->indices = {0, 0, 1, 1};  // more indices than points, as several mo_vec entries on one cluster produce (cpp:644-648)
    pcl::ExtractIndices<P> ex;
    ex.setInputCloud(c);
    ex.setIndices(idx);
    ex.setNegative(true);
    Cloud out;
    ex.filter(out);
    pcl::PointIndices::Ptr dup(new pcl::PointIndices);
    dup->indices = {1, 1};  // duplicates within the size limit are harmless
    ex.setIndices(dup);
    Cloud out2;
    ex.filter(out2);
    char msg[96];
    std::snprintf(msg, sizeof msg, "(A18: overflow -> %zu points (want 0); duplicates -> %zu points (want 2))", out.points.size(), out2.points.size());
    report("extract_overflow", out.points.empty() && out2.points.size() == 2, msg);
}

int main() {
    std::printf("PCL %d.%d.%d\n", PCL_MAJOR_VERSION, PCL_MINOR_VERSION, PCL_REVISION_VERSION);
    probe_passthrough();
    probe_cropbox();
    probe_cluster_radius();
    probe_cluster_order();
    probe_centroid();
    probe_transform();
    probe_octree_anchor();
    probe_voxelgrid();
    probe_reciprocal();
    probe_extract_overflow();
    std::printf("%d probe(s) failed\n", g_fail);
    return g_fail;
}
