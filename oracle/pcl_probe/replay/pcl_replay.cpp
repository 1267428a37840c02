// pcl_replay.cpp — runs the UNMODIFIED reference class (src/MovingObjectRemoval.cpp of prabinrath/dynamicslamtool,
// compiled from where it lies; nothing of it is copied into this repository) over the seeded C1 fixture sequence and
// writes tests/golden/pcl_c1.json in the schema of tests/golden/make_golden.py. With that file present,
// tests/test_golden.py::test_*_reproduces_pcl_golden compare the oracle and the CUDA path with the real
// PCL 1.8 / FLANN / tf arithmetic: this is what turns "parity unpinned" into pinned.
//
// Needs ROS melodic + PCL 1.8 (docker image ros:melodic-perception; see ../run_in_docker.sh). NOT part of this
// repository's build: it has never been compiled here (no ROS, no PCL in the build image).
//
//   pcl_replay <MOR_config.txt> <out.json> [frames=12] [scenario=1] [seed=1]
//
// The reference keeps its frame state private; the fixture needs cluster sizes, centroids, flags and mo_vec, so the
// header is included with `private` opened up. Test infrastructure only.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#define private public
#define class struct
#include "MOR/MovingObjectRemoval.h"  // the reference's header (include path: <reference>/include)
#undef class
#undef private

#include "mor_synth.h"  // this repository's seeded generator (dynamicslamtool_b200/csrc)

static uint32_t crc32_bytes(const void* data, size_t n) {  // zlib's CRC-32 (what tests/helpers.py:crc computes)
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    uint32_t c = 0xFFFFFFFFu;
    const uint8_t* p = static_cast<const uint8_t*>(data);
    for (size_t i = 0; i < n; i++) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

static void json_ints(FILE* f, const char* key, const std::vector<long long>& v) {
    std::fprintf(f, "   \"%s\": [", key);
    for (size_t i = 0; i < v.size(); i++) std::fprintf(f, "%s%lld", i ? ", " : "", v[i]);
    std::fprintf(f, "]");
}

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: pcl_replay <MOR_config.txt> <out.json> [frames] [scenario] [seed]\n"); return 2; }
    const std::string config = argv[1], out_path = argv[2];
    const int frames = argc > 3 ? std::atoi(argv[3]) : 12;
    const int scenario = argc > 4 ? std::atoi(argv[4]) : 1;
    const uint64_t seed = argc > 5 ? std::strtoull(argv[5], nullptr, 10) : 1;

    ros::init(argc, argv, "pcl_replay");
    ros::NodeHandle nh;  // the reference's constructor takes one (VISUALIZE publishers); needs a running roscore
    MovingObjectRemoval mor(nh, config, 4, 3);

    mor_synth* syn = nullptr;
    if (mor_synth_create(scenario, seed, &syn)) return 3;
    uint32_t maxp = 0, nominal = 0; double hz = 0;
    mor_synth_info(syn, &maxp, &nominal, &hz);
    std::vector<float> xyzi((size_t)maxp * 4);

    FILE* f = std::fopen(out_path.c_str(), "w");
    if (!f) return 4;
    std::fprintf(f, "{\n \"scenario\": %d, \"seed\": %llu, \"config\": \"config/MOR_config.txt\", \"n_bad\": 4, \"n_good\": 3,\n"
                    " \"source\": \"prabinrath/dynamicslamtool src/MovingObjectRemoval.cpp, PCL %d.%d.%d\",\n \"frames\": [\n",
                 scenario, (unsigned long long)seed, PCL_MAJOR_VERSION, PCL_MINOR_VERSION, PCL_REVISION_VERSION);
    for (int fr = 0; fr < frames; fr++) {
        uint32_t n = 0; double pose7[7];
        if (mor_synth_frame(syn, (uint32_t)fr, xyzi.data(), maxp, &n, pose7, 1)) return 5;
        // the frame as the driver would publish it: PointXYZI cloud -> PCLPointCloud2 (what pcl_conversions::toPCL yields, external_sync_test.cpp:12-13)
        pcl::PointCloud<pcl::PointXYZI> pc;
        pc.resize(n);
        for (uint32_t i = 0; i < n; i++) { pc[i].x = xyzi[i * 4]; pc[i].y = xyzi[i * 4 + 1]; pc[i].z = xyzi[i * 4 + 2]; pc[i].intensity = xyzi[i * 4 + 3]; }
        pc.width = n; pc.height = 1; pc.is_dense = false;
        pcl::PCLPointCloud2 cloud;
        pcl::toPCLPointCloud2(pc, cloud);
        geometry_msgs::Pose pose;
        pose.position.x = pose7[0]; pose.position.y = pose7[1]; pose.position.z = pose7[2];
        pose.orientation.x = pose7[3]; pose.orientation.y = pose7[4]; pose.orientation.z = pose7[5]; pose.orientation.w = pose7[6];

        mor.pushRawCloudAndPose(cloud, pose);
        const bool have = mor.filterCloud(cloud, "/filtered");

        const MovingObjectDetectionCloud& cb = *mor.cb;
        const size_t NC = cb.cloud->size(), NG = cb.gp_indices ? cb.gp_indices->size() : 0, K = cb.clusters.size();
        std::vector<int32_t> cid(NC, -1);
        std::vector<long long> sizes, roots, flags, conf;
        for (size_t k = 0; k < K; k++) {
            const std::vector<int>& ind = cb.cluster_indices[k].indices;
            long long mn = ind.empty() ? -1 : ind[0];
            for (int i : ind) { cid[(size_t)i] = (int32_t)k; if (i < mn) mn = i; }
            sizes.push_back((long long)ind.size());
            roots.push_back(mn);
        }
        for (size_t k = 0; k < cb.detection_results.size(); k++) flags.push_back(cb.detection_results[k] ? 1 : 0);
        while (flags.size() < K) flags.push_back(0);  // the first frame has no detection results
        for (const MovingObjectCentroid& m : mor.mo_vec) conf.push_back(m.confidence);
        const size_t nout = have ? (size_t)mor.output.width * mor.output.height : 0;

        std::fprintf(f, "  {\n   \"N\": %u, \"NT\": %zu, \"NC\": %zu, \"NG\": %zu, \"K\": %zu, \"NMO\": %zu, \"NOUT\": %zu,\n", n, NC + NG, NC, NG, K,
                     mor.mo_vec.size(), nout);
        std::fprintf(f, "   \"crc_input\": %u, \"crc_cluster_id\": %u, \"crc_output\": %u,\n", crc32_bytes(xyzi.data(), (size_t)n * 16),
                     crc32_bytes(cid.data(), cid.size() * 4), have ? crc32_bytes(mor.output.data.data(), mor.output.data.size()) : crc32_bytes("", 0));
        json_ints(f, "cluster_root", roots); std::fprintf(f, ",\n");
        json_ints(f, "cluster_size", sizes); std::fprintf(f, ",\n");
        json_ints(f, "flags", flags); std::fprintf(f, ",\n");
        json_ints(f, "mo_conf", conf); std::fprintf(f, ",\n");
        std::fprintf(f, "   \"centroids\": [");
        for (size_t k = 0; k < cb.centroid_collection->size(); k++) {
            const pcl::PointXYZ& c = (*cb.centroid_collection)[k];
            std::fprintf(f, "%s[%.9g, %.9g, %.9g]", k ? ", " : "", c.x, c.y, c.z);
        }
        std::fprintf(f, "],\n   \"pose\": [%.17g, %.17g, %.17g, %.17g, %.17g, %.17g, %.17g]\n  }%s\n", pose7[0], pose7[1], pose7[2], pose7[3], pose7[4], pose7[5],
                     pose7[6], fr + 1 < frames ? "," : "");
        std::printf("frame %d: N %u NC %zu NG %zu K %zu mo %zu out %zu\n", fr, n, NC, NG, K, mor.mo_vec.size(), nout);
    }
    std::fprintf(f, " ]\n}\n");
    std::fclose(f);
    mor_synth_destroy(syn);
    return 0;
}
