// mov_harness.cpp — ROS-free stand-in for the reference node `mov_e` (src/external_sync_test.cpp:7-41).
//
// The reference callback converts the incoming message, calls pushRawCloudAndPose + filterCloud, publishes
// `output` and prints the milliseconds the pair took (external_sync_test.cpp:9-20). This harness does the same
// with frames from the seeded synthetic generator instead of ROS topics, and adds a CRC of every output cloud so
// runs can be compared across implementations.
//
//   mov_harness <MOR_config.txt> <scenario 1..4> <seed> <frames> [n_bad=4] [n_good=3] [--quiet]
#include <malloc.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../dynamicslamtool_b200/csrc/mor_synth.h"
#include "../include/MOR/MovingObjectRemoval.h"

static uint32_t crc32_buf(const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; }
        init = true;
    }
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

int main(int argc, char** argv) {
    if (argc < 5) { std::fprintf(stderr, "usage: %s <config> <scenario> <seed> <frames> [n_bad] [n_good] [--quiet]\n", argv[0]); return 2; }
    const std::string cfg = argv[1];
    const int scenario = std::atoi(argv[2]);
    const uint64_t seed = std::strtoull(argv[3], nullptr, 10);
    const int frames = std::atoi(argv[4]);
    const int n_bad = argc > 5 && argv[5][0] != '-' ? std::atoi(argv[5]) : 4, n_good = argc > 6 && argv[6][0] != '-' ? std::atoi(argv[6]) : 3;
    bool quiet = false;
    for (int i = 5; i < argc; i++) quiet |= !std::strcmp(argv[i], "--quiet");

    // The callback allocates and frees ~6 MB of cloud buffers per frame (its local PCLPointCloud2, as in the
    // reference). With glibc's defaults those are mapped and unmapped every time (0.25 ms of page faults per MB on
    // the bench host); keep them in the heap instead.
    mallopt(M_MMAP_THRESHOLD, 32 << 20);
    mallopt(M_TRIM_THRESHOLD, 512 << 20);

    mor_synth* syn = nullptr;
    if (mor_synth_create(scenario, seed, &syn)) { std::fprintf(stderr, "bad scenario\n"); return 2; }
    uint32_t maxp = 0;
    mor_synth_info(syn, &maxp, nullptr, nullptr);
    mor_limits lim{};
    lim.max_points = maxp;
    ros::NodeHandle nh;
    MovingObjectRemoval mor(nh, cfg, n_bad, n_good, 0, &lim);  // mor.reset(new MovingObjectRemoval(nh, "...MOR_config.txt", 4, 3))

    pcl::PCLPointCloud2 cloud;  // what pcl_conversions::toPCL(*input, cloud) would hand over: 16-byte x,y,z,intensity records
    cloud.height = 1; cloud.point_step = 16; cloud.is_dense = 1;
    const char* names[4] = {"x", "y", "z", "intensity"};
    for (int i = 0; i < 4; i++) { pcl::PCLPointField f; f.name = names[i]; f.offset = 4 * i; f.datatype = 7; f.count = 1; cloud.fields.push_back(f); }
    std::vector<float> buf((size_t)maxp * 4);
    std::vector<double> ms;
    double stage_ms[3] = {0, 0, 0};  // callback copy, pushRawCloudAndPose, filterCloud (frames after the fifth)
    for (int f = 0; f < frames; f++) {
        uint32_t n = 0;
        double p7[7];
        if (mor_synth_frame(syn, (uint32_t)f, buf.data(), maxp, &n, p7, 8)) { std::fprintf(stderr, "generator failed\n"); return 1; }
        cloud.width = n; cloud.row_step = 16 * n;
        cloud.data.assign((const uint8_t*)buf.data(), (const uint8_t*)buf.data() + (size_t)n * 16);
        geometry_msgs::Pose pose;
        pose.position.x = p7[0]; pose.position.y = p7[1]; pose.position.z = p7[2];
        pose.orientation.x = p7[3]; pose.orientation.y = p7[4]; pose.orientation.z = p7[5]; pose.orientation.w = p7[6];

        const auto t0 = std::chrono::steady_clock::now();
        pcl::PCLPointCloud2 work = cloud;  // the callback's own copy (external_sync_test.cpp:11-12)
        const auto t1 = std::chrono::steady_clock::now();
        mor.pushRawCloudAndPose(work, pose);
        const auto t2 = std::chrono::steady_clock::now();
        const bool ok = mor.filterCloud(work, "/filtered");
        const auto t3 = std::chrono::steady_clock::now();
        const double dt = std::chrono::duration<double, std::milli>(t3 - t0).count();
        if (f >= 5) {
            stage_ms[0] += std::chrono::duration<double, std::milli>(t1 - t0).count();
            stage_ms[1] += std::chrono::duration<double, std::milli>(t2 - t1).count();
            stage_ms[2] += std::chrono::duration<double, std::milli>(t3 - t2).count();
        }
        if (!ok) { std::fprintf(stderr, "frame %d: %s\n", f, mor_status_string(mor.lastStatus())); return 1; }
        ms.push_back(dt);
        if (!quiet) std::printf("frame %d in %u out %u crc %08x ms %.3f\n", f, n, mor.output.width, crc32_buf(mor.output.data.data(), mor.output.data.size()), dt);
    }
    if (ms.size() > 5) {
        std::vector<double> s(ms.begin() + 5, ms.end());
        std::sort(s.begin(), s.end());
        double sum = 0;
        for (double v : s) sum += v;
        std::printf("summary frames %zu mean_ms %.3f p50_ms %.3f p99_ms %.3f fps %.1f\n", s.size(), sum / s.size(), s[s.size() / 2], s[(size_t)(s.size() * 0.99)], 1e3 * s.size() / sum);
        std::printf("stages copy_ms %.3f push_ms %.3f filter_ms %.3f\n", stage_ms[0] / s.size(), stage_ms[1] / s.size(), stage_ms[2] / s.size());
    }
    mor_synth_destroy(syn);
    return 0;
}
