// mor_config.cpp — MOR_config.txt parser of the product library (host C++).
//
// Grammar and key set of MovingObjectRemoval::setVariables (reference src/MovingObjectRemoval.cpp:698-864):
//   * a line starting with '#' or shorter than 3 characters is skipped (cpp:712);
//   * the key is everything before the first ':'; the value is every later character that is not ':'
//     (cpp:718-733) - no whitespace trimming;
//   * floats via std::stof, cluster sizes via std::stol, method_choice via std::stoi,
//     opc_normalization_factor via std::stof truncated into an int (cpp:843).
// Differences, all on error paths: status codes instead of exit(0); a key that never
// appears is an error instead of an uninitialised member; method_choice outside {1,2} is an error.
// Extension keys (absent from the reference): ground_mode, gp_planarity, gp_bin_width.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>

#include "../../include/mor_b200.h"

namespace {

enum Kind { F32, I64, I32, I32_FROM_F32, STR };

struct KeySpec {
    const char* name;
    Kind kind;
    size_t offset;
    bool required;
};

#define MOR_KEY(field, kind, req) {#field, kind, offsetof(mor_config, field), req}
const KeySpec kKeys[] = {
    MOR_KEY(method_choice, I32, true),
    MOR_KEY(output_topic, STR, true),
    MOR_KEY(debug_topic, STR, true),
    MOR_KEY(marker_topic, STR, true),
    MOR_KEY(input_pointcloud_topic, STR, true),
    MOR_KEY(input_odometry_topic, STR, true),
    MOR_KEY(output_fid, STR, true),
    MOR_KEY(debug_fid, STR, true),
    MOR_KEY(ec_distance_threshold, F32, true),
    MOR_KEY(min_cluster_size, I64, true),
    MOR_KEY(max_cluster_size, I64, true),
    MOR_KEY(gp_leaf, F32, true),
    MOR_KEY(bin_gap, F32, true),
    MOR_KEY(gp_limit, F32, true),
    MOR_KEY(trim_x, F32, true),
    MOR_KEY(trim_y, F32, true),
    MOR_KEY(trim_z, F32, true),
    MOR_KEY(pde_lb, F32, true),
    MOR_KEY(pde_ub, F32, true),
    MOR_KEY(pde_distance_threshold, F32, true),
    MOR_KEY(opc_normalization_factor, I32_FROM_F32, true),
    MOR_KEY(volume_constraint, F32, true),
    MOR_KEY(leave_off_distance, F32, true),
    MOR_KEY(catch_up_distance, F32, true),
    MOR_KEY(ground_mode, I32, false),
    MOR_KEY(gp_planarity, F32, false),
    MOR_KEY(gp_bin_width, F32, false),
};
constexpr int kNumKeys = sizeof(kKeys) / sizeof(kKeys[0]);

}  // namespace

extern "C" int mor_parse_config(const char* config_path, mor_config* out) {
    if (!config_path || !out) return MOR_ERR_ARG;
    std::ifstream file(config_path);
    if (!file.is_open()) return MOR_ERR_CONFIG_OPEN;

    mor_config cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    bool seen[kNumKeys] = {};
    std::string line;
    while (std::getline(file, line)) {
        if (line.empty() || line[0] == '#' || line.size() < 3) continue;
        const size_t colon = line.find(':');
        std::string key = line.substr(0, colon), value;
        if (colon != std::string::npos)
            for (size_t i = colon + 1; i < line.size(); i++)
                if (line[i] != ':') value.push_back(line[i]);

        int which = -1;
        for (int k = 0; k < kNumKeys; k++)
            if (key == kKeys[k].name) { which = k; break; }
        if (which < 0) return MOR_ERR_CONFIG_KEY;

        char* field = reinterpret_cast<char*>(&cfg) + kKeys[which].offset;
        try {
            switch (kKeys[which].kind) {
                case F32: { float v = std::stof(value); std::memcpy(field, &v, sizeof v); } break;
                case I64: { int64_t v = std::stol(value); std::memcpy(field, &v, sizeof v); } break;
                case I32: { int32_t v = std::stoi(value); std::memcpy(field, &v, sizeof v); } break;
                case I32_FROM_F32: { int32_t v = static_cast<int32_t>(std::stof(value)); std::memcpy(field, &v, sizeof v); } break;
                case STR: std::snprintf(field, 64, "%s", value.c_str()); break;
            }
        } catch (...) {
            return MOR_ERR_CONFIG_VALUE;
        }
        seen[which] = true;
    }
    for (int k = 0; k < kNumKeys; k++) {
        if (seen[k]) continue;
        if (kKeys[k].required) return MOR_ERR_CONFIG_MISSING;
        if (!std::strcmp(kKeys[k].name, "ground_mode")) cfg.ground_mode = MOR_GROUND_CROP;
        if (!std::strcmp(kKeys[k].name, "gp_planarity")) cfg.gp_planarity = 0.01f;
        if (!std::strcmp(kKeys[k].name, "gp_bin_width")) cfg.gp_bin_width = cfg.gp_leaf;
    }
    if (cfg.method_choice != 1 && cfg.method_choice != 2) return MOR_ERR_CONFIG_VALUE;
    if (cfg.ground_mode < MOR_GROUND_CROP || cfg.ground_mode > MOR_GROUND_VOXEL_EIGEN) return MOR_ERR_CONFIG_VALUE;
    if (cfg.method_choice == 2 && cfg.opc_normalization_factor == 0) return MOR_ERR_CONFIG_VALUE;
    *out = cfg;
    return MOR_OK;
}

extern "C" const char* mor_status_string(int status) {
    switch (status) {
        case MOR_OK: return "MOR_OK";
        case MOR_ERR_CONFIG_OPEN: return "MOR_ERR_CONFIG_OPEN: could not open the config file";
        case MOR_ERR_CONFIG_KEY: return "MOR_ERR_CONFIG_KEY: invalid parameter found in config file";
        case MOR_ERR_CONFIG_VALUE: return "MOR_ERR_CONFIG_VALUE: unparsable or out-of-domain value";
        case MOR_ERR_CONFIG_MISSING: return "MOR_ERR_CONFIG_MISSING: a required key never appears";
        case MOR_ERR_ARG: return "MOR_ERR_ARG";
        case MOR_ERR_CAPACITY: return "MOR_ERR_CAPACITY";
        case MOR_ERR_CUDA: return "MOR_ERR_CUDA";
        case MOR_ERR_STATE: return "MOR_ERR_STATE";
    }
    return "MOR_ERR_UNKNOWN";
}
