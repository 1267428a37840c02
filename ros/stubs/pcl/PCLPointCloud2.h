#pragma once
#include <cstdint>
#include <string>
#include <vector>
namespace pcl {
struct PCLHeader { uint32_t seq = 0; uint64_t stamp = 0; std::string frame_id; };
struct PCLPointField {
    std::string name; uint32_t offset = 0; uint8_t datatype = 0; uint32_t count = 0;
    enum PointFieldTypes { INT8 = 1, UINT8 = 2, INT16 = 3, UINT16 = 4, INT32 = 5, UINT32 = 6, FLOAT32 = 7, FLOAT64 = 8 };
};
struct PCLPointCloud2 {
    PCLHeader header; uint32_t height = 0, width = 0; std::vector<PCLPointField> fields;
    uint8_t is_bigendian = 0; uint32_t point_step = 0, row_step = 0; std::vector<uint8_t> data; uint8_t is_dense = 0;
};
}  // namespace pcl
