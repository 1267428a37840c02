#pragma once
#include <pcl/PCLPointCloud2.h>
#include <sensor_msgs/PointCloud2.h>
namespace pcl_conversions {
void toPCL(const sensor_msgs::PointCloud2& in, pcl::PCLPointCloud2& out);
void fromPCL(const pcl::PCLPointCloud2& in, sensor_msgs::PointCloud2& out);
}
