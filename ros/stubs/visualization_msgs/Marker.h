#pragma once
#include <geometry_msgs/Pose.h>
#include <std_msgs/Header.h>
namespace std_msgs { struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; }; }
namespace visualization_msgs {
struct Marker {
    enum { ARROW = 0, CUBE = 1, SPHERE = 2 }; enum { ADD = 0, MODIFY = 0, DELETE = 2 };
    std_msgs::Header header; std::string ns; int32_t id = 0; int32_t type = 0; int32_t action = 0;
    geometry_msgs::Pose pose; geometry_msgs::Vector3 scale; std_msgs::ColorRGBA color; ros::Duration lifetime;
};
}  // namespace visualization_msgs
