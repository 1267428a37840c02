#pragma once
#include <geometry_msgs/Pose.h>
#include <std_msgs/Header.h>
#include <boost/shared_ptr.hpp>
namespace nav_msgs {
struct Odometry { std_msgs::Header header; std::string child_frame_id; geometry_msgs::PoseWithCovariance pose; };
typedef boost::shared_ptr<Odometry const> OdometryConstPtr;
}  // namespace nav_msgs
