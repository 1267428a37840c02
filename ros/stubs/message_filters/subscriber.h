#pragma once
#include <ros/ros.h>
namespace message_filters {
template <class M> class Subscriber { public: Subscriber(ros::NodeHandle&, const std::string&, uint32_t) {} };
}
