#pragma once
#include <cstdint>
namespace message_filters { namespace sync_policies {
template <class M0, class M1> struct ApproximateTime { explicit ApproximateTime(uint32_t) {} };
} }
