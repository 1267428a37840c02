#pragma once
#include <functional>
namespace message_filters {
template <class Policy> class Synchronizer {
public:
    template <class F0, class F1> Synchronizer(const Policy&, F0&, F1&) {}
    template <class C> void registerCallback(const C&) {}
};
}
