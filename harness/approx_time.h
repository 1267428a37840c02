// approx_time.h — message_filters::sync_policies::ApproximateTime for the ROS-free harness (host only).
//
// The reference node pairs `/velodyne_points` with `/camera/odom/sample` through
// message_filters::Synchronizer<ApproximateTime<PointCloud2, Odometry>>(queue 10) (src/external_sync_test.cpp:31-35):
// the two topics carry independent time stamps and rates, and the callback sees the pairs that policy selects. This
// is that policy's algorithm for recorded streams (the adaptive algorithm of ROS message_filters, approximate_time.h:
// a candidate set is the heads of all queues; the topic whose head is latest is the pivot; the earliest head is moved
// aside while that shrinks the set's time span; a candidate is published once no later set can beat it; queue
// overflow drops the oldest message of the topic and restarts the search), with the policy's defaults: age penalty 0.1,
// no maximum interval, inter-message lower bounds 0. Messages are (stamp in ns, payload index).
#pragma once
#include <cstdint>
#include <deque>
#include <functional>
#include <vector>

namespace replay {

template <int N>
class ApproximateTime {
public:
    struct Msg { int64_t stamp; int index; };
    using Callback = std::function<void(const Msg (&)[N])>;

    ApproximateTime(uint32_t queue_size, Callback cb) : queue_size_(queue_size), cb_(std::move(cb)) {
        for (int i = 0; i < N; i++) has_dropped_[i] = false;
    }
    void set_age_penalty(double p) { age_penalty_ = p; }

    // a message arrives on topic i (callers feed the topics in arrival order, as a bag player would)
    void add(int i, int64_t stamp, int index) {
        deques_[i].push_back(Msg{stamp, index});
        if (deques_[i].size() == 1u) {
            ++non_empty_;
            if (non_empty_ == N) process();
        }
        if (deques_[i].size() + past_[i].size() > queue_size_) {
            // cancel the ongoing candidate search and drop the oldest message of the offending topic
            non_empty_ = 0;
            for (int t = 0; t < N; t++) recover_all(t);
            deques_[i].pop_front();
            has_dropped_[i] = true;
            if (pivot_ != kNoPivot) {
                pivot_ = kNoPivot;
                process();
            }
        }
    }
    uint64_t dropped() const { return dropped_; }

private:
    static constexpr int kNoPivot = -1;
    uint32_t queue_size_;
    Callback cb_;
    std::deque<Msg> deques_[N];
    std::vector<Msg> past_[N];
    bool has_dropped_[N];
    int non_empty_ = 0, pivot_ = kNoPivot;
    Msg candidate_[N];
    int64_t candidate_start_ = 0, candidate_end_ = 0, pivot_time_ = 0;
    double age_penalty_ = 0.1;
    uint64_t dropped_ = 0;

    void delete_front(int i) { deques_[i].pop_front(); if (deques_[i].empty()) --non_empty_; }
    void move_front_to_past(int i) { past_[i].push_back(deques_[i].front()); deques_[i].pop_front(); if (deques_[i].empty()) --non_empty_; }
    void make_candidate() {  // the heads of all queues; everything moved aside before is older than a better candidate: gone for good
        for (int i = 0; i < N; i++) { candidate_[i] = deques_[i].front(); past_[i].clear(); }
    }
    void recover(int i, size_t n) {  // the last n messages of past_[i] go back to the front of the queue
        while (n-- > 0) { deques_[i].push_front(past_[i].back()); past_[i].pop_back(); }
        if (!deques_[i].empty()) ++non_empty_;
    }
    void recover_all(int i) { recover(i, past_[i].size()); }
    void recover_and_delete(int i) {
        while (!past_[i].empty()) { deques_[i].push_front(past_[i].back()); past_[i].pop_back(); }
        deques_[i].pop_front();  // the message that went out with the candidate
        if (!deques_[i].empty()) ++non_empty_;
    }
    void publish_candidate() {
        cb_(candidate_);
        pivot_ = kNoPivot;
        non_empty_ = 0;
        for (int i = 0; i < N; i++) recover_and_delete(i);
    }
    int64_t virtual_time(int i) const {
        if (deques_[i].empty()) {
            const int64_t lower = past_[i].back().stamp;  // + inter_message_lower_bounds_[i] (default 0)
            return lower > pivot_time_ ? lower : pivot_time_;
        }
        return deques_[i].front().stamp;
    }
    void bounds(bool virt, int& start_i, int64_t& start_t, int& end_i, int64_t& end_t) const {
        start_i = end_i = 0;
        start_t = end_t = virt ? virtual_time(0) : deques_[0].front().stamp;
        for (int i = 1; i < N; i++) {
            const int64_t t = virt ? virtual_time(i) : deques_[i].front().stamp;
            if (t < start_t) { start_t = t; start_i = i; }
            if (t > end_t) { end_t = t; end_i = i; }
        }
    }
    bool not_better(int64_t end_t, int64_t than) const { return (double)(end_t - candidate_end_) * (1.0 + age_penalty_) >= (double)than; }

    void process() {
        while (non_empty_ == N) {
            int start_i, end_i;
            int64_t start_t, end_t;
            bounds(false, start_i, start_t, end_i, end_t);
            for (int i = 0; i < N; i++)
                if (i != end_i) has_dropped_[i] = false;
            if (pivot_ == kNoPivot) {
                if (has_dropped_[end_i]) { delete_front(start_i); ++dropped_; continue; }  // not a good pivot
                make_candidate();
                candidate_start_ = start_t; candidate_end_ = end_t;
                pivot_ = end_i; pivot_time_ = end_t;
                move_front_to_past(start_i);
            } else if (not_better(end_t, start_t - candidate_start_)) {
                move_front_to_past(start_i);
            } else {
                make_candidate();
                candidate_start_ = start_t; candidate_end_ = end_t;
                move_front_to_past(start_i);
            }
            if (start_i == pivot_) {
                publish_candidate();  // every candidate of this pivot has been seen
            } else if (not_better(end_t, pivot_time_ - candidate_start_)) {
                publish_candidate();  // provably optimal: any later set contains [pivot_time, end_time]
            } else if (non_empty_ < N) {
                // a queue ran empty: try to prove optimality with the earliest times its next message could have
                size_t moves[N];
                for (int i = 0; i < N; i++) moves[i] = 0;
                while (true) {
                    bounds(true, start_i, start_t, end_i, end_t);
                    if (not_better(end_t, pivot_time_ - candidate_start_)) { publish_candidate(); break; }
                    if (!not_better(end_t, start_t - candidate_start_)) {  // an optimistic candidate would be better: wait
                        non_empty_ = 0;
                        for (int i = 0; i < N; i++) recover(i, moves[i]);
                        break;
                    }
                    move_front_to_past(start_i);
                    moves[start_i]++;
                }
            }
        }
    }
};

}  // namespace replay
