// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or executed from the product path.
//
// CPU restatement of CanvasSmooth (reference @ v1.40.0):
//   CanvasCommon/Utilities.cs:767-791        MedianFilter — the streaming loop, statement for statement (a sorted window,
//                                             a queue of the values still inside it; a median is emitted once the window
//                                             holds halfWindow + 1 values, the oldest value leaves when it holds
//                                             2 * halfWindow + 1, and the tail shrinks the window back to halfWindow + 1)
//   CanvasSmooth/CanvasSmooth.cs:66-77        RepeatedMedianFilter: half windows 1 .. maxHalfWindowSize, each on the last output
// Pinned by CanvasTest/TestUtilities.cs:195-206 (TestMedianFilter), which also fixes SortedList<float>.Median(): mean of the
// two middle elements, in float.
#include <cstdint>
#include <deque>
#include <set>
#include <vector>

#include "oracle.h"
#include "ref_stats.hpp"

namespace {

struct DotnetLessF {
    bool operator()(float a, float b) const { return ora::dotnet_less<float>(a, b); }
};

float window_median(const std::multiset<float, DotnetLessF>& w) {
    const size_t n = w.size();
    auto it = w.begin();
    std::advance(it, (n - 1) / 2);
    if (n & 1) return *it;
    const float a = *it;
    ++it;
    return (a + *it) / 2.0f;
}

std::vector<float> median_filter(const std::vector<float>& values, uint32_t half) {
    const size_t boundary = (size_t)half + 1, window_size = (size_t)half * 2 + 1;
    std::multiset<float, DotnetLessF> window;
    std::deque<float> previous;
    std::vector<float> out;
    for (float v : values) {
        if (window.size() >= window_size && !previous.empty()) {
            window.erase(window.find(previous.front()));
            previous.pop_front();
        }
        window.insert(v);
        if (window.size() >= boundary) out.push_back(window_median(window));
        previous.push_back(v);
    }
    while (window.size() > boundary && !previous.empty()) {
        window.erase(window.find(previous.front()));
        previous.pop_front();
        out.push_back(window_median(window));
    }
    return out;
}

}  // namespace

extern "C" int64_t ora_median_filter(int64_t n, const float* in, uint32_t half_window, float* out) {
    std::vector<float> r = median_filter(std::vector<float>(in, in + n), half_window);
    for (size_t i = 0; i < r.size(); i++) out[i] = r[i];
    return (int64_t)r.size();
}

extern "C" int64_t ora_repeated_median_filter(int64_t n, const float* in, uint32_t max_half_window, float* out) {
    std::vector<float> cur(in, in + n);
    for (uint32_t h = 1; h <= max_half_window; h++) cur = median_filter(cur, h);
    for (size_t i = 0; i < cur.size(); i++) out[i] = cur[i];
    return (int64_t)cur.size();
}
