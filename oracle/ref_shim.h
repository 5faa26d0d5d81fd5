// ref_shim.h - the few names the Boost-free, torch-free part of the reference's games/game_helpers.cpp needs
// (TEST INFRASTRUCTURE; written for oracle/build_ref.py, which compiles lines 15-66, 146-156 and 191-279 of the
// reference file where it lies into oracle/_ref/libgame_ref.so).  Nothing here computes anything: `point` with
// bg::get<>, the two RaceTrack members update_players reads, and a minimal stand-in for at::Tensor / accessor.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <utility>
#include <vector>

namespace bg {
struct point_xy { float x, y; };
template <int I> inline float get(const point_xy& p) { return I == 0 ? p.x : p.y; }
}  // namespace bg
using point = bg::point_xy;

struct RaceTrack {              // game_helpers.cpp:85-107 minus the Boost members (not used by update_players)
    std::vector<point> left;
    std::vector<point> right;
    size_t length;
};

namespace at {
enum ScalarType { kByte, kFloat, kLong };
template <typename T, int N> struct Accessor {
    T* data; const int64_t* strides;
    Accessor<T, N - 1> operator[](int64_t i) const { return Accessor<T, N - 1>{data + i * strides[0], strides + 1}; }
};
template <typename T> struct Accessor<T, 1> {
    T* data; const int64_t* strides;
    T& operator[](int64_t i) const { return data[i * strides[0]]; }
};
struct Tensor {
    void* data = nullptr;
    std::vector<int64_t> sizes, strides;
    std::vector<uint8_t> owned;
    int64_t size(int d) const { return sizes[d]; }
    // (a tensor made by empty() owns its bytes; copies of it must not point at the original's buffer)
    void* ptr() const { return owned.empty() ? data : const_cast<uint8_t*>(owned.data()); }
    template <typename T, size_t N> Accessor<T, (int)N> accessor() const { return Accessor<T, (int)N>{static_cast<T*>(ptr()), strides.data()}; }
};
inline Tensor empty(std::initializer_list<size_t> shape, ScalarType t) {
    Tensor r;
    size_t n = 1;
    for (size_t s : shape) { r.sizes.push_back((int64_t)s); n *= s; }
    r.strides.assign(r.sizes.size(), 1);
    for (int d = (int)r.sizes.size() - 2; d >= 0; --d) r.strides[d] = r.strides[d + 1] * r.sizes[d + 1];
    r.owned.assign(n * (t == kByte ? 1 : t == kFloat ? 4 : 8) + 8, 0);
    return r;
}
inline Tensor wrap(void* data, std::initializer_list<int64_t> shape, std::initializer_list<int64_t> strides) {
    Tensor r;
    r.data = data;
    r.sizes.assign(shape);
    r.strides.assign(strides);
    return r;
}
}  // namespace at
namespace torch { inline at::ScalarType CPU(at::ScalarType t) { return t; } }
