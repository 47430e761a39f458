// smallmat.h — fixed-size column-major matrices for the host-side mirror classes (Eigen is not a dependency here).
// Only storage and element access: all algebra of the hot path runs on the GPU.
#pragma once
#include <cstddef>

namespace ilqg {
template <class T, int R, int C>
struct Mat {
    static constexpr int Rows = R, Cols = C;
    T d[R * C];
    Mat() { for (int i = 0; i < R * C; i++) d[i] = T(0); }
    T& operator()(int i, int j) { return d[i + (std::size_t)j * R]; }
    const T& operator()(int i, int j) const { return d[i + (std::size_t)j * R]; }
    T& operator()(int i) { return d[i]; }
    const T& operator()(int i) const { return d[i]; }
    T* data() { return d; }
    const T* data() const { return d; }
    void setZero() { for (int i = 0; i < R * C; i++) d[i] = T(0); }
};
// non-owning column-major view (the role of Eigen::Map in the reference)
template <class T, int R, int C>
struct MatMap {
    T* p;
    explicit MatMap(T* ptr) : p(ptr) {}
    T& operator()(int i, int j) { return p[i + (std::size_t)j * R]; }
    const T& operator()(int i, int j) const { return p[i + (std::size_t)j * R]; }
    T& operator()(int i) { return p[i]; }
    const T& operator()(int i) const { return p[i]; }
    T* data() { return p; }
};
}  // namespace ilqg
