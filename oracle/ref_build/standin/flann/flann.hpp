// TEST INFRASTRUCTURE (oracle/_ref build only).  Stand-in for the one FLANN type src/matcher.cpp names
// (flann::Matrix<double>, a non-owning row-major view: flann/util/matrix.h).  The kd-tree itself lives behind
// Frame::SearchRadius, which ../mock/frame.h restates as the exhaustive radius search it is equivalent to.
#ifndef REF_STANDIN_FLANN_HPP_
#define REF_STANDIN_FLANN_HPP_
#include <cstddef>
namespace flann {
template <class T> class Matrix {
public:
    Matrix() : rows(0), cols(0), data_(nullptr) {}
    Matrix(T *data, size_t rows_, size_t cols_) : rows(rows_), cols(cols_), data_(data) {}
    T *operator[](size_t r) const { return data_ + r * cols; }
    T *ptr() const { return data_; }
    size_t rows, cols;
private:
    T *data_;
};
}  // namespace flann
#endif
