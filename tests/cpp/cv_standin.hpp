// Minimal stand-ins for the OpenCV types include/sfe_adapter.hpp touches, for boxes without OpenCV
// C++ headers.  Layout-compatible cv::KeyPoint (28 B POD), a ref-counted-less cv::Mat with the members
// the adapter and the reference's callers use (data, rows, cols, step, type(), empty(), ptr(i), row(i),
// create(), release()).  InputArray / OutputArray collapse to (const) Mat&.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
#define CV_8U 0
#define CV_8UC1 0
namespace cv {
struct Point2f { float x, y; };
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id; };
class Mat {
public:
    Mat() {}
    Mat(int r, int c, int /*type*/, void *d, size_t s = 0) : data((uint8_t *)d), rows(r), cols(c), step(s ? s : (size_t)c) {}
    void create(int r, int c, int /*type*/) {
        own = std::shared_ptr<uint8_t>(new uint8_t[(size_t)r * c], std::default_delete<uint8_t[]>());
        data = own.get(); rows = r; cols = c; step = (size_t)c;
    }
    void release() { own.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    uint8_t *ptr(int i = 0) const { return data + (size_t)i * step; }
    Mat row(int i) const { Mat m(1, cols, CV_8U, ptr(i), step); m.own = own; return m; }
    Mat getMat() const { return *this; }
    uint8_t *data = nullptr;
    int rows = 0, cols = 0;
    size_t step = 0;
private:
    std::shared_ptr<uint8_t> own;
};
typedef const Mat &InputArray;
typedef Mat &OutputArray;
inline Mat noArray() { return Mat(); }
}  // namespace cv
