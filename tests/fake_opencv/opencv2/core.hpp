// Minimal stand-in for <opencv2/core.hpp>, used ONLY by tests/test_cpp_shim.py to compile
// include/video/Stabilizer.h in a container without OpenCV headers.  It provides the
// handful of cv::Mat / cv::Rect members the shim touches; it is not part of the product.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>
#define CV_8UC3 16
namespace cv {
struct Rect {
    int x = 0, y = 0, width = 0, height = 0;
    Rect() = default;
    Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {}
};
class Mat {
public:
    int rows = 0, cols = 0;
    uint8_t* data = nullptr;
    size_t step = 0;
    Mat() = default;
    Mat(int r, int c, int type) : rows(r), cols(c), step((size_t)c * 3), type_(type), buf_(std::make_shared<std::vector<uint8_t>>((size_t)r * c * 3)) {
        data = buf_->data();
    }
    int type() const { return type_; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    Mat operator()(const Rect& r) const {
        Mat m = *this;
        m.rows = r.height; m.cols = r.width;
        m.data = data + (size_t)r.y * step + (size_t)r.x * 3;
        return m;
    }
private:
    int type_ = CV_8UC3;
    std::shared_ptr<std::vector<uint8_t>> buf_;
};
}  // namespace cv
