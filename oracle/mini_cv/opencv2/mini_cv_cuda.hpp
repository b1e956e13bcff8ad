// mini_cv_cuda.hpp — the cv::cuda:: slice that /root/reference/src/RollCorrection.cpp and AutoZoomCrop.cpp use.
// TEST INFRASTRUCTURE ONLY.  Those two reference components are GPU-only (cv::cuda::*, no CPU branch).  This container
// has neither OpenCV's CUDA modules nor a GPU, so each cv::cuda:: call is served by the CPU function of the same OpenCV
// build (the cv2 wheel) through the mini_cv callback table:
//     cuda::resize -> cv::resize          cuda::cvtColor -> cv::cvtColor        cuda::threshold -> cv::threshold
//     CannyEdgeDetector -> cv::Canny      HoughLinesDetector -> cv::HoughLines  cuda::remap -> cv::remap
//     createMorphologyFilter -> cv::morphologyEx                                cuda::warpAffine -> cv::warpAffine
// cuda::buildWarpAffineMaps has no CPU twin; it is restated here from OpenCV's cudawarping sources (invertAffineTransform in
// double, coefficients narrowed to float, map = c0*x + c1*y + c2 in float).  Known residuals against a real CUDA OpenCV:
// its resize / remap / warpAffine interpolate in float (the CPU functions in fixed point), its Hough transform returns
// the lines in the order the atomics landed (the CPU function sorts by votes), and nvcc may contract the map arithmetic into
// FMAs.  DESIGN.md section 2 states this residual.
#ifndef MINI_CV_CUDA_HPP
#define MINI_CV_CUDA_HPP

#include "opencv.hpp"
#include <iostream>

namespace cv {

inline Mat getStructuringElement(int shape, Size ksize) {
    Mat k(ksize.height, ksize.width, CV_8UC1);
    mini_cv_mat d = k.view();
    mini_cv_check(ops()->structuring_element(shape, &d), "getStructuringElement");
    return k;
}
inline Mat getRotationMatrix2D(Point2f center, double angle, double scale) {
    double m[6];
    mini_cv_check(ops()->rotation_matrix(center.x, center.y, angle, scale, m), "getRotationMatrix2D");
    Mat M(2, 3, CV_64FC1);
    for (int i = 0; i < 6; i++) M.at<double>(i / 3, i % 3) = m[i];
    return M;
}
inline void invertAffineTransform(const Mat &M, Mat &iM) {
    // imgwarp.cpp, double branch
    double m[6];
    for (int i = 0; i < 6; i++) m[i] = M.getElem(i / 3, i % 3);
    double D = m[0] * m[4] - m[1] * m[3];
    D = D != 0 ? 1. / D : 0;
    double A11 = m[4] * D, A22 = m[0] * D, A12 = -m[1] * D, A21 = -m[3] * D;
    double b1 = -A11 * m[2] - A12 * m[5];
    double b2 = -A21 * m[2] - A22 * m[5];
    iM.create(2, 3, CV_64FC1);
    double r[6] = {A11, A12, b1, A21, A22, b2};
    for (int i = 0; i < 6; i++) iM.at<double>(i / 3, i % 3) = r[i];
}
inline void drawContours(Mat &image, const std::vector<std::vector<Point>> &contours, int idx, const Scalar &color, int thickness = 1) {
    std::vector<int> pts, lens;
    for (auto &c : contours) {
        lens.push_back((int)c.size());
        for (auto &p : c) { pts.push_back(p.x); pts.push_back(p.y); }
    }
    mini_cv_mat d = image.view();
    mini_cv_check(ops()->draw_contours(&d, pts.data(), lens.data(), (int)contours.size(), idx, color.val, thickness), "drawContours");
}

namespace cuda {

class Stream {
public:
    void waitForCompletion() {}
    static Stream &Null() { static Stream s; return s; }
};

class GpuMat : public Mat {
public:
    GpuMat() {}
    GpuMat(const Mat &m) : Mat(m) {}
    GpuMat(int r, int c, int type) : Mat(r, c, type) {}
    void upload(const Mat &m) { m.copyTo(*this); }
    void upload(const Mat &m, Stream &) { m.copyTo(*this); }
    void download(Mat &m) const { Mat::copyTo(m); }
    void download(Mat &m, Stream &) const { Mat::copyTo(m); }
    GpuMat operator()(const Rect &roi) const { return GpuMat(Mat(*this, roi)); }
    GpuMat clone() const { return GpuMat(Mat::clone()); }
};

inline void resize(const GpuMat &src, GpuMat &dst, Size dsize, double fx = 0, double fy = 0, int interp = INTER_LINEAR, Stream & = Stream::Null()) {
    Mat d;
    cv::resize(src, d, dsize, fx, fy, interp);
    dst = GpuMat(d);
}
inline void cvtColor(const GpuMat &src, GpuMat &dst, int code, int dcn = 0, Stream & = Stream::Null()) {
    Mat d;
    cv::cvtColor(src, d, code, dcn);
    dst = GpuMat(d);
}
inline double threshold(const GpuMat &src, GpuMat &dst, double thresh, double maxval, int type, Stream & = Stream::Null()) {
    Mat d;
    cv::threshold(src, d, thresh, maxval, type);
    dst = GpuMat(d);
    return thresh;
}

class CannyEdgeDetector {
public:
    CannyEdgeDetector(double lo, double hi, int ap, bool l2) : low(lo), high(hi), aperture(ap), l2grad(l2) {}
    void detect(const GpuMat &image, GpuMat &edges, Stream & = Stream::Null()) {
        if (image.empty()) throw Exception("Canny: empty image");
        Mat out(image.rows, image.cols, CV_8UC1);
        mini_cv_mat s = image.view(), d = out.view();
        mini_cv_check(ops()->canny(&s, &d, low, high, aperture, l2grad ? 1 : 0), "Canny");
        edges = GpuMat(out);
    }
    double low, high;
    int aperture;
    bool l2grad;
};
inline Ptr<CannyEdgeDetector> createCannyEdgeDetector(double low, double high, int aperture = 3, bool L2gradient = false) {
    return std::make_shared<CannyEdgeDetector>(low, high, aperture, L2gradient);
}

// cv::cuda::HoughLinesDetector::detect returns a 2 x N CV_32FC2 matrix: row 0 = (rho, theta), row 1 = the votes (one int in the
// first 4 bytes of each element).  The reference reads linesMat.total() Vec2f's from it, i.e. BOTH rows; the vote row decodes
// to (tiny denormal, 0) -> angleDeg = -90, which its angle filter drops unless angleFilterMin <= -90.  Reproduced here.
class HoughLinesDetector {
public:
    HoughLinesDetector(float r, float t, int thr, bool sort, int maxl) : rho(r), theta(t), threshold(thr), doSort(sort), maxLines(maxl) {}
    void detect(const GpuMat &edges, GpuMat &lines, Stream & = Stream::Null()) {
        std::vector<float> rt((size_t)maxLines * 2);
        std::vector<int> votes(maxLines);
        int n = 0;
        mini_cv_mat e = edges.view();
        mini_cv_check(ops()->hough_lines(&e, rho, theta, threshold, maxLines, rt.data(), votes.data(), maxLines, &n), "HoughLines");
        if (n == 0) { lines.release(); return; }
        Mat out = Mat::zeros(2, n, CV_32FC2);
        for (int i = 0; i < n; i++) {
            out.at<Vec2f>(0, i) = Vec2f(rt[2 * i], rt[2 * i + 1]);
            std::memcpy(out.ptr(1) + (size_t)i * 8, &votes[i], 4);
        }
        lines = GpuMat(out);
    }
    float rho, theta;
    int threshold;
    bool doSort;
    int maxLines;
};
inline Ptr<HoughLinesDetector> createHoughLinesDetector(float rho, float theta, int threshold, bool doSort = false, int maxLines = 4096) {
    return std::make_shared<HoughLinesDetector>(rho, theta, threshold, doSort, maxLines);
}

inline void buildWarpAffineMaps(const Mat &M, bool inverse, Size dsize, GpuMat &xmap, GpuMat &ymap, Stream & = Stream::Null()) {
    if (M.rows != 2 || M.cols != 3) throw Exception("buildWarpAffineMaps: M must be 2x3");
    float c[6];
    Mat src = M;
    if (!inverse) { Mat iM; invertAffineTransform(M, iM); src = iM; }
    for (int i = 0; i < 6; i++) c[i] = (float)src.getElem(i / 3, i % 3);
    Mat mx(dsize.height, dsize.width, CV_32FC1), my(dsize.height, dsize.width, CV_32FC1);
    for (int y = 0; y < dsize.height; y++) {
        float *px = mx.ptr<float>(y), *py = my.ptr<float>(y);
        const float fy = (float)y;
        for (int x = 0; x < dsize.width; x++) {
            const float fx = (float)x;
            volatile float a = c[0] * fx, b = c[1] * fy, d = c[3] * fx, e = c[4] * fy;   // no contraction: each product rounded
            px[x] = (a + b) + c[2];
            py[x] = (d + e) + c[5];
        }
    }
    xmap = GpuMat(mx);
    ymap = GpuMat(my);
}
inline void remap(const GpuMat &src, GpuMat &dst, const GpuMat &xmap, const GpuMat &ymap, int interp, int borderMode = BORDER_CONSTANT,
                  Scalar = Scalar(), Stream & = Stream::Null()) {
    Mat out(xmap.rows, xmap.cols, src.type());
    mini_cv_mat s = src.view(), d = out.view(), mx = xmap.view(), my = ymap.view();
    mini_cv_check(ops()->remap(&s, &d, &mx, &my, interp, borderMode), "remap");
    dst = GpuMat(out);
}
inline void warpAffine(const GpuMat &src, GpuMat &dst, const Mat &M, Size dsize, int flags = INTER_LINEAR, int borderMode = BORDER_CONSTANT,
                       Scalar borderValue = Scalar(), Stream & = Stream::Null()) {
    Mat d;
    cv::warpAffine(src, d, M, dsize, flags, borderMode, borderValue);
    dst = GpuMat(d);
}

class Filter {
public:
    Filter(int o, int sh, Mat k) : op(o), kernel(k) { (void)sh; }
    void apply(const GpuMat &src, GpuMat &dst, Stream & = Stream::Null()) {
        Mat out(src.rows, src.cols, src.type());
        mini_cv_mat s = src.view(), d = out.view(), k = kernel.view();
        mini_cv_check(ops()->morphology(&s, &d, op, &k), "morphologyEx");
        dst = GpuMat(out);
    }
    int op;
    Mat kernel;
};
inline Ptr<Filter> createMorphologyFilter(int op, int srcType, const Mat &kernel) { return std::make_shared<Filter>(op, srcType, kernel); }

}  // namespace cuda
}  // namespace cv
#endif
