// TEST INFRASTRUCTURE (oracle/_ref build only).  The OpenCV entry points src/orb_extractor.cpp calls
// (FAST :809,814; resize :1120; copyMakeBorder :1122,1127; GaussianBlur :1086; fastAtan2 :103), implemented by the
// closed-form models of oracle/orb_oracle.c that tests/test_oracle_vs_cv2.py pins bit-for-bit against cv2.
// Anything outside what the reference passes (other kernel sizes, types, border modes) aborts instead of guessing.
#include "opencv2/core/core.hpp"

#include <cstdio>

#include "../orb_oracle.h"

namespace cv {

static void need(bool ok, const char *what) {
    if (!ok) {
        std::fprintf(stderr, "cv stand-in: unsupported use: %s\n", what);
        std::abort();
    }
}

const _OutputArray &noArray() {
    static _OutputArray none;
    return none;
}

static thread_local std::vector<FastCall> t_log;
static thread_local std::vector<Mat> t_blur_log;
static thread_local bool t_log_on = false;
std::vector<FastCall> &fastCallLog() { return t_log; }
std::vector<Mat> &blurCallLog() { return t_blur_log; }
void fastCallLogEnable(bool on) { t_log_on = on; }

// cv::FAST(image, keypoints, threshold, nonmax) = FastFeatureDetector TYPE_9_16: keypoints in raster order,
// KeyPoint(x, y, 7.f, -1, score) (features2d/src/fast.cpp, FAST_t<16>)
void FAST(InputArray image, std::vector<KeyPoint> &keypoints, int threshold, bool nonmaxSuppression) {
    const Mat img = image.getMat();
    need(img.type() == CV_8UC1 && nonmaxSuppression, "FAST on 8-bit gray with non-max suppression");
    keypoints.clear();
    if (!img.empty()) {
        int cap = img.rows * img.cols / 4 + 16;
        std::vector<float> xyr((size_t)cap * 3);
        const int n = orc_fast_nms(img.data, img.cols, img.rows, (int)(size_t)img.step, threshold, xyr.data(), cap);
        need(n <= cap, "FAST result capacity");
        for (int i = 0; i < n; i++) keypoints.push_back(KeyPoint(xyr[3 * i], xyr[3 * i + 1], 7.f, -1, xyr[3 * i + 2]));
    }
    if (t_log_on) {
        FastCall c = {img.data, (size_t)img.step, img.cols, img.rows, threshold, keypoints};
        t_log.push_back(c);
    }
}

void resize(InputArray src_, OutputArray dst_, Size dsize, double fx, double fy, int interpolation) {
    const Mat src = src_.getMat();
    need(src.type() == CV_8UC1 && interpolation == INTER_LINEAR && fx == 0 && fy == 0 && dsize.width > 0 && dsize.height > 0,
         "resize(8-bit gray, dsize, INTER_LINEAR)");
    dst_.create(dsize, src.type());  // keeps an ROI of matching size (mvImagePyramid[level] is a view into `temp`)
    Mat &dst = dst_.getMatRef();
    orc_resize_linear_u8(src.data, src.cols, src.rows, (int)(size_t)src.step, dst.data, dst.cols, dst.rows, (int)(size_t)dst.step);
}

int borderInterpolate(int p, int len, int borderType) {
    need((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101, "BORDER_REFLECT_101");
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// copyMakeBorder: dst(y, x) = src(reflect101(y - top), reflect101(x - left)).  Works when src is the interior view of
// dst (src/orb_extractor.cpp:1122): interior pixels map to themselves and every border pixel reads an interior one.
// Without BORDER_ISOLATED OpenCV would read real pixels around a sub-matrix src; the reference's level-0 input is
// the caller's whole image, so that case does not arise here.
void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int borderType) {
    const Mat src = src_.getMat();
    need(src.type() == CV_8UC1, "copyMakeBorder on 8-bit gray");
    dst_.create(src.rows + top + bottom, src.cols + left + right, src.type());
    Mat &dst = dst_.getMatRef();
    for (int y = 0; y < dst.rows; y++) {
        const uchar *s = src.ptr(borderInterpolate(y - top, src.rows, borderType));
        uchar *d = dst.ptr(y);
        if (d + left != s) std::memmove(d + left, s, (size_t)src.cols);
        for (int x = 0; x < left; x++) d[x] = s[borderInterpolate(x - left, src.cols, borderType)];
        for (int x = 0; x < right; x++) d[left + src.cols + x] = s[borderInterpolate(src.cols + x, src.cols, borderType)];
    }
}

void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sigmaX, double sigmaY, int borderType) {
    const Mat src = src_.getMat();
    need(src.type() == CV_8UC1 && ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2 && borderType == BORDER_REFLECT_101,
         "GaussianBlur(8-bit gray, 7x7, sigma 2, BORDER_REFLECT_101)");
    const Mat in = src.clone();  // the reference blurs in place
    dst_.create(src.rows, src.cols, src.type());
    Mat &dst = dst_.getMatRef();
    orc_gaussian7_s2_u8(in.data, in.cols, in.rows, (int)(size_t)in.step, dst.data, (int)(size_t)dst.step);
    if (t_log_on) t_blur_log.push_back(dst.clone());
}

float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

// KeyPointsFilter::retainBest (features2d/src/keypoint.cpp): keep the npoints strongest responses plus ties with the
// weakest kept one.  Only ComputeKeyPointsOld calls it and extract() never calls that (src/orb_extractor.cpp:1057).
void KeyPointsFilter::retainBest(std::vector<KeyPoint> &keypoints, int npoints) {
    if (npoints < 0 || (int)keypoints.size() <= npoints) return;
    if (npoints == 0) { keypoints.clear(); return; }
    std::nth_element(keypoints.begin(), keypoints.begin() + npoints - 1, keypoints.end(),
                     [](const KeyPoint &a, const KeyPoint &b) { return a.response > b.response; });
    const float ambiguous = keypoints[(size_t)npoints - 1].response;
    std::vector<KeyPoint>::iterator end = std::partition(keypoints.begin() + npoints, keypoints.end(),
                                                         [ambiguous](const KeyPoint &k) { return k.response >= ambiguous; });
    keypoints.resize((size_t)(end - keypoints.begin()));
}

}  // namespace cv
