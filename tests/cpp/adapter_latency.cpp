// Latency of the reference-shaped calls through include/sfe_adapter.hpp, from C++ (no Python in the loop):
//   three calls : ORBextractor::extract(left), extract(right), StereoMatch(frame)   -- src/frame.cpp:47,388, src/pipeline.cpp:248
//   one call    : ORBextractor::extractStereo(left, right, ...)                     -- the same work in one round trip
// usage: adapter_latency left.raw right.raw w h [iterations]   -> prints microseconds per stereo pair, and checks that both
// ways return the same bytes.
#define SFE_ADAPTER_CV_STANDIN
#include "cv_standin.hpp"
#include "../../include/sfe_adapter.hpp"

#include <chrono>
#include <cstdio>
#include <fstream>

struct Camera { int w, h; int GetWidth() const { return w; } int GetHeight() const { return h; } };
struct Frame {
    std::vector<cv::KeyPoint> keypoints_, r_keypoints_;
    cv::Mat descriptions_, r_descriptions_;
    std::vector<int> stereo_correspond_;
    Camera cam;
    const std::vector<cv::KeyPoint> &GetKeypoints() const { return keypoints_; }
    const std::vector<cv::KeyPoint> &GetRightKeypoints() const { return r_keypoints_; }
    const cv::Mat GetDescription(int i) const { return descriptions_.row(i); }
    const cv::Mat GetRightDescription(int i) const { return r_descriptions_.row(i); }
    void SetStereoCorrespond(const std::vector<int> &c) { stereo_correspond_ = c; }
    const Camera *GetCamera() const { return &cam; }
};

int main(int argc, char **argv) {
    if (argc < 5) return 2;
    const int w = atoi(argv[3]), h = atoi(argv[4]), iters = argc > 5 ? atoi(argv[5]) : 300;
    std::vector<uint8_t> L((size_t)w * h), R((size_t)w * h);
    std::ifstream(argv[1], std::ios::binary).read((char *)L.data(), L.size());
    std::ifstream(argv[2], std::ios::binary).read((char *)R.data(), R.size());
    try {
        ORB_SLAM2::ORBextractor extractor(2000, 1.2f, 8, 20, 7);
        cv::Mat ml(h, w, CV_8UC1, L.data()), mr(h, w, CV_8UC1, R.data());
        Frame f, g;
        f.cam = g.cam = Camera{w, h};
        auto three = [&]() {
            extractor.extract(ml, cv::noArray(), f.keypoints_, f.descriptions_);
            extractor.extract(mr, cv::noArray(), f.r_keypoints_, f.r_descriptions_);
            sfe_adapter::StereoMatch(&f);
        };
        auto one = [&]() {
            extractor.extractStereo(ml, mr, g.keypoints_, g.descriptions_, g.r_keypoints_, g.r_descriptions_, g.stereo_correspond_);
        };
        for (int i = 0; i < 20; i++) { three(); one(); }
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < iters; i++) three();
        auto t1 = std::chrono::steady_clock::now();
        for (int i = 0; i < iters; i++) one();
        auto t2 = std::chrono::steady_clock::now();
        const bool same = f.keypoints_.size() == g.keypoints_.size() && f.stereo_correspond_ == g.stereo_correspond_ &&
                          !std::memcmp(f.keypoints_.data(), g.keypoints_.data(), f.keypoints_.size() * 28) &&
                          !std::memcmp(f.descriptions_.data, g.descriptions_.data, f.keypoints_.size() * 32) &&
                          f.r_keypoints_.size() == g.r_keypoints_.size() &&
                          !std::memcmp(f.r_descriptions_.data, g.r_descriptions_.data, f.r_keypoints_.size() * 32);
        std::printf("three_calls_us=%.1f one_call_us=%.1f same=%d nl=%zu nr=%zu\n",
                    std::chrono::duration<double, std::micro>(t1 - t0).count() / iters,
                    std::chrono::duration<double, std::micro>(t2 - t1).count() / iters, (int)same, f.keypoints_.size(), f.r_keypoints_.size());
        return same ? 0 : 1;
    } catch (const std::exception &e) {
        std::printf("exception: %s\n", e.what());
        return 1;
    }
}
