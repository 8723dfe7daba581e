// Drives include/sfe_adapter.hpp exactly the way the reference's frame.cpp / matcher.cpp callers do,
// with mock Frame / StereoFrame / Mappoint / SE3Quat classes exposing the reference's accessor names.
// usage: adapter_test left.raw right.raw w h  -> prints counts and FNV-1a checksums of the results
#define SFE_ADAPTER_CV_STANDIN
#include "cv_standin.hpp"
#include "../../include/sfe_adapter.hpp"

#include <array>
#include <cstdio>
#include <fstream>

struct Vec3 { double v[3]; double operator[](int i) const { return v[i]; } };
struct Quat { double x_, y_, z_, w_; double x() const { return x_; } double y() const { return y_; } double z() const { return z_; } double w() const { return w_; } };
struct SE3 { Quat r; Vec3 t; const Quat &rotation() const { return r; } const Vec3 &translation() const { return t; } };  // g2o::SE3Quat's accessors
struct KMat { double k[3][3]; double operator()(int r, int c) const { return k[r][c]; } };
struct DVec { double d[4]; double operator()(int i) const { return d[i]; } };
struct Camera {
    KMat K; DVec D; int w, h;
    const KMat &GetK() const { return K; } const DVec &GetD() const { return D; }
    int GetWidth() const { return w; } int GetHeight() const { return h; }
};
struct Mappoint {
    Vec3 X; cv::Mat desc;
    Vec3 GetXw() const { return X; } cv::Mat GetDescription() const { return desc; }
};
struct Frame {  // the accessors of reference include/frame.h:42-173 that the hot path uses
    std::vector<cv::KeyPoint> keypoints_, r_keypoints_;
    cv::Mat descriptions_, r_descriptions_;
    std::vector<int> stereo_correspond_;
    Camera cam;
    const std::vector<cv::KeyPoint> &GetKeypoints() const { return keypoints_; }
    const std::vector<cv::KeyPoint> &GetRightKeypoints() const { return r_keypoints_; }
    const cv::Mat GetDescription(int i) const { return descriptions_.row(i); }
    const cv::Mat GetRightDescription(int i) const { return r_descriptions_.row(i); }
    void SetStereoCorrespond(const std::vector<int> &c) { stereo_correspond_ = c; }
    int GetIndex(const Mappoint *) const { return -1; }
    const Camera *GetCamera() const { return &cam; }
};

static uint64_t fnv(const void *p, size_t n, uint64_t h = 1469598103934665603ull) {
    const uint8_t *b = (const uint8_t *)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

int main(int argc, char **argv) {
    if (argc < 5) return 2;
    const int w = atoi(argv[3]), h = atoi(argv[4]);
    std::vector<uint8_t> L((size_t)w * h), R((size_t)w * h);
    std::ifstream(argv[1], std::ios::binary).read((char *)L.data(), L.size());
    std::ifstream(argv[2], std::ios::binary).read((char *)R.data(), R.size());
    try {
        ORB_SLAM2::ORBextractor extractor(2000, 1.2f, 8, 20, 7);  // src/pipeline.cpp:46-50
        Frame f;
        f.cam = Camera{{{{718.856, 0, 607.1928}, {0, 718.856, 185.2157}, {0, 0, 1}}}, {{0, 0, 0, 0}}, w, h};
        extractor.extract(cv::Mat(h, w, CV_8UC1, L.data()), cv::noArray(), f.keypoints_, f.descriptions_);       // frame.cpp:47
        extractor.extract(cv::Mat(h, w, CV_8UC1, R.data()), cv::noArray(), f.r_keypoints_, f.r_descriptions_);   // frame.cpp:388
        sfe_adapter::StereoMatch(&f);                                                                             // pipeline.cpp:248
        // map points = stereo-matched keypoints back-projected with the KITTI intrinsics, identity pose
        std::vector<Mappoint> pts;
        std::set<Mappoint *> mps;
        for (size_t i = 0; i < f.keypoints_.size(); i++) {
            int j = f.stereo_correspond_[i];
            if (j < 0) continue;
            double dx = f.keypoints_[i].pt.x - f.r_keypoints_[j].pt.x;
            if (dx <= 0) continue;
            double z = 718.856 * 0.5371657 / dx;
            pts.push_back(Mappoint{{{(f.keypoints_[i].pt.x - 607.1928) / 718.856 * z, (f.keypoints_[i].pt.y - 185.2157) / 718.856 * z, z}},
                                   f.descriptions_.row((int)i)});
        }
        for (auto &p : pts) mps.insert(&p);
        SE3 T{{0, 0, 0, 1}, {{0, 0, 0}}};
        std::map<int, Mappoint *> m = sfe_adapter::ProjectionMatch(mps, T, &f, 50.);                              // posetracker.cpp:186
        size_t self = 0;
        for (auto &kv : m) self += kv.second->desc.data == f.descriptions_.ptr(kv.first);
        cv::Mat empty_desc; std::vector<cv::KeyPoint> none;
        extractor.extract(cv::Mat(), cv::noArray(), none, empty_desc);  // empty image: silent return
        if (argc > 5) {  // Frame::ComputeBoW, src/frame.cpp:419-427
            sfe_adapter::Vocabulary voc;
            if (!voc.loadFromTextFile(argv[5])) throw std::runtime_error("vocabulary text file rejected");
            std::vector<cv::Mat> vdesc;
            for (int j = 0; j < f.descriptions_.rows; j++) vdesc.push_back(f.descriptions_.row(j));
            std::map<unsigned, double> bowvec;
            std::map<unsigned, std::vector<unsigned>> featvec;
            voc.transform(vdesc, bowvec, featvec, 4);
            uint64_t hb = 1469598103934665603ull, hf = hb;
            for (auto &kv : bowvec) { hb = fnv(&kv.first, 4, hb); hb = fnv(&kv.second, 8, hb); }
            for (auto &kv : featvec) { hf = fnv(&kv.first, 4, hf); hf = fnv(kv.second.data(), kv.second.size() * 4, hf); }
            printf("bow=%016llx bown=%zu fv=%016llx ", (unsigned long long)hb, bowvec.size(), (unsigned long long)hf);
        }
        printf("nl=%zu nr=%zu kps=%016llx desc=%016llx stereo=%016llx proj=%zu self=%zu dd=%d\n", f.keypoints_.size(),
               f.r_keypoints_.size(), (unsigned long long)fnv(f.keypoints_.data(), f.keypoints_.size() * 28),
               (unsigned long long)fnv(f.descriptions_.data, (size_t)f.descriptions_.rows * 32),
               (unsigned long long)fnv(f.stereo_correspond_.data(), f.stereo_correspond_.size() * 4), m.size(), self,
               ORB_SLAM2::ORBextractor::DescriptorDistance(f.descriptions_.row(0), f.descriptions_.row(1)));
    } catch (const std::exception &e) {
        printf("exception: %s\n", e.what());
        return 1;
    }
    return 0;
}
