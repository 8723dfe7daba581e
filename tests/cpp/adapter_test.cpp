// Drives include/sfe_adapter.hpp exactly the way the reference's frame.cpp / matcher.cpp callers do,
// with mock Frame / StereoFrame / Mappoint / SE3Quat classes exposing the reference's accessor names.
// usage: adapter_test left.raw right.raw w h  -> prints counts and FNV-1a checksums of the results
#define SFE_ADAPTER_CV_STANDIN
#include "cv_standin.hpp"
#include "../../include/sfe_adapter.hpp"

#include <algorithm>
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
        // ---- the multi-GPU entry points of include/sfe.h from C++: one process drives every visible GPU (up to 4); the
        // sharded results must equal the unsharded calls on GPU 0
        {
            using sfe_adapter::check;
            int ndev = 0;
            check(sfe_device_count(&ndev), "sfe_device_count");
            const int nd = std::min(ndev, 4);
            std::vector<int> devs(nd);
            for (int i = 0; i < nd; i++) devs[i] = i;
            std::vector<sfe_comm *> comms(nd, nullptr);
            check(sfe_comm_create_local(devs.data(), nd, comms.data()), "sfe_comm_create_local");
            std::vector<sfe_matcher *> ms(nd, nullptr);
            for (int i = 0; i < nd; i++) check(sfe_matcher_create(i, &ms[i]), "sfe_matcher_create");
            // brute-force top-2: 100003 pseudo-random rows, 64 queries = rows with a few flipped bits
            const int64_t rows = 100003;
            const int nq = 64;
            std::vector<uint8_t> db((size_t)rows * 32), qs((size_t)nq * 32);
            uint64_t x = 88172645463325252ull;
            for (auto &b : db) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; b = (uint8_t)(x >> 24); }
            for (int i = 0; i < nq; i++) {
                std::memcpy(&qs[(size_t)i * 32], &db[(size_t)(i * 1543 % rows) * 32], 32);
                for (int k = 0; k < i % 9; k++) qs[(size_t)i * 32 + (k * 7) % 32] ^= (uint8_t)(1u << (k % 8));
            }
            sfe_db *whole = nullptr;
            check(sfe_db_create(ms[0], db.data(), rows, 0, &whole), "sfe_db_create");
            std::vector<int32_t> want((size_t)nq * 4), got((size_t)nq * 4);
            check(sfe_knn2(ms[0], whole, qs.data(), nq, want.data()), "sfe_knn2");
            std::vector<sfe_db *> shard(nd, nullptr);
            std::vector<void *> dq(nd, nullptr), dout(nd, nullptr);
            for (int i = 0; i < nd; i++) {
                const int64_t a = rows * i / nd, b = rows * (i + 1) / nd;
                check(sfe_db_create(ms[i], db.data() + (size_t)a * 32, b - a, a, &shard[i]), "sfe_db_create(shard)");
                check(sfe_device_alloc(i, &dq[i], qs.size()), "sfe_device_alloc");
                check(sfe_device_alloc(i, &dout[i], got.size() * 4), "sfe_device_alloc");
                check(sfe_copy_to_device(i, dq[i], qs.data(), qs.size()), "sfe_copy_to_device");
            }
            for (int i = 0; i < nd; i++) check(sfe_knn2_sharded(ms[i], comms[i], shard[i], (const uint8_t *)dq[i], nq, (int32_t *)dout[i]), "sfe_knn2_sharded");
            bool knn_ok = true;
            for (int i = 0; i < nd; i++) {
                check(sfe_matcher_wait(ms[i]), "sfe_matcher_wait");
                check(sfe_copy_to_host(i, got.data(), dout[i], got.size() * 4), "sfe_copy_to_host");
                knn_ok = knn_ok && got == want;
            }
            // ProjectionMatch of the same map points, sharded by points, against the resident frame
            std::vector<double> xw;
            std::vector<uint8_t> mdesc;
            for (auto &p : pts) {
                xw.push_back(p.X[0]); xw.push_back(p.X[1]); xw.push_back(p.X[2]);
                mdesc.insert(mdesc.end(), p.desc.data, p.desc.data + 32);
            }
            const int np = (int)pts.size(), mk = (int)f.keypoints_.size();
            const sfe_se3 ident = {0, 0, 0, 1, 0, 0, 0};
            sfe_camera cam = {718.856, 718.856, 607.1928, 185.2157, {0, 0, 0, 0}, w, h};
            std::vector<uint8_t> kdesc((size_t)mk * 32);
            for (int i = 0; i < mk; i++) std::memcpy(&kdesc[(size_t)i * 32], f.descriptions_.ptr(i), 32);
            std::vector<int32_t> pwant(mk), pgot(mk);
            check(sfe_projection_match_se3(ms[0], xw.data(), mdesc.data(), nullptr, np, &ident, &cam, (const sfe_keypoint *)f.keypoints_.data(),
                                           kdesc.data(), mk, 50., 0.5, pwant.data(), nullptr), "sfe_projection_match_se3");
            bool proj_ok = true;
            std::vector<sfe_frame *> fr(nd, nullptr);
            std::vector<void *> dx(nd, nullptr), dd(nd, nullptr), dt(nd, nullptr);
            for (int i = 0; i < nd; i++) {
                const int a = (int)((int64_t)np * i / nd), b = (int)((int64_t)np * (i + 1) / nd);
                check(sfe_frame_create(ms[i], (const sfe_keypoint *)f.keypoints_.data(), kdesc.data(), mk, &cam, &fr[i]), "sfe_frame_create");
                check(sfe_device_alloc(i, &dx[i], (size_t)std::max(b - a, 1) * 24), "sfe_device_alloc");
                check(sfe_device_alloc(i, &dd[i], (size_t)std::max(b - a, 1) * 32), "sfe_device_alloc");
                check(sfe_device_alloc(i, &dt[i], (size_t)mk * 4), "sfe_device_alloc");
                if (b > a) {
                    check(sfe_copy_to_device(i, dx[i], xw.data() + (size_t)a * 3, (size_t)(b - a) * 24), "sfe_copy_to_device");
                    check(sfe_copy_to_device(i, dd[i], mdesc.data() + (size_t)a * 32, (size_t)(b - a) * 32), "sfe_copy_to_device");
                }
            }
            for (int i = 0; i < nd; i++) {
                const int a = (int)((int64_t)np * i / nd), b = (int)((int64_t)np * (i + 1) / nd);
                check(sfe_projection_match_sharded(ms[i], comms[i], fr[i], (const double *)dx[i], (const uint8_t *)dd[i], nullptr, b - a, a, &ident,
                                                   50., 0.5, (int32_t *)dt[i], nullptr), "sfe_projection_match_sharded");
            }
            for (int i = 0; i < nd; i++) {
                check(sfe_matcher_wait(ms[i]), "sfe_matcher_wait");
                check(sfe_copy_to_host(i, pgot.data(), dt[i], (size_t)mk * 4), "sfe_copy_to_host");
                proj_ok = proj_ok && pgot == pwant;
                check(sfe_comm_status(comms[i], nullptr), "sfe_comm_status");
            }
            printf("sharded_gpus=%d sharded_knn=%s sharded_proj=%s ", nd, knn_ok ? "ok" : "DIFF", proj_ok ? "ok" : "DIFF");
            for (int i = 0; i < nd; i++) {
                sfe_frame_destroy(fr[i]); sfe_db_destroy(shard[i]);
                sfe_device_free(i, dq[i]); sfe_device_free(i, dout[i]); sfe_device_free(i, dx[i]); sfe_device_free(i, dd[i]); sfe_device_free(i, dt[i]);
                sfe_comm_destroy(comms[i]); sfe_matcher_destroy(ms[i]);
            }
            sfe_db_destroy(whole);
            if (!knn_ok || !proj_ok) throw std::runtime_error("sharded entry points differ from the unsharded ones");
        }
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
