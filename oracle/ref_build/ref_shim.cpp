// TEST INFRASTRUCTURE (oracle/_ref build only).  C entry points around the reference's OWN translation units
// (/root/reference/src/orb_extractor.cpp, matcher.cpp, camera.cpp, compiled unmodified by ./Makefile) so that the
// tests can run the reference's code on the same inputs as oracle/orb_oracle.c and the CUDA path.
//   * ORBextractor is driven through its public surface (constructor, extract, getters, mvImagePyramid) and, for the
//     quadtree alone, through a derived class that re-exports the protected DistributeOctTree;
//   * StereoMatch / ProjectionMatch are the reference's functions, fed by the mock Frame / Mappoint of ./mock and the
//     reference's real Camera;
//   * heap mode: the reference orders equally-full quadtree nodes by their HEAP ADDRESS (orb_extractor.cpp:684).
//     REF_HEAP_MONOTONIC serves std::list<ExtractorNode> nodes from a bump arena, so address order = creation order
//     (which is the oracle's declared rule T1); REF_HEAP_MALLOC leaves them to glibc malloc, i.e. whatever the
//     reference does on this machine.  Bump mode is not thread-safe (tests are single-threaded).
#include <sys/mman.h>

#include <atomic>
#include <thread>
#include <cstdio>
#include <new>

#include "camera.h"
#include "frame.h"
#include "mappoint.h"
#include "matcher.h"
#include "orb_extractor.h"

bool Frame::IsInFrame(const Eigen::Vector2d &uv) const { return camera_->IsInImage(uv); }  // src/frame.cpp:465-467

// ---- heap control --------------------------------------------------------------------------------------------------
namespace {
const size_t kNodeBytes = sizeof(std::_List_node<ORB_SLAM2::ExtractorNode>);
const size_t kArenaBytes = (size_t)1 << 30;  // address space only (MAP_NORESERVE); an extraction touches a few MB
char *g_arena = nullptr;
std::atomic<size_t> g_used(0);
std::atomic<int> g_monotonic(0);
std::atomic<long> g_arena_allocs(0);

void *arena_alloc(size_t n) {
    if (!g_arena) {
        void *p = mmap(nullptr, kArenaBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) std::abort();
        g_arena = (char *)p;
    }
    n = (n + 15) & ~(size_t)15;
    const size_t at = g_used.fetch_add(n);
    if (at + n > kArenaBytes) {
        std::fprintf(stderr, "_ref: bump arena exhausted\n");
        std::abort();
    }
    g_arena_allocs++;
    return g_arena + at;
}
inline bool in_arena(const void *p) { return g_arena && (const char *)p >= g_arena && (const char *)p < g_arena + kArenaBytes; }
void arena_reset() {
    if (g_arena && g_used.load() > 0) madvise(g_arena, (g_used.load() + 4095) & ~(size_t)4095, MADV_DONTNEED);
    g_used = 0;
}
}  // namespace

void *operator new(size_t n) {
    if (g_monotonic.load(std::memory_order_relaxed) && n == kNodeBytes) return arena_alloc(n);
    void *p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void *operator new[](size_t n) {
    void *p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void operator delete(void *p) noexcept { if (p && !in_arena(p)) std::free(p); }
void operator delete[](void *p) noexcept { if (p && !in_arena(p)) std::free(p); }
void operator delete(void *p, size_t) noexcept { if (p && !in_arena(p)) std::free(p); }
void operator delete[](void *p, size_t) noexcept { if (p && !in_arena(p)) std::free(p); }

// ---- extractor -----------------------------------------------------------------------------------------------------
namespace {
struct Tap : public ORB_SLAM2::ORBextractor {
    Tap(int n, float s, int l, int i, int m) : ORB_SLAM2::ORBextractor(n, s, l, i, m) {}
    using ORB_SLAM2::ORBextractor::DistributeOctTree;
    using ORB_SLAM2::ORBextractor::mnFeaturesPerLevel;
    using ORB_SLAM2::ORBextractor::umax;
    std::vector<std::vector<float> > cands;  // per level, x y response (window-relative) of the last extract
    std::vector<cv::Mat> blurred;            // per level, GaussianBlur output of the last extract (empty = not blurred)
};
const int kEdge = 19;  // EDGE_THRESHOLD, src/orb_extractor.cpp:74
}  // namespace

namespace cv {  // blur outputs are local to extract(); cv_standin.cpp's GaussianBlur reports them here
std::vector<Mat> &blurCallLog();
}

extern "C" {

typedef struct { float x, y, size, angle, response; int32_t octave, class_id; } ref_keypoint;
typedef struct { double fx, fy, cx, cy; double d[4]; int width, height; } ref_camera;
static_assert(sizeof(ref_keypoint) == sizeof(cv::KeyPoint), "cv::KeyPoint layout");

enum { REF_HEAP_MALLOC = 0, REF_HEAP_MONOTONIC = 1 };
void ref_set_heap_mode(int mode) { g_monotonic = mode == REF_HEAP_MONOTONIC; }
long ref_arena_allocations(void) { return g_arena_allocs.load(); }
int ref_list_node_bytes(void) { return (int)kNodeBytes; }

void *ref_extractor_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th) {
    return new Tap(nfeatures, scale_factor, nlevels, ini_th, min_th);
}
void ref_extractor_destroy(void *h) { delete (Tap *)h; }

void ref_extractor_tables(void *h, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2, int *per_level, int *umax16) {
    Tap *ex = (Tap *)h;
    const std::vector<float> a = ex->GetScaleFactors(), b = ex->GetInverseScaleFactors(), c = ex->GetScaleSigmaSquares(),
                             d = ex->GetInverseScaleSigmaSquares();
    for (int i = 0; i < ex->GetLevels(); i++) {
        if (scale) scale[i] = a[i];
        if (inv_scale) inv_scale[i] = b[i];
        if (sigma2) sigma2[i] = c[i];
        if (inv_sigma2) inv_sigma2[i] = d[i];
        if (per_level) per_level[i] = ex->mnFeaturesPerLevel[i];
    }
    if (umax16) for (int i = 0; i < 16; i++) umax16[i] = ex->umax[i];
}

// ORBextractor::extract on one image, exactly as src/frame.cpp:47 calls it.  Returns the keypoint count, or -1 when
// more than cap.
int ref_extract(void *h, const uint8_t *img, int w, int hgt, int stride, ref_keypoint *kps, uint8_t *desc, int cap) {
    Tap *ex = (Tap *)h;
    if (g_monotonic.load()) arena_reset();
    cv::fastCallLog().clear();
    cv::blurCallLog().clear();
    cv::fastCallLogEnable(true);
    cv::Mat image(hgt, w, CV_8UC1, (void *)img, (size_t)stride);
    std::vector<cv::KeyPoint> keypoints;
    cv::Mat descriptors;
    ex->extract(image, cv::noArray(), keypoints, descriptors);
    cv::fastCallLogEnable(false);
    // sort the recorded cv::FAST calls into levels by the buffer they read
    const int L = ex->GetLevels();
    ex->cands.assign((size_t)L, std::vector<float>());
    for (const cv::FastCall &c : cv::fastCallLog()) {
        for (int l = 0; l < L; l++) {
            const cv::Mat &P = ex->mvImagePyramid[l];
            if (P.empty()) continue;
            const size_t step = P.step;
            const uchar *lo = P.data - (size_t)kEdge * step - kEdge, *hi = lo + (size_t)(P.rows + 2 * kEdge) * step;
            if (c.data < lo || c.data >= hi) continue;
            const ptrdiff_t off = c.data - P.data;
            const int y0 = (int)(off / (ptrdiff_t)step), x0 = (int)(off % (ptrdiff_t)step);
            for (const cv::KeyPoint &k : c.keypoints) {  // + (j*wCell, i*hCell) = + (cell origin - 16): orb_extractor.cpp:822-823
                ex->cands[l].push_back(k.pt.x + (float)(x0 - (kEdge - 3)));
                ex->cands[l].push_back(k.pt.y + (float)(y0 - (kEdge - 3)));
                ex->cands[l].push_back(k.response);
            }
            break;
        }
    }
    ex->blurred.assign((size_t)L, cv::Mat());
    {
        size_t next = 0;
        for (int l = 0; l < L && next < cv::blurCallLog().size(); l++) {
            const cv::Mat &b = cv::blurCallLog()[next];
            if (!ex->mvImagePyramid[l].empty() && b.rows == ex->mvImagePyramid[l].rows && b.cols == ex->mvImagePyramid[l].cols) {
                bool has = false;
                for (const cv::KeyPoint &k : keypoints) has = has || k.octave == l;
                if (has) ex->blurred[l] = b, next++;
            }
        }
    }
    const int n = (int)keypoints.size();
    if (n > cap) return -1;
    if (n) std::memcpy(kps, keypoints.data(), sizeof(ref_keypoint) * (size_t)n);
    for (int i = 0; i < n; i++) std::memcpy(desc + (size_t)32 * i, descriptors.ptr(i), 32);
    return n;
}

int ref_level_size(void *h, int level, int *w, int *hgt) {
    const cv::Mat &P = ((Tap *)h)->mvImagePyramid[level];
    *w = P.cols; *hgt = P.rows;
    return P.empty() ? -1 : 0;
}
// level plane; ring = 0: the w x h image, ring = 1: the (w+38) x (h+38) buffer with the reflected border (:1115-1128)
int ref_get_level(void *h, int level, int ring, uint8_t *out) {
    const cv::Mat &P = ((Tap *)h)->mvImagePyramid[level];
    if (P.empty()) return -1;
    const size_t step = P.step;
    const int e = ring ? kEdge : 0;
    for (int y = -e; y < P.rows + e; y++) std::memcpy(out + (size_t)(y + e) * (P.cols + 2 * e), P.data + (ptrdiff_t)y * (ptrdiff_t)step - e, (size_t)(P.cols + 2 * e));
    return 0;
}
int ref_get_blur(void *h, int level, uint8_t *out) {
    const cv::Mat &B = ((Tap *)h)->blurred[level];
    if (B.empty()) return -1;
    for (int y = 0; y < B.rows; y++) std::memcpy(out + (size_t)y * B.cols, B.ptr(y), (size_t)B.cols);
    return 0;
}
int ref_get_candidates(void *h, int level, float *xyr, int cap) {
    const std::vector<float> &c = ((Tap *)h)->cands[level];
    const int n = (int)(c.size() / 3);
    if (n <= cap && n) std::memcpy(xyr, c.data(), c.size() * sizeof(float));
    return n;
}
// ORBextractor::DistributeOctTree (src/orb_extractor.cpp:539-763) on a caller-supplied candidate list
int ref_distribute(void *h, const float *xyr, int n, int min_x, int max_x, int min_y, int max_y, int n_want, int level,
                   float *out_xyr, int cap) {
    Tap *ex = (Tap *)h;
    if (g_monotonic.load()) arena_reset();
    std::vector<cv::KeyPoint> in;
    for (int i = 0; i < n; i++) in.push_back(cv::KeyPoint(xyr[3 * i], xyr[3 * i + 1], 7.f, -1, xyr[3 * i + 2]));
    const std::vector<cv::KeyPoint> out = ex->DistributeOctTree(in, min_x, max_x, min_y, max_y, n_want, level);
    const int m = (int)out.size();
    for (int i = 0; i < m && i < cap; i++) {
        out_xyr[3 * i] = out[i].pt.x; out_xyr[3 * i + 1] = out[i].pt.y; out_xyr[3 * i + 2] = out[i].response;
    }
    return m;
}

int ref_hamming256(const void *a, const void *b) {  // include/orb_extractor.h:87-103
    return ORB_SLAM2::ORBextractor::DescriptorDistance(cv::Mat(1, 32, CV_8U, (void *)a), cv::Mat(1, 32, CV_8U, (void *)b));
}
float ref_fast_atan2(float y, float x) { return cv::fastAtan2(y, x); }

// ---- camera (the reference's src/camera.cpp) ---------------------------------------------------------------------------
static Camera *make_camera(const ref_camera *c) {
    Eigen::Matrix<double, 3, 3> K;
    K(0, 0) = c->fx; K(1, 1) = c->fy; K(0, 2) = c->cx; K(1, 2) = c->cy; K(2, 2) = 1.;
    Eigen::VectorXd D(4);
    for (int i = 0; i < 4; i++) D(i) = c->d[i];
    return new Camera(K, D, c->width, c->height);
}
int ref_camera_project(const ref_camera *c, const double xc[3], double uv[2]) {
    std::unique_ptr<Camera> cam(make_camera(c));
    const Eigen::Vector2d p = cam->Project(Eigen::Vector3d(xc[0], xc[1], xc[2]));
    uv[0] = p[0]; uv[1] = p[1];
    return cam->IsInImage(p) ? 1 : 0;
}
void ref_normalized_undistort(const ref_camera *c, const ref_keypoint *kps, int n, double *xy) {
    std::unique_ptr<Camera> cam(make_camera(c));
    for (int i = 0; i < n; i++) {  // as src/frame.cpp:52-56
        const Eigen::Vector3d nuv = cam->NormalizedUndistort(Eigen::Vector2d(kps[i].x, kps[i].y));
        xy[2 * i] = nuv[0]; xy[2 * i + 1] = nuv[1];
    }
}
// g2o::SE3Quat(q, t) * x; qt = {qx, qy, qz, qw, tx, ty, tz}; q_held = the rotation the SE3Quat holds after its constructor
void ref_se3_apply(const double qt[7], const double *x, int n, double *out, double q_held[4]) {
    const g2o::SE3Quat T(Eigen::Quaterniond(qt[3], qt[0], qt[1], qt[2]), Eigen::Vector3d(qt[4], qt[5], qt[6]));
    if (q_held) { q_held[0] = T.rotation().x(); q_held[1] = T.rotation().y(); q_held[2] = T.rotation().z(); q_held[3] = T.rotation().w(); }
    for (int i = 0; i < n; i++) {
        const Eigen::Vector3d y = T * Eigen::Vector3d(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
        out[3 * i] = y[0]; out[3 * i + 1] = y[1]; out[3 * i + 2] = y[2];
    }
}

// ---- matchers (the reference's src/matcher.cpp) ------------------------------------------------------------------------
static void fill_frame(Frame *f, const ref_keypoint *kps, const uint8_t *desc, int n) {
    f->keypoints_.resize((size_t)n);
    if (n) std::memcpy(f->keypoints_.data(), kps, sizeof(ref_keypoint) * (size_t)n);
    f->descriptions_ = cv::Mat(n, 32, CV_8U, (void *)desc);
    f->mappoints_.assign((size_t)n, nullptr);
}

void ref_stereo_match(const ref_keypoint *kl, const uint8_t *dl, int nl, const ref_keypoint *kr, const uint8_t *dr, int nr,
                      const ref_camera *c, int *out_idx) {
    std::unique_ptr<Camera> cam(make_camera(c));
    StereoFrame f;
    f.camera_ = cam.get();
    fill_frame(&f, kl, dl, nl);
    f.r_keypoints_.resize((size_t)nr);
    if (nr) std::memcpy(f.r_keypoints_.data(), kr, sizeof(ref_keypoint) * (size_t)nr);
    f.r_descriptions_ = cv::Mat(nr, 32, CV_8U, (void *)dr);
    StereoMatch(&f);
    for (int i = 0; i < nl; i++) out_idx[i] = f.stereo_correspond_[(size_t)i];
}

// Map points live in one array, so the std::set<Mappoint*> the reference iterates (pointer order) visits them in
// caller order (the oracle's rule T3).  kp_to_query[j] = index of the map point matched to keypoint j, or -1.
void ref_projection_match(const double *xw, const uint8_t *mp_desc, const uint8_t *skip, int n, const double qt[7],
                          const ref_camera *c, const ref_keypoint *kps, const uint8_t *kp_desc, int m, double radius,
                          int *kp_to_query) {
    std::unique_ptr<Camera> cam(make_camera(c));
    Frame f;
    f.camera_ = cam.get();
    fill_frame(&f, kps, kp_desc, m);
    std::vector<Mappoint> mps((size_t)n);
    std::set<Mappoint *> all;
    for (int i = 0; i < n; i++) {
        mps[(size_t)i].xw_ = Eigen::Vector3d(xw[3 * i], xw[3 * i + 1], xw[3 * i + 2]);
        mps[(size_t)i].desc_ = cv::Mat(1, 32, CV_8U, (void *)(mp_desc + (size_t)32 * i));
        all.insert(&mps[(size_t)i]);
        if (skip && skip[i]) f.in_frame_.insert(&mps[(size_t)i]);
    }
    const g2o::SE3Quat Tcw(Eigen::Quaterniond(qt[3], qt[0], qt[1], qt[2]), Eigen::Vector3d(qt[4], qt[5], qt[6]));
    const std::map<int, Mappoint *> matches = ProjectionMatch(all, Tcw, &f, radius);
    for (int j = 0; j < m; j++) kp_to_query[j] = -1;
    for (std::map<int, Mappoint *>::const_iterator it = matches.begin(); it != matches.end(); ++it)
        kp_to_query[it->first] = (int)(it->second - mps.data());
}

// ---- CPU baseline driver (bench.py --impl reference / cpu_baseline) -----------------------------------------------------
// The reference's front end on `count` consecutive stereo frames, nthreads worker threads (the reference itself is one
// serial tracking thread, src/pipeline.cpp:143-147; threads here only keep all host cores busy with independent frames):
//   per frame      ORBextractor::extract(left), extract(right)  (src/frame.cpp:47,388), StereoMatch (src/pipeline.cpp:248)
//   per frame >= 1 ProjectionMatch of the previous frame's stereo points with an identity motion prior, r = radius
//                  (src/posetracker.cpp:186); the points come from StereoFrame::GetDepth, restated here from
//                  src/frame.cpp:391-409 (its translation unit needs FLANN / DBoW2) on the real Camera::NormalizedUndistort.
// glibc heap (as the reference runs).  Returns the number of stereo matches.
struct SeqFrame {
    StereoFrame f;
    cv::Mat dl, dr;
    std::vector<Eigen::Vector2d> nrm;
};

int64_t ref_stereo_sequence(const uint8_t *left, const uint8_t *right, int count, int w, int h, int nthreads, int nfeatures,
                            float scale_factor, int nlevels, int ini_th, int min_th, const ref_camera *c, double baseline,
                            double radius, int64_t *total_kps, int64_t *total_tracked) {
    const int was = g_monotonic.exchange(0);
    std::unique_ptr<Camera> cam(make_camera(c));
    std::vector<SeqFrame> fr((size_t)count);
    std::atomic<int> next(0);
    std::atomic<int64_t> kps(0), stereo(0), tracked(0);
    if (nthreads < 1) nthreads = 1;
    auto phase1 = [&]() {
        ORB_SLAM2::ORBextractor ex(nfeatures, scale_factor, nlevels, ini_th, min_th);
        for (int i = next++; i < count; i = next++) {
            SeqFrame &s = fr[(size_t)i];
            s.f.camera_ = cam.get();
            cv::Mat L(h, w, CV_8UC1, (void *)(left + (size_t)i * w * h)), R(h, w, CV_8UC1, (void *)(right + (size_t)i * w * h));
            ex.extract(L, cv::noArray(), s.f.keypoints_, s.dl);
            ex.extract(R, cv::noArray(), s.f.r_keypoints_, s.dr);
            s.f.descriptions_ = s.dl;
            s.f.r_descriptions_ = s.dr;
            s.f.mappoints_.assign(s.f.keypoints_.size(), nullptr);
            for (size_t k = 0; k < s.f.keypoints_.size(); k++) {  // src/frame.cpp:52-56
                const Eigen::Vector3d nuv = cam->NormalizedUndistort(Eigen::Vector2d(s.f.keypoints_[k].pt.x, s.f.keypoints_[k].pt.y));
                s.nrm.push_back(nuv.head<2>());
            }
            StereoMatch(&s.f);
            kps += (int64_t)(s.f.keypoints_.size() + s.f.r_keypoints_.size());
            for (size_t k = 0; k < s.f.stereo_correspond_.size(); k++) stereo += s.f.stereo_correspond_[k] >= 0;
        }
    };
    auto phase2 = [&]() {
        const g2o::SE3Quat Tcw(Eigen::Quaterniond(1, 0, 0, 0), Eigen::Vector3d(0, 0, 0));
        for (int i = next++; i < count; i = next++) {
            if (i == 0) continue;
            const SeqFrame &p = fr[(size_t)i - 1];
            std::vector<Mappoint> mps;
            mps.reserve(p.f.keypoints_.size());
            for (size_t k = 0; k < p.f.keypoints_.size(); k++) {  // StereoFrame::GetDepth, src/frame.cpp:391-409
                const int j = p.f.stereo_correspond_[k];
                if (j < 0) continue;
                const double dx = p.f.keypoints_[k].pt.x - p.f.r_keypoints_[(size_t)j].pt.x;
                if (dx < 0.) continue;
                const double depth = c->fx * baseline / dx;
                Mappoint mp;
                const Eigen::Vector2d xy = p.nrm[k] * depth;
                mp.xw_ = Eigen::Vector3d(xy[0], xy[1], depth);
                mp.desc_ = p.dl.row((int)k);
                mps.push_back(mp);
            }
            std::set<Mappoint *> all;
            for (size_t k = 0; k < mps.size(); k++) all.insert(&mps[k]);
            tracked += (int64_t)ProjectionMatch(all, Tcw, &fr[(size_t)i].f, radius).size();
        }
    };
    for (int ph = 0; ph < 2; ph++) {
        next = 0;
        std::vector<std::thread> th;
        for (int t = 1; t < nthreads; t++) th.push_back(ph == 0 ? std::thread(phase1) : std::thread(phase2));
        if (ph == 0) phase1(); else phase2();
        for (size_t t = 0; t < th.size(); t++) th[t].join();
    }
    if (total_kps) *total_kps = kps.load();
    if (total_tracked) *total_tracked = tracked.load();
    g_monotonic = was;
    return stereo.load();
}

}  // extern "C"
