// TEST INFRASTRUCTURE (oracle/_ref build only).  Mock of the reference's include/frame.h with exactly the members
// src/matcher.cpp calls, same names and signatures (frame.h:50-51,78,94,96-99,107,156-159), and the behaviour of
// src/frame.cpp:465-467 (IsInFrame -> Camera::IsInImage, the REAL src/camera.cpp), :205-209 (GetDescription = row view),
// :170-178 (SearchRadius).  The per-frame FLANN kd-tree (frame.cpp:59-68) is replaced by the exhaustive search it is
// an index for: flann::KDTreeSingleIndex::radiusSearch(points, indices, dists, r*r, SearchParams()) returns every
// point whose L2<double> squared distance (dx*dx, then + dy*dy) is < r*r (RadiusResultSet::addPoint: `dist < radius`),
// sorted by distance (SearchParams().sorted = true).  Keypoints enter as (double)kpt.pt.x/y (frame.cpp:62-66).
#ifndef FRAME_H_
#define FRAME_H_
#include <flann/flann.hpp>
#include "stdafx.h"
class Mappoint;
class Camera;
class Frame {
public:
    virtual ~Frame() {}
    int GetIndex(const Mappoint *mp) const { return in_frame_.count(mp) ? 0 : -1; }
    const std::vector<cv::KeyPoint> &GetKeypoints() const { return keypoints_; }
    const std::vector<Mappoint *> &GetVecMappoints() const { return mappoints_; }
    bool IsInFrame(const Eigen::Vector2d &uv) const;
    const cv::Mat GetDescription(int i) const { return descriptions_.row(i); }
    std::vector<std::vector<int> > SearchRadius(const flann::Matrix<double> &points, double radius) const {
        // candidates from a 64-px bucket grid (a pure accelerator, like the kd-tree it stands in for); the test and the order
        // are the exhaustive search's
        const int B = 64;
        if (grid_.empty() && !keypoints_.empty()) {
            float mx = 0, my = 0;
            for (size_t j = 0; j < keypoints_.size(); j++) { mx = std::max(mx, keypoints_[j].pt.x); my = std::max(my, keypoints_[j].pt.y); }
            gw_ = (int)mx / B + 1; gh_ = (int)my / B + 1;
            grid_.assign((size_t)gw_ * gh_, std::vector<int>());
            for (size_t j = 0; j < keypoints_.size(); j++) grid_[cell(keypoints_[j].pt.x, keypoints_[j].pt.y, B)].push_back((int)j);
        }
        std::vector<std::vector<int> > indices(points.rows);
        const double r2 = radius * radius;
        std::vector<std::pair<double, int> > hits;
        for (size_t q = 0; q < points.rows && !grid_.empty(); q++) {
            hits.clear();
            const double u = points[q][0], v = points[q][1];
            const int cx0 = clampi((int)std::floor((u - radius) / B), gw_), cx1 = clampi((int)std::floor((u + radius) / B), gw_);
            const int cy0 = clampi((int)std::floor((v - radius) / B), gh_), cy1 = clampi((int)std::floor((v + radius) / B), gh_);
            for (int cy = cy0; cy <= cy1; cy++)
                for (int cx = cx0; cx <= cx1; cx++) {
                    const std::vector<int> &c = grid_[(size_t)cy * gw_ + cx];
                    for (size_t t = 0; t < c.size(); t++) {
                        const int j = c[t];
                        const double dx = u - (double)keypoints_[j].pt.x, dy = v - (double)keypoints_[j].pt.y;
                        double d = 0.;
                        d += dx * dx;
                        d += dy * dy;
                        if (d < r2) hits.push_back(std::make_pair(d, j));
                    }
                }
            std::sort(hits.begin(), hits.end());
            for (size_t k = 0; k < hits.size(); k++) indices[q].push_back(hits[k].second);
        }
        return indices;
    }
    const Camera *GetCamera() const { return camera_; }

    static int clampi(int c, int n) { return c < 0 ? 0 : (c >= n ? n - 1 : c); }
    int cell(float x, float y, int B) const { return clampi((int)y / B, gh_) * gw_ + clampi((int)x / B, gw_); }
    mutable std::vector<std::vector<int> > grid_;
    mutable int gw_ = 0, gh_ = 0;

    std::vector<cv::KeyPoint> keypoints_;
    cv::Mat descriptions_;
    std::vector<Mappoint *> mappoints_;
    std::set<const Mappoint *> in_frame_;
    const Camera *camera_ = nullptr;
};
class StereoFrame : public Frame {
public:
    const std::vector<cv::KeyPoint> &GetRightKeypoints() const { return r_keypoints_; }
    const cv::Mat GetRightDescription(int i) const { return r_descriptions_.row(i); }
    void SetStereoCorrespond(const std::vector<int> &correspond) { stereo_correspond_ = correspond; }
    std::vector<int> stereo_correspond_;
    std::vector<cv::KeyPoint> r_keypoints_;
    cv::Mat r_descriptions_;
};
#endif
