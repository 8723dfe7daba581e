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
        std::vector<std::vector<int> > indices(points.rows);
        const double r2 = radius * radius;
        std::vector<std::pair<double, int> > hits;
        for (size_t q = 0; q < points.rows; q++) {
            hits.clear();
            for (size_t j = 0; j < keypoints_.size(); j++) {
                const double dx = points[q][0] - (double)keypoints_[j].pt.x, dy = points[q][1] - (double)keypoints_[j].pt.y;
                double d = 0.;
                d += dx * dx;
                d += dy * dy;
                if (d < r2) hits.push_back(std::make_pair(d, (int)j));
            }
            std::sort(hits.begin(), hits.end());
            for (size_t k = 0; k < hits.size(); k++) indices[q].push_back(hits[k].second);
        }
        return indices;
    }
    const Camera *GetCamera() const { return camera_; }

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
