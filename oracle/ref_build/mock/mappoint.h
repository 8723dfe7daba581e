// TEST INFRASTRUCTURE (oracle/_ref build only).  Mock of the reference's include/mappoint.h: the two accessors
// src/matcher.cpp calls (GetXw :150, GetDescription :168), same signatures (mappoint.h:43-45).
#ifndef MAPPOINT_H_
#define MAPPOINT_H_
#include "stdafx.h"
class Mappoint {
public:
    Mappoint() {}
    Eigen::Vector3d GetXw() const { return xw_; }
    cv::Mat GetDescription() const { return desc_; }
    Eigen::Vector3d xw_;
    cv::Mat desc_;  // 1 x 32 CV_8U view
};
#endif
