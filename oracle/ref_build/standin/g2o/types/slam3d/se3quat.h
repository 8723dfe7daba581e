// TEST INFRASTRUCTURE (oracle/_ref build only).  Stand-in for g2o::SE3Quat (g2o is built from source by the reference,
// CMakeLists.txt:33-37, and is not in this image).  Restates the published class (g2o/types/slam3d/se3quat.h):
// a unit quaternion _r + translation _t; the constructor normalises the rotation (w >= 0, unit norm);
// operator*(Vector3D v) = _t + _r * v with Eigen's Quaternion * vector product.
#ifndef REF_STANDIN_G2O_SE3QUAT_H_
#define REF_STANDIN_G2O_SE3QUAT_H_
#include <Eigen/Core>
namespace g2o {
class SE3Quat {
public:
    SE3Quat() {}
    SE3Quat(const Eigen::Quaterniond &q, const Eigen::Vector3d &t) : _r(q), _t(t) { normalizeRotation(); }
    const Eigen::Vector3d &translation() const { return _t; }
    const Eigen::Quaterniond &rotation() const { return _r; }
    Eigen::Vector3d operator*(const Eigen::Vector3d &v) const { return _t + _r * v; }
    Eigen::Vector3d map(const Eigen::Vector3d &xyz) const { return _r * xyz + _t; }
    void normalizeRotation() {
        if (_r.w() < 0) _r.negate();
        _r.normalize();
    }
protected:
    Eigen::Quaterniond _r;
    Eigen::Vector3d _t;
};
}  // namespace g2o
#endif
