// stand-in for ../tracker/camera.hpp (absent): pinhole camera with an optional rectangle of valid pixels
#pragma once
#include <memory>
#include <string>
#include <Eigen/Dense>
namespace tracker {
class Camera {
public:
    double fx = 1, fy = 1, cx = 0, cy = 0;
    double vx0 = -1e300, vy0 = -1e300, vx1 = 1e300, vy1 = 1e300;   // valid region [vx0, vx1) x [vy0, vy1)
    virtual ~Camera() = default;
    virtual bool isValidPixel(double x, double y) const { return x >= vx0 && x < vx1 && y >= vy0 && y < vy1; }
    bool isValidPixel(const Eigen::Vector2d &p) const { return isValidPixel(p.x(), p.y()); }
    virtual bool rayToPixel(const Eigen::Vector3d &ray, Eigen::Vector2d &pix) const {
        if (!(ray.z() > 0)) return false;
        pix = Eigen::Vector2d(fx * (ray.x() / ray.z()) + cx, fy * (ray.y() / ray.z()) + cy);
        return true;
    }
    virtual bool pixelToRay(const Eigen::Vector2d &pix, Eigen::Vector3d &ray) const {
        ray = Eigen::Vector3d((pix.x() - cx) / fx, (pix.y() - cy) / fy, 1.0).normalized();
        return true;
    }
    virtual std::string serialize() const { return "pinhole"; }
    static std::shared_ptr<const Camera> deserialize(const std::string &) { return std::make_shared<Camera>(); }
};
}  // namespace tracker
