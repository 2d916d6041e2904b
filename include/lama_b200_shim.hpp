// lama_b200_shim.hpp -- header-only C++ shim that re-creates the reference's front-end classes on top of the C-ABI.
//
// A maintainer of iris-ua/iris_lama would add this file next to include/lama/pf_slam2d.h and let
// lama::PFSlam2D / Slam2D / Loc2D forward to it.  It is templated on the caller's
// point-cloud and pose types so that it compiles without Eigen: any cloud with `.points` (elements indexable
// [0..2]), `.sensor_origin_` (indexable [0..2]) and `.sensor_orientation_` (with x(), y(), z(), w()) works,
// i.e. lama::PointCloudXYZ (include/lama/types.h:111-120); any pose with x(), y(), rotation() works, i.e.
// lama::Pose2D (include/lama/pose2d.h:42-78).
#pragma once

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "lama_b200.h"

namespace lama_b200_shim {

inline void check(int rc)
{
    if (rc != LAMA_OK) throw std::runtime_error(std::string("lama_b200: ") + lama_last_error());
}

template <typename Cloud>
struct FlatCloud {
    std::vector<double> pts;
    double origin[3], quat[4];
    explicit FlatCloud(const Cloud& c)
    {
        pts.reserve(c.points.size() * 3);
        for (const auto& p : c.points) { pts.push_back(p[0]); pts.push_back(p[1]); pts.push_back(p[2]); }
        for (int i = 0; i < 3; ++i) origin[i] = c.sensor_origin_[i];
        quat[0] = c.sensor_orientation_.x(); quat[1] = c.sensor_orientation_.y();
        quat[2] = c.sensor_orientation_.z(); quat[3] = c.sensor_orientation_.w();
    }
};

// lama::PFSlam2D (include/lama/pf_slam2d.h:187-232)
class PFSlam2D {
public:
    using Options = lama_pf_options;
    static Options defaults(uint32_t particles)
    {
        Options o;
        check(lama_pf_options_default(&o));
        o.particles = particles;
        return o;
    }
    explicit PFSlam2D(const Options& o) { check(lama_pf_create(&o, &h_)); }
    ~PFSlam2D() { lama_pf_destroy(h_); }
    PFSlam2D(const PFSlam2D&) = delete;
    PFSlam2D& operator=(const PFSlam2D&) = delete;

    template <typename Pose>
    void setPrior(const Pose& prior)  // pf_slam2d.cpp:146-149
    {
        const double xyr[3] = {prior.x(), prior.y(), prior.rotation()};
        check(lama_pf_set_prior(h_, xyr));
    }
    // bool update(const PointCloudXYZ::Ptr& surface, const Pose2D& odometry, double timestamp)  pf_slam2d.h:199
    template <typename CloudPtr, typename Pose>
    bool update(const CloudPtr& surface, const Pose& odometry, double timestamp)
    {
        FlatCloud<typename std::remove_reference<decltype(*surface)>::type> f(*surface);
        const double odom[3] = {odometry.x(), odometry.y(), odometry.rotation()};
        int did = 0;
        check(lama_pf_update(h_, f.pts.data(), (int)(f.pts.size() / 3), f.origin, f.quat, odom, timestamp, &did));
        return did != 0;
    }
    void getPose(double xyr[3]) const { check(lama_pf_get_pose(h_, xyr)); }       // pf_slam2d.cpp:332-336
    size_t getBestParticleIdx() const { int i = 0; check(lama_pf_get_best_particle(h_, &i)); return (size_t)i; }
    double getNeff() const { double v = 0; check(lama_pf_get_neff(h_, &v)); return v; }
    // Map::write of a particle's occupancy (kind 0) / distance (kind 1) map: a reference .sdm file (map.cpp:490-529)
    void writeMap(int particle, int kind, const std::string& path) const { check(lama_pf_write_map(h_, particle, kind, path.c_str())); }
    // the grey image PFSlam2D::saveOccImage hands to sdm::export_to_png (pf_slam2d.cpp:338-342, export.cpp:46-73), row-major
    std::vector<uint8_t> occImage(int& width, int& height) const
    {
        int dims[2] = {0, 0};
        const int best = (int)getBestParticleIdx();
        check(lama_pf_export_image(h_, best, 0, nullptr, 0, dims));
        std::vector<uint8_t> img((size_t)dims[0] * dims[1]);
        if (!img.empty()) check(lama_pf_export_image(h_, best, 0, img.data(), img.size(), dims));
        width = dims[0]; height = dims[1];
        return img;
    }
    lama_pf* handle() const { return h_; }

private:
    lama_pf* h_ = nullptr;
};

// lama::Slam2D (include/lama/slam2d.h:128-161)
class Slam2D {
public:
    using Options = lama_slam_options;
    static Options defaults() { Options o; check(lama_slam_options_default(&o)); return o; }
    explicit Slam2D(const Options& o) { check(lama_slam_create(&o, &h_)); }
    ~Slam2D() { lama_slam_destroy(h_); }
    Slam2D(const Slam2D&) = delete;
    Slam2D& operator=(const Slam2D&) = delete;
    template <typename Pose>
    void setPose(const Pose& p) { const double xyr[3] = {p.x(), p.y(), p.rotation()}; check(lama_slam_set_pose(h_, xyr)); }
    template <typename CloudPtr, typename Pose>
    bool update(const CloudPtr& surface, const Pose& odometry, double timestamp)  // slam2d.h:134
    {
        FlatCloud<typename std::remove_reference<decltype(*surface)>::type> f(*surface);
        const double odom[3] = {odometry.x(), odometry.y(), odometry.rotation()};
        int did = 0;
        check(lama_slam_update(h_, f.pts.data(), (int)(f.pts.size() / 3), f.origin, f.quat, odom, timestamp, &did));
        return did != 0;
    }
    void getPose(double xyr[3]) const { check(lama_slam_get_pose(h_, xyr)); }
    uint32_t getNumberOfProcessedCells() const { uint32_t n = 0; check(lama_slam_get_processed_cells(h_, &n)); return n; }
    void writeMap(int kind, const std::string& path) const { check(lama_slam_write_map(h_, kind, path.c_str())); }
    lama_slam* handle() const { return h_; }

protected:
    lama_slam* h_ = nullptr;
};

// lama::LidarOdometry2D (include/lama/lidar_odometry_2d.h:45-75): the Slam2D handle in its lidar-odometry mode
class LidarOdometry2D {
public:
    struct Options {
        double resolution;
        uint32_t max_iter;
        Options() : resolution(0.05), max_iter(100) {}   // lidar_odometry_2d.h:62-68
    };
    explicit LidarOdometry2D(const Options& o = Options())
    {
        lama_slam_options s;
        check(lama_slam_options_default(&s));
        s.lidar_odometry = 1; s.resolution = o.resolution; s.max_iter = o.max_iter;
        check(lama_slam_create(&s, &h_));
    }
    ~LidarOdometry2D() { lama_slam_destroy(h_); }
    LidarOdometry2D(const LidarOdometry2D&) = delete;
    LidarOdometry2D& operator=(const LidarOdometry2D&) = delete;
    template <typename CloudPtr>
    bool update(const CloudPtr& surface, double timestamp)  // lidar_odometry_2d.h:73
    {
        FlatCloud<typename std::remove_reference<decltype(*surface)>::type> f(*surface);
        int did = 0;
        check(lama_slam_update(h_, f.pts.data(), (int)(f.pts.size() / 3), f.origin, f.quat, nullptr, timestamp, &did));
        return did != 0;
    }
    void getOdom(double xyr[3]) const { check(lama_slam_get_pose(h_, xyr)); }   // the public member `odom`

private:
    lama_slam* h_ = nullptr;
};

// lama::Loc2D (include/lama/loc2d.h:103-130); the caller fills distance_map() through lama_dm_add_obstacles + lama_dm_update
class Loc2D {
public:
    using Options = lama_loc_options;
    static Options defaults() { Options o; check(lama_loc_options_default(&o)); return o; }
    void Init(const Options& o) { check(lama_loc_create(&o, &h_)); }
    ~Loc2D() { lama_loc_destroy(h_); }
    lama_dm* distance_map() { lama_dm* d = nullptr; check(lama_loc_distance_map(h_, &d)); return d; }
    template <typename Pose>
    void setPose(const Pose& p) { const double xyr[3] = {p.x(), p.y(), p.rotation()}; check(lama_loc_set_pose(h_, xyr)); }
    template <typename CloudPtr, typename Pose>
    bool update(const CloudPtr& surface, const Pose& odometry, double timestamp, bool force_update = false)  // loc2d.h:113
    {
        FlatCloud<typename std::remove_reference<decltype(*surface)>::type> f(*surface);
        const double odom[3] = {odometry.x(), odometry.y(), odometry.rotation()};
        int did = 0;
        check(lama_loc_update(h_, f.pts.data(), (int)(f.pts.size() / 3), f.origin, f.quat, odom, timestamp, force_update ? 1 : 0, &did));
        return did != 0;
    }
    void getPose(double xyr[3]) const { check(lama_loc_get_pose(h_, xyr)); }
    void getCovar(double cov[9]) const { check(lama_loc_get_covar(h_, cov)); }
    double getRMSE() const { double v = 0; check(lama_loc_get_rmse(h_, &v)); return v; }
    void triggerGlobalLocalization() { check(lama_loc_trigger_global_localization(h_)); }   // loc2d.cpp:194-197
    void readOccupancyMap(const std::string& path) { check(lama_loc_occupancy_read(h_, path.c_str())); }    // occupancy_map->read(path)
    void readDistanceMap(const std::string& path) { check(lama_dm_read(distance_map(), path.c_str())); }    // distance_map->read(path)

private:
    lama_loc* h_ = nullptr;
};

}  // namespace lama_b200_shim
