// lama_oracle.hpp -- CPU restatement of the LaMa particle-filter SLAM hot path.
//
// *** TEST INFRASTRUCTURE ONLY ***  This file is the parity ORACLE for the
// B200 kernels under iris_lama_b200/csrc.  Nothing in the product path may
// include, link or call it; only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py do.
//
// PARITY UNPINNED: the reference (iris-ua/iris_lama @ fd60e55) ships no tests,
// fixtures or golden vectors, and it cannot be compiled here (it needs Eigen 3.3,
// which is neither installed nor vendored).  This restatement follows the cited
// reference lines literally in plain C++17 (no Eigen) and is pinned only by the
// independent checks in tests/ (brute-force EDT, finite differences, numpy
// MT19937, hand-enumerated Bresenham, closed-form SE2 identities).
//
// Citations are `path:line` relative to /root/reference.  libstdc++ <random>
// and <queue> are used as-is so RNG streams and heap tie order equal a
// reference build made with this toolchain.
#pragma once

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <functional>
#include <limits>
#include <memory>
#include <mutex>
#include <queue>
#include <random>
#include <stdexcept>
#include <thread>
#include <unordered_map>
#include <vector>

namespace orc {

// ----------------------------------------------------------------------------------------------
// Lie groups: SO2 / SE2 as in the vendored Sophus headers.
// ----------------------------------------------------------------------------------------------
constexpr double kLieEps = 1e-10;  // include/lama/sophus/sophus.hpp:37-39

struct SO2 {
    double c = 1.0, s = 0.0;  // unit complex (so2.hpp:474-476)

    // so2.hpp:246-255 -- throws on a (near) zero complex number.
    void normalize()
    {
        double len = std::sqrt(c * c + s * s);
        if (len < kLieEps) throw std::runtime_error("Complex number is (near) zero!");
        c /= len;
        s /= len;
    }
    // so2.hpp:506-509 -- the (real, imag) constructor normalises.
    static SO2 from_complex(double re, double im)
    {
        SO2 r;
        r.c = re;
        r.s = im;
        r.normalize();
        return r;
    }
    // so2.hpp:322-324 (exp) through so2.hpp:537-539 (angle ctor).
    static SO2 exp(double theta) { return from_complex(std::cos(theta), std::sin(theta)); }
    // so2.hpp:401-404
    double log() const { return std::atan2(s, c); }
    // so2.hpp:192-194
    SO2 inverse() const { return from_complex(c, -s); }
    // so2.hpp:167-176 + :275-278 : multiply, then renormalise.
    void mul_assign(const SO2& o)
    {
        double lr = c, li = s;
        c = lr * o.c - li * o.s;
        s = lr * o.s + li * o.c;
        normalize();
    }
    // so2.hpp:262-266
    void rotate(double px, double py, double& ox, double& oy) const
    {
        ox = c * px - s * py;
        oy = s * px + c * py;
    }
};

struct SE2 {
    SO2 r;
    double tx = 0.0, ty = 0.0;

    SE2() = default;
    SE2(const SO2& rot, double x, double y) : r(rot), tx(x), ty(y) {}
    // se2.hpp:648-651
    SE2(double theta, double x, double y) : r(SO2::exp(theta)), tx(x), ty(y) {}

    // se2.hpp:153-157 (fastMultiply) + :262-265 (operator*=)
    void mul_assign(const SE2& o)
    {
        double dx, dy;
        r.rotate(o.tx, o.ty, dx, dy);
        tx += dx;
        ty += dy;
        r.mul_assign(o.r);
    }
    SE2 operator*(const SE2& o) const
    {
        SE2 res(*this);
        res.mul_assign(o);
        return res;
    }
    // se2.hpp:163-167
    SE2 inverse() const
    {
        SO2 ir = r.inverse();
        double nx = tx * -1.0, ny = ty * -1.0, ox, oy;
        ir.rotate(nx, ny, ox, oy);
        return SE2(ir, ox, oy);
    }
    // se2.hpp:389-412
    static SE2 exp(const double a[3])
    {
        double theta = a[2];
        SO2 so2      = SO2::exp(theta);
        double sin_theta_by_theta, one_minus_cos_theta_by_theta;
        if (std::abs(theta) < kLieEps) {
            double theta_sq              = theta * theta;
            sin_theta_by_theta           = 1. - (1. / 6.) * theta_sq;
            one_minus_cos_theta_by_theta = 0.5 * theta - (1. / 24.) * theta * theta_sq;
        } else {
            sin_theta_by_theta           = so2.s / theta;
            one_minus_cos_theta_by_theta = (1. - so2.c) / theta;
        }
        return SE2(so2, sin_theta_by_theta * a[0] - one_minus_cos_theta_by_theta * a[1],
                   one_minus_cos_theta_by_theta * a[0] + sin_theta_by_theta * a[1]);
    }
};

// src/pose2d.cpp:41-131
struct Pose2D {
    SE2 state;
    Pose2D() = default;
    Pose2D(double x, double y, double rot) : state(rot, x, y) {}
    explicit Pose2D(const SE2& s) : state(s) {}
    Pose2D plus(const Pose2D& o) const { return Pose2D(state * o.state); }             // :76-79
    Pose2D minus(const Pose2D& o) const { return Pose2D(state.inverse() * o.state); }  // :81-84
    void plus_assign(const Pose2D& o) { state.mul_assign(o.state); }                   // :86-90
    double x() const { return state.tx; }
    double y() const { return state.ty; }
    double rotation() const { return state.r.log(); }  // :117-120
    double xy_norm() const { return std::sqrt(state.tx * state.tx + state.ty * state.ty); }
};

// ----------------------------------------------------------------------------------------------
// Random numbers: one process-global mt19937, a fresh distribution per call (src/random.cpp:38-73)
// ----------------------------------------------------------------------------------------------
struct Random {
    std::mt19937 gen;
    void seed(uint32_t s) { gen.seed(s); }
    double uniform()
    {
        std::uniform_real_distribution<double> d(0.0, 1.0);
        return d(gen);
    }
    double normal(double stddev)
    {
        std::normal_distribution<double> d(0.0, stddev);
        return d(gen);
    }
};

// ----------------------------------------------------------------------------------------------
// 3-D affine transform helpers standing in for the Eigen Affine3d algebra used at
// src/match_surface_2d.cpp:49-58, src/pf_slam2d.cpp:397-403,444-452.
// ----------------------------------------------------------------------------------------------
struct Affine3 {
    double l[3][3];
    double t[3];
    void apply(const double p[3], double out[3]) const
    {
        for (int i = 0; i < 3; ++i) out[i] = ((l[i][0] * p[0] + l[i][1] * p[1]) + l[i][2] * p[2]) + t[i];
    }
};

struct PointCloud {
    std::vector<double> pts;                  // xyz AoS, N x 3 (include/lama/types.h:111-120)
    double origin[3] = {0, 0, 0};             // sensor_origin_
    double quat[4]   = {0, 0, 0, 1};          // sensor_orientation_ as (x,y,z,w)
    size_t size() const { return pts.size() / 3; }
};

// Translation3d(origin) * Quaterniond  (Eigen quaternion -> rotation matrix formula)
inline Affine3 moving_tf(const PointCloud& pc)
{
    const double x = pc.quat[0], y = pc.quat[1], z = pc.quat[2], w = pc.quat[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    Affine3 a;
    a.l[0][0] = 1 - (tyy + tzz); a.l[0][1] = txy - twz;       a.l[0][2] = txz + twy;
    a.l[1][0] = txy + twz;       a.l[1][1] = 1 - (txx + tzz); a.l[1][2] = tyz - twx;
    a.l[2][0] = txz - twy;       a.l[2][1] = tyz + twx;       a.l[2][2] = 1 - (txx + tyy);
    for (int i = 0; i < 3; ++i) a.t[i] = pc.origin[i];
    return a;
}

// Translation3d(x,y,0) * AngleAxisd(theta, UnitZ)   (Eigen AngleAxis::toRotationMatrix, axis = z)
inline Affine3 fixed_tf(double x, double y, double theta)
{
    const double s = std::sin(theta), c = std::cos(theta);
    Affine3 a;
    a.l[0][0] = c;  a.l[0][1] = -s; a.l[0][2] = 0;
    a.l[1][0] = s;  a.l[1][1] = c;  a.l[1][2] = 0;
    a.l[2][0] = 0;  a.l[2][1] = 0;  a.l[2][2] = (1 - c) + c;
    a.t[0] = x; a.t[1] = y; a.t[2] = 0.0;
    return a;
}

inline Affine3 compose(const Affine3& f, const Affine3& m)
{
    Affine3 r;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) r.l[i][j] = (f.l[i][0] * m.l[0][j] + f.l[i][1] * m.l[1][j]) + f.l[i][2] * m.l[2][j];
        r.t[i] = ((f.l[i][0] * m.t[0] + f.l[i][1] * m.t[1]) + f.l[i][2] * m.t[2]) + f.t[i];
    }
    return r;
}

// ----------------------------------------------------------------------------------------------
// Sparse-dense map: patch hash + dense patches with "known" bitmask and copy-on-write sharing.
// include/lama/sdm/map.h:68-198, src/sdm/map.cpp:42-107,198-227,371-455,
// include/lama/sdm/container.h:102-123,167-183, include/lama/cow_ptr.h:96-114.
// ----------------------------------------------------------------------------------------------
constexpr uint64_t kUniversalConstant = 2642244;  // map.h:68

struct Vec3u {
    uint32_t x, y, z;
    bool operator==(const Vec3u& o) const { return x == o.x && y == o.y && z == o.z; }
};

template <typename Cell>
struct Patch {
    std::vector<Cell> cells;      // calloc'd block (container.cpp:78-95): zero initialised
    std::vector<uint64_t> mask;   // 1 bit per cell (container.cpp:39-43)
    explicit Patch(uint32_t volume) : cells(volume), mask((volume + 63) / 64, 0)
    {
        std::memset(static_cast<void*>(cells.data()), 0, sizeof(Cell) * volume);
    }
    bool is_on(uint32_t i) const { return 0 != (mask[i >> 6] & (uint64_t(1) << (i & 63))); }
    void set_on(uint32_t i) { mask[i >> 6] |= (uint64_t(1) << (i & 63)); }
};

template <typename Cell>
class SparseMap {
public:
    using PatchT   = Patch<Cell>;
    using PatchPtr = std::shared_ptr<PatchT>;

    double resolution, scale;
    uint32_t patch_length, patch_volume, log2dim;
    double offset;  // (UNIVERSAL_CONSTANT >> 1) * patch_length, map.cpp:55-58
    std::unordered_map<uint64_t, PatchPtr> patches;
    uint64_t detach_count = 0;  // deep patch copies performed by copy-on-write (work counter D)
    mutable uint64_t prev_idx_ = ~0ull;          // map.h:374-375
    mutable PatchPtr* prev_patch_ = nullptr;

    SparseMap(double res, uint32_t patch_size)
        : resolution(res), scale(1.0 / res), patch_length(1u << ((int)std::log2(patch_size)))
    {
        patch_volume = patch_length * patch_length;  // 2-D maps only (is_3d == false)
        log2dim      = (uint32_t)std::log2(patch_length);
        offset       = (double)((kUniversalConstant >> 1) * patch_length);
    }
    // COW copy: every patch pointer is shared (map.cpp:96-97); the patch cache is not carried over.
    SparseMap(const SparseMap& o)
        : resolution(o.resolution), scale(o.scale), patch_length(o.patch_length), patch_volume(o.patch_volume), log2dim(o.log2dim), offset(o.offset),
          patches(o.patches), detach_count(o.detach_count)
    {
    }

    // Map::memory (map.cpp:115-125) with the reference's sizes: key 8 + sizeof(COWPtr<Container>) 56 (shared_ptr + std::mutex, cow_ptr.h:117-118) +
    // pointer 8 per table entry, Container::memory() = volume * sizeof(cell type of the reference) shared between the owners
    size_t memory(uint32_t ref_cell_bytes) const
    {
        double total = 0.0;
        for (auto& kv : patches) {
            total += 8 + 56 + 8;
            total += (double)((size_t)patch_volume * ref_cell_bytes) / (double)kv.second.use_count();
        }
        return (size_t)total;
    }
    // map.h:137-138 : tf_ * p with tf_ = Translation(adjust*patch_length) * Scaling(scale)
    void w2m_nocast(const double p[3], double m[3]) const
    {
        for (int i = 0; i < 3; ++i) m[i] = p[i] * scale + offset;
    }
    // map.h:125-126
    Vec3u w2m(const double p[3]) const
    {
        double m[3];
        w2m_nocast(p, m);
        return Vec3u{(uint32_t)(m[0] + 0.5), (uint32_t)(m[1] + 0.5), (uint32_t)(m[2] + 0.5)};
    }
    // map.h:147-148
    void m2w(const Vec3u& c, double p[3]) const
    {
        p[0] = ((double)c.x - offset) / scale;
        p[1] = ((double)c.y - offset) / scale;
        p[2] = ((double)c.z - offset) / scale;
    }
    // map.h:153-161 (2-D branch)
    uint64_t m2p(const Vec3u& c) const { return (uint64_t)(c.x >> log2dim) * kUniversalConstant + (uint64_t)(c.y >> log2dim); }
    // map.h:182-189 (MASK3D == 0)
    uint32_t m2c(const Vec3u& c) const
    {
        const uint32_t mask = (1u << log2dim) - 1;
        return (c.x & mask) | ((c.y & mask) << log2dim);
    }
    // map.h:166-177
    Vec3u p2m(uint64_t idx) const
    {
        return Vec3u{(uint32_t)((idx / kUniversalConstant) << log2dim), (uint32_t)((idx % kUniversalConstant) << log2dim), 0};
    }

    // ---- persistence (src/sdm/map.cpp:490-575, header include/lama/sdm/map.h:72-75,95-103, patch payload container.cpp:143-176) ----
    struct IOHeader {  // map.h:95-103, natural alignment: 32 bytes
        uint32_t magic;
        uint16_t version;
        uint32_t cell_size;
        uint32_t patch_length;
        size_t num_patches;
        float resolution;
        bool is_3d;
    };
    static constexpr uint32_t kMagic = 0x6d64732e;   // map.h:72
    static constexpr uint16_t kIoVersion = 0x0103;   // map.h:75
    // `params` = what the concrete map's writeParameters emits (DynamicDistanceMap: max_sqdist_, dynamic_distance_map.cpp:200-203)
    bool write(const std::string& filename, const void* params, size_t nparams) const
    {
        std::ofstream f(filename.c_str(), std::ios::out | std::ios::binary | std::ios::trunc);
        if (!f.is_open()) return false;
        IOHeader header;
        std::memset(&header, 0, sizeof(header));
        header.magic = kMagic; header.version = kIoVersion; header.cell_size = (uint32_t)sizeof(Cell); header.patch_length = patch_length;
        header.num_patches = patches.size(); header.resolution = (float)resolution; header.is_3d = false;
        f.write((const char*)&header, sizeof(IOHeader));
        if (!f) return false;
        if (nparams) f.write((const char*)params, (std::streamsize)nparams);
        for (auto it = patches.begin(); it != patches.end(); ++it) {
            f.write((const char*)&it->first, sizeof(uint64_t));
            f.write((const char*)it->second->cells.data(), (std::streamsize)(sizeof(Cell) * patch_volume));   // container.cpp:160
            f.write((const char*)it->second->mask.data(), (std::streamsize)(sizeof(uint64_t) * it->second->mask.size()));
        }
        f.close();
        return true;
    }
    bool read(const std::string& filename, void* params, size_t nparams)
    {
        std::ifstream f(filename.c_str(), std::ios::in | std::ios::binary);
        if (!f.is_open()) return false;
        IOHeader header;
        f.read((char*)&header, sizeof(IOHeader));
        if (!f) return false;
        if (header.magic != kMagic || header.version != kIoVersion) return false;
        if (header.cell_size != sizeof(Cell) || header.is_3d) return false;
        resolution   = header.resolution;   // the file holds a float (map.cpp:548)
        scale        = 1.0 / resolution;
        patch_length = header.patch_length;
        patch_volume = patch_length * patch_length;
        log2dim      = (uint32_t)std::log2(patch_length);
        offset       = (double)((kUniversalConstant >> 1) * patch_length);
        if (nparams) f.read((char*)params, (std::streamsize)nparams);
        prev_idx_ = ~0ull; prev_patch_ = nullptr;
        for (size_t i = 0; i < header.num_patches; ++i) {
            uint64_t idx;
            f.read((char*)&idx, sizeof(idx));
            if (!f) return false;
            auto it = patches.insert(std::make_pair(idx, std::make_shared<PatchT>(patch_volume))).first;
            f.read((char*)it->second->cells.data(), (std::streamsize)(sizeof(Cell) * patch_volume));
            f.read((char*)it->second->mask.data(), (std::streamsize)(sizeof(uint64_t) * it->second->mask.size()));
        }
        return true;
    }
    // Map::bounds in cells (map.cpp:139-157): patch granular; false when the map is empty
    bool bounds(Vec3u& mn, Vec3u& mx) const
    {
        if (patches.empty()) return false;
        mn = Vec3u{0xffffffffu, 0xffffffffu, 0}; mx = Vec3u{0, 0, 0};
        for (auto& kv : patches) {
            Vec3u a = p2m(kv.first);
            mn.x = std::min(mn.x, a.x); mn.y = std::min(mn.y, a.y);
            mx.x = std::max(mx.x, a.x); mx.y = std::max(mx.y, a.y);
        }
        mx.x += patch_length; mx.y += patch_length;
        return true;
    }
    // Map::deletePatchAt (map.cpp:465-488)
    bool delete_patch_at(const Vec3u& c)
    {
        auto it = patches.find(m2p(c));
        if (it == patches.end()) return false;
        patches.erase(it);
        prev_idx_ = ~0ull; prev_patch_ = nullptr;   // the reference's cache holds a pointer into the erased node as well; it is never
                                                    // dereferenced before the next lookup of another patch in the code paths restated here
        return true;
    }
    // Map::visit_all_patches (map.cpp:361-367): the anchor cell of every patch
    template <typename F>
    void visit_all_patches(F&& walker) const
    {
        for (auto& kv : patches) walker(p2m(kv.first));
    }
    // Map::visit_all_cells (map.cpp:352-359): every cell whose mask bit is on
    template <typename F>
    void visit_all_cells(F&& walker) const
    {
        for (auto& kv : patches) {
            Vec3u a = p2m(kv.first);
            for (uint32_t ci = 0; ci < patch_volume; ++ci)
                if (kv.second->is_on(ci)) walker(Vec3u{a.x + (ci & (patch_length - 1)), a.y + (ci >> log2dim), 0});
        }
    }

    // Mutable access (map.cpp:371-412): allocate-on-touch, detach shared patch, set the known bit.  The
    // one-entry patch cache mirrors prev_idx_ / prev_patch_ (map.cpp:400-409): it only saves hash lookups.
    Cell* get(const Vec3u& c)
    {
        uint64_t idx = m2p(c);
        if (prev_idx_ != idx || prev_patch_ == nullptr) {
            auto it = patches.find(idx);
            if (it == patches.end()) it = patches.emplace(idx, std::make_shared<PatchT>(patch_volume)).first;
            prev_idx_   = idx;
            prev_patch_ = &it->second;  // unordered_map never moves its nodes
        }
        PatchPtr& p = *prev_patch_;
        if (p.use_count() > 1) {  // cow_ptr.h:104-114
            p = std::make_shared<PatchT>(*p);
            ++detach_count;
        }
        uint32_t ci = m2c(c);
        if (!p->is_on(ci)) p->set_on(ci);  // container.h:102-106
        return &p->cells[ci];
    }
    // Const access (map.cpp:414-455, container.h:119-123): null if patch absent or bit off.
    const Cell* get(const Vec3u& c) const
    {
        uint64_t idx = m2p(c);
        if (prev_idx_ != idx) {
            auto it = patches.find(idx);
            prev_idx_ = idx;
            if (it == patches.end()) {
                prev_patch_ = nullptr;
                return nullptr;
            }
            prev_patch_ = const_cast<PatchPtr*>(&it->second);
        } else if (prev_patch_ == nullptr) {
            return nullptr;
        }
        uint32_t ci = m2c(c);
        if (!(*prev_patch_)->is_on(ci)) return nullptr;
        return &(*prev_patch_)->cells[ci];
    }

    // Integer Bresenham, both endpoints excluded (map.cpp:198-227).
    template <typename F>
    static void compute_ray(const Vec3u& from, const Vec3u& to, F&& cb)
    {
        if (from == to) return;
        int64_t err[3]   = {0, 0, 0};
        int64_t coord[3] = {(int64_t)from.x, (int64_t)from.y, (int64_t)from.z};
        int64_t delta[3] = {(int64_t)to.x - coord[0], (int64_t)to.y - coord[1], (int64_t)to.z - coord[2]};
        int64_t step[3];
        for (int j = 0; j < 3; ++j) {
            step[j]  = delta[j] < 0 ? -1 : 1;
            delta[j] = delta[j] < 0 ? -delta[j] : delta[j];
        }
        int n = (int)std::max(delta[0], std::max(delta[1], delta[2]));
        for (int i = 0; i < n - 1; ++i) {
            for (int j = 0; j < 3; ++j) err[j] += delta[j];
            for (int j = 0; j < 3; ++j) {
                if ((err[j] << 1) < n) continue;
                coord[j] += step[j];
                err[j] -= n;
            }
            cb(Vec3u{(uint32_t)coord[0], (uint32_t)coord[1], (uint32_t)coord[2]});
        }
    }

    size_t num_patches() const { return patches.size(); }
};

// ----------------------------------------------------------------------------------------------
// Occupancy maps
// ----------------------------------------------------------------------------------------------
struct FreqCell {  // include/lama/sdm/frequency_occupancy_map.h:43-46
    uint16_t occupied;
    uint16_t visited;
};

class FrequencyOccupancyMap : public SparseMap<FreqCell> {
public:
    using SparseMap<FreqCell>::SparseMap;
    static constexpr double occ_thresh = 0.25;  // frequency_occupancy_map.cpp:38
    static double prob(const FreqCell& f)       // :40-45
    {
        if (f.visited == 0) return occ_thresh;
        return ((double)f.occupied) / ((double)f.visited);
    }
    bool set_free(const Vec3u& c)  // :65-74
    {
        FreqCell* cell = get(c);
        bool free      = prob(*cell) < occ_thresh;
        cell->visited++;
        if (free) return false;
        return prob(*cell) < occ_thresh;
    }
    bool set_occupied(const Vec3u& c)  // :81-91
    {
        FreqCell* cell = get(c);
        bool occupied  = prob(*cell) > occ_thresh;
        cell->occupied++;
        cell->visited++;
        if (occupied) return false;
        return prob(*cell) > occ_thresh;
    }
    bool is_free(const Vec3u& c) const      // :115-121
    {
        const FreqCell* cell = static_cast<const SparseMap<FreqCell>*>(this)->get(c);
        return cell != nullptr && prob(*cell) < occ_thresh;
    }
    bool is_occupied(const Vec3u& c) const  // :128-134
    {
        const FreqCell* cell = static_cast<const SparseMap<FreqCell>*>(this)->get(c);
        return cell != nullptr && prob(*cell) > occ_thresh;
    }
    bool is_unknown(const Vec3u& c) const   // :141-147
    {
        const FreqCell* cell = static_cast<const SparseMap<FreqCell>*>(this)->get(c);
        return cell == nullptr || cell->visited == 0;
    }
    double get_probability(const Vec3u& c) const  // :166-172
    {
        const FreqCell* cell = static_cast<const SparseMap<FreqCell>*>(this)->get(c);
        return cell == nullptr ? occ_thresh : prob(*cell);
    }
    uint64_t ray_cells = 0;  // work counter C (cells visited by ray casts incl. hit cells)
};

struct ProbCell {  // include/lama/sdm/probabilistic_occupancy_map.h (prob_tag)
    float prob;
};

class ProbabilisticOccupancyMap : public SparseMap<ProbCell> {
public:
    static float logods(const float& p) { return std::log(p / (1.0 - p)); }  // probabilistic_occupancy_map.cpp:43-46
    double miss_, hit_, clamp_min_, clamp_max_, occ_thresh_;
    ProbabilisticOccupancyMap(double res, uint32_t patch_size) : SparseMap<ProbCell>(res, patch_size)
    {
        miss_       = logods(0.4);   // :53-59
        hit_        = logods(0.7);
        clamp_min_  = logods(0.12);
        clamp_max_  = logods(0.97);
        occ_thresh_ = 0.0 * logods(0.5);
    }
    bool set_free(const Vec3u& c)  // :82-91
    {
        ProbCell* cell = get(c);
        bool free      = cell->prob < occ_thresh_;
        cell->prob     = std::max(cell->prob + miss_, clamp_min_);
        if (free) return false;
        return cell->prob < occ_thresh_;
    }
    bool set_occupied(const Vec3u& c)  // :98-107
    {
        ProbCell* cell = get(c);
        bool occupied  = cell->prob > occ_thresh_;
        cell->prob     = std::min(cell->prob + hit_, clamp_max_);
        if (occupied) return false;
        return cell->prob > occ_thresh_;
    }
    bool is_free(const Vec3u& c) const      // :130-136
    {
        const ProbCell* cell = static_cast<const SparseMap<ProbCell>*>(this)->get(c);
        return cell != nullptr && cell->prob < occ_thresh_;
    }
    bool is_occupied(const Vec3u& c) const  // :143-149
    {
        const ProbCell* cell = static_cast<const SparseMap<ProbCell>*>(this)->get(c);
        return cell != nullptr && cell->prob > occ_thresh_;
    }
    bool is_unknown(const Vec3u& c) const   // :156-162
    {
        const ProbCell* cell = static_cast<const SparseMap<ProbCell>*>(this)->get(c);
        return cell == nullptr || cell->prob == occ_thresh_;
    }
    static float prob_of(const float& logods) { return 1.0 - 1.0 / (1.0 + std::exp(logods)); }  // :38-41
    double get_probability(const Vec3u& c) const  // :169-175
    {
        const ProbCell* cell = static_cast<const SparseMap<ProbCell>*>(this)->get(c);
        return cell == nullptr ? prob_of(occ_thresh_) : prob_of(cell->prob);
    }
    uint64_t ray_cells = 0;
};

// src/sdm/simple_occupancy_map.cpp:36-149 : int8 tri-state cell (-1 free, 0 unknown, 1 occupied); Loc2D's static map
struct SimpleCell {
    int8_t v;
};
class SimpleOccupancyMap : public SparseMap<SimpleCell> {
public:
    using SparseMap<SimpleCell>::SparseMap;
    bool set_free(const Vec3u& c) { SimpleCell* cell = get(c); if (cell->v == -1) return false; cell->v = -1; return true; }      // :51-58
    bool set_occupied(const Vec3u& c) { SimpleCell* cell = get(c); if (cell->v == 1) return false; cell->v = 1; return true; }    // :66-73
    bool set_unknown(const Vec3u& c) { SimpleCell* cell = get(c); if (cell->v == 0) return false; cell->v = 0; return true; }     // :81-88
    bool is_free(const Vec3u& c) const                                                                                             // :96-102
    {
        const SimpleCell* cell = static_cast<const SparseMap<SimpleCell>*>(this)->get(c);
        return cell != nullptr && cell->v == -1;
    }
    bool is_free_world(const double p[3]) const { return is_free(w2m(p)); }                                                         // :91-94
    // Map::bounds (map.cpp:119-138, map.h:208-212): patch-granular, in world coordinates
    bool bounds_world(double mn[3], double mx[3]) const
    {
        if (patches.empty()) return false;
        uint32_t lo[2] = {0xffffffffu, 0xffffffffu}, hi[2] = {0, 0};
        for (auto& kv : patches) {
            Vec3u a = p2m(kv.first);
            lo[0] = std::min(lo[0], a.x); lo[1] = std::min(lo[1], a.y);
            hi[0] = std::max(hi[0], a.x); hi[1] = std::max(hi[1], a.y);
        }
        hi[0] += patch_length; hi[1] += patch_length;
        m2w(Vec3u{lo[0], lo[1], 0}, mn);
        m2w(Vec3u{hi[0], hi[1], patch_length}, mx);
        return true;
    }
};

// ----------------------------------------------------------------------------------------------
// Dynamic distance map (Lau et al. dynamic brushfire, 4-neighbourhood in 2-D)
// include/lama/sdm/dynamic_distance_map.h:48-104, src/sdm/dynamic_distance_map.cpp:66-330
// ----------------------------------------------------------------------------------------------
struct DistCell {  // dynamic_distance_map.h:48-53 (10 bytes)
    int16_t ox, oy, oz;
    uint16_t sqdist;
    bool valid_obstacle;
    bool is_queued;
};

class DynamicDistanceMap : public SparseMap<DistCell> {
public:
    uint32_t max_sqdist_ = 100;  // dynamic_distance_map.cpp:38
    uint64_t processed_total = 0;
    size_t peak_queue = 0;  // instrumentation: largest heap size seen

    // Tie-order experiment hook (NOT reference behaviour): when nonzero, entries of equal
    // priority pop in a pseudo-random order instead of libstdc++ heap order.  Call only while
    // both queues are empty.
    uint32_t shuffle_ties = 0;
    void set_shuffle(uint32_t s)
    {
        shuffle_ties = s;
        lower_       = Queue(ComparePrio{s != 0});
        raise_       = Queue(ComparePrio{s != 0});
    }

    DynamicDistanceMap(double res, uint32_t patch_size) : SparseMap<DistCell>(res, patch_size) {}

    void set_max_distance(double d)  // :149-153
    {
        max_sqdist_ = (uint32_t)std::ceil(d * scale);
        max_sqdist_ *= max_sqdist_;
    }
    double max_distance() const { return std::sqrt((double)max_sqdist_) * resolution; }  // :155-158

    // :140-147
    double distance(const Vec3u& c) const
    {
        const DistCell* cell = static_cast<const SparseMap<DistCell>*>(this)->get(c);
        if (cell == nullptr || !cell->valid_obstacle) return std::sqrt((double)max_sqdist_) * resolution;
        return std::sqrt((double)cell->sqdist) * resolution;
    }

    // :66-92 (2-D branch); grad may be null.
    double distance(const double p[3], double* grad) const
    {
        double m[3];
        w2m_nocast(p, m);
        Vec3u d{(uint32_t)m[0], (uint32_t)m[1], (uint32_t)m[2]};
        double mu0 = m[0] - (double)d.x, mu1 = m[1] - (double)d.y;
        double nu0 = 1.0 - mu0, nu1 = 1.0 - mu1;
        double v0 = distance(d);
        double v1 = distance(Vec3u{d.x + 1, d.y, d.z});
        double v2 = distance(Vec3u{d.x, d.y + 1, d.z});
        double v3 = distance(Vec3u{d.x + 1, d.y + 1, d.z});
        double dist = v0 * nu0 * nu1 + v1 * nu1 * mu0 + v2 * nu0 * mu1 + v3 * mu0 * mu1;
        if (grad) {
            grad[0] = -((v0 - v1) * nu1 + (v2 - v3) * mu1) * scale;
            grad[1] = -((v0 - v2) * nu0 + (v1 - v3) * mu0) * scale;
            grad[2] = 0;
        }
        return dist;
    }

    void add_obstacle(const Vec3u& loc)  // :212-226
    {
        DistCell* cell = get(loc);
        if (cell->valid_obstacle && cell->sqdist == 0) return;
        cell->sqdist = 0;
        cell->ox = cell->oy = cell->oz = 0;
        cell->valid_obstacle = true;
        cell->is_queued      = true;
        push(lower_, 0, loc);
    }
    void remove_obstacle(const Vec3u& loc)  // :228-242
    {
        DistCell* cell = get(loc);
        if (!(cell->valid_obstacle && cell->sqdist == 0)) return;
        cell->sqdist = 0;
        cell->ox = cell->oy = cell->oz = 0;
        cell->valid_obstacle = false;
        cell->is_queued      = true;
        push(raise_, 0, loc);
    }

    uint32_t update()  // :160-197
    {
        uint32_t processed = 0;
        while (!raise_.empty()) {
            Vec3u loc = raise_.top().loc;
            raise_.pop();
            DistCell* cur = get(loc);
            ++processed;
            raise(loc, *cur);
        }
        while (!lower_.empty()) {
            Vec3u loc = lower_.top().loc;
            lower_.pop();
            DistCell* cur = get(loc);
            ++processed;
            if (cur->valid_obstacle) {
                Vec3u obs{(uint32_t)((int64_t)loc.x + cur->ox), (uint32_t)((int64_t)loc.y + cur->oy), (uint32_t)((int64_t)loc.z + cur->oz)};
                const DistCell* o = get(obs);
                if (o->sqdist == 0) lower(loc, *cur);
            }
        }
        processed_total += processed;
        return processed;
    }

    size_t queued() const { return raise_.size() + lower_.size(); }

private:
    struct QEntry {  // queue_pair_t, dynamic_distance_map.h:90
        int prio;
        Vec3u loc;
        uint32_t salt;  // only used by the tie-order experiment
    };
    struct ComparePrio {  // dynamic_distance_map.h:92-95 -- compares .first only
        bool shuffle = false;
        bool operator()(const QEntry& l, const QEntry& r) const
        {
            if (shuffle && l.prio == r.prio) return l.salt > r.salt;
            return l.prio > r.prio;
        }
    };
    using Queue = std::priority_queue<QEntry, std::vector<QEntry>, ComparePrio>;
    Queue lower_, raise_;
    uint32_t salt_state_ = 12345;

    void push(Queue& q, int prio, const Vec3u& loc)
    {
        uint32_t salt = 0;
        if (shuffle_ties) {
            salt_state_ = salt_state_ * 1664525u + 1013904223u + shuffle_ties;
            salt        = salt_state_ >> 8;
        }
        q.push(QEntry{prio, loc, salt});
        peak_queue = std::max(peak_queue, q.size());
    }

    static constexpr int kDelta[4][2] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}};  // :40-43

    void raise(const Vec3u& loc, DistCell& current)  // :244-279
    {
        for (int i = 0; i < 4; ++i) {
            Vec3u nl{(uint32_t)((int64_t)loc.x + kDelta[i][0]), (uint32_t)((int64_t)loc.y + kDelta[i][1]), loc.z};
            DistCell* nb = get(nl);
            if (nb->is_queued || !nb->valid_obstacle) continue;
            Vec3u obs{(uint32_t)((int64_t)nl.x + nb->ox), (uint32_t)((int64_t)nl.y + nb->oy), (uint32_t)((int64_t)nl.z + nb->oz)};
            // NB: get() may rehash `patches` but never moves patch storage, so `nb`/`current` stay valid.
            const DistCell* o = get(obs);
            if (!o->valid_obstacle) {
                push(raise_, nb->sqdist, nl);
                nb->sqdist = 0;
                nb->ox = nb->oy = nb->oz = 0;
                nb->valid_obstacle = false;
                nb->is_queued      = true;
            } else if (!nb->is_queued) {
                push(lower_, nb->sqdist, nl);
                nb->is_queued = true;
            }
        }
        current.is_queued = false;
    }

    void lower(const Vec3u& loc, DistCell& current)  // :281-330
    {
        if (!current.is_queued) return;
        for (int i = 0; i < 4; ++i) {
            int64_t dx = kDelta[i][0], dy = kDelta[i][1];
            // only update away from the obstacle (:296)
            if (dx * (int64_t)current.ox > 0 || dy * (int64_t)current.oy > 0) continue;
            int64_t nx = (int64_t)loc.x + dx, ny = (int64_t)loc.y + dy, nz = (int64_t)loc.z;
            Vec3u nl{(uint32_t)nx, (uint32_t)ny, (uint32_t)nz};
            DistCell* nb = get(nl);
            int64_t ox = (int64_t)loc.x + current.ox, oy = (int64_t)loc.y + current.oy, oz = (int64_t)loc.z + current.oz;
            int64_t ddx = nx - ox, ddy = ny - oy, ddz = nz - oz;
            uint32_t new_sqdist = (uint32_t)(ddx * ddx + ddy * ddy + ddz * ddz);
            uint32_t cmp_sqdist = nb->valid_obstacle ? nb->sqdist : max_sqdist_;
            bool overwrite      = new_sqdist < cmp_sqdist;
            if (!overwrite && new_sqdist == nb->sqdist) {
                Vec3u nobs{(uint32_t)(nx + nb->ox), (uint32_t)(ny + nb->oy), (uint32_t)(nz + nb->oz)};
                const DistCell* o = get(nobs);
                if (!nb->valid_obstacle || !(o->valid_obstacle && o->sqdist == 0)) overwrite = true;
            }
            if (overwrite) {
                push(lower_, (int)new_sqdist, nl);
                nb->sqdist         = (uint16_t)new_sqdist;
                nb->valid_obstacle = true;
                nb->ox = (int16_t)(ox - nx);
                nb->oy = (int16_t)(oy - ny);
                nb->oz = (int16_t)(oz - nz);
                nb->is_queued = true;
            }
        }
        current.is_queued = false;
    }
};

// ----------------------------------------------------------------------------------------------
// Image content of sdm::export_to_png (src/sdm/export.cpp:46-96).  The PNG encoding itself is a library call
// (stb) in the reference; what is restated here is the grey image it is given: width = bounds x extent,
// height = bounds y extent, pixel (u, v) at data[u + v * width] (include/lama/image.h:79-80).
// ----------------------------------------------------------------------------------------------
template <class Occ>
inline std::vector<uint8_t> occupancy_image(const Occ& occ, uint32_t& w, uint32_t& h)
{
    Vec3u mn, mx;
    w = h = 0;
    if (!occ.bounds(mn, mx)) return {};
    w = mx.x - mn.x; h = mx.y - mn.y;
    std::vector<uint8_t> img((size_t)w * h, 90);                       // export.cpp:55
    occ.visit_all_cells([&](const Vec3u& c) {
        uint8_t& px = img[(size_t)(c.x - mn.x) + (size_t)(c.y - mn.y) * w];
        if (occ.is_free(c)) px = 255;                                   // :64-69
        else if (occ.is_occupied(c)) px = 0;
        else px = 127;
    });
    return img;
}
template <class Dm>
inline std::vector<uint8_t> distance_image(const Dm& dm, uint32_t& w, uint32_t& h)
{
    Vec3u mn, mx;
    w = h = 0;
    if (!dm.bounds(mn, mx)) return {};
    w = mx.x - mn.x; h = mx.y - mn.y;
    std::vector<uint8_t> img((size_t)w * h, 127);                      // export.cpp:83
    dm.visit_all_cells([&](const Vec3u& c) {
        img[(size_t)(c.x - mn.x) + (size_t)(c.y - mn.y) * w] = (uint8_t)(dm.distance(c) * 255 / dm.max_distance());   // :91
    });
    return img;
}

// ----------------------------------------------------------------------------------------------
// Scan-to-map residual problem (src/match_surface_2d.cpp:42-122)
// ----------------------------------------------------------------------------------------------
struct MatchSurface2D {
    const DynamicDistanceMap* surface;
    const PointCloud* scan;
    SE2 state;
    uint64_t evals = 0;  // residual evaluations (work counter E)

    MatchSurface2D(const DynamicDistanceMap* dm, const PointCloud* pc, const SE2& est) : surface(dm), scan(pc), state(est) {}

    // :42-90 ; J is row-major N x 3 when non-null
    void eval(std::vector<double>& r, std::vector<double>* J)
    {
        Affine3 tf = compose(fixed_tf(state.tx, state.ty, state.r.log()), moving_tf(*scan));
        const size_t n = scan->size();
        r.resize(n);
        if (J) J->resize(n * 3);
        double hit[3], grad[3];
        for (size_t i = 0; i < n; ++i) {
            tf.apply(&scan->pts[3 * i], hit);
            hit[2] = 0.0;
            r[i]   = surface->distance(hit, grad);
            if (J) {
                (*J)[3 * i + 0] = grad[0];
                (*J)[3 * i + 1] = grad[1];
                (*J)[3 * i + 2] = grad[1] * hit[0] - grad[0] * hit[1];
            }
        }
        ++evals;
    }
    // :118-122
    void update(const double h[3]) { state = SE2::exp(h) * state; }
};

// ----------------------------------------------------------------------------------------------
// Robust weights (src/nlls/robust_cost.cpp:36-82)
// ----------------------------------------------------------------------------------------------
struct RobustCost {
    enum Kind { Unit = 0, Cauchy = 1, Huber = 2, Tukey = 3, TDist = 4 } kind = Unit;
    double param = 0;
    double value(double x) const
    {
        switch (kind) {
        case Cauchy: { double c = 1.0 / (param * param); return 1.0 / (1.0 + x * x * c); }                    // :62-73
        case Huber: return (x < param) ? 1.0 : (param / std::fabs(x));                                          // :75-82
        case Tukey: { double bb = param * param, xx = x * x; if (xx <= bb) { double w = 1.0 - xx / bb; return w * w; } return 0.0; }  // :41-55
        case TDist: return ((param + 1.0f) / (param + (x * x)));                                                // :57-64
        default: return 1.0;
        }
    }
};

// ----------------------------------------------------------------------------------------------
// 3x3 dense helpers standing in for Eigen (gauss_newton.cpp:55-66, solver.cpp:133-150)
// ----------------------------------------------------------------------------------------------
// Solve A h = b for symmetric A (lower triangle referenced) with an unpivoted LDL^T; Eigen's
// LDLT pivots on the largest diagonal entry, which agrees to rounding for SPD normal equations.
inline void ldlt_solve3(const double A[9], const double b[3], double h[3])
{
    // pivoted LDLT (symmetric pivoting on max |diag|) to mirror Eigen::LDLT behaviour.
    double M[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[i][j] = (i >= j) ? A[i * 3 + j] : A[j * 3 + i];
    int perm[3] = {0, 1, 2};
    double L[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, D[3] = {0, 0, 0};
    // work on a permuted copy
    for (int k = 0; k < 3; ++k) {
        int piv = k;
        double best = std::fabs(M[k][k]);
        for (int i = k + 1; i < 3; ++i)
            if (std::fabs(M[i][i]) > best) { best = std::fabs(M[i][i]); piv = i; }
        if (piv != k) {
            std::swap(perm[k], perm[piv]);
            for (int j = 0; j < 3; ++j) std::swap(M[k][j], M[piv][j]);
            for (int i = 0; i < 3; ++i) std::swap(M[i][k], M[i][piv]);
            for (int j = 0; j < k; ++j) std::swap(L[k][j], L[piv][j]);
        }
        D[k] = M[k][k];
        for (int i = k + 1; i < 3; ++i) L[i][k] = (D[k] != 0.0) ? M[i][k] / D[k] : 0.0;
        for (int i = k + 1; i < 3; ++i)
            for (int j = k + 1; j < 3; ++j) M[i][j] -= L[i][k] * D[k] * L[j][k];
    }
    double pb[3], y[3], z[3];
    for (int i = 0; i < 3; ++i) pb[i] = b[perm[i]];
    for (int i = 0; i < 3; ++i) {
        y[i] = pb[i];
        for (int j = 0; j < i; ++j) y[i] -= L[i][j] * y[j];
    }
    for (int i = 0; i < 3; ++i) y[i] = (D[i] != 0.0) ? y[i] / D[i] : 0.0;
    for (int i = 2; i >= 0; --i) {
        z[i] = y[i];
        for (int j = i + 1; j < 3; ++j) z[i] -= L[j][i] * z[j];
    }
    for (int i = 0; i < 3; ++i) h[perm[i]] = z[i];
}

inline bool inverse3(const double A[9], double inv[9])
{
    double M[3][6];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { M[i][j] = A[i * 3 + j]; M[i][3 + j] = (i == j) ? 1.0 : 0.0; }
    for (int k = 0; k < 3; ++k) {
        int piv = k;
        for (int i = k + 1; i < 3; ++i)
            if (std::fabs(M[i][k]) > std::fabs(M[piv][k])) piv = i;
        if (M[piv][k] == 0.0) return false;
        if (piv != k) for (int j = 0; j < 6; ++j) std::swap(M[k][j], M[piv][j]);
        double d = M[k][k];
        for (int j = 0; j < 6; ++j) M[k][j] /= d;
        for (int i = 0; i < 3; ++i) {
            if (i == k) continue;
            double f = M[i][k];
            for (int j = 0; j < 6; ++j) M[i][j] -= f * M[k][j];
        }
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) inv[i * 3 + j] = M[i][3 + j];
    return true;
}

// Jacobi eigen-decomposition of a symmetric 3x3 (for the SVD fallback of solver.cpp:141-148).
inline void sym_eig3(const double A[9], double w[3], double V[9])
{
    double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[i][j] = A[i * 3 + j];
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0.0) continue;
                double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                double t = ((theta >= 0) ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) { double akp = a[k][p], akq = a[k][q]; a[k][p] = c * akp - s * akq; a[k][q] = s * akp + c * akq; }
                for (int k = 0; k < 3; ++k) { double apk = a[p][k], aqk = a[q][k]; a[p][k] = c * apk - s * aqk; a[q][k] = s * apk + c * aqk; }
                for (int k = 0; k < 3; ++k) { double vkp = v[k][p], vkq = v[k][q]; v[k][p] = c * vkp - s * vkq; v[k][q] = s * vkp + c * vkq; }
            }
    }
    for (int i = 0; i < 3; ++i) { w[i] = a[i][i]; for (int j = 0; j < 3; ++j) V[i * 3 + j] = v[i][j]; }
}

// ----------------------------------------------------------------------------------------------
// Solver: strategies (gauss_newton.cpp:38-91, levenberg_marquardt.cpp:39-107) and the
// iteration control of solver.cpp:53-158.
// ----------------------------------------------------------------------------------------------
struct Strategy {
    enum Kind { GaussNewton = 0, LevenbergMarquardt = 1 } kind = GaussNewton;
    double eps1 = 1e-4, eps2 = 1e-4, tau = 1e-4;
    // state
    bool stop_ = false;
    double chi2_ = 0, mu_ = -1, v_ = 2.0;
    double g_[3] = {0, 0, 0}, h_[3] = {0, 0, 0};

    void reset() { stop_ = false; mu_ = -1; v_ = 2.0; }
    bool stop() const { return stop_; }

    static void normal_eq(const std::vector<double>& r, const std::vector<double>& J, double g[3], double A[9], double& chi2)
    {
        const size_t n = r.size();
        g[0] = g[1] = g[2] = 0;
        for (int i = 0; i < 9; ++i) A[i] = 0;
        chi2 = 0;
        for (size_t i = 0; i < n; ++i) {
            const double* j = &J[3 * i];
            for (int a = 0; a < 3; ++a) g[a] += j[a] * r[i];
            chi2 += r[i] * r[i];
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) A[a * 3 + b] += j[a] * j[b];
        }
    }

    // returns false when the step must not be applied (stop raised inside step()).
    void step(const std::vector<double>& r, const std::vector<double>& J, double h[3])
    {
        double A[9];
        normal_eq(r, J, g_, A, chi2_);
        double max_abs_g = std::max(std::fabs(g_[0]), std::max(std::fabs(g_[1]), std::fabs(g_[2])));
        if (max_abs_g < eps1) {
            stop_ = true;
            h[0] = h[1] = h[2] = 0;
            return;
        }
        if (kind == LevenbergMarquardt) {
            if (mu_ < 0) mu_ = tau * std::max(A[0], std::max(A[4], A[8]));
            A[0] += mu_; A[4] += mu_; A[8] += mu_;
        }
        double ng[3] = {-g_[0], -g_[1], -g_[2]};
        ldlt_solve3(A, ng, h);
        h_[0] = h[0]; h_[1] = h[1]; h_[2] = h[2];
        double max_abs_h = std::max(std::fabs(h[0]), std::max(std::fabs(h[1]), std::fabs(h[2])));
        if (max_abs_h < eps2) stop_ = true;
    }

    bool valid(const std::vector<double>& ur)
    {
        if (stop_) return true;
        double n2 = 0;
        for (double v : ur) n2 += v * v;
        double dF = chi2_ - n2;
        if (kind == GaussNewton) {  // gauss_newton.cpp:75-86
            if (dF > 0) return true;
            stop_ = true;
            return false;
        }
        // levenberg_marquardt.cpp:83-102
        double dL = 0;
        for (int i = 0; i < 3; ++i) dL += h_[i] * (mu_ * h_[i] - g_[i]);
        dL *= 0.5;
        if (dL > 0.0 && dF > 0.0) {
            mu_ = mu_ * std::max(1.0 / 3.0, 1 - std::pow(2 * (dF / dL) - 1, 3));
            v_  = 2.0;
            return true;
        }
        mu_ = mu_ * v_;
        v_  = 2 * v_;
        return false;
    }
};

struct SolverOptions {
    uint32_t max_iterations = 100;
    Strategy strategy;
    RobustCost robust;
};

struct SolveStats {
    uint32_t iterations = 0;
    uint32_t evals      = 0;
};

// solver.cpp:53-131 ; cov (3x3 row-major) optional.
inline SolveStats solve(const SolverOptions& opt, MatchSurface2D& problem, double* cov)
{
    std::vector<double> r, ur, J;
    double h[3];
    Strategy strategy = opt.strategy;
    strategy.reset();
    bool valid    = true;
    uint32_t iter = 0;
    uint64_t e0   = problem.evals;
    while (!strategy.stop() && iter < opt.max_iterations) {
        if (valid) {
            problem.eval(r, &J);
            for (size_t i = 0; i < r.size(); ++i) {
                double w = std::sqrt(opt.robust.value(r[i]));
                r[i] *= w;
                J[3 * i + 0] *= w; J[3 * i + 1] *= w; J[3 * i + 2] *= w;
            }
        }
        strategy.step(r, J, h);
        if (strategy.stop()) break;
        problem.update(h);
        problem.eval(ur, nullptr);
        for (size_t i = 0; i < ur.size(); ++i) {
            double w = std::sqrt(opt.robust.value(ur[i]));
            ur[i] *= w;
        }
        valid = strategy.valid(ur);
        if (!valid) {
            double nh[3] = {-h[0], -h[1], -h[2]};
            problem.update(nh);
        }
        ++iter;
    }
    if (cov) {
        problem.eval(r, &J);
        for (size_t i = 0; i < r.size(); ++i) {
            double w = std::sqrt(opt.robust.value(r[i]));
            J[3 * i + 0] *= w; J[3 * i + 1] *= w; J[3 * i + 2] *= w;
        }
        // solver.cpp:133-150 : (J^T J)^-1 when J has full column rank, else thin-SVD pseudo-inverse.
        double g[3], A[9], chi2;
        Strategy::normal_eq(r, J, g, A, chi2);
        double w[3], V[9];
        sym_eig3(A, w, V);
        double wmax = std::max(w[0], std::max(w[1], w[2])), wmin = std::min(w[0], std::min(w[1], w[2]));
        double thr  = std::numeric_limits<double>::epsilon() * (double)std::max<size_t>(r.size(), 3);
        bool full   = wmax > 0 && std::sqrt(std::max(wmin, 0.0)) > thr * std::sqrt(wmax);
        if (!(full && inverse3(A, cov))) {
            double f[3];
            for (int i = 0; i < 3; ++i) {
                double sv = std::sqrt(std::max(w[i], 0.0));
                f[i] = (std::fabs(sv) > 1.e-3) ? 1.0 / (sv * sv) : 3.0;
            }
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                    double s = 0;
                    for (int k = 0; k < 3; ++k) s += V[i * 3 + k] * f[k] * V[j * 3 + k];
                    cov[i * 3 + j] = s;
                }
        }
    }
    SolveStats st;
    st.iterations = iter;
    st.evals      = (uint32_t)(problem.evals - e0);
    return st;
}

// ----------------------------------------------------------------------------------------------
// GraphSlam2D's loop-closure front end (src/graph_slam2d.cpp:283-392)
// ----------------------------------------------------------------------------------------------
// MatchSurface2D::error (src/match_surface_2d.cpp:92-116)
inline double match_error(const DynamicDistanceMap& dm, const PointCloud& pc, const SE2& state)
{
    Affine3 tf = compose(fixed_tf(state.tx, state.ty, state.r.log()), moving_tf(pc));
    double sq = 0, hit[3];
    for (size_t i = 0; i < pc.size(); ++i) {
        tf.apply(&pc.pts[3 * i], hit);
        const double d = dm.distance(dm.w2m(hit));
        sq += d * d;
    }
    return std::sqrt(sq / (double)pc.size());
}
// findLoopClosureCandidates (:283-313): nanoflann radius search (squared L2, results sorted by distance) over the first n - ignore key poses
inline std::vector<int> find_loop_closure_candidates(const std::vector<double>& key_xy, int ignore_n, const double query[2], double radius, size_t max_candidates)
{
    std::vector<std::pair<int, double>> results;
    const int n = (int)(key_xy.size() / 2) - ignore_n;
    for (int i = 0; i < n; ++i) {
        const double d0 = query[0] - key_xy[2 * i], d1 = query[1] - key_xy[2 * i + 1];
        const double d2 = d0 * d0 + d1 * d1;
        if (d2 < radius * radius) results.push_back({i, d2});
    }
    std::sort(results.begin(), results.end(), [](const std::pair<int, double>& a, const std::pair<int, double>& b) { return a.second < b.second; });
    if (results.size() > max_candidates) results.erase(results.begin() + (long)max_candidates, results.end());   // :298-301
    std::vector<int> out;
    for (auto& c : results) out.push_back(c.first);
    return out;
}
inline SolverOptions loop_solver(uint32_t max_iter)
{
    SolverOptions so;
    so.max_iterations = max_iter;
    so.strategy.kind  = Strategy::GaussNewton;   // :324
    so.robust.kind    = RobustCost::Huber;       // :325
    so.robust.param   = 0.15;
    return so;
}
// correlateCandidateScan (:315-355)
inline double correlate_candidate_scan(const DynamicDistanceMap& dm, const PointCloud& cloud, const Pose2D& ref_pose, const Pose2D& candidate_pose, Pose2D& between)
{
    MatchSurface2D ms0(&dm, &cloud, candidate_pose.state);
    MatchSurface2D ms1(&dm, &cloud, Pose2D(ref_pose.x(), ref_pose.y(), candidate_pose.rotation()).state);
    solve(loop_solver(1), ms0, nullptr);
    const double rmse0 = match_error(dm, cloud, ms0.state);
    solve(loop_solver(1), ms1, nullptr);
    const double rmse1 = match_error(dm, cloud, ms1.state);
    MatchSurface2D* msp = rmse0 < rmse1 ? &ms0 : &ms1;
    solve(loop_solver(100), *msp, nullptr);
    between = Pose2D(msp->state).minus(ref_pose);
    return match_error(dm, cloud, msp->state);
}
// coarseSearchAndCorrelateCandidateScan (:357-392)
inline double coarse_correlate_candidate_scan(const DynamicDistanceMap& map, const PointCloud& ref_cloud, const PointCloud& cloud, const Pose2D& ref_pose,
                                              const Pose2D& candidate_pose, Pose2D& between)
{
    Affine3 tf = compose(fixed_tf(ref_pose.x(), ref_pose.y(), ref_pose.state.r.log()), moving_tf(ref_cloud));
    DynamicDistanceMap dm(0.25, 32);
    dm.set_max_distance(2.5);
    double hit[3];
    for (size_t i = 0; i < ref_cloud.size(); ++i) {
        tf.apply(&ref_cloud.pts[3 * i], hit);
        dm.add_obstacle(dm.w2m(hit));
    }
    dm.update();
    MatchSurface2D ms(&dm, &cloud, candidate_pose.state);
    solve(loop_solver(100), ms, nullptr);
    ms.surface = &map;
    solve(loop_solver(100), ms, nullptr);
    between = Pose2D(ms.state).minus(ref_pose);
    return match_error(map, cloud, ms.state);
}

// ----------------------------------------------------------------------------------------------
// Thread pool used by the timed CPU baseline: one task per particle per phase, wait() barrier
// (src/thread_pool.cpp:52-114, src/pf_slam2d.cpp:254-266,292-302).
// ----------------------------------------------------------------------------------------------
class ThreadPool {
public:
    explicit ThreadPool(size_t n)
    {
        if (n == 0) n = std::thread::hardware_concurrency();
        for (size_t i = 0; i < n; ++i) workers_.emplace_back([this] { run(); });
    }
    ~ThreadPool()
    {
        {
            std::unique_lock<std::mutex> l(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    void enqueue(std::function<void()> f)
    {
        {
            std::unique_lock<std::mutex> l(m_);
            tasks_.push(std::move(f));
            ++pending_;
        }
        cv_.notify_one();
    }
    void wait()
    {
        std::unique_lock<std::mutex> l(m_);
        done_cv_.wait(l, [this] { return pending_ == 0; });
    }
    size_t size() const { return workers_.size(); }

private:
    void run()
    {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [this] { return stop_ || !tasks_.empty(); });
                if (stop_ && tasks_.empty()) return;
                f = std::move(tasks_.front());
                tasks_.pop();
            }
            f();
            {
                std::unique_lock<std::mutex> l(m_);
                if (--pending_ == 0) done_cv_.notify_all();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::queue<std::function<void()>> tasks_;
    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    size_t pending_ = 0;
    bool stop_      = false;
};

// ----------------------------------------------------------------------------------------------
// Map update shared by PFSlam2D::updateParticleMaps (pf_slam2d.cpp:439-509) and
// Slam2D::updateMaps (slam2d.cpp:247-321).
// ----------------------------------------------------------------------------------------------
struct MapUpdateCounters {
    uint64_t ray_cells = 0;   // C
    uint64_t dm_pops   = 0;   // W
};

// `lo_ray`: LidarOdometry2D's own ray shortening (lidar_odometry_2d.cpp:103-110: `if (ray_length >= 1.0) start = hit - AB / ray_length`)
// instead of the truncated_ray / truncated_range logic; `hit_min` / `hit_max`: the surface AABB of the transient-map code
// (slam2d.cpp:264-267,303-306, lidar_odometry_2d.cpp:96-99,113-114).
template <typename OccMap>
inline uint32_t update_maps(OccMap& occ, DynamicDistanceMap& dm, const PointCloud& surface, const Pose2D& pose,
                            double truncated_ray, double truncated_range, MapUpdateCounters* ctr, bool lo_ray = false, double* hit_min = nullptr,
                            double* hit_max = nullptr)
{
    if (hit_min)
        for (int k = 0; k < 3; ++k) { hit_min[k] = std::numeric_limits<double>::max(); hit_max[k] = -std::numeric_limits<double>::max(); }
    Affine3 tf = compose(fixed_tf(pose.x(), pose.y(), pose.rotation()), moving_tf(surface));
    const double wso[3] = {tf.t[0], tf.t[1], tf.t[2]};
    const size_t n = surface.size();
    uint64_t cells = 0;
    for (size_t i = 0; i < n; ++i) {
        double start[3] = {wso[0], wso[1], wso[2]};
        double hit[3], AB[3] = {0, 0, 0};
        tf.apply(&surface.pts[3 * i], hit);
        double ray_length = 1.0;
        bool mark_hit     = true;
        if (truncated_range > 0.0) {
            for (int k = 0; k < 3; ++k) AB[k] = hit[k] - start[k];
            ray_length = std::sqrt(AB[0] * AB[0] + AB[1] * AB[1] + AB[2] * AB[2]);
            if (truncated_range < ray_length) {
                for (int k = 0; k < 3; ++k) hit[k] = start[k] + AB[k] / ray_length * truncated_range;
                mark_hit = false;
            }
        }
        if (mark_hit && truncated_ray > 0.0) {
            if (truncated_range == 0.0) {
                for (int k = 0; k < 3; ++k) AB[k] = hit[k] - start[k];
                ray_length = std::sqrt(AB[0] * AB[0] + AB[1] * AB[1] + AB[2] * AB[2]);
            }
            if (truncated_ray < ray_length)
                for (int k = 0; k < 3; ++k) start[k] = hit[k] - AB[k] / ray_length * truncated_ray;
        }
        if (lo_ray) {
            for (int k = 0; k < 3; ++k) AB[k] = hit[k] - start[k];
            ray_length = std::sqrt(AB[0] * AB[0] + AB[1] * AB[1] + AB[2] * AB[2]);
            if (ray_length >= 1.0)
                for (int k = 0; k < 3; ++k) start[k] = hit[k] - AB[k] / ray_length;
        }
        Vec3u mhit = occ.w2m(hit);
        if (hit_min)
            for (int k = 0; k < 3; ++k) { hit_min[k] = std::min(hit_min[k], hit[k]); hit_max[k] = std::max(hit_max[k], hit[k]); }
        if (mark_hit) {
            ++cells;
            if (occ.set_occupied(mhit)) dm.add_obstacle(mhit);
        }
        OccMap::compute_ray(occ.w2m(start), mhit, [&](const Vec3u& c) {
            ++cells;
            if (occ.set_free(c)) dm.remove_obstacle(c);
        });
    }
    uint32_t processed = dm.update();
    if (ctr) {
        ctr->ray_cells += cells;
        ctr->dm_pops += processed;
    }
    return processed;
}

// Transient map (slam2d.cpp:323-379 with `stretch` 2, lidar_odometry_2d.cpp:130-181 with `stretch` 1): keep only the patches whose
// AABB meets the AABB of the latest surface, made symmetric around the pose and grown by twice the distance map's reach.
// AABB::testIntersection: include/lama/aabb.h:65-72.
template <typename OccMap>
inline size_t transient_prune(OccMap& occ, DynamicDistanceMap& dm, double px, double py, double mn[3], double mx[3], double stretch)
{
    mn[2] = mx[2] = 0;
    const double xdist = std::max(px - mn[0], mx[0] - px) * stretch;
    const double ydist = std::max(py - mn[1], mx[1] - py) * stretch;
    mn[0] = px - xdist; mn[1] = py - ydist;
    mx[0] = px + xdist; mx[1] = py + ydist;
    double ac[3], ah[3];
    for (int k = 0; k < 3; ++k) { ah[k] = (mx[k] - mn[k]) * 0.5; ac[k] = mn[k] + ah[k]; }       // AABB(min, max), aabb.h:50-55
    for (int k = 0; k < 3; ++k) ah[k] += 2.0 * dm.max_distance();
    std::vector<Vec3u> to_remove;
    dm.visit_all_patches([&](const Vec3u& origin) {
        const uint32_t length = occ.patch_length;
        double ws[3], we[3];
        occ.m2w(origin, ws);
        occ.m2w(Vec3u{origin.x + length, origin.y + length, origin.z}, we);   // origin + Vector3ui(length, length, 0.0)
        ws[2] = we[2] = 0.0;
        bool hitb = true;
        for (int k = 0; k < 3; ++k) {
            const double bh = (we[k] - ws[k]) * 0.5, bc = ws[k] + bh;
            hitb = hitb && (std::abs(ac[k] - bc) <= (ah[k] + bh));
        }
        if (!hitb) to_remove.push_back(origin);
    });
    for (auto& c : to_remove) {
        occ.delete_patch_at(c);
        dm.delete_patch_at(c);
    }
    return to_remove.size();
}

// ----------------------------------------------------------------------------------------------
// PFSlam2D (include/lama/pf_slam2d.h:132-185, src/pf_slam2d.cpp:106-138,178-312,365-574)
// ----------------------------------------------------------------------------------------------
struct PFOptions {
    uint32_t particles      = 1;
    double srr = 0.1, str = 0.2, stt = 0.1, srt = 0.2;
    double meas_sigma       = 0.05;
    double meas_sigma_gain  = 3;
    double trans_thresh     = 0.5;
    double rot_thresh       = 0.5;
    double l2_max           = 0.5;
    double truncated_ray    = 0.0;
    double truncated_range  = 0.0;
    double resolution       = 0.05;
    uint32_t patch_size     = 32;
    uint32_t max_iter       = 100;
    int32_t threads         = -1;
    uint32_t seed           = 0;
};

struct Particle {  // pf_slam2d.h:67-86
    double weight = 0, normalized_weight = 0, weight_sum = 0;
    Pose2D pose;
    std::vector<Pose2D> poses;
    std::shared_ptr<DynamicDistanceMap> dm;
    std::shared_ptr<FrequencyOccupancyMap> occ;
};

struct ScanCounters {
    uint64_t evals = 0, ray_cells = 0, dm_pops = 0, detached = 0, gn_iters = 0;
    int resampled = 0;
};

class PFSlam2D {
public:
    PFOptions opt;
    Random rng;
    std::vector<Particle> particles[2];
    int cur = 0;
    bool has_first_scan = false;
    Pose2D odom, pose;  // pose = prior
    double acc_trans = 0, acc_rot = 0, neff = 0;
    const PointCloud* surface = nullptr;
    std::unique_ptr<ThreadPool> pool;
    std::vector<int32_t> last_sample_idx;
    ScanCounters last, total;
    double t_solve = 0, t_norm = 0, t_resample = 0, t_map = 0;  // Summary buckets (pf_slam2d.h:88-129)
    uint32_t shuffle_ties = 0;

    explicit PFSlam2D(const PFOptions& o) : opt(o)
    {
        if (opt.threads > 1) pool.reset(new ThreadPool(opt.threads));  // pf_slam2d.cpp:123-128
        if (opt.seed == 0) opt.seed = std::random_device{}();        // :131-132
        rng.seed(opt.seed);                                           // :134
    }
    void set_prior(const Pose2D& p) { pose = p; }

    bool update(const PointCloud& pc, const Pose2D& odometry)  // :178-312
    {
        using clk = std::chrono::steady_clock;
        surface = &pc;
        last    = ScanCounters();
        const uint32_t P = opt.particles;
        if (!has_first_scan) {
            odom = odometry;
            particles[0].assign(P, Particle());
            cur = 0;
            Particle& p0 = particles[0][0];
            p0.poses.push_back(pose);
            p0.pose = pose;
            p0.dm.reset(new DynamicDistanceMap(opt.resolution, opt.patch_size));
            p0.dm->set_max_distance(opt.l2_max);
            if (shuffle_ties) p0.dm->set_shuffle(shuffle_ties);
            p0.occ.reset(new FrequencyOccupancyMap(opt.resolution, opt.patch_size));
            update_particle_maps(&p0);
            for (uint32_t i = 1; i < P; ++i) {
                Particle& pi = particles[0][i];
                pi.poses.push_back(pose);
                pi.pose = pose;
                pi.dm.reset(new DynamicDistanceMap(*p0.dm));
                pi.occ.reset(new FrequencyOccupancyMap(*p0.occ));
            }
            has_first_scan = true;
            accumulate();
            return true;
        }
        // 1. predict
        Pose2D odelta = odom.minus(odometry);
        odom          = odometry;
        for (uint32_t i = 0; i < P; ++i) draw_from_motion(odelta, particles[cur][i].pose);
        acc_trans += odelta.xy_norm();
        acc_rot += std::fabs(odelta.rotation());
        if (acc_trans <= opt.trans_thresh && acc_rot <= opt.rot_thresh) return false;
        acc_trans = 0;
        acc_rot   = 0;
        // 2. scan matching
        auto t0 = clk::now();
        for_each_particle([this](Particle* p) { scan_match(p); });
        auto t1 = clk::now();
        // 3. normalize
        normalize();
        auto t2 = clk::now();
        // 4. resample
        last_sample_idx.clear();
        if (neff < (opt.particles * 0.5)) {
            resample();
            last.resampled = 1;
        }
        auto t3 = clk::now();
        // 5. maps
        for_each_particle([this](Particle* p) { update_particle_maps(p); });
        auto t4 = clk::now();
        t_solve += std::chrono::duration<double>(t1 - t0).count();
        t_norm += std::chrono::duration<double>(t2 - t1).count();
        t_resample += std::chrono::duration<double>(t3 - t2).count();
        t_map += std::chrono::duration<double>(t4 - t3).count();
        accumulate();
        return true;
    }

    size_t best_particle_idx() const  // :314-330
    {
        size_t best = 0;
        double ws   = particles[cur][0].weight_sum;
        for (uint32_t i = 1; i < opt.particles; ++i)
            if (ws < particles[cur][i].weight_sum) {
                ws   = particles[cur][i].weight_sum;
                best = i;
            }
        return best;
    }

private:
    std::mutex ctr_mutex_;

    template <typename F>
    void for_each_particle(F&& f)
    {
        const uint32_t P = opt.particles;
        if (pool) {
            for (uint32_t i = 0; i < P; ++i) {
                Particle* p = &particles[cur][i];
                pool->enqueue([f, p] { f(p); });
            }
            pool->wait();
        } else {
            for (uint32_t i = 0; i < P; ++i) f(&particles[cur][i]);
        }
    }
    void accumulate()
    {
        total.evals += last.evals;
        total.ray_cells += last.ray_cells;
        total.dm_pops += last.dm_pops;
        total.detached += last.detached;
        total.gn_iters += last.gn_iters;
        total.resampled += last.resampled;
    }

    void draw_from_motion(const Pose2D& delta, Pose2D& p)  // :365-391
    {
        double sigma, x, y, yaw;
        double sxy = 0.3 * opt.stt;
        sigma = opt.stt * std::fabs(delta.x()) + opt.str * std::fabs(delta.rotation()) + sxy * std::fabs(delta.y());
        x     = delta.x() + rng.normal(sigma);
        sigma = opt.stt * std::fabs(delta.y()) + opt.str * std::fabs(delta.rotation()) + sxy * std::fabs(delta.x());
        y     = delta.y() + rng.normal(sigma);
        sigma = opt.srr * std::fabs(delta.rotation()) + opt.srt * delta.xy_norm();
        yaw   = delta.rotation() + rng.normal(sigma);
        yaw   = std::fmod(yaw, 2 * M_PI);
        if (yaw > M_PI) yaw -= 2 * M_PI;
        p.plus_assign(Pose2D(x, y, yaw));
    }

    double calculate_likelihood(const Particle& p)  // :393-414
    {
        Affine3 tf = compose(fixed_tf(p.pose.x(), p.pose.y(), p.pose.rotation()), moving_tf(*surface));
        const size_t n    = surface->size();
        double likelihood = 0;
        double hit[3];
        for (size_t i = 0; i < n; ++i) {
            tf.apply(&surface->pts[3 * i], hit);
            double dist = p.dm->distance(hit, nullptr);
            likelihood += -(dist * dist) / opt.meas_sigma;
        }
        return likelihood;
    }

    void scan_match(Particle* p)  // :416-437
    {
        MatchSurface2D ms(p->dm.get(), surface, p->pose.state);
        SolverOptions so;
        so.max_iterations = opt.max_iter;
        so.strategy.kind  = Strategy::GaussNewton;
        so.robust.kind    = RobustCost::Cauchy;
        so.robust.param   = 0.15;
        SolveStats st     = solve(so, ms, nullptr);
        p->pose.state     = ms.state;
        p->poses.push_back(p->pose);
        double l = calculate_likelihood(*p);
        p->weight_sum += l;
        p->weight += l;
        std::unique_lock<std::mutex> lk(ctr_mutex_);
        last.evals += st.evals + 1;
        last.gn_iters += st.iterations;
    }

    void update_particle_maps(Particle* p)  // :439-509
    {
        MapUpdateCounters c;
        uint64_t d0 = p->dm->detach_count + p->occ->detach_count;
        update_maps(*p->occ, *p->dm, *surface, p->pose, opt.truncated_ray, opt.truncated_range, &c);
        uint64_t d1 = p->dm->detach_count + p->occ->detach_count;
        std::unique_lock<std::mutex> lk(ctr_mutex_);
        last.ray_cells += c.ray_cells;
        last.dm_pops += c.dm_pops;
        last.detached += d1 - d0;
    }

    void normalize()  // :511-535
    {
        const uint32_t P = opt.particles;
        auto& ps         = particles[cur];
        double gain      = 1.0 / (opt.meas_sigma_gain * opt.particles);
        double max_l     = ps[0].weight;
        for (uint32_t i = 1; i < P; ++i)
            if (max_l < ps[i].weight) max_l = ps[i].weight;
        double sum = 0;
        for (uint32_t i = 0; i < P; ++i) {
            ps[i].normalized_weight = std::exp(gain * (ps[i].weight - max_l));
            sum += ps[i].normalized_weight;
        }
        neff = 0;
        for (uint32_t i = 0; i < P; ++i) {
            ps[i].normalized_weight /= sum;
            neff += ps[i].normalized_weight * ps[i].normalized_weight;
        }
        neff = 1.0 / neff;
    }

    void resample()  // :537-574
    {
        const uint32_t P = opt.particles;
        std::vector<int32_t> sample_idx(P);
        double interval = 1.0 / (double)P;
        double target   = interval * rng.uniform();
        double cw       = 0.0;
        uint32_t n      = 0;
        for (size_t i = 0; i < P; ++i) {
            cw += particles[cur][i].normalized_weight;
            while (cw > target) {
                if (n < P) sample_idx[n] = (int32_t)i;  // guard: the reference can overrun here
                ++n;
                target += interval;
            }
        }
        int ps = 1 - cur;
        particles[ps].assign(P, Particle());
        for (size_t i = 0; i < P; ++i) {
            uint32_t idx          = sample_idx[i];
            particles[ps][i]      = particles[cur][idx];
            particles[ps][i].weight     = 0.0;
            particles[ps][i].weight_sum = particles[cur][idx].weight_sum;
            particles[ps][i].dm.reset(new DynamicDistanceMap(*particles[cur][idx].dm));
            particles[ps][i].occ.reset(new FrequencyOccupancyMap(*particles[cur][idx].occ));
        }
        particles[cur].clear();
        cur             = ps;
        last_sample_idx = sample_idx;
    }
};

// ----------------------------------------------------------------------------------------------
// Slam2D (include/lama/slam2d.h:91-125, src/slam2d.cpp:92-121,143-198,247-321)
// ----------------------------------------------------------------------------------------------
struct SlamOptions {
    double trans_thresh = 0.5, rot_thresh = 0.5, l2_max = 0.5;
    double truncated_ray = 0.0, truncated_range = 0.0;
    double resolution = 0.05;
    uint32_t patch_size = 32, max_iter = 100;
    int strategy = 0;  // 0 = "gn", 1 = "lm" (slam2d.cpp:226-233)
    bool transient_map = false;  // slam2d.h:122
};

// OccMap = FrequencyOccupancyMap is the reference's Slam2D (slam2d.cpp:97); OccMap = ProbabilisticOccupancyMap is the
// same front end over the log-odds map (the combination LidarOdometry2D uses, lidar_odometry_2d.cpp:46), kept so that
// row a18 of SURVEY 8(a) has a complete update path to be checked against.
template <typename OccMap>
class Slam2DT {
public:
    SlamOptions opt;
    DynamicDistanceMap dm;
    OccMap occ;
    SolverOptions so;
    Pose2D pose, odom;
    bool has_first_scan = false;
    uint32_t processed_cells = 0;
    uint64_t removed_patches = 0;
    ScanCounters last, total;

    explicit Slam2DT(const SlamOptions& o) : opt(o), dm(o.resolution, o.patch_size), occ(o.resolution, o.patch_size)
    {
        dm.set_max_distance(o.l2_max);
        so.max_iterations = o.max_iter;
        so.strategy.kind  = o.strategy == 1 ? Strategy::LevenbergMarquardt : Strategy::GaussNewton;
        so.robust.kind    = RobustCost::Cauchy;
        so.robust.param   = 0.15;
    }
    bool update(const PointCloud& pc, const Pose2D& odometry)  // :143-198
    {
        last = ScanCounters();
        if (!has_first_scan) {
            odom = odometry;
            update_maps_(pc);
            has_first_scan = true;
            return true;
        }
        Pose2D odelta = odom.minus(odometry);
        Pose2D ppose  = pose.plus(odelta);
        if (odelta.xy_norm() <= opt.trans_thresh && std::abs(odelta.rotation()) <= opt.rot_thresh) return false;
        pose = ppose;
        odom = odometry;
        MatchSurface2D ms(&dm, &pc, pose.state);
        SolveStats st = solve(so, ms, nullptr);
        pose.state    = ms.state;
        last.evals += st.evals;
        last.gn_iters += st.iterations;
        update_maps_(pc);
        return true;
    }

private:
    void update_maps_(const PointCloud& pc)
    {
        MapUpdateCounters c;
        double mn[3], mx[3];
        processed_cells = update_maps(occ, dm, pc, pose, opt.truncated_ray, opt.truncated_range, &c, false, opt.transient_map ? mn : nullptr,
                                      opt.transient_map ? mx : nullptr);
        if (opt.transient_map) removed_patches += transient_prune(occ, dm, pose.x(), pose.y(), mn, mx, 2.0);   // slam2d.cpp:329-379
        last.ray_cells += c.ray_cells;
        last.dm_pops += c.dm_pops;
        total.evals += last.evals;
        total.ray_cells += last.ray_cells;
        total.dm_pops += last.dm_pops;
        total.gn_iters += last.gn_iters;
    }
};
using Slam2D     = Slam2DT<FrequencyOccupancyMap>;
using Slam2DProb = Slam2DT<ProbabilisticOccupancyMap>;

// ----------------------------------------------------------------------------------------------
// LidarOdometry2D (include/lama/lidar_odometry_2d.h:45-75, src/lidar_odometry_2d.cpp:42-181)
// ----------------------------------------------------------------------------------------------
class LidarOdometry2D {
public:
    DynamicDistanceMap dm;
    ProbabilisticOccupancyMap occ;
    SolverOptions so;
    Pose2D odom, map_update_odom;
    bool has_first_scan = false;
    uint64_t removed_patches = 0, map_updates = 0;
    ScanCounters last;

    explicit LidarOdometry2D(double resolution = 0.05, uint32_t max_iter = 100) : dm(resolution, 32), occ(resolution, 32)
    {
        dm.set_max_distance(1.0);                 // :45
        so.max_iterations = max_iter;             // :48-50
        so.strategy.kind  = Strategy::GaussNewton;
        so.robust.kind    = RobustCost::Cauchy;
        so.robust.param   = 0.15;
    }
    bool update(const PointCloud& pc)             // :59-83
    {
        last = ScanCounters();
        if (!has_first_scan) {
            update_maps_(pc);
            has_first_scan = true;
            return true;
        }
        MatchSurface2D ms(&dm, &pc, odom.state);
        SolveStats st = solve(so, ms, nullptr);
        odom.state    = ms.state;
        last.evals += st.evals;
        last.gn_iters += st.iterations;
        Pose2D odelta = map_update_odom.minus(odom);   // map_update_odom - odom, pose2d.cpp:81-84
        if (odelta.xy_norm() > 0.1 || std::abs(odelta.rotation()) > 0.5) {
            update_maps_(pc);
            map_update_odom = odom;
        }
        return true;
    }

private:
    void update_maps_(const PointCloud& pc)       // :85-181
    {
        MapUpdateCounters c;
        double mn[3], mx[3];
        update_maps(occ, dm, pc, odom, 0.0, 0.0, &c, true, mn, mx);
        removed_patches += transient_prune(occ, dm, odom.x(), odom.y(), mn, mx, 1.0);
        last.ray_cells += c.ray_cells;
        last.dm_pops += c.dm_pops;
        ++map_updates;
    }
};

// ----------------------------------------------------------------------------------------------
// Loc2D match path (src/loc2d.cpp:46-108,126-192); global localisation / sampling covariance are
// "next" rows and not restated here.
// ----------------------------------------------------------------------------------------------
struct LocOptions {
    double trans_thresh = 0.5, rot_thresh = 0.5, l2_max = 1.0, resolution = 0.05;
    uint32_t patch_size = 32, max_iter = 100;
    int strategy = 0;
    uint32_t gloc_particles = 3000, gloc_iters = 10;  // loc2d.cpp:53-55
    double gloc_thresh = 0.15, cov_blend = 0.0;       // :55,57
};

class Loc2D {
public:
    LocOptions opt;
    DynamicDistanceMap dm;      // public `distance_map`, filled by the caller (loc2d.h:103-104)
    SimpleOccupancyMap occ;     // public `occupancy_map`
    SolverOptions so;
    Pose2D pose, odom;
    bool has_first_scan = false;
    double rmse = 0;
    double cov[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    SolveStats last_stats;
    Random rng;                 // stands in for the process-global generator (src/random.cpp:38-39)
    bool do_global_localization = false;
    uint32_t gloc_cur_iter = 0;
    double cov_blend;
    std::vector<std::pair<double, double>> sampling_steps;
    uint64_t gloc_evals = 0;

    explicit Loc2D(const LocOptions& o) : opt(o), dm(o.resolution, o.patch_size), occ(o.resolution, o.patch_size)
    {
        dm.set_max_distance(o.l2_max);
        so.max_iterations = o.max_iter;
        so.strategy.kind  = o.strategy == 1 ? Strategy::LevenbergMarquardt : Strategy::GaussNewton;
        so.robust.kind    = RobustCost::Cauchy;
        so.robust.param   = 0.15;
        cov_blend = std::max(std::min(o.cov_blend, 1.0), 0.0);  // loc2d.cpp:90
        // sampling steps cache, loc2d.cpp:93-107
        const double sstep = dm.resolution;
        sampling_steps.push_back({0.0, 0.0});
        for (int i = 1; i <= 20; ++i) {
            sampling_steps.push_back({i * sstep, 0.0});
            sampling_steps.push_back({0.0, i * sstep});
            sampling_steps.push_back({-i * sstep, 0.0});
            sampling_steps.push_back({0.0, -i * sstep});
            sampling_steps.push_back({i * sstep, i * sstep});
            sampling_steps.push_back({-i * sstep, i * sstep});
            sampling_steps.push_back({i * sstep, -i * sstep});
            sampling_steps.push_back({-i * sstep, -i * sstep});
        }
    }
    void set_pose(const Pose2D& p)  // loc2d.h:117-118
    {
        pose           = p;
        has_first_scan = false;
    }
    void trigger_global_localization() { do_global_localization = true; }  // loc2d.cpp:194-197

    bool update(const PointCloud& pc, const Pose2D& odometry, bool force_update)  // :126-192
    {
        if (!has_first_scan) {
            odom           = odometry;
            has_first_scan = true;
            if (!force_update) return true;
            MatchSurface2D ms(&dm, &pc, pose.state);
            std::vector<double> res;
            ms.eval(res, nullptr);
            rmse = rmse_of(res, pc.size());
        }
        Pose2D odelta = odom.minus(odometry);
        Pose2D ppose  = pose.plus(odelta);
        if (!force_update && !(odelta.xy_norm() > opt.trans_thresh || std::abs(odelta.rotation()) > opt.rot_thresh)) return false;
        pose = ppose;
        odom = odometry;
        if (do_global_localization) {  // :154-166
            if (gloc_cur_iter < opt.gloc_iters) {
                gloc_cur_iter++;
                global_localization(pc);
            } else {
                do_global_localization = false;
                gloc_cur_iter          = 0;
            }
        }
        MatchSurface2D ms(&dm, &pc, pose.state);
        last_stats = solve(so, ms, cov);
        pose.state = ms.state;
        if (cov_blend > 0.0) add_sampling_covariance(pc);  // :175-176
        std::vector<double> res;
        ms.eval(res, nullptr);
        rmse = rmse_of(res, pc.size());
        if (do_global_localization && rmse < opt.gloc_thresh) {  // :182-188
            do_global_localization = false;
            gloc_cur_iter          = 0;
        }
        return true;
    }

    // loc2d.cpp:199-247
    void add_sampling_covariance(const PointCloud& pc)
    {
        double K[4] = {0, 0, 0, 0}, u[2] = {0, 0}, sl = 0;
        Affine3 mtf = moving_tf(pc);
        const size_t num_points = pc.size();
        const size_t step = std::max(num_points / 100, size_t(1));
        const double rot = pose.rotation();
        for (size_t i = 0; i < sampling_steps.size(); ++i) {
            const double x = pose.x() + sampling_steps[i].first, y = pose.y() + sampling_steps[i].second;
            Affine3 tf = compose(fixed_tf(x, y, rot), mtf);
            double l = 0.0;
            double hit[3];
            for (size_t k = 0; k < num_points; k += step) {
                tf.apply(&pc.pts[3 * k], hit);
                double dist = dm.distance(dm.w2m(hit));  // nearest cell, no interpolation
                double e = std::exp(-(dist * dist) / 0.01);
                l += e * e * e;
            }
            K[0] = K[0] + x * x * l; K[1] = K[1] + x * y * l; K[2] = K[2] + y * x * l; K[3] = K[3] + y * y * l;
            u[0] = u[0] + x * l; u[1] = u[1] + y * l;
            sl = sl + l;
        }
        const double a = 1.0 / sl, b = 1.0 / (sl * sl);
        const double sc[4] = {a * K[0] - b * u[0] * u[0], a * K[1] - b * u[0] * u[1], a * K[2] - b * u[1] * u[0], a * K[3] - b * u[1] * u[1]};
        const double alpha = cov_blend;
        cov[0] = alpha * sc[0] + (1.0 - alpha) * cov[0]; cov[1] = alpha * sc[1] + (1.0 - alpha) * cov[1];
        cov[3] = alpha * sc[2] + (1.0 - alpha) * cov[3]; cov[4] = alpha * sc[3] + (1.0 - alpha) * cov[4];
    }

    // loc2d.cpp:249-286
    void global_localization(const PointCloud& pc)
    {
        double mn[3], mx[3];
        if (!occ.bounds_world(mn, mx)) return;
        const double diff[2] = {mx[0] - mn[0], mx[1] - mn[1]};
        double best_error = std::numeric_limits<double>::max();
        for (uint32_t i = 0; i < opt.gloc_particles; ++i) {
            double x, y, a;
            for (;;) {
                x = mn[0] + rng.uniform() * diff[0];
                y = mn[1] + rng.uniform() * diff[1];
                const double p[3] = {x, y, 0.0};
                if (!occ.is_free_world(p)) continue;
                a = rng.uniform() * 2 * M_PI - M_PI;
                break;
            }
            Pose2D p(x, y, a);
            MatchSurface2D ms(&dm, &pc, p.state);
            std::vector<double> res;
            ms.eval(res, nullptr);
            ++gloc_evals;
            double error = 0;
            for (double v : res) error += v * v;
            if (error < best_error) {
                best_error = error;
                pose       = p;
            }
        }
    }

private:
    static double rmse_of(const std::vector<double>& r, size_t n)
    {
        double s = 0;
        for (double v : r) s += v * v;
        return std::sqrt(s / ((double)(n - 1)));  // :178-180 (size_t arithmetic as in the source)
    }
};

}  // namespace orc
