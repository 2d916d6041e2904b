"""The header-only C++ binding of INTEGRATION.md compiles against the C-ABI with reference-shaped (Eigen-free) types."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include <array>
#include <cstdio>
#include "lama_b200_shim.hpp"
struct Q { double x() const {return 0;} double y() const {return 0;} double z() const {return 0;} double w() const {return 1;} };
struct Cloud { std::vector<std::array<double,3>> points; std::array<double,3> sensor_origin_{}; Q sensor_orientation_; };
struct Pose { double x_, y_, r_; double x() const {return x_;} double y() const {return y_;} double rotation() const {return r_;} };
int main() {
  auto o = lama_b200_shim::PFSlam2D::defaults(4);
  if (o.particles != 4 || o.max_iter != 100) return 2;
  try {
    lama_b200_shim::PFSlam2D pf(o);
    auto c = std::make_shared<Cloud>(); c->points.push_back({1,0,0});
    pf.setPrior(Pose{0,0,0});
    bool u = pf.update(c, Pose{0,0,0}, 0.0);
    std::printf("updated %d\n", (int)u);
    lama_b200_shim::LidarOdometry2D lo;            // instantiate every template of the header
    lo.update(c, 0.0);
    double xyr[3]; lo.getOdom(xyr);
    lama_b200_shim::Loc2D loc; loc.Init(lama_b200_shim::Loc2D::defaults()); loc.triggerGlobalLocalization();
  } catch (const std::exception& e) { std::printf("%s\n", e.what()); }
  return 0; }
'''


def test_shim_compiles_and_links(tmp_path):
    src = tmp_path / "shim.cpp"
    src.write_text(SRC)
    exe = tmp_path / "shim"
    lib_dir = os.path.join(ROOT, "iris_lama_b200")
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", lib_dir, "-llama_b200",
                           f"-Wl,-rpath,{lib_dir}"])
    out = subprocess.check_output([str(exe)]).decode()
    assert "updated 1" in out or "no CUDA device" in out      # loud failure without a GPU, never a CPU fallback


RUN_SRC = r'''
// A reference-side caller in C++: PFSlam2D through the header-only shim over scans read from a file; prints pose, best particle, Neff per scan.
#include <array>
#include <cstdio>
#include <memory>
#include <vector>
#include "lama_b200_shim.hpp"
struct Q { double x() const {return 0;} double y() const {return 0;} double z() const {return 0;} double w() const {return 1;} };
struct Cloud { std::vector<std::array<double,3>> points; std::array<double,3> sensor_origin_{}; Q sensor_orientation_; };
struct Pose { double x_, y_, r_; double x() const {return x_;} double y() const {return y_;} double rotation() const {return r_;} };
int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 3;
  int hdr[3];                                       // scans, beams, particles
  if (std::fread(hdr, 4, 3, f) != 3) return 4;
  const int T = hdr[0], N = hdr[1];
  auto o = lama_b200_shim::PFSlam2D::defaults((uint32_t)hdr[2]);
  o.trans_thresh = 0.05; o.rot_thresh = 0.05; o.seed = 5; o.meas_sigma_gain = 0.02;
  try {
    lama_b200_shim::PFSlam2D pf(o);
    std::vector<double> scan((size_t)N * 3); double odom[3], prior[3];
    if (std::fread(prior, 8, 3, f) != 3) return 5;
    pf.setPrior(Pose{prior[0], prior[1], prior[2]});
    for (int t = 0; t < T; ++t) {
      if (std::fread(scan.data(), 8, scan.size(), f) != scan.size() || std::fread(odom, 8, 3, f) != 3) return 6;
      auto c = std::make_shared<Cloud>();
      for (int i = 0; i < N; ++i) c->points.push_back({scan[3 * i], scan[3 * i + 1], scan[3 * i + 2]});
      const bool u = pf.update(c, Pose{odom[0], odom[1], odom[2]}, (double)t);
      double xyr[3]; pf.getPose(xyr);
      std::printf("%d %d %a %a %a %zu %a\n", t, (int)u, xyr[0], xyr[1], xyr[2], pf.getBestParticleIdx(), pf.getNeff());
    }
  } catch (const std::exception& e) { std::printf("EXC %s\n", e.what()); return 7; }
  return 0; }
'''


import pytest  # noqa: E402


@pytest.mark.gpu
def test_cpp_caller_through_the_shim_on_the_gpu(tmp_path, gpu_api, po, synth):
    """The reference-side C++ binding of INTEGRATION.md on a real device: a compiled caller (no Python in its process) gets the poses the ctypes
    mirror gets (bit for bit: same library, same seed) and the oracle's within 1e-9."""
    import numpy as np
    P, T, N = 12, 12, 180
    ds = synth.make_dataset("room", T, n_beams=N)
    data = tmp_path / "scans.bin"
    with open(data, "wb") as f:
        f.write(np.array([T, N, P], np.int32).tobytes())
        f.write(np.asarray(ds.truth[0], np.float64).tobytes())
        for t in range(T):
            f.write(np.ascontiguousarray(ds.scans[t], np.float64).tobytes())
            f.write(np.asarray(ds.odom[t], np.float64).tobytes())
    src = tmp_path / "caller.cpp"
    src.write_text(RUN_SRC)
    exe = tmp_path / "caller"
    lib_dir = os.path.join(ROOT, "iris_lama_b200")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", lib_dir, "-llama_b200",
                           f"-Wl,-rpath,{lib_dir}"])
    lines = subprocess.check_output([str(exe), str(data)]).decode().strip().splitlines()
    assert len(lines) == T and not any(l.startswith("EXC") for l in lines)
    kw = dict(trans_thresh=0.05, rot_thresh=0.05, seed=5, meas_sigma_gain=0.02)
    g = gpu_api.PFSlam2D(gpu_api.PFSlam2D.Options(P, **kw))
    o = po.PFSlam2D(po.PFOptions.defaults(P, **kw))
    g.setPrior(*ds.truth[0]); o.set_prior(*ds.truth[0])
    n_res = 0
    for t in range(T):
        ug, uo = g.update(ds.scans[t], ds.odom[t], timestamp=float(t)), o.update(ds.scans[t], ds.odom[t])
        n_res += int(len(o.last_resample()) > 0)
        tok = lines[t].split()
        pose = np.array([float.fromhex(v) for v in tok[2:5]])
        assert int(tok[0]) == t and bool(int(tok[1])) == ug == uo
        assert (pose == np.asarray(g.getPose())).all() and int(tok[5]) == g.getBestParticleIdx() and float.fromhex(tok[6]) == g.getNeff()
        so = o.particles()[0][o.best()]                                          # state (c, s, x, y) of the best particle = getPose()
        assert abs(pose[0] - so[2]) < 1e-9 and abs(pose[1] - so[3]) < 1e-9 and abs(pose[2] - np.arctan2(so[1], so[0])) < 1e-9 and int(tok[5]) == o.best()
    assert n_res >= 1
