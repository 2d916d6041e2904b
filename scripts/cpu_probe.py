"""What does this host give the CPU arm?  Prints the CPU facts (logical CPUs, affinity, cgroup quota, load) and the
thread scaling of the oracle's thread pool in BOTH regimes (exploration scans 3-10, steady state after `pre` scans)."""
import json, os, sys, time
sys.path.insert(0, '.')
from oracle import pyoracle as po
from iris_lama_b200 import synth


def host_facts():
    f = {"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)), "loadavg": os.getloadavg()}
    for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us", "/sys/fs/cgroup/cpuset.cpus.effective"):
        try:
            f[p] = open(p).read().strip()
        except Exception:
            pass
    try:
        lines = open("/proc/cpuinfo").read().splitlines()
        f["model"] = next(l.split(":", 1)[1].strip() for l in lines if l.startswith("model name"))
        f["sockets"] = len({l.split(":")[1] for l in lines if l.startswith("physical id")})
        f["cores_per_socket"] = next(int(l.split(":")[1]) for l in lines if l.startswith("cpu cores"))
    except Exception:
        pass
    return f


if __name__ == "__main__":
    pre = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    P = 256
    print(json.dumps(host_facts()), flush=True)
    ds = synth.make_dataset("loop", pre + 24, n_beams=1080)
    for th in [int(a) for a in sys.argv[2:]] or [8, 16, 32, 64, 128]:
        if th > (os.cpu_count() or 1):
            continue
        o = po.PFSlam2D(po.PFOptions.defaults(P, trans_thresh=0.05, rot_thresh=0.05, seed=42, threads=th))
        o.set_prior(*ds.truth[0])
        for t in range(3):
            o.update(ds.scans[t], ds.odom[t])
        t0 = time.perf_counter()
        for t in range(3, 11):
            o.update(ds.scans[t], ds.odom[t])
        explore = 8 / (time.perf_counter() - t0)
        for t in range(11, pre):
            o.update(ds.scans[t], ds.odom[t])
        t0 = time.perf_counter()
        for t in range(pre, pre + 20):
            o.update(ds.scans[t], ds.odom[t])
        steady = 20 / (time.perf_counter() - t0)
        print(json.dumps({"threads": th, "explore_scans_per_s": round(explore, 2), "steady_scans_per_s": round(steady, 2), "loadavg": os.getloadavg()[0]}), flush=True)
