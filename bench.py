#!/usr/bin/env python
"""bench.py -- scans/sec of the PFSlam2D hot path at 256 particles x 1080 beams (BASELINE.json metric).

A "step" is one PFSlam2D::update() of one synthetic 1080-beam scan (predict -> scan matching of every particle ->
normalise / resample -> ray-cast + distance-map update of every particle).  Synthetic data, fp64 arithmetic over
packed u32 map cells.  N GPUs: particles shard over ranks (weak in scans, strong in particles: the SAME 256-particle
filter is split, so `scaling` is "strong").

  python bench.py --gpus 1 --steps K --warmup W            # this framework
  python bench.py --impl reference --steps K --warmup W    # the CPU restatement of the reference on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PARTICLES = 256
BEAMS = 1080
WORLD = "loop"   # 30 m x 30 m room with four pillars, rounded-square loop (BASELINE.json configs[3] world family)
METRIC = "scans/sec at 256 particles x 1080 beams"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)["hbm_gbs"], "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def make_data(n_scans):
    from iris_lama_b200 import synth
    return synth.make_dataset(WORLD, n_scans, n_beams=BEAMS)


def pf_options_kwargs():
    # reference defaults (pf_slam2d.h:132-185) except the gates so that every scan updates, and a fixed seed
    return dict(trans_thresh=0.05, rot_thresh=0.05, seed=42)


# --------------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference's thread-pool path (the reference cannot be built: no Eigen)
# --------------------------------------------------------------------------------------------------------
def run_cpu(ds, first, steps, warmup, threads):
    from oracle import pyoracle as po
    o = po.PFSlam2D(po.PFOptions.defaults(PARTICLES, threads=threads, **pf_options_kwargs()))
    o.set_prior(*ds.truth[0])
    o.update(ds.scans[0], ds.odom[0])
    for t in range(1, first):
        o.update(ds.scans[t], ds.odom[t])
    for t in range(first, first + warmup):
        o.update(ds.scans[t], ds.odom[t])
    t0 = time.perf_counter()
    n = 0
    for t in range(first + warmup, first + warmup + steps):
        n += int(o.update(ds.scans[t], ds.odom[t]))
    dt = time.perf_counter() - t0
    return n / dt, dt, o


def pick_threads(ds):
    """The reference's thread pool does not scale to every core of a large host (measured on the 128-core GPU box:
    best around 16-32 threads); give the CPU arm its best thread count, found on a few early scans."""
    from oracle import pyoracle as po
    ncpu = os.cpu_count() or 1
    best, best_v = 1, 0.0
    for th in sorted({t for t in (8, 16, 32, 64, ncpu) if t <= ncpu} | {min(ncpu, 8)}):
        o = po.PFSlam2D(po.PFOptions.defaults(PARTICLES, threads=th, **pf_options_kwargs()))
        o.set_prior(*ds.truth[0])
        for t in range(3):
            o.update(ds.scans[t], ds.odom[t])
        t0 = time.perf_counter()
        for t in range(3, 7):
            o.update(ds.scans[t], ds.odom[t])
        v = 4.0 / (time.perf_counter() - t0)
        if v > best_v:
            best, best_v = th, v
    return best


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup, pre = args.steps, args.warmup, args.prebuild
    ds = make_data(max(8, 1 + pre + warmup + steps))
    threads = pick_threads(ds)
    val, dt, _ = run_cpu(ds, 1 + pre, steps, warmup, threads)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "scans/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": 1000.0 * dt / max(steps, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": f"PFSlam2D {PARTICLES} particles x {BEAMS} beams, 0.05 m grid, synthetic 30 m loop room",
                                            "particles": PARTICLES, "beams": BEAMS},
            "cpu_baseline": {"value": val, "unit": "scans/s", "cores": threads, "kind": "port",
                             "sample": f"{steps} scans after {pre} map-building + {warmup} warm-up scans of the same workload; oracle restatement (the reference needs Eigen, absent), "
                                       f"g++ -O3 -march=x86-64-v3, one task per particle per phase on {threads} threads"},
            "e2e": {"value": val, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------
def gpu_arm(args):
    import torch
    from iris_lama_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if api.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: the lama_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        from iris_lama_b200.distributed import ShardedPFSlam2D
        dist.init_process_group("nccl", device_id=dev)

    steps, warmup, pre = args.steps, args.warmup, args.prebuild
    # scan 0 initialises the maps, scans 1 .. pre build them (untimed, the filter leaves the exploration phase that
    # only covers the first ~3 % of the 5 000-scan loop), then W warm-up and K timed scans; every pass replays the SAME scans
    n_scans = max(8, 1 + pre + warmup + steps)
    ds = make_data(n_scans)

    stream = torch.cuda.Stream(device=dev)   # the engine launches on this stream, so torch CUDA events see its kernels

    def new_pf(timing):
        opts = api.PFSlam2D.Options(PARTICLES, device=local_rank, timing=int(timing), shard_rank=rank, shard_count=world,
                                    stream=stream.cuda_stream, **pf_options_kwargs())
        pf = api.PFSlam2D(opts)
        pf.setPrior(*ds.truth[0])
        return pf

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_pass(staged, timing=False):
        """one fresh filter over scans 0 .. W+K; returns (device seconds of the K timed steps, wall seconds, pf, extras)"""
        pf = new_pf(timing)
        sh = ShardedPFSlam2D(pf, PARTICLES, device=dev) if world > 1 else None
        if staged and sh is None:
            pf.stageScans(ds.scans)
            step = lambda t: pf.updateStaged(t, ds.odom[t])
        elif sh is not None:
            step = lambda t: sh.update(ds.scans[t], ds.odom[t])
        else:
            step = lambda t: pf.update(ds.scans[t], ds.odom[t])   # the public call with HOST buffers
        step(0)
        for t in range(1, 1 + pre + warmup):
            step(t)
        sampler = ClockSampler(local_rank)
        barrier()
        pf.traffic(reset=True)
        _, tot0 = pf.counters()
        if rank == 0 and staged:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        w0 = time.perf_counter()
        n_upd = 0
        for t in range(1 + pre + warmup, 1 + pre + warmup + steps):
            n_upd += int(step(t))
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - w0
        dt = max_over_ranks(e0.elapsed_time(e1) * 1e-3)   # device timeline of the launching stream, max over ranks
        clocks = sampler.stop() if (rank == 0 and staged) else None
        _, tot1 = pf.counters()
        return dict(dt=dt, wall=wall, pf=pf, clocks=clocks, n_upd=n_upd, work={k: tot1[k] - tot0[k] for k in tot1})

    # ---- pass 1: `value` -- scans resident in HBM (single GPU: staged scans; sharded: host scans, see config) ----
    r1 = timed_pass(staged=True)
    dt_value, wall_value, clocks, n_upd = r1["dt"], r1["wall"], r1["clocks"], r1["n_upd"]
    _, launches = r1["pf"].kernelTimes()
    gpu_launches = int(sum(launches.values()))
    work = r1["work"]
    del r1

    # ---- pass 2: `e2e` -- same metric, same scans, through the public API with host buffers -----------------------
    r2 = timed_pass(staged=False)
    dt_e2e = r2["dt"]
    h2d, d2h = r2["pf"].traffic()
    del r2

    # ---- pass 3: per-kernel CUDA-event durations for the roofline (timing mode adds event records) ------------
    roofline = None
    kernel_ms = None
    if world == 1:
        r3 = timed_pass(staged=True, timing=True)
        ms, ln = r3["pf"].kernelTimes()
        d = r3["work"]
        peak, peak_kind = load_peaks()
        # algorithmic bytes: match E*N*(4 cells x 2 B); ray C*(4 B read + 4 B write); brushfire W*(5x8 B read + 4x8 B write)
        by = {"k_match": d["evals"] * BEAMS * 8.0, "k_raycast": d["ray_cells"] * 8.0, "k_brushfire": d["dm_pops"] * 72.0}
        tm = {"k_match": ms["match_ms"], "k_raycast": ms["raycast_ms"], "k_brushfire": ms["brushfire_ms"]}
        kernel_ms = {kk: tm[kk] / steps for kk in tm}
        dom = max(tm, key=tm.get)
        ach = by[dom] / (tm[dom] * 1e-3) / 1e9 if tm[dom] > 0 else 0.0
        traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of that kernel (per launch)
        tpath = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(dom, {}).get("dram_bytes_per_launch")
        roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "peak_kind": peak_kind, "algorithmic_bytes_per_launch": by[dom] / steps, "avg_launch_ms": tm[dom] / steps,
                    "all_kernels": {kk: {"GBps": (by[kk] / (tm[kk] * 1e-3) / 1e9 if tm[kk] > 0 else 0.0), "ms_per_step": tm[kk] / steps,
                                         "bytes_per_step": by[kk] / steps} for kk in tm}}
        del r3

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload ----------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        threads = pick_threads(ds)
        cpu_steps = max(4, min(steps, args.cpu_steps))
        val, dtc, _ = run_cpu(ds, 1 + pre + warmup - 2, cpu_steps, 2, threads)   # same scans as the GPU's timed region
        cpu = {"value": val, "unit": "scans/s", "cores": threads, "kind": "port",
               "sample": f"{cpu_steps} scans (after {pre} map-building scans) of the same workload, oracle restatement, thread pool of {threads} "
                         f"(best of 8/16/32/64/{os.cpu_count()} threads on this host)"}

    if rank == 0:
        value = steps / dt_value
        line = {"metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": 1000.0 * dt_value / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": f"PFSlam2D {PARTICLES} particles x {BEAMS} beams, 0.05 m grid, l2_max 0.5, GN+Cauchy(0.15), synthetic 30 m loop room",
                           "particles": PARTICLES, "beams": BEAMS, "parallelism": f"particles sharded over {world} GPU(s)",
                           "l2": "per-scan working set (~290 MB of touched map patches over 256 particles) exceeds the 126 MB L2; no explicit flush",
                           "value_inputs": "scans staged in HBM" if world == 1 else "host scans (sharded path)",
                           "updates_in_timed_region": n_upd, "prebuild_scans": pre,
                           "timer": "CUDA events on the launching stream around the K steps (every step enqueues match + map update at once and "
                                    "the host waits for the match results only, for its normalise/resample logic; the roofline kernel "
                                    "times come from a separate pass with per-kernel event records); host wall clock of the same "
                                    "region: %.3f s" % wall_value},
                "clocks": clocks,
                "e2e": {"value": steps / dt_e2e, "unit": "scans/s", "h2d_bytes_per_step": h2d / steps, "d2h_bytes_per_step": d2h / steps},
                "gpu_launches": gpu_launches,
                "counters_per_step": {k: v / steps for k, v in work.items()}}
        if roofline:
            line["roofline"] = roofline
            line["kernel_ms_per_step"] = kernel_ms
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--cpu-steps", type=int, default=40)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--prebuild", type=int, default=300, help="untimed scans that build the map before warm-up (both arms)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
