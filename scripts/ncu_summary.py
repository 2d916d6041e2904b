"""Markdown summary + DRAM traffic json of an `ncu --set full --import-source on` report holding one launch of each hot kernel.
usage: python scripts/ncu_summary.py report.ncu-rep out.md out_traffic.json "title line"
"""
import csv, json, subprocess, sys

rep, out_md, out_json, title = sys.argv[1:5]
KERNELS = ["k_raycast", "k_brushfire", "k_match", "k_ray_pull", "k_ray_setup"]
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed.sum",
           "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "l1tex__t_sector_hit_rate.pct", "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
name_col = hdr.index("Kernel Name")
to_bytes = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
to_us = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}
md = ["# " + title, ""]
traffic = {}
for k in KERNELS:
    r = next((r for r in rows[2:] if k in r[name_col]), None)
    if r is None:
        continue
    md += ["## " + k, "```"]
    for m in METRICS:
        if m in hdr:
            md.append("%s = %s %s" % (m, r[hdr.index(m)], units[hdr.index(m)]))
    rd, wr, du = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    traffic[k] = {"dram_bytes_per_launch": float(r[rd]) * to_bytes[units[rd]] + float(r[wr]) * to_bytes[units[wr]],
                  "duration_us_under_ncu": float(r[du]) * to_us[units[du]]}
    md.append("```")
    src = subprocess.run([sys.executable, "scripts/ncu_lines.py", rep, k, "14"], capture_output=True, text=True).stdout
    md += ["per source line (samples, warp instructions, top stall reasons):", "```", src.rstrip(), "```", ""]
open(out_md, "w").write("\n".join(md) + "\n")
commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
traffic["commit"] = commit + " (HEAD when the summary was written; the library of the capture was built from it)"
json.dump(traffic, open(out_json, "w"), indent=1)
print(json.dumps(traffic))
