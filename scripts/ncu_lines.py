"""Per-CUDA-line instruction / sample distribution of one kernel of an ncu report (sorted by instructions)."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; items = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] != "": items.append((cur, r))
si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
tot_s = sum(int(r[si] or 0) for _, r in items); tot_i = sum(int(r[ii] or 0) for _, r in items)
print("samples", tot_s, "warp instructions", tot_i)
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
for f, r in items:
    for i, h in stall_cols: agg[h] = agg.get(h, 0) + int(r[i] or 0)
print(", ".join(f"{k[6:]} {100*v/max(tot_s,1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for f, r in sorted(items, key=lambda fr: -int(fr[1][ii] or 0))[:top]:
    s = int(r[si] or 0); i = int(r[ii] or 0)
    st = sorted(((int(r[c] or 0), h[6:]) for c, h in stall_cols), reverse=True)[:2]
    print(f"{100*s/tot_s:5.1f}% smp {100*i/tot_i:5.1f}% ins  {f}:{r[0]:>4} [{st[0][1]} {st[0][0]}, {st[1][1]} {st[1][0]}] {r[1].strip()[:100]}")
