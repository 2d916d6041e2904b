"""SASS opcode histogram per kernel of the built library: python scripts/sass_hist.py iris_lama_b200/liblama_b200.so out.md
(the mnemonics that prove the sm_100a features: UBLKCP = bulk TMA copy, SYNCS = mbarrier, REDG = fire-and-forget reduction, ATOMG, LDS / STS, SHFL, BAR)"""
import collections, re, subprocess, sys

so, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
kern, hist = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        full = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        mk = re.search(r"(k_\w+(?:<[^>]*>)?)\s*\(", full)
        kern = mk.group(1) if mk else full[:60]
        hist.setdefault(kern, collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ["UBLKCP", "SYNCS", "REDG", "RED", "ATOMG", "ATOMS", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "VOTE", "DADD", "DMUL", "DFMA", "IMAD", "LOP3", "IADD3", "UTCMMA", "HMMA"]
md = ["# SASS opcode histogram of `%s` (cubins: %s), commit %s" % (so, ", ".join(arch), commit), "",
      "Columns: instructions whose mnemonic STARTS with the name (e.g. `REDG.E.ADD.STRONG.GPU` counts under REDG).  No tensor-core opcode (`UTC*MMA`, `HMMA`) is", 
      "expected: the path has no GEMM (SURVEY 8(d)).", "",
      "| kernel | total | " + " | ".join(KEY) + " |", "|---|---|" + "---|" * len(KEY)]
for k, h in hist.items():
    tot = sum(h.values())
    row = []
    for key in KEY:
        row.append(str(sum(v for op, v in h.items() if op == key or op.startswith(key + "."))))
    md.append("| `%s` | %d | %s |" % (k, tot, " | ".join(row)))
md += ["", "Full mnemonics of the sm_100a-specific instructions:", "```"]
for k, h in hist.items():
    sp = {op: v for op, v in h.items() if op.startswith(("UBLKCP", "SYNCS", "REDG", "ATOMG", "UTMA", "ELECT", "FENCE"))}
    if sp:
        md.append(k + ": " + ", ".join("%s x%d" % kv for kv in sorted(sp.items())))
md.append("```")
open(out, "w").write("\n".join(md) + "\n")
print("\n".join(md[:12]))
