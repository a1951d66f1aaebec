"""every mbarrier wait site of a kernel (SASS TRYWAIT + its NANOSLEEP) with the share of warp-samples spent there"""
import csv, io, subprocess, sys, re
rep, kid = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr, data = rows[0], rows[1:]
col = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
print("total samples", tot)
for i, r in enumerate(data):
    src = r[col["Source"]]
    if "TRYWAIT" in src:
        n = sum(int(data[j][col["# Samples"]] or 0) for j in range(i, min(i + 6, len(data))))
        m = re.search(r"\+(0x[0-9a-f]+)\]", src)
        print(f"{i:5d} off {m.group(1) if m else '?':>8s}  execs {int(r[col['Instructions Executed']] or 0):9d}  samples {100*n/tot:5.2f}%")
# instruction mix by opcode
mix = {}
for r in data:
    op = r[col["Source"]].strip().split()
    if not op: continue
    o = op[1] if op[0].startswith("@") else op[0]
    o = o.split(".")[0]
    mix[o] = mix.get(o, 0) + int(r[col["Instructions Executed"]] or 0)
ti = sum(mix.values())
print("instruction mix:", {k: f"{100*v/ti:.1f}%" for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:22]})
