"""Top stall sites of one kernel from an ncu report's source page (SASS view).
    python tools/ncu_src_top.py report.ncu-rep <kernel-id> [top]"""
import csv, io, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr, data = rows[0], rows[1:]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
tot_inst = sum(int(r[col["Instructions Executed"]] or 0) for r in data)
print(f"total samples {tot}, instructions executed {tot_inst}")
agg = {s: sum(int(r[col[s]] or 0) for r in data) for s in stalls}
print("stall totals:", {k: f"{100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.005 * tot})
idx = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]] or 0))[:top]
for i in sorted(idx):
    r = data[i]
    n = int(r[col["# Samples"]] or 0)
    why = sorted(((int(r[col[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"{i:5d} {100*n/tot:5.2f}%  inst {int(r[col['Instructions Executed']] or 0):9d}  {r[col['Source']].strip()[:70]:70s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
