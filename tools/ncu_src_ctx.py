import csv, io, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
ranges = [tuple(map(int, a.split("-"))) for a in sys.argv[3:]]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr, data = rows[0], rows[1:]
col = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
for lo, hi in ranges:
    print(f"--- {lo}-{hi}")
    for i in range(lo, hi + 1):
        r = data[i]
        print(f"{i:5d} {100*int(r[col['# Samples']] or 0)/tot:5.2f}% {int(r[col['Instructions Executed']] or 0):9d}  {r[col['Source']].strip()[:90]}")
