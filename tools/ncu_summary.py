"""Summarise an ncu report (read on the CPU box): one CSV row per kernel launch with duration, DRAM bytes,
tensor-pipe / DRAM / L2 utilisation, plus totals.  Used to fill `roofline.traffic` in bench.py and to commit the
evidence under profiles/.

    python tools/ncu_summary.py gpurun_out/r01_step_gemms.ncu-rep profiles/r01_ncu_step_gemms.csv
"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "sm__cycles_elapsed.avg"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def main(rep, out_csv):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = [["id", "kernel", "grid", "block"] + METRICS]
    tot_us = tot_bytes = 0.0
    for r in data:
        vals = []
        for m in METRICS:
            v = float(r[col[m]].replace(",", "")) if r[col[m]] else 0.0
            u = units[col[m]]
            if m.startswith("dram__bytes") or m == "gpu__time_duration.sum":
                v *= UNIT.get(u, 1.0)
            vals.append(v)
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
        out.append([r[col["ID"]], name, r[col["Grid Size"]], r[col["Block Size"]]] + [f"{v:.6g}" for v in vals])
        tot_us += vals[0]
        tot_bytes += vals[1] + vals[2]
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["# ncu --set full --clock-control none (cold-cache, serialised launches: compare shares, not absolutes); "
                    "time in us, DRAM in bytes"])
        w.writerows(out)
        w.writerow(["# total", "", "", "", f"{tot_us:.6g}", f"{tot_bytes:.6g}"])
    print(f"{len(data)} launches, {tot_us / 1e3:.3f} ms under ncu, DRAM traffic {tot_bytes / 1e9:.3f} GB")
    return tot_bytes


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
