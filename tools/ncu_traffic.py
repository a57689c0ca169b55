"""DRAM bytes of one force-pass launch from an `ncu --set full` report -> profiles/r2_force_pass_traffic.json
(bench.py's roofline.traffic).  Usage: python tools/ncu_traffic.py <report.ncu-rep> <n> <group> <a_in> <a_out>"""
import csv, json, subprocess, sys
rep, n, group, a_in, a_out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), float(sys.argv[5])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot = 0.0
for name, unit, val in zip(h, u, v):
    if name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(val) * scale[unit]
import os
head = os.environ.get("GPLUM_HEAD") or subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
out = {"dram_bytes": int(tot), "n": n, "group": group, "a_in": a_in, "a_out": a_out, "kernel": "force_pass_kernel<2, false>",
       "source": "ncu --set full --clock-control none, one launch inside bench.py (cold L2: ncu flushes caches between replays)",
       "commit": head}
json.dump(out, open("profiles/r2_force_pass_traffic.json", "w"), indent=1)
print(out)
