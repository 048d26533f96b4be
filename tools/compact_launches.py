"""ncu `--metrics gpu__time_duration.sum --csv` launch list -> compact CSV (launch, kernel, grid, block, us)."""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
names = rows[hdr]
ix = {k: names.index(k) for k in ("ID", "Kernel Name", "Grid Size", "Block Size", "Metric Value", "Metric Unit")}
w = csv.writer(sys.stdout, lineterminator="\n")
w.writerow(["launch", "kernel", "grid", "block", "gpu__time_duration.sum_us"])
for r in rows[hdr + 1:]:
    if len(r) <= ix["Metric Value"] or r[ix["Kernel Name"]].startswith("void at::"):
        continue
    k = r[ix["Kernel Name"]].split("(")[0]
    k = re.sub(r"^void ", "", k).replace("vcl::<unnamed>::", "").replace("vcl::", "")
    v = float(r[ix["Metric Value"]].replace(",", ""))
    v = v / 1000 if r[ix["Metric Unit"]] == "ns" else (v * 1000 if r[ix["Metric Unit"]] == "ms" else v)
    w.writerow([r[ix["ID"]], k, r[ix["Grid Size"]], r[ix["Block Size"]], round(v, 2)])
