"""Dev tool: dram bytes per launch of a kernel class out of an `ncu --set full` report ->
profiles/r01_ncu_traffic.json (read by bench.py's roofline.traffic).
    python tools/ncu_traffic.py <report.ncu-rep> <workload> <kernel label> [algorithmic bytes per launch]"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
rep, workload, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
alg = float(sys.argv[4]) if len(sys.argv) > 4 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]
units = rows[1]
ir, iw, it, ik = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum"), h.index("Kernel Name")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
launches = []
for r in rows[2:]:
    launches.append({"kernel": r[ik][:60], "dram_read": float(r[ir]) * scale[units[ir]], "dram_write": float(r[iw]) * scale[units[iw]],
                     "time_" + units[it]: float(r[it])})
tot = sum(x["dram_read"] + x["dram_write"] for x in launches)
p = REPO / "profiles" / "r01_ncu_traffic.json"
db = json.loads(p.read_text()) if p.exists() else {}
db[f"{workload}:{kernel}"] = {"dram_bytes_per_launch": tot / len(launches), "launches_captured": len(launches),
                              "algorithmic_bytes_per_launch": alg, "report": Path(rep).name, "launches": launches}
p.write_text(json.dumps(db, indent=1))
print(json.dumps(db[f"{workload}:{kernel}"], indent=1)[:1500])
