"""Writes profiles/ncu_traffic.json (dram bytes per launch of the three gather kernels) from an `ncu --set full`
report of tools/dfaust_layer_run.py; bench.py copies the dominant kernel's figure into roofline.traffic."""
import csv
import json
import os
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(v.replace(",", "")) * mult


res = {"source": os.path.basename(rep), "command": "python tools/dfaust_layer_run.py seg_head 3 1"}
for r in data:
    name = r[col["Kernel Name"]]
    if "k_edge_tc<" in name or "k_edge_row_tc<" in name:
        key = "k_edge_tc"
    elif "k_agg_tc<" in name:
        key = "k_agg_tc_tr" if ", 1, " in name.split("<")[1] else "k_agg_tc_fwd"
    else:
        continue
    rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
    wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    res[key] = {"kernel": name, "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                "time_us": float(r[col["gpu__time_duration.sum"]])}
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles",
                                 "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(res, indent=1))
