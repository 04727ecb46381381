"""Turn ncu captures of the traversal kernel (scripts/profile_search.py --bench-data 1 under
`ncu --set full`) into profiles/r2_traffic.json, stamped with the kernel sources' hash so that bench.py
quotes `roofline.traffic` only while the kernel it times is the kernel that was captured.

usage: make_traffic_json.py <config>:<efSearch>=<file.ncu-rep> [...]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_sha  # noqa: E402

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
out = {"kernel_src_sha": kernel_source_sha(), "how": "ncu --set full --clock-control none; dram__bytes_read.sum + "
       "dram__bytes_write.sum of ONE beam_kernel launch over the bench.py batch (scripts/profile_search.py --bench-data 1)",
       "entries": {}}
path = os.environ.get("TRAFFIC_JSON", os.path.join(ROOT, "profiles", "r2_traffic.json"))
if os.path.exists(path):
    old = json.load(open(path))
    if old.get("kernel_src_sha") == out["kernel_src_sha"]:
        out["entries"] = old.get("entries", {})
for arg in sys.argv[1:]:
    key, rep = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    best = None
    for vals in rows[2:]:
        get = lambda name: float(vals[hdr.index(name)].replace(",", "")) * UNIT[units[hdr.index(name)]]
        tot = get("dram__bytes_read.sum") + get("dram__bytes_write.sum")
        if best is None or tot > best[0]:   # the batch launch, not a warm-up on fewer queries
            best = (tot, vals[hdr.index("Kernel Name")], float(vals[hdr.index("gpu__time_duration.sum")].replace(",", "")),
                    units[hdr.index("gpu__time_duration.sum")])
    out["entries"][key] = {"dram_bytes_per_launch": int(best[0]), "kernel": best[1], "ncu_duration": f"{best[2]} {best[3]}"}
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
