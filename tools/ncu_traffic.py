"""DRAM bytes per launch of the profiled kernel out of `ncu --page raw --csv` files -> profiles/r2_traffic.json, keyed by
the bench.py workload name and stamped with the hash of the kernel sources (bench.py reports `roofline.traffic` only while
the sources still hash to what was profiled).
usage: python tools/ncu_traffic.py KEY=raw.csv[=caveat] [...]     e.g. C5=profiles/r2b_c5full_raw.csv
       python tools/ncu_traffic.py --restamp "why"     after tools/sass_identity.sh has shown that the profiled kernels'
       SASS did not change although a hashed source file did: every entry gets the current hash and the reason."""
import csv
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)

out_path = os.path.join(ROOT, "profiles", "r2_traffic.json")
doc = json.load(open(out_path)) if os.path.exists(out_path) else {}
if len(sys.argv) >= 3 and sys.argv[1] == "--restamp":
    for entry in doc.values():
        entry.setdefault("restamped", []).append({"from": entry["source_sha256_16"], "to": bench.kernel_source_hash(), "why": sys.argv[2]})
        entry["source_sha256_16"] = bench.kernel_source_hash()
    json.dump(doc, open(out_path, "w"), indent=1, sort_keys=True)
    sys.exit(0)
for arg in sys.argv[1:]:
    key, path = arg.split("=", 1)
    note = None
    if "=" in path:      # KEY=raw.csv=free-text caveat carried into bench.py's traffic_source
        path, note = path.split("=", 1)
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    r = rows[2]

    def val(name):
        i = hdr.index(name)
        v = float(r[i].replace(",", ""))
        u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0,
                    "ms": 1e3, "msecond": 1e3}.get(u, 1)
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    doc[key] = {"kernel": r[hdr.index("Kernel Name")], "dram_bytes_read": rd, "dram_bytes_write": wr,
                "dram_bytes_per_launch": rd + wr, "duration_us_under_ncu": val("gpu__time_duration.sum"),
                "source": os.path.relpath(path, ROOT) + " (ncu --set full --clock-control none --cache-control none, one warm launch)",
                "source_sha256_16": bench.kernel_source_hash()}
    if note:
        doc[key]["source"] += "; " + note
    print(key, doc[key]["kernel"][:60], f"{(rd + wr) / 1e6:.1f} MB")
json.dump(doc, open(out_path, "w"), indent=1, sort_keys=True)
