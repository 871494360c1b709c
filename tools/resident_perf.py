"""C2 (one 47-atom ligand x 3 grids of 208x278x231, one evaluation per MD step): launch-per-step against the resident
evaluator, through the C ABI (ctypes loop) and through the platform plugin (C++ step loop). Prints bench.py's C2 object.
    python tools/resident_perf.py [steps]"""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)
import openmmgridforce_b200 as gf  # noqa: E402

dev = gf.Device(0)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
for _ in range(2):
    print(json.dumps(bench.run_single_ligand(gf, dev, steps=steps)))
