"""bench.py's host-side pieces (no GPU): the algorithmic-bytes figure, the workload description and the argument contract
the driver relies on (--gpus/--steps/--warmup/--impl), and that only the CPU legs touch oracle/."""
import importlib.util
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_algorithmic_bytes_per_evaluation():
    b = _bench()
    assert b.b_alg(1) == 88.0                      # 32 + 4 + 52/G
    assert abs(b.b_alg(3) - 53.333333333333336) < 1e-12
    assert b.b_alg(1, precision=1) == 124.0        # double mode: 72 + 52/G
    evals = b.REPLICAS_TOTAL * b.N_ATOMS * b.N_GRIDS
    assert evals == 9240576 and abs(evals * b.b_alg(3) - 492.8e6) < 0.1e6
    assert abs(b.b_alg(3, forces=False) - (36.0 + 28.0 / 3)) < 1e-12      # energy only: no force bytes


def test_workload_config_names_the_baseline_config():
    """configs[4] as BASELINE.json names it: 65,536 replicas IN TOTAL, sharded over the GPUs (strong scaling) by default;
    the same dictionary for both arms (nothing arm-specific in it)."""
    b = _bench()
    cfg = b.workload_config(8)
    assert "configs[4]" in cfg["workload"] and cfg["replicas_total"] == 65536 and cfg["replicas_per_gpu"] == 8192
    assert cfg["scaling"] == "strong" and cfg["grid_points"] == [192, 192, 192] and cfg["evals_per_step"] == 9240576
    assert "model" not in cfg
    weak = b.workload_config(8, "weak")
    assert weak["scaling"] == "weak" and weak["replicas_total"] == 8 * 65536
    assert b.workload_config(1, "weak") == b.workload_config(1, "strong")      # one GPU: the same run


def test_window_accumulator_rotation():
    """The K-step window's energy accumulators: consecutive steps never share one, and the last step's differs from the
    first's (whose buffer the last launch clears for the next window / graph replay)."""
    b = _bench()
    for steps in range(2, 40):
        a = b.DeviceLoop.accumulators(steps)
        assert len(a) == steps and all(0 <= x < 3 for x in a)
        assert all(a[j] != a[j + 1] for j in range(steps - 1)) and a[-1] != a[0]


def test_traffic_is_only_reported_for_the_profiled_sources(tmp_path, monkeypatch):
    b = _bench()
    val, why = b.measured_traffic("no_such_workload")
    assert val is None and why


def test_committed_traffic_file_matches_the_kernel_sources_and_captures():
    """profiles/r2_traffic.json is what bench.py reports as roofline.traffic: every entry must carry the hash of the kernel
    sources as they are in the tree (a stale capture is dropped, not reported), point at a committed ncu export and hold
    that export's DRAM bytes; the physical roofline helper refuses lower-bound captures."""
    import csv
    import json
    b = _bench()
    doc = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
    assert {"C5", "C3", "C5_double", "C5_energy_only"} <= set(doc)
    for key, entry in doc.items():
        assert entry["source_sha256_16"] == b.kernel_source_hash(), key
        raw = os.path.join(ROOT, entry["source"].split(" ")[0])
        rows = list(csv.reader(open(raw)))
        hdr, units, r = rows[0], rows[1], rows[2]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        total = sum(float(r[hdr.index(n)]) * scale[units[hdr.index(n)]] for n in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        assert abs(total - entry["dram_bytes_per_launch"]) <= 1e-6 * total, key
        val, src = b.measured_traffic(key)
        assert val == entry["dram_bytes_per_launch"]
    c5 = doc["C5"]["dram_bytes_per_launch"]
    assert abs(c5 / (9240576 * b.b_alg(3)) - 1.0) < 0.03          # the full launch moves what the algorithm needs
    assert b.physical_roofline(c5, doc["C5"]["source"], 75.0, 6541.8)["frac"] > 0.9
    assert b.physical_roofline(doc["C4"]["dram_bytes_per_launch"], doc["C4"]["source"], 6.5, 6541.8) is None


def test_oracle_is_only_reached_from_the_cpu_legs():
    src = open(os.path.join(ROOT, "bench.py")).read()
    users = [m.start() for m in re.finditer(r"from oracle import bindings", src)]
    assert len(users) == 2
    for pos in users:      # both imports sit inside cpu_baseline() / reference_arm()
        head = src[:pos]
        fn = re.findall(r"\ndef (\w+)\(", head)[-1]
        assert fn in ("cpu_baseline", "reference_arm"), fn
    pkg = os.path.join(ROOT, "openmmgridforce_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower() or f in ("plugin_driver.cpp",) or "oracle/" not in text, f


def test_reference_arm_other_ranks_exit_quietly(monkeypatch, capsys):
    b = _bench()
    monkeypatch.setenv("RANK", "3")

    class Args:
        gpus, steps, warmup, scaling = 4, 1, 3, "strong"
    b.reference_arm(Args())
    assert capsys.readouterr().out == ""
