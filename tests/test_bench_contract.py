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
    evals = b.REPLICAS_PER_GPU * b.N_ATOMS * b.N_GRIDS
    assert evals == 9240576 and abs(evals * b.b_alg(3) - 492.8e6) < 0.1e6


def test_workload_config_names_the_baseline_config():
    b = _bench()
    cfg = b.workload_config(8)
    assert "configs[4]" in cfg["workload"] and cfg["replicas_total"] == 8 * 65536 and cfg["grid_points"] == [192, 192, 192]
    assert "model" not in cfg


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
        gpus, steps, warmup = 4, 1, 3
    b.reference_arm(Args())
    assert capsys.readouterr().out == ""
