#!/usr/bin/env python
"""bench.py — atom-grid evaluations/s (energy + force) of the GridForce path on B200.

    python bench.py --gpus 1 --steps K --warmup W                 # this repo's CUDA path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                          # the reference's own CPU kernel, all host cores

Workload (config.workload): BASELINE.json configs[4] — ligand replicas x 47 atoms x 3 grids of 192^3 — with 65,536
replicas PER GPU (weak scaling: rank g evaluates its own 65,536-replica batch; grids replicated; no collective on the
force path; one NCCL all-gather of per-replica energies per step). At N=1 that is configs[4] itself on one GPU. The
single-GPU configs C3 (1M atoms x 256^3) and C4 (4096 replicas x 3 grids) are timed in the same run and reported
under "other_workloads" (N=1 only).

A "step" = one evaluation of the whole batch: positions -> per-replica energies + forces for every atom on every grid.
  value      device-resident inputs, CUDA-event time on the launching stream, max over ranks
  e2e        the same through gfb_kernel_execute_host with pinned HOST buffers (H2D of positions and D2H of forces and
             energies inside the timed region)
  roofline   algorithmic bytes per evaluation (36 + 52/G bytes, G grids per atom) x evaluations per launch /
             average launch duration, against the measured HBM copy bandwidth in MEASURED_PEAKS.json
Only the cpu_baseline leg and --impl reference load anything from oracle/ (the test-only CPU oracle).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "atom-grid evals/s (E+F)"
UNIT = "evals/s"
REPLICAS_PER_GPU = 65536
N_ATOMS = 47
N_GRIDS = 3
GRID_N = 192
C5_DRAM_BYTES_PER_LAUNCH = 495.4e6     # measured once per change with ncu --set full (profiles/README.md, r1c)


def b_alg(n_grids, precision=0):
    """Algorithmic bytes per atom-grid evaluation (SURVEY.md §8d / DESIGN.md §4)."""
    return (36.0 + 52.0 / n_grids) if precision == 0 else (72.0 + 52.0 / n_grids)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 100 ms from just before the timed device loop until the last
    GPU measurement of the run (device loop, e2e loop, and at N=1 the C3/C4 loops)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.tmp,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.split(",") for r in open(self.tmp.name).read().splitlines() if r.count(",") >= 8]
        os.unlink(self.tmp.name)
        if not rows:
            return out
        sm = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in rows:
            for name, v in zip(names, r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=float(rows[0][2]), reasons=sorted(reasons), samples=len(rows))
        return out


def bind_to_gpu_cpus(gpu_index):
    """Restrict this process to the CPU cores NVML reports as local to the GPU (what `nvidia-smi topo -m` prints), so that
    its pinned buffers are allocated on that NUMA node: with 8 ranks streaming 2 x 74 MB per step each, host memory on the
    wrong socket halves the PCIe rate. Returns a short description for the JSON line; never fatal."""
    if os.environ.get("GFB_BIND_CPUS", "1") == "0":
        return "off"
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus ({min(cpus)}-{max(cpus)})"
        return "no local cpus reported"
    except Exception as exc:
        return f"unavailable ({type(exc).__name__})"


def pinned_array(shape, dtype=np.float64):
    """numpy view over page-locked host memory (torch is the allocator; no torch type crosses the C ABI)."""
    import torch
    t = torch.empty(tuple(shape), dtype={np.float64: torch.float64, np.int64: torch.int64}[dtype], pin_memory=True)
    return t.numpy(), t


# ----------------------------------------------------------------------------------------------------------------
# device-timed loops
# ----------------------------------------------------------------------------------------------------------------
def time_device_steps(torch, gf, kern, pos_sets, n_replicas, n_atoms, steps, warmup, stream, post_step=None,
                      force_mode=None, energy_bufs_out=None, final_step=None, barrier=None):
    """K launches on `stream`, rotating through pos_sets (and matching force/energy buffers). Returns
    (seconds, launches) with CUDA events recorded on the launching stream."""
    force_mode = gf.FORCE_FIXED_ADD if force_mode is None else force_mode
    dev = pos_sets[0].device
    n = n_replicas * n_atoms
    stride = ((n + 31) // 32) * 32
    bufs = []
    for _ in pos_sets:
        d_f = torch.zeros(3 * stride, dtype=torch.int64, device=dev)
        bufs.append(d_f)
    # four per-replica energy accumulators: step i adds into e[i % 4] and zero-fills e[(i + 1) % 4] in the same launch;
    # the energy gather of step i (N > 1) reads e[i % 4] on its own stream while steps i+1, i+2 run, and is waited for
    # (on the host: it finished long before) ahead of step i+3, whose launch clears e[i % 4] again.
    d_e3 = [torch.zeros(n_replicas, dtype=torch.float64, device=dev) for _ in range(4)]
    if energy_bufs_out is not None:
        energy_bufs_out.extend(d_e3)
    pending = {}
    torch.cuda.synchronize()

    def step(i):
        s = i % len(pos_sets)
        d_f, d_e, d_next = bufs[s], d_e3[i % 4], d_e3[(i + 1) % 4]
        if i - 3 in pending:
            pending.pop(i - 3).wait()
        kern.execute_device(n_replicas, n_atoms, pos_sets[s].data_ptr(), d_e.data_ptr(), None, d_f.data_ptr(), force_mode,
                            stride, None, stream.cuda_stream, d_energies_clear=d_next.data_ptr())
        if post_step is not None:
            work = post_step(d_e)
            if work is not None:
                pending[i] = work

    def drain():
        for key in sorted(pending):
            pending.pop(key).wait()

    with torch.cuda.stream(stream):
        for i in range(warmup):
            step(i)
        drain()
        if final_step is not None:   # untimed: the collective's first call sets up NCCL's connections (milliseconds)
            final_step(d_e3[(warmup - 1) % 4]).wait()
        stream.synchronize()
        if barrier is not None:      # all ranks enter the timed region together (barrier + synchronize on both sides)
            barrier()
        l0 = gf.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            step(warmup + i)
        drain()                  # the last gathers are inside the timed region
        if final_step is not None:   # the one gather of a "final" run: after the last step, before the clock stops
            stream.wait_event(final_step(d_e3[(warmup + steps - 1) % 4]).ev)     # the clock stops after the gather
        e1.record(stream)
        stream.synchronize()
        if barrier is not None:
            barrier()
        launches = gf.launch_count() - l0
    return e0.elapsed_time(e1) * 1e-3, launches, bufs


def time_e2e_steps(gf, kern, pos_host, forces_host, energies_host, steps, warmup):
    """Public host API, pinned host buffers, wall clock around synchronous calls (each returns with results on the host)."""
    for _ in range(warmup):
        kern.execute_host(pos_host, forces=forces_host, energies_out=energies_host)
    t0 = time.perf_counter()
    for _ in range(steps):
        kern.execute_host(pos_host, forces=forces_host, energies_out=energies_host)
    return time.perf_counter() - t0


def run_single_ligand(gf, dev, steps=2000):
    """configs[1]: one 47-atom ligand in three 208x278x231 grids, one evaluation per MD step through the host API
    (positions in, energy + forces out, synchronous). OpenMM's integrator is not available here, so the figure is the
    grid-force-limited upper bound: steps/s x 4 fs (example/input.json:24)."""
    from openmmgridforce_b200 import workloads as W
    w = W.c2_single_ligand()
    grids = [gf.Grid(dev, w.counts, w.spacing, w.origin, v, gf.PRECISION_MIXED) for v in w.grids]
    kern = gf.Kernel(dev, grids, w.scaling, oob_k=w.oob_k)
    pos_h, _t1 = pinned_array(w.pos.shape)
    pos_h[...] = w.pos
    f_h, _t2 = pinned_array(w.pos.shape)
    e_h, _t3 = pinned_array((1,))
    secs = time_e2e_steps(gf, kern, pos_h, f_h, e_h, steps, 50)
    kern.close()
    for g in grids:
        g.close()
    sps = steps / secs
    # The same step as an OpenMM integrator would issue it: System with three GridForce objects on the B200 platform
    # plugin (libOpenMMGridForceB200.so), Context::calcForcesAndEnergy timed from C++. The platform evaluates the three
    # forces of a step in ONE launch (step fusion).
    plugin = None
    try:
        import openmmgridforce_b200.gridforceplugin as gfp
        system = gfp.System()
        for _ in range(w.n_atoms):
            system.addParticle(1.0)
        for g in range(w.n_grids):
            f = gfp.GridForce()
            f.addGridCounts(*w.counts)
            f.addGridSpacing(*w.spacing)
            f.setGridOrigin(*w.origin)
            f.setGridValues(w.grids[g])
            f.setScalingFactors(w.scaling[g])
            f.setOutOfBoundsRestraint(w.oob_k[g])
            f.setForceGroup(g)
            system.addForce(f)
        ctx = gfp.Context(system, gfp.Platform.getPlatformByName("B200"))
        ctx.setPositions(w.pos.reshape(-1, 3))
        ctx.timeEvaluations(200)
        psecs, _e = ctx.timeEvaluations(steps)
        plugin = {"us_per_step": psecs * 1e6, "steps_per_s": 1.0 / psecs, "grid_force_limited_ns_per_day": 4e-6 * 86400.0 / psecs,
                  "api": "B200 platform plugin: Context::calcForcesAndEnergy over 3 GridForces, C++ step loop"}
        del ctx
    except Exception as exc:      # reported, not fatal: the C-ABI figure above stands on its own
        plugin = {"error": repr(exc)}
    return {"workload": w.name, "us_per_step": secs / steps * 1e6, "steps_per_s": sps, "openmm_plugin_path": plugin,
            "grid_force_limited_ns_per_day": sps * 4e-6 * 86400.0, "value": w.evals * sps, "unit": UNIT,
            "note": "latency-bound: one launch per step on host-mapped memory (15.6 us per call from C++, the rest is ctypes); "
                    "upper bound on MD ns/day at 4 fs"}


def run_other_workload(torch, gf, dev, tdev, stream, name, steps, warmup, peak_gbs, l2_gbs):
    from openmmgridforce_b200 import workloads as W
    if name == "C3":
        w = W.c3_million_atoms()
        n_sets = 8          # 8 x (24 MB positions + 24 MB forces + touched grid lines) >> 126 MB L2
        rng = np.random.default_rng(99)
        length = w.spacing[0] * (w.counts[0] - 1)
        sets = [w.pos] + [rng.uniform(0.0, 0.999 * length, size=w.pos.shape) for _ in range(n_sets - 1)]
    else:
        w = W.c4_batched_replicas()
        sets = [w.pos] + [W.ligand_replicas(w.n_replicas, W.ligand47()[0].mean(axis=0), seed=W.SEED + 10 + i,
                                            escape_shift=(1.0, 0.0, 0.0)) for i in range(15)]
    grids = [gf.Grid(dev, w.counts, w.spacing, w.origin, v, gf.PRECISION_MIXED) for v in w.grids]
    kern = gf.Kernel(dev, grids, w.scaling, oob_k=w.oob_k)
    kern.set_launch_overlap(True)
    pos_sets = [torch.from_numpy(np.ascontiguousarray(p)).to(tdev) for p in sets]
    secs, launches, _ = time_device_steps(torch, gf, kern, pos_sets, w.n_replicas, w.n_atoms, steps, warmup, stream)
    # e2e on the first set
    pos_h, _t1 = pinned_array(w.pos.shape)
    pos_h[...] = w.pos
    f_h, _t2 = pinned_array(w.pos.shape)
    e_h, _t3 = pinned_array((w.n_replicas,))
    e2e_steps = max(3, min(steps, 20))
    e2e_secs = time_e2e_steps(gf, kern, pos_h, f_h, e_h, e2e_steps, 5)
    rate = w.evals * steps / secs
    ach = rate * b_alg(w.n_grids) / 1e9
    out = {"workload": w.name, "value": rate, "unit": UNIT, "us_per_step": secs / steps * 1e6,
           "l2": f"rotating {len(pos_sets)} position/force sets (aggregate footprint > L2)",
           "roofline": {"bound": "hbm", "achieved": ach, "peak": peak_gbs, "unit": "GB/s", "frac": ach / peak_gbs,
                        "traffic": None, "bytes_per_eval": b_alg(w.n_grids)},
           "roofline_l2_gather": {"achieved": ach, "peak": l2_gbs, "unit": "GB/s", "frac": ach / l2_gbs},
           "e2e": {"value": w.evals * e2e_steps / e2e_secs, "unit": UNIT, "h2d_bytes_per_step": int(w.pos.nbytes),
                   "d2h_bytes_per_step": int(w.pos.nbytes + 8 * w.n_replicas)}}
    kern.close()
    for g in grids:
        g.close()
    return out


# ----------------------------------------------------------------------------------------------------------------
# CPU legs (the only users of oracle/)
# ----------------------------------------------------------------------------------------------------------------
def _cpu_oracle_for_sample(bindings, w, n_sample):
    """One Context holding the sample's atoms (replicas flattened), G GridForces — evaluated by the reference's own
    kernel when oracle/_ref is present, else by the C restatement."""
    n = n_sample * w.n_atoms
    pos = np.ascontiguousarray(w.pos[:n_sample].reshape(n, 3))
    scaling = np.tile(w.scaling, (1, n_sample))
    if bindings.ref_available():
        ref = bindings.RefOracle(n, w.counts, w.spacing, w.origin, w.grids, scaling, oob_k=w.oob_k)
        return "reference", (lambda reps: ref.execute_repeat(pos, reps)), n * w.n_grids
    port = bindings.PortOracle(w.counts, w.spacing, w.origin, w.grids, scaling, oob_k=w.oob_k)

    def run(reps):
        for _ in range(reps):
            for g in range(w.n_grids):
                port.execute(pos, g)
    return "port", run, n * w.n_grids


def cpu_baseline(w, target_seconds=12.0, n_sample=2048):
    """Single-threaded (the reference is single-threaded as written) on a bounded sample of the same workload."""
    from oracle import bindings
    bindings.build()
    n_sample = min(n_sample, w.n_replicas)
    kind, run, evals = _cpu_oracle_for_sample(bindings, w, n_sample)
    run(3)                                      # warm-up (also gets the reference's debug prints out of the way)
    t0 = time.perf_counter()
    run(2)
    per = (time.perf_counter() - t0) / 2
    reps = max(3, int(target_seconds / max(per, 1e-6)))
    t0 = time.perf_counter()
    run(reps)
    secs = time.perf_counter() - t0
    return {"value": evals * reps / secs, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"first {n_sample} replicas ({n_sample * w.n_atoms} atoms x {w.n_grids} grids), {reps} passes, {secs:.1f} s"}


def reference_arm(args):
    """The reference's own CPU implementation on all host threads: one Context per thread over disjoint replica shards."""
    from oracle import bindings
    from openmmgridforce_b200 import workloads as W
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    bindings.build()
    threads = os.cpu_count() or 1
    per_thread = 256
    w = W.c5_sharded_replicas(n_replicas=REPLICAS_PER_GPU, n=GRID_N, n_local=threads * per_thread)
    runs = []
    kind = "port"
    for t in range(threads):
        sub = W.Workload(w.name, w.counts, w.spacing, w.origin, w.grids, w.scaling,
                         w.pos[t * per_thread:(t + 1) * per_thread], w.oob_k, w.inv_power)
        kind, run, evals = _cpu_oracle_for_sample(bindings, sub, per_thread)
        runs.append(run)
    evals_per_pass = threads * per_thread * w.n_atoms * w.n_grids
    runs[0](3)                                  # serial warm-up: the reference's static debug counters are not thread-safe

    def one_step(reps):
        th = [threading.Thread(target=r, args=(reps,)) for r in runs]
        for t in th:
            t.start()
        for t in th:
            t.join()

    one_step(1)
    t0 = time.perf_counter()
    one_step(2)
    per = (time.perf_counter() - t0) / 2
    reps = max(1, int(2.0 / max(per, 1e-6)))    # ~2 s of wall clock per step
    for _ in range(args.warmup):
        one_step(reps)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step(reps)
    secs = time.perf_counter() - t0
    value = evals_per_pass * reps * args.steps / secs
    sample = (f"{threads} threads x {per_thread} replicas x {w.n_atoms} atoms x {w.n_grids} grids, {reps} passes per step "
              f"(ctypes releases the GIL; one reference Context per thread)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {"workload": f"configs[4]: ligand replicas x {N_ATOMS} atoms x {N_GRIDS} grids of {GRID_N}^3, "
                        f"{REPLICAS_PER_GPU} replicas per GPU (N=1 is configs[4] on one GPU)",
            "replicas_per_gpu": REPLICAS_PER_GPU, "replicas_total": REPLICAS_PER_GPU * n_gpus, "atoms_per_replica": N_ATOMS,
            "grids": N_GRIDS, "grid_points": [GRID_N] * 3, "precision": "mixed", "parallelism": f"replica-sharded x{n_gpus}",
            "l2": "inputs larger than L2: each step streams 74 MB of positions + 74 MB of forces per GPU and gathers from "
                  "3 grids; no L2 flush between steps",
            "launch_overlap": "programmatic dependent launch: a step's blocks may fetch their (static) inputs during the "
                              "previous step's tail and wait for it before their first write (gfb_kernel_set_launch_overlap)",
            "energy_gather": "N>1: per-replica energies of every rank gathered on every rank; config.energy_gather_when = final "
                             "(once, after the last step, inside the timed region: BASELINE.json north_star) or every (after "
                             "every step, overlapping the next two steps); config.energy_gather_mode = nccl "
                             "(all_gather_into_tensor) or peer-put (copy-engine puts over NVLink into symmetric memory + "
                             "signal barrier)"}


# ----------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip C3/C4 and the CPU baseline (N=1 only)")
    ap.add_argument("--energy-gather", default=os.environ.get("GFB_ENERGY_GATHER_WHEN", "final"), choices=["final", "every"],
                    help="N>1: gather per-replica energies once, after the last step (what BASELINE.json's north_star asks "
                         "for; inside the timed region), or after every step (overlapping the next steps)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import openmmgridforce_b200 as gf
    from openmmgridforce_b200 import workloads as W

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    tdev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_cpus(local_rank)     # pinned host buffers are then first-touched on the GPU's own NUMA node
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=tdev)

    dev = gf.Device(local_rank)         # raises if the CUDA library or an sm_100 GPU is missing: no fallback
    stream = torch.cuda.Stream(device=tdev)
    peak_gbs, peak_src = measured_peaks()

    # this rank's batch: its own 65,536 replica poses (same grids and scaling factors on every rank)
    w = W.c5_sharded_replicas(n_replicas=REPLICAS_PER_GPU, n=GRID_N, pose_seed=W.SEED + rank)
    grids = [gf.Grid(dev, w.counts, w.spacing, w.origin, v, gf.PRECISION_MIXED) for v in w.grids]
    kern = gf.Kernel(dev, grids, w.scaling, oob_k=w.oob_k)
    # back-to-back evaluations of resident poses: a launch may start while the previous one's tail is still running (PDL);
    # its inputs are not produced by that previous launch. Host-path (e2e) launches are unaffected.
    # (with a gather after EVERY step the overlap is switched off: the two together measured slower, 88.2 vs 84.7 us at N=2)
    kern.set_launch_overlap(os.environ.get("GFB_BENCH_PDL", "1") != "0" and (world == 1 or args.energy_gather == "final"))
    d_pos = torch.from_numpy(w.pos).to(tdev)

    from openmmgridforce_b200 import sharding

    # The one collective (N > 1): per-replica energies of every rank, every step, overlapping the next step's kernel.
    #   "nccl"      torch.distributed all_gather_into_tensor, asynchronous (default: what the north star names);
    #   "peer-put"  GFB_ENERGY_GATHER=peer-put: each rank copies its 512 KB into its slot of every peer's buffer with the
    #               copy engines over NVLink (gfb_peer_put on torch symmetric-memory buffers) and then passes a signal-pad
    #               barrier, so no SM is taken from the evaluation kernel running alongside. Measured on the same boxes:
    #               N=8 807 vs 791 G evals/s, N=2 209 vs 213 — within box-to-box noise, so NCCL stays the default.
    gather_mode = "none"
    counter = [0]
    if world > 1:
        gather_mode = os.environ.get("GFB_ENERGY_GATHER", "nccl")
        if gather_mode == "peer-put":
            try:
                import torch.distributed._symmetric_memory as symm
                sym_buf = symm.empty(2 * world * REPLICAS_PER_GPU, dtype=torch.float64, device=tdev)
                sym_hdl = symm.rendezvous(sym_buf, dist.group.WORLD)
                peer_ptrs = [int(p) for p in sym_hdl.buffer_ptrs]
            except Exception as exc:      # no symmetric memory on this box/build: use NCCL, and say so
                print(f"[bench] symmetric memory unavailable ({exc!r}); energy gather falls back to NCCL", file=sys.stderr)
                gather_mode = "nccl"
        if gather_mode == "nccl":
            gathered2 = [torch.empty(world * REPLICAS_PER_GPU, dtype=torch.float64, device=tdev) for _ in range(2)]
        gather_stream = torch.cuda.Stream(device=tdev)

    class _GatherDone:
        """Completion of a gather issued on gather_stream: .wait() makes the launching stream wait for it (what an NCCL
        Work's wait() does); .ev is the CUDA event."""
        def __init__(self, ev):
            self.ev = ev

        def wait(self):
            torch.cuda.current_stream().wait_event(self.ev)

    def post_step(d_e, force=False):
        if world == 1 or (args.energy_gather == "final" and not force):
            return None
        counter[0] += 1
        b = counter[0] % 2
        if gather_mode == "nccl" and not force:
            return dist.all_gather_into_tensor(gathered2[b], d_e, async_op=True)     # NCCL's own stream; Work.wait()
        ready = torch.cuda.Event()
        ready.record(stream)                       # the step's kernel has produced d_e
        gather_stream.wait_event(ready)
        if gather_mode == "nccl":
            with torch.cuda.stream(gather_stream):
                dist.all_gather_into_tensor(gathered2[b], d_e)      # enqueued behind gather_stream; returns at once
        else:
            dev.peer_put(d_e.data_ptr(), peer_ptrs, (b * world + rank) * REPLICAS_PER_GPU * 8, REPLICAS_PER_GPU * 8,
                         first_peer=rank + 1, stream=gather_stream.cuda_stream)
            with torch.cuda.stream(gather_stream):
                sym_hdl.barrier(channel=b)         # every rank's puts of this step have landed everywhere
        done = torch.cuda.Event()
        done.record(gather_stream)
        return _GatherDone(done)

    def gathered_view(b):
        if gather_mode == "nccl":
            return gathered2[b]
        return sym_buf.view(2, world * REPLICAS_PER_GPU)[b]

    l2_gbs = dev.bench_sector_gather(32 << 20, 1 << 24, 10) if rank == 0 else 0.0

    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    energy_bufs = []
    secs, launches, bufs = time_device_steps(torch, gf, kern, [d_pos], REPLICAS_PER_GPU, N_ATOMS, args.steps, args.warmup, stream,
                                             post_step=post_step, energy_bufs_out=energy_bufs,
                                             final_step=(lambda d_e: post_step(d_e, force=True))
                                             if world > 1 and args.energy_gather == "final" else None,
                                             barrier=(lambda: (dist.barrier(), torch.cuda.synchronize())) if world > 1 else None)
    torch.cuda.synchronize()
    if world > 1:
        # outside the timed region: the last step's gathered energies must equal a plain blocking NCCL all-gather of them
        last = (args.warmup + args.steps - 1) % 4
        check = torch.empty(world * REPLICAS_PER_GPU, dtype=torch.float64, device=tdev)
        dist.all_gather_into_tensor(check, energy_bufs[last])
        if not torch.equal(check, gathered_view(counter[0] % 2)):
            raise SystemExit(f"rank {rank}: energy gather ({gather_mode}) does not match NCCL all_gather")
    per_rank_us = [secs / args.steps * 1e6]
    if world > 1:
        t = torch.tensor([secs], dtype=torch.float64, device=tdev)
        all_t = torch.empty(world, dtype=torch.float64, device=tdev)
        dist.all_gather_into_tensor(all_t, t)          # every rank's own device-timed loop: the value uses the MAX
        per_rank_us = [float(x) / args.steps * 1e6 for x in all_t.tolist()]
        secs_max = float(all_t.max().item())
        dist.barrier()
    else:
        secs_max = secs

    # e2e: public host API with pinned buffers, every rank on its own batch
    pos_h, _k1 = pinned_array(w.pos.shape)
    pos_h[...] = w.pos
    f_h, _k2 = pinned_array(w.pos.shape)
    e_h, _k3 = pinned_array((REPLICAS_PER_GPU,))
    e2e_steps = max(3, min(args.steps, 20))
    if world > 1:
        dist.barrier()
    e2e_secs = time_e2e_steps(gf, kern, pos_h, f_h, e_h, e2e_steps, 5)
    if world > 1:
        t = torch.tensor([e2e_secs], dtype=torch.float64, device=tdev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_secs = float(t.item())

    extras = {}
    if rank == 0 and world == 1 and not args.no_extras:
        for name in ("C3", "C4"):
            extras[name] = run_other_workload(torch, gf, dev, tdev, stream, name, max(20, min(args.steps, 200)), args.warmup,
                                              peak_gbs, l2_gbs)
        extras["C2"] = run_single_ligand(gf, dev)
    clocks = sampler.stop()

    evals_step_rank = REPLICAS_PER_GPU * N_ATOMS * N_GRIDS
    value = evals_step_rank * world * args.steps / secs_max
    kernel_us = secs / args.steps * 1e6        # this rank's average step (N=1: exactly the kernel's launch-to-launch time)
    ach = evals_step_rank * b_alg(N_GRIDS) / (kernel_us * 1e-6) / 1e9

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": secs_max / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 interpolation, f64 index/energy, i64 fixed-point forces", "data": "synthetic",
                "config": dict(workload_config(world), energy_gather_mode=gather_mode,
                               energy_gather_when=args.energy_gather if world > 1 else "none"),
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak_gbs, "unit": "GB/s", "frac": ach / peak_gbs,
                             "traffic": C5_DRAM_BYTES_PER_LAUNCH, "traffic_unit": "bytes per launch",
                             "traffic_source": "profiles/r1c_c5_lines_warm_raw.csv: dram__bytes_read.sum 397.1 MB + "
                                               "dram__bytes_write.sum 98.3 MB (ncu --set full, this kernel, this workload)",
                             "peak_source": peak_src, "kernel": "gf_eval_lines_kernel<3 grids, FIXED_ADD> (one 128-byte record per cell)",
                             "bytes_per_eval": b_alg(N_GRIDS), "evals_per_launch": evals_step_rank,
                             "launch_us": kernel_us},
                "roofline_l2_gather": {"achieved": ach, "peak": l2_gbs, "unit": "GB/s", "frac": ach / l2_gbs if l2_gbs else None,
                                       "peak_source": "gfb_bench_sector_gather: random 32-byte sectors over 32 MB, this run"},
                "e2e": {"value": evals_step_rank * world * e2e_steps / e2e_secs, "unit": UNIT,
                        "h2d_bytes_per_step": int(w.pos.nbytes), "d2h_bytes_per_step": int(w.pos.nbytes + 8 * REPLICAS_PER_GPU),
                        "api": "gfb_kernel_execute_host (pinned host positions in, forces + energies out; H2D by copy engine in 8 chunks, "
                               "forces stored by the kernels straight into the pinned host buffer)"},
                "gpu_launches": int(launches), "clocks": clocks,
                "per_rank_us_per_step": [round(x, 2) for x in per_rank_us], "cpu_binding": numa}
        if world == 1 and not args.no_extras:
            line["other_workloads"] = extras
            line["cpu_baseline"] = cpu_baseline(w)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
