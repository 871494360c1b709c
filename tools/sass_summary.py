"""SASS evidence for the key kernels (no GPU needed): python tools/sass_summary.py > profiles/r2_sass_summary.txt
Counts the memory / atomic / PDL / FP64 mnemonics of `cuobjdump -sass lib/libgridforce_b200.so` per kernel."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "openmmgridforce_b200", "lib", "libgridforce_b200.so")
WANT = {   # round-2 names (the lines kernel gained the PERSIST flag; substring match on the demangled-free symbol)
    "gf_eval_lines_kernelILi3ELi2ELi0ELb0ELb0ELb0E": "gf_eval_lines_kernel<3 grids, FIXED_ADD, red, batched, PERSIST=0> — the C5 step (one block per tile)",
    "gf_eval_lines_kernelILi3ELi2ELi0ELb0ELb0ELb1E": "gf_eval_lines_kernel<3 grids, FIXED_ADD, red, batched, PERSIST=1> — small launches under launch overlap (tile-striding)",
    "gf_eval_lines_kernelILi1ELi2ELi1ELb1ELb0ELb0E": "gf_eval_lines_kernel<1 grid, FIXED_ADD, prefetch, single> — the C3 step",
    "gf_eval_lines_kernelILi3ELi3ELi0ELb0ELb0ELb0E": "gf_eval_lines_kernel<3 grids, F32_STORE> — host-path chunks (FP32 forces stored to pinned host memory)",
    "gf_eval_lines_kernelILi3ELi4ELi0ELb0ELb0ELb0E": "gf_eval_lines_kernel<3 grids, energy only>",
    "gf_eval_lines_f64_kernelILi3ELi2ELb0ELb0E": "gf_eval_lines_f64_kernel<3 grids, FIXED_ADD> — DOUBLE 256-byte records",
    "gf_eval_bspline_kernelILb0E": "gf_eval_bspline_kernel<batched> — MIXED cubic B-spline records",
    "gf_eval_bspline_f64_kernelILb0E": "gf_eval_bspline_f64_kernel<batched> — DOUBLE cubic B-spline records",
    "gf_gather_ll_kernel": "gf_gather_ll_kernel — the energy gather as one flag-in-data kernel over NVLink peer mappings",
    "gf_gather_push_kernel": "gf_gather_push_kernel — peer stores + arrival flags",
    "gf_gather_wait_kernel": "gf_gather_wait_kernel — waits for all ranks' flags, copies the gathered array out",
    "gf_rendezvous_kernel": "gf_rendezvous_kernel — device-side rendezvous of all ranks",
    "gf_resident_kernelIfE": "gf_resident_kernel<float> — resident evaluator (one ligand per MD step), MIXED: packets polled and sent with 16-byte .SYS accesses, no MEMBAR",
    "gf_eval_kernelIdLi5E": "gf_eval_kernel<double, POINTS> — tricubic Hermite (interpolation method 2), DOUBLE",
}
PATS = ["MEMBAR", "LDGSTS.E.BYPASS.128", "LDG.E.ELL2.256", "LDG.E.LTC64B.ELL2.256", "LDG.E.128.STRONG.SYS", "STG.E.128.STRONG.SYS", "STG.E.64.STRONG.SYS", "LDG.E.64.STRONG.SYS", "LDS.64", "LDG.E.EF.128", "LDG.E.EF.64", "LDS.128", "STS", "STG.E.128", "STG.E.64",
        "REDG.E.ADD.64", "REDG.E.ADD.F64", "CCTL", "PREEXIT", "ACQBULK", "LDGDEPBAR", "DEPBAR", "DFMA", "DMUL", "DADD",
        "F2F.F64.F32", "FFMA", "SHFL", "BAR.SYNC", "WARPSYNC", "CALL"]
NOTE = {"LDGSTS.E.BYPASS.128": "cp.async.cg 16 B (one granule of a 128-byte record / a brick row)",
        "LDG.E.ELL2.256": "ld.global.nc.L2::evict_last.v8.f32 (32-byte stencil)", "LDG.E.EF.128": "ld.global.cs.v2.f64 (streamed positions)",
        "REDG.E.ADD.64": "red.global.add.u64 (fixed-point forces)", "REDG.E.ADD.F64": "red.global.add.f64 (energies)",
        "PREEXIT": "griddepcontrol.launch_dependents", "ACQBULK": "griddepcontrol.wait", "CCTL": "prefetch.global.L2 (force lines)"}
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
print("cuobjdump -sass openmmgridforce_b200/lib/libgridforce_b200.so (sm_100a), instruction counts per kernel\n")
seen = set()
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    key = next((k for k in WANT if k in name), None)
    if key is None or key in seen:
        continue
    seen.add(key)
    ins = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f)
    c = collections.Counter()
    for i in ins:
        for p in PATS:
            if i.startswith(p):
                c[p] += 1
    print(WANT[key])
    print(f"  {name}: {len(ins)} SASS instructions")
    for p in PATS:
        if c[p]:
            print(f"    {p:22s} x{c[p]:<4d} {NOTE.get(p, '')}")
    print()
