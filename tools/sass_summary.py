"""SASS evidence for the key kernels (no GPU needed): python tools/sass_summary.py > profiles/r1b_sass_summary.txt
Counts the memory / atomic / PDL / FP64 mnemonics of `cuobjdump -sass lib/libgridforce_b200.so` per kernel."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "openmmgridforce_b200", "lib", "libgridforce_b200.so")
WANT = {
    "_ZN3gfb20gf_eval_lines_kernelILi3ELi2ELi0ELb0ELb0EEEvNS_10EvalParamsE": "gf_eval_lines_kernel<3 grids, FIXED_ADD, red, batched> — the C5/C4 step",
    "_ZN3gfb20gf_eval_lines_kernelILi1ELi2ELi1ELb1ELb0EEEvNS_10EvalParamsE": "gf_eval_lines_kernel<1 grid, FIXED_ADD, prefetch, single> — the C3 step",
    "_ZN3gfb20gf_eval_lines_kernelILi3ELi0ELi0ELb0ELb0EEEvNS_10EvalParamsE": "gf_eval_lines_kernel<3 grids, F64_STORE> — host-path chunks (forces stored to pinned host memory)",
    "_ZN3gfb22gf_eval_bspline_kernelILi2ELb0EEEvNS_10EvalParamsE": "gf_eval_bspline_kernel<FIXED_ADD, batched> — cubic B-spline bricks",
    "_ZN3gfb14gf_eval_kernelIdLi1ELi3ELb1ELi2ELb0EEEvNS_10EvalParamsE": "gf_eval_kernel<double, CELLS, 3 grids, FIXED_ADD> — DOUBLE precision",
}
PATS = ["LDGSTS.E.BYPASS.128", "LDG.E.ELL2.256", "LDG.E.EF.128", "LDG.E.EF.64", "LDS.128", "STS", "STG.E.128", "STG.E.64",
        "REDG.E.ADD.64", "REDG.E.ADD.F64", "CCTL", "PREEXIT", "ACQBULK", "LDGDEPBAR", "DEPBAR", "DFMA", "DMUL", "DADD",
        "F2F.F64.F32", "FFMA", "SHFL", "BAR.SYNC", "WARPSYNC", "CALL"]
NOTE = {"LDGSTS.E.BYPASS.128": "cp.async.cg 16 B (one granule of a 128-byte record / a brick row)",
        "LDG.E.ELL2.256": "ld.global.nc.L2::evict_last.v8.f32 (32-byte stencil)", "LDG.E.EF.128": "ld.global.cs.v2.f64 (streamed positions)",
        "REDG.E.ADD.64": "red.global.add.u64 (fixed-point forces)", "REDG.E.ADD.F64": "red.global.add.f64 (energies)",
        "PREEXIT": "griddepcontrol.launch_dependents", "ACQBULK": "griddepcontrol.wait", "CCTL": "prefetch.global.L2 (force lines)"}
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
print("cuobjdump -sass openmmgridforce_b200/lib/libgridforce_b200.so (sm_100a), instruction counts per kernel\n")
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    if name not in WANT:
        continue
    ins = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f)
    c = collections.Counter()
    for i in ins:
        for p in PATS:
            if i.startswith(p):
                c[p] += 1
    print(WANT[name])
    print(f"  {name}: {len(ins)} SASS instructions")
    for p in PATS:
        if c[p]:
            print(f"    {p:22s} x{c[p]:<4d} {NOTE.get(p, '')}")
    print()
