#!/bin/bash
# Are the PROFILED kernels' machine code unchanged between a git revision and the working tree?
#   tools/sass_identity.sh [REV]      (default HEAD)
# profiles/r2_traffic.json is stamped with a hash of the kernel sources (bench.py: KERNEL_SOURCES) and bench.py drops its
# measured-traffic figures when the sources change. When an edit to a shared header cannot touch the profiled kernels
# (e.g. a new interpolation branch of the general kernel in gf_kernels.cuh), this script proves it: it compiles the
# translation units of gf_eval_lines_kernel (1 and 3 grids) and gf_eval_lines_f64_kernel from both trees and compares
# the SASS. Only then may the traffic file be re-stamped:  python tools/ncu_traffic.py --restamp "<why>"
set -e
REV=${1:-HEAD}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$ROOT/build/sasscheck
rm -rf "$W" && mkdir -p "$W/old"
cd "$ROOT"
for f in $(git ls-files openmmgridforce_b200/csrc include); do mkdir -p "$W/old/$(dirname $f)"; git show $REV:$f > "$W/old/$f" 2>/dev/null || true; done
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
unit() {   # name, source, extra flags
    nvcc $F -I"$W/old/include" -I"$W/old/openmmgridforce_b200/csrc" $3 -cubin -o "$W/old_$1.cubin" "$W/old/openmmgridforce_b200/csrc/$2" &
    nvcc $F -Iinclude -Iopenmmgridforce_b200/csrc $3 -cubin -o "$W/new_$1.cubin" "openmmgridforce_b200/csrc/$2" &
}
unit l3 gf_launch_lines.cu -DGFB_LINES_NG=3
unit l1 gf_launch_lines.cu -DGFB_LINES_NG=1
unit f64 gf_launch_records_f64.cu ""
wait
rc=0
for k in l3 l1 f64; do
    for t in old new; do cuobjdump -sass "$W/${t}_$k.cubin" | grep -v '^\s*//\|Fatbin\|=====' > "$W/${t}_$k.sass"; done
    if cmp -s "$W/old_$k.sass" "$W/new_$k.sass"; then echo "$k: SASS identical to $REV ($(wc -l < "$W/new_$k.sass") lines)"; else echo "$k: SASS DIFFERS from $REV"; rc=1; fi
done
exit $rc
