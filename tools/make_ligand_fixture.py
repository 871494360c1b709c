"""Extracts the 47-atom ligand fixture (coordinates, charges) from the reference's AMBER test files into
openmmgridforce_b200/data/ligand47.json. Run in the build container (needs /root/reference); the JSON is
committed because /root/reference does not exist on the GPU box.

Sources: python/prmtopcrd/ligand.trans.inpcrd (Angstrom, 12.7 fixed width, 6 per line) and
python/prmtopcrd/ligand.prmtop %FLAG CHARGE (AMBER internal units: e * 18.2223), as used by
python/tests/test_grid_force.py:117-138.
"""
import json
import os
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"


def read_inpcrd(path):
    lines = open(path).read().splitlines()
    n = int(lines[1].split()[0])
    vals = []
    for ln in lines[2:]:
        vals += [float(ln[i:i + 12]) for i in range(0, len(ln.rstrip()), 12)]
    return n, vals[:3 * n]


def read_flag(path, flag):
    out, on = [], False
    for ln in open(path):
        if ln.startswith("%FLAG"):
            on = ln.split()[1] == flag
            continue
        if on and not ln.startswith("%FORMAT"):
            out += ln.split()
    return out


n, xyz = read_inpcrd(os.path.join(REF, "python/prmtopcrd/ligand.trans.inpcrd"))
charges = [float(v) / 18.2223 for v in read_flag(os.path.join(REF, "python/prmtopcrd/ligand.prmtop"), "CHARGE")]
assert n == 47 and len(charges) == 47
doc = {
    "source": "jimtufts/openmmgridforce python/prmtopcrd/ligand.{trans.inpcrd,prmtop}",
    "units": {"positions": "nm", "charges": "e"},
    "positions_nm": [[round(xyz[3 * i + k] * 0.1, 8) for k in range(3)] for i in range(n)],
    "charges_e": [round(c, 8) for c in charges],
}
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "openmmgridforce_b200", "data", "ligand47.json")
json.dump(doc, open(dst, "w"), indent=0)
print("wrote", os.path.normpath(dst))
