"""TEST INFRASTRUCTURE — ctypes bindings of the two parity oracles. Import only from tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__ (smoke check, oracle build). Never from the product.

  PortOracle  oracle/liboracle_port.so   — the plain-C restatement (gridforce_oracle.c)
  RefOracle   oracle/_ref/liboracle_ref.so — the reference's own unmodified kernel behind ref_driver.cpp
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_PATH = os.path.join(HERE, "liboracle_port.so")
REF_PATH = os.path.join(HERE, "_ref", "liboracle_ref.so")


def build(quiet=True):
    """make -C oracle all: always (re)builds the C port; builds _ref only where /root/reference exists."""
    subprocess.run(["make", "-C", HERE, "all"], check=True, stdout=subprocess.DEVNULL if quiet else None)


class _Grid(C.Structure):
    _fields_ = [("counts", C.c_int * 3), ("spacing", C.c_double * 3), ("origin", C.c_double * 3),
                ("vals", C.POINTER(C.c_double)), ("inv_power", C.c_double), ("oob_k", C.c_double),
                ("interp_method", C.c_int)]


CLASS_DTYPE = np.dtype([("inside", np.int32), ("cell", np.int32, (3,))])


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class PortOracle:
    """G grids acting on the same A atoms (C restatement)."""

    def __init__(self, counts, spacing, origin, grids, scaling, oob_k=None, inv_power=None, interpolation_method=0):
        if not os.path.exists(PORT_PATH):
            build()
        self.lib = C.CDLL(PORT_PATH)
        self.lib.gfo_execute.restype = C.c_double
        self.lib.gfo_execute.argtypes = [C.POINTER(_Grid), C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int),
                                         C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p]
        self.lib.gfo_execute_batched.restype = None
        self.lib.gfo_execute_batched.argtypes = [C.POINTER(_Grid), C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int,
                                                 C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
        self._vals = [_f64(g).ravel() for g in grids]
        self.n_grids = len(self._vals)
        self.scaling = _f64(scaling).reshape(self.n_grids, -1)
        self.n_atoms = self.scaling.shape[1]
        oob_k = oob_k if oob_k is not None else [10000.0] * self.n_grids
        inv_power = inv_power if inv_power is not None else [0.0] * self.n_grids
        self.grids = (_Grid * self.n_grids)()
        for g in range(self.n_grids):
            self.grids[g].counts = (C.c_int * 3)(*counts)
            self.grids[g].spacing = (C.c_double * 3)(*spacing)
            self.grids[g].origin = (C.c_double * 3)(*origin)
            self.grids[g].vals = _dp(self._vals[g])
            self.grids[g].inv_power = inv_power[g]
            self.grids[g].oob_k = oob_k[g]
            self.grids[g].interp_method = interpolation_method

    def execute(self, pos, grid=0, ligand_atoms=None, classify=False):
        """One GridForce, one Context. pos [P,3]. Returns (E, forces[A,3], cls or None)."""
        pos = _f64(pos)
        forces = np.zeros((self.n_atoms, 3))
        cls = np.zeros(self.n_atoms, dtype=CLASS_DTYPE) if classify else None
        la = np.ascontiguousarray(ligand_atoms, dtype=np.int32) if ligand_atoms is not None else None
        e = self.lib.gfo_execute(C.byref(self.grids[grid]), _dp(self.scaling[grid]), self.n_atoms,
                                 la.ctypes.data_as(C.POINTER(C.c_int)) if la is not None else None, _dp(pos), _dp(forces),
                                 cls.ctypes.data_as(C.c_void_p) if classify else None)
        return e, forces, cls

    def execute_batched(self, pos, n_threads=1, want_forces=True):
        """pos [R,A,3]. Returns (grid_energies [R,G], forces [R,A,3] or None)."""
        pos = _f64(pos)
        r = pos.shape[0]
        forces = np.empty((r, self.n_atoms, 3)) if want_forces else None
        en = np.empty((r, self.n_grids))
        self.lib.gfo_execute_batched(self.grids, self.n_grids, _dp(self.scaling), r, self.n_atoms, _dp(pos),
                                     _dp(forces) if want_forces else None, _dp(en), n_threads)
        return en, forces


def ref_available():
    return os.path.exists(REF_PATH)


def ref_save_file(mode, path, counts, spacing, origin, values, grid_type="", inv_power=0.0, inv_power_mode=0):
    """mode 0: the reference's GridForce::saveToFile; mode 1: its GridData::saveToFile."""
    lib = C.CDLL(REF_PATH)
    lib.oracle_ref_last_error.restype = C.c_char_p
    v = _f64(values).ravel()
    rc = lib.oracle_ref_save_file(mode, os.fsencode(path), (C.c_int * 3)(*counts), (C.c_double * 3)(*spacing),
                                  (C.c_double * 3)(*origin), _dp(v), C.c_longlong(v.size), grid_type.encode(),
                                  C.c_double(inv_power), inv_power_mode)
    if rc:
        raise RuntimeError(lib.oracle_ref_last_error().decode())


def ref_load_file(path, capacity):
    """The reference's GridForce::loadFromFile -> (counts, spacing, origin, values, inv_power, inv_power_mode)."""
    lib = C.CDLL(REF_PATH)
    lib.oracle_ref_last_error.restype = C.c_char_p
    counts, spacing, origin = (C.c_int * 3)(), (C.c_double * 3)(), (C.c_double * 3)()
    vals = np.empty(capacity)
    ip, mode = C.c_double(0), C.c_int(0)
    rc = lib.oracle_ref_load_file(os.fsencode(path), counts, spacing, origin, _dp(vals), C.c_longlong(capacity), C.byref(ip),
                                  C.byref(mode))
    if rc:
        raise RuntimeError(lib.oracle_ref_last_error().decode())
    n = counts[0] * counts[1] * counts[2]
    return tuple(counts), tuple(spacing), tuple(origin), vals[:n].reshape(tuple(counts)), ip.value, mode.value


class RefOracle:
    """The reference's own ReferenceCalcGridForceKernel: one System with P particles and G GridForces."""

    def __init__(self, n_particles, counts, spacing, origin, grids, scaling, oob_k=None, inv_power=None,
                 ligand_atoms=None, interpolation_method=0):
        if not ref_available():
            raise RuntimeError(f"{REF_PATH} not built (needs /root/reference; see oracle/Makefile)")
        lib = C.CDLL(REF_PATH)
        lib.oracle_ref_create.restype = C.c_void_p
        lib.oracle_ref_create.argtypes = [C.c_int]
        lib.oracle_ref_add_grid.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                            C.POINTER(C.c_double), C.c_longlong, C.POINTER(C.c_double), C.c_int,
                                            C.POINTER(C.c_int), C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
        lib.oracle_ref_finalize.argtypes = [C.c_void_p]
        lib.oracle_ref_execute.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double),
                                           C.POINTER(C.c_double)]
        lib.oracle_ref_execute_repeat.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]
        lib.oracle_ref_destroy.argtypes = [C.c_void_p]
        lib.oracle_ref_last_error.restype = C.c_char_p
        self.lib = lib
        self.n_particles = n_particles
        self.h = lib.oracle_ref_create(n_particles)
        scaling = _f64(scaling)
        if scaling.ndim == 1:
            scaling = scaling.reshape(1, -1)
        n_grids = len(grids)
        oob_k = oob_k if oob_k is not None else [10000.0] * n_grids
        inv_power = inv_power if inv_power is not None else [0.0] * n_grids
        la = np.ascontiguousarray(ligand_atoms, dtype=np.int32) if ligand_atoms is not None else None
        for g in range(n_grids):
            v = _f64(grids[g]).ravel()
            rc = lib.oracle_ref_add_grid(self.h, (C.c_int * 3)(*counts), (C.c_double * 3)(*spacing), (C.c_double * 3)(*origin),
                                         _dp(v), v.size, _dp(scaling[g]), scaling.shape[1],
                                         la.ctypes.data_as(C.POINTER(C.c_int)) if la is not None else None,
                                         la.size if la is not None else 0, inv_power[g], oob_k[g], interpolation_method, g)
            if rc:
                raise RuntimeError(lib.oracle_ref_last_error().decode())
        if lib.oracle_ref_finalize(self.h):
            raise RuntimeError(lib.oracle_ref_last_error().decode())

    def execute(self, pos, groups=0xFFFFFFFF):
        """pos [P,3] -> (energy summed over the forces in `groups`, forces [P,3])"""
        pos = _f64(pos)
        assert pos.shape == (self.n_particles, 3)
        e = C.c_double(0.0)
        f = np.zeros((self.n_particles, 3))
        if self.lib.oracle_ref_execute(self.h, _dp(pos), C.c_int(groups & 0x7FFFFFFF if groups != 0xFFFFFFFF else -1), C.byref(e), _dp(f)):
            raise RuntimeError(self.lib.oracle_ref_last_error().decode())
        return e.value, f

    def execute_repeat(self, pos, reps):
        pos = _f64(pos)
        e = C.c_double(0.0)
        if self.lib.oracle_ref_execute_repeat(self.h, _dp(pos), reps, C.byref(e)):
            raise RuntimeError(self.lib.oracle_ref_last_error().decode())
        return e.value

    def close(self):
        if self.h:
            self.lib.oracle_ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


GRID_TYPE_CODES = {"charge": 1, "ljr": 2, "lja": 3}


def port_generate_grid(counts, spacing, origin, grid_type, pos, charges, sigmas, epsilons, grid_cap=41840.0, n_threads=1):
    """C restatement of the reference's generateGrid -> [nx, ny, nz] float64."""
    if not os.path.exists(PORT_PATH):
        build()
    lib = C.CDLL(PORT_PATH)
    pos, q, sg, ep = _f64(pos), _f64(charges), _f64(sigmas), _f64(epsilons)
    out = np.empty(tuple(counts))
    lib.gfo_generate_grid.restype = None
    lib.gfo_generate_grid((C.c_int * 3)(*counts), (C.c_double * 3)(*spacing), (C.c_double * 3)(*origin),
                          GRID_TYPE_CODES[grid_type], pos.shape[0], _dp(pos), _dp(q), _dp(sg), _dp(ep), C.c_double(grid_cap),
                          _dp(out), n_threads)
    return out


def ref_generate_grid(counts, spacing, origin, grid_type, pos, charges, sigmas, epsilons, grid_cap=41840.0):
    """The reference's own auto-generation path (Context creation on a System with a NonbondedForce)."""
    lib = C.CDLL(REF_PATH)
    lib.oracle_ref_last_error.restype = C.c_char_p
    pos, q, sg, ep = _f64(pos), _f64(charges), _f64(sigmas), _f64(epsilons)
    out = np.empty(tuple(counts))
    rc = lib.oracle_ref_generate_grid((C.c_int * 3)(*counts), (C.c_double * 3)(*spacing), (C.c_double * 3)(*origin),
                                      grid_type.encode(), pos.shape[0], _dp(pos), _dp(q), _dp(sg), _dp(ep),
                                      C.c_double(grid_cap), _dp(out))
    if rc:
        raise RuntimeError(lib.oracle_ref_last_error().decode())
    return out


def port_inv_power_transform(values, inv_power):
    """C restatement of GridForce::applyInvPowerTransformation -> new array."""
    if not os.path.exists(PORT_PATH):
        build()
    lib = C.CDLL(PORT_PATH)
    out = np.array(values, dtype=np.float64, order="C", copy=True)
    lib.gfo_inv_power_transform.restype = None
    lib.gfo_inv_power_transform(_dp(out), C.c_size_t(out.size), C.c_double(inv_power))
    return out


def ref_inv_power_transform(values, inv_power):
    """The reference's own GridForce::applyInvPowerTransformation on a RUNTIME-mode grid -> (new array, mode after)."""
    lib = C.CDLL(REF_PATH)
    lib.oracle_ref_last_error.restype = C.c_char_p
    out = np.array(values, dtype=np.float64, order="C", copy=True)
    counts = (C.c_int * 3)(*out.shape) if out.ndim == 3 else (C.c_int * 3)(out.size, 1, 1)
    mode = C.c_int(-1)
    rc = lib.oracle_ref_apply_inv_power(counts, _dp(out), C.c_longlong(out.size), C.c_double(inv_power), C.byref(mode))
    if rc:
        raise RuntimeError(lib.oracle_ref_last_error().decode())
    return out, mode.value
