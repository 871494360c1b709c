"""ctypes binding of include/gridforce_b200.h (one Python method per C entry point).

Mirrors what a SWIG/cgo/JNI stub over the same header would do; see INTEGRATION.md for the SWIG form the
reference's python/gridforceplugin.i would carry. numpy arrays are passed as raw host pointers, torch CUDA
tensors as raw device pointers (``tensor.data_ptr()``): no torch type crosses the boundary.
"""
import ctypes as C
import os

import numpy as np

PRECISION_MIXED = 0
PRECISION_DOUBLE = 1
FORCE_F64_STORE = 0
FORCE_F64_ADD = 1
FORCE_FIXED_ADD = 2
FORCE_F32_STORE = 3
COMM_ID_BYTES = 128
IPC_HANDLE_BYTES = 64
LAYOUT_AUTO, LAYOUT_CELLS, LAYOUT_ROWS, LAYOUT_PAIRS, LAYOUT_BSPLINE, LAYOUT_POINTS, LAYOUT_HERMITE, LAYOUT_BSPLINE_POINTS = 0, 1, 2, 3, 4, 5, 6, 7
LAYOUT_NAMES = {0: "auto", 1: "cells", 2: "rows", 3: "pairs", 4: "bspline", 5: "points", 6: "hermite", 7: "bspline_points"}
MAX_GRIDS = 8

_LIB = None


class GridForceB200Error(RuntimeError):
    """Raised for every non-zero status from the C ABI (the reference maps OpenMMException to RuntimeError
    the same way, python/gridforceplugin.i:49-59)."""


class GridFileHeader(C.Structure):
    _fields_ = [("counts", C.c_int * 3), ("spacing", C.c_double * 3), ("origin", C.c_double * 3), ("grid_type", C.c_int),
                ("inv_power", C.c_double), ("inv_power_mode", C.c_int), ("deriv_count", C.c_uint),
                ("data_offset", C.c_ulonglong)]


class _Props(C.Structure):
    _fields_ = [("name", C.c_char * 128), ("cc_major", C.c_int), ("cc_minor", C.c_int), ("sm_count", C.c_int),
                ("l2_bytes", C.c_int), ("total_mem_bytes", C.c_size_t)]


CLASS_DTYPE = np.dtype([("inside", np.int32), ("cell", np.int32, (3,))])

# name -> (restype, argtypes); must list every GFB_API symbol of include/gridforce_b200.h
_vp, _i, _ll, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_size_t
_pi, _pd = C.POINTER(C.c_int), C.POINTER(C.c_double)
SIGNATURES = {
    "gfb_version": (_i, []),
    "gfb_last_error": (C.c_char_p, []),
    "gfb_device_count": (_i, [_pi]),
    "gfb_device_open": (_i, [_i, C.POINTER(_vp)]),
    "gfb_device_close": (_i, [_vp]),
    "gfb_device_get_props": (_i, [_vp, C.POINTER(_Props)]),
    "gfb_device_synchronize": (_i, [_vp]),
    "gfb_grid_create": (_i, [_vp, _pi, _pd, _pd, _vp, _sz, _i, _i, C.POINTER(_vp)]),
    "gfb_grid_create_from_device": (_i, [_vp, _pi, _pd, _pd, _vp, _sz, _i, _i, C.POINTER(_vp)]),
    "gfb_grid_layout": (_i, [_vp]),
    "gfb_gridfile_read_header": (_i, [C.c_char_p, C.POINTER(GridFileHeader)]),
    "gfb_gridfile_read_values": (_i, [C.c_char_p, _vp, _sz]),
    "gfb_gridfile_write": (_i, [C.c_char_p, C.POINTER(GridFileHeader), _vp, _sz, _i]),
    "gfb_grid_generate": (_i, [_vp, _pi, _pd, _pd, _i, _i, _vp, _vp, _vp, _vp, C.c_double, _vp, _i, _i, C.POINTER(_vp)]),
    "gfb_grid_create_from_file": (_i, [_vp, C.c_char_p, _i, _i, C.POINTER(_vp), C.POINTER(GridFileHeader)]),
    "gfb_inv_power_transform": (_i, [_vp, _vp, _sz, C.c_double, _i]),
    "gfb_grid_destroy": (_i, [_vp]),
    "gfb_grid_release_cells": (_i, [_vp]),
    "gfb_grid_device_bytes": (_sz, [_vp]),
    "gfb_kernel_create": (_i, [_vp, _i, C.POINTER(_vp), _i, _vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    "gfb_kernel_destroy": (_i, [_vp]),
    "gfb_kernel_update_parameters": (_i, [_vp, _vp, _vp]),
    "gfb_kernel_set_energy_slots": (_i, [_vp, _vp, _i]),
    "gfb_kernel_eval_path": (_i, [_vp]),
    "gfb_kernel_set_launch_overlap": (_i, [_vp, _i]),
    "gfb_kernel_set_resident": (_i, [_vp, _i, _ll]),
    "gfb_kernel_resident_stop": (_i, [_vp]),
    "gfb_kernel_resident_launches": (_ll, [_vp]),
    "gfb_kernel_resident_timeline": (_i, [_vp, _vp]),
    "gfb_kernel_execute_host": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _i]),
    "gfb_kernel_execute_device": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _i, _ll, _vp, _vp, _vp]),
    "gfb_kernel_sort_atoms": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "gfb_kernel_classify_host": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "gfb_forces_fixed_to_f64": (_i, [_vp, _vp, _ll, _ll, _vp, _vp]),
    "gfb_peer_put": (_i, [_vp, _vp, C.POINTER(_vp), _i, _sz, _sz, _i, _vp]),
    "gfb_launch_count": (C.c_ulonglong, []),
    "gfb_bench_sector_gather": (_i, [_vp, _sz, _ll, _i, _pd]),
    "gfb_bench_host_copy": (_i, [_vp, _sz, _i, _pd]),
    "gfb_host_register": (_i, [_vp, _sz]),
    "gfb_host_unregister": (_i, [_vp]),
    "gfb_kernel_request_atom_energies": (_i, [_vp, _i]),
    "gfb_kernel_get_atom_energies": (_i, [_vp, _vp, _sz]),
    "gfb_graph_begin": (_i, [_vp, _vp]),
    "gfb_graph_end": (_i, [_vp, _vp, C.POINTER(_vp)]),
    "gfb_graph_launch": (_i, [_vp, _vp]),
    "gfb_graph_destroy": (_i, [_vp]),
    "gfb_comm_unique_id": (_i, [_vp]),
    "gfb_comm_create": (_i, [_vp, _i, _i, _vp, C.POINTER(_vp)]),
    "gfb_comm_destroy": (_i, [_vp]),
    "gfb_comm_all_gather": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "gfb_comm_gather_alloc": (_i, [_vp, _sz, _vp]),
    "gfb_comm_gather_attach": (_i, [_vp, _vp]),
    "gfb_kernel_execute_device_gather": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i, _ll, _vp, _vp, _sz, _vp]),
    "gfb_comm_gather_wait": (_i, [_vp, _vp, _vp]),
    "gfb_comm_gather_push": (_i, [_vp, _vp, _sz, _sz, _vp]),
    "gfb_comm_gather_status": (_i, [_vp]),
    "gfb_comm_gather": (_i, [_vp, _vp, _sz, _sz, _vp, _vp]),
    "gfb_comm_rendezvous": (_i, [_vp, _i, _vp]),
    "gfb_comm_rendezvous_release": (_i, [_vp]),
    "gfb_multi_create": (_i, [_i, _pi, C.POINTER(_vp)]),
    "gfb_multi_destroy": (_i, [_vp]),
    "gfb_multi_num_devices": (_i, [_vp]),
    "gfb_multi_add_grid": (_i, [_vp, _pi, _pd, _pd, _vp, _sz, _i, _i]),
    "gfb_multi_build": (_i, [_vp, _i, _vp, _vp, _vp]),
    "gfb_multi_execute_host": (_i, [_vp, _i, _vp, _vp, _vp, _i]),
    "gfb_multi_upload": (_i, [_vp, _i, _vp]),
    "gfb_multi_step": (_i, [_vp, _i]),
    "gfb_multi_download": (_i, [_vp, _i, _vp, _vp]),
}


def library_path():
    if os.environ.get("GFB_LIB_PATH"):      # A/B builds of the same library (tuning experiments)
        return os.environ["GFB_LIB_PATH"]
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libgridforce_b200.so")


def load_library():
    """dlopen the in-tree CUDA library. There is deliberately no fallback."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise GridForceB200Error(
                f"{path} is missing: build it with `make lib` (or __graft_entry__.build()); "
                "there is no CPU/PyTorch fallback for the GridForce path")
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def _check(rc):
    if rc != 0:
        raise GridForceB200Error(load_library().gfb_last_error().decode() or f"gridforce_b200 status {rc}")


def launch_count():
    return int(load_library().gfb_launch_count())


def _host_f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _ptr(a):
    """Raw pointer of a numpy array / int address / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def read_grid_file(path):
    """V3 OMGRID file -> (header dict, values [nx, ny, nz] float64). Host only."""
    lib = load_library()
    h = GridFileHeader()
    _check(lib.gfb_gridfile_read_header(os.fsencode(path), C.byref(h)))
    vals = np.empty(tuple(h.counts), dtype=np.float64)
    _check(lib.gfb_gridfile_read_values(os.fsencode(path), _ptr(vals), vals.size))
    return _header_dict(h), vals


def write_grid_file(path, counts, spacing, origin, values, grid_type=0, inv_power=0.0, inv_power_mode=0, with_trailer=False):
    """Writes the reference's V3 format: GridForce::saveToFile bytes, or GridData::saveToFile bytes with_trailer."""
    h = GridFileHeader()
    h.counts = (C.c_int * 3)(*[int(c) for c in counts])
    h.spacing = (C.c_double * 3)(*[float(c) for c in spacing])
    h.origin = (C.c_double * 3)(*[float(c) for c in origin])
    h.grid_type, h.inv_power, h.inv_power_mode = int(grid_type), float(inv_power), int(inv_power_mode)
    v = _host_f64(values).ravel()
    _check(load_library().gfb_gridfile_write(os.fsencode(path), C.byref(h), _ptr(v), v.size, 1 if with_trailer else 0))


def _header_dict(h):
    return {"counts": tuple(h.counts), "spacing": tuple(h.spacing), "origin": tuple(h.origin), "grid_type": h.grid_type,
            "inv_power": h.inv_power, "inv_power_mode": h.inv_power_mode, "deriv_count": h.deriv_count,
            "data_offset": h.data_offset}


class Device:
    """gfb_device: one GPU."""

    def __init__(self, ordinal=0):
        self._h = C.c_void_p()
        self.ordinal = ordinal
        _check(load_library().gfb_device_open(ordinal, C.byref(self._h)))

    @staticmethod
    def count():
        n = C.c_int(0)
        _check(load_library().gfb_device_count(C.byref(n)))
        return n.value

    def props(self):
        p = _Props()
        _check(load_library().gfb_device_get_props(self._h, C.byref(p)))
        return {"name": p.name.decode(), "cc": (p.cc_major, p.cc_minor), "sm_count": p.sm_count,
                "l2_bytes": p.l2_bytes, "total_mem_bytes": p.total_mem_bytes}

    def synchronize(self):
        _check(load_library().gfb_device_synchronize(self._h))

    def bench_sector_gather(self, nbytes, n_loads, reps=10):
        out = C.c_double(0.0)
        _check(load_library().gfb_bench_sector_gather(self._h, nbytes, n_loads, reps, C.byref(out)))
        return out.value

    def bench_host_copy(self, nbytes=64 << 20, reps=5):
        """Pinned-host <-> device copy bandwidth on this GPU's link: (h2d, d2h, both directions together) in GB/s."""
        out = (C.c_double * 3)()
        _check(load_library().gfb_bench_host_copy(self._h, nbytes, reps, out))
        return tuple(out)

    def inv_power_transform(self, values, inv_power, device_ptr=None, n=None):
        """GridForce::applyInvPowerTransformation on the GPU: G -> sign(G)|G|^(1/inv_power). Host array in -> new host
        array out; or, with device_ptr and n, in place on device doubles."""
        if device_ptr is not None:
            _check(load_library().gfb_inv_power_transform(self._h, _ptr(int(device_ptr)), n, float(inv_power), 1))
            return None
        out = np.array(values, dtype=np.float64, order="C", copy=True)
        _check(load_library().gfb_inv_power_transform(self._h, _ptr(out), out.size, float(inv_power), 0))
        return out

    def peer_put(self, d_src, peer_ptrs, dst_offset, nbytes, first_peer=0, stream=0):
        """Copy-engine put of nbytes from local device memory into every peer buffer at dst_offset (gfb_peer_put)."""
        arr = (C.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        _check(load_library().gfb_peer_put(self._h, _ptr(int(d_src)), arr, len(peer_ptrs), dst_offset, nbytes, first_peer,
                                           _ptr(stream or None)))

    def fixed_to_f64(self, d_fixed, stride, n, d_out, stream=0):
        _check(load_library().gfb_forces_fixed_to_f64(self._h, _ptr(d_fixed), stride, n, _ptr(d_out), _ptr(stream or None)))

    def close(self):
        if self._h:
            load_library().gfb_device_close(self._h)
            self._h = C.c_void_p()


class Grid:
    """gfb_grid: counts/spacing/origin + values (x-major, z fastest), repacked cell-major on the GPU."""

    def __init__(self, device, counts, spacing, origin, values, precision=PRECISION_MIXED, layout=None, device_ptr=None):
        if layout is None:      # GFB_LAYOUT env var: experiment knob for the benches (auto|cells|rows|pairs)
            layout = {v: k for k, v in LAYOUT_NAMES.items()}[os.environ.get("GFB_LAYOUT", "auto")]
        self.device = device
        self.counts = tuple(int(c) for c in counts)
        self.spacing = tuple(float(s) for s in spacing)
        self.origin = tuple(float(o) for o in origin)
        self.precision = precision
        self._h = C.c_void_p()
        cn = (C.c_int * 3)(*self.counts)
        sp = (C.c_double * 3)(*self.spacing)
        og = (C.c_double * 3)(*self.origin)
        lib = load_library()
        if device_ptr is not None:
            n = int(np.prod(self.counts))
            _check(lib.gfb_grid_create_from_device(device._h, cn, sp, og, C.c_void_p(device_ptr), n, precision, layout, C.byref(self._h)))
        else:
            v = _host_f64(values).ravel()
            _check(lib.gfb_grid_create(device._h, cn, sp, og, _ptr(v), v.size, precision, layout, C.byref(self._h)))
        self.layout = int(lib.gfb_grid_layout(self._h))

    @classmethod
    def from_file(cls, device, path, precision=PRECISION_MIXED, layout=LAYOUT_AUTO):
        """gfb_grid_create_from_file: V3 OMGRID file -> pinned staging -> HBM -> on-device repack."""
        self = cls.__new__(cls)
        self.device, self.precision, self._h = device, precision, C.c_void_p()
        h = GridFileHeader()
        lib = load_library()
        _check(lib.gfb_grid_create_from_file(device._h, os.fsencode(path), precision, layout, C.byref(self._h), C.byref(h)))
        self.header = _header_dict(h)
        self.counts, self.spacing, self.origin = self.header["counts"], self.header["spacing"], self.header["origin"]
        self.layout = int(lib.gfb_grid_layout(self._h))
        return self

    @classmethod
    def generate(cls, device, counts, spacing, origin, grid_type, pos, charges=None, sigmas=None, epsilons=None,
                 grid_cap=41840.0, precision=PRECISION_MIXED, layout=LAYOUT_AUTO, want_values=True, want_grid=True):
        """gfb_grid_generate: grid from receptor atoms ("charge" | "ljr" | "lja"). Returns (Grid or None, values or None)."""
        code = {"charge": 1, "ljr": 2, "lja": 3}.get(grid_type)
        if code is None:
            raise GridForceB200Error(f"GridForce: Invalid grid type '{grid_type}'. Must be 'charge', 'ljr', or 'lja'")
        pos = _host_f64(pos)
        q = _host_f64(charges) if charges is not None else None
        sg = _host_f64(sigmas) if sigmas is not None else None
        ep = _host_f64(epsilons) if epsilons is not None else None
        vals = np.empty(tuple(int(c) for c in counts), dtype=np.float64) if want_values else None
        handle = C.c_void_p()
        lib = load_library()
        _check(lib.gfb_grid_generate(device._h, (C.c_int * 3)(*[int(c) for c in counts]), (C.c_double * 3)(*spacing),
                                     (C.c_double * 3)(*origin), code, pos.shape[0], _ptr(pos), _ptr(q), _ptr(sg), _ptr(ep),
                                     float(grid_cap), _ptr(vals), precision, layout, C.byref(handle) if want_grid else None))
        grid = None
        if want_grid:
            grid = cls.__new__(cls)
            grid.device, grid.precision, grid._h = device, precision, handle
            grid.counts, grid.spacing, grid.origin = tuple(counts), tuple(spacing), tuple(origin)
            grid.layout = int(lib.gfb_grid_layout(handle))
        return grid, vals

    @property
    def device_bytes(self):
        return int(load_library().gfb_grid_device_bytes(self._h))

    def release_cells(self):
        """gfb_grid_release_cells: free the per-grid packed cells once only record kernels read this grid."""
        _check(load_library().gfb_grid_release_cells(self._h))

    def close(self):
        if self._h:
            load_library().gfb_grid_destroy(self._h)
            self._h = C.c_void_p()


class Kernel:
    """gfb_kernel: the state CalcGridForceKernel::initialize captures, for G grids acting on the same atoms."""

    def __init__(self, device, grids, scaling, particles=None, inv_power=None, oob_k=None):
        self.device = device
        self.grids = list(grids)
        g = len(self.grids)
        sc = _host_f64(scaling)
        if sc.ndim == 1:
            sc = sc.reshape(1, -1)
        if sc.shape[0] != g:
            raise GridForceB200Error(f"scaling has {sc.shape[0]} rows for {g} grids")
        self.n_grids = g
        self.n_atoms = sc.shape[1]
        self.n_slots = 1
        self._scaling = sc
        ok = _host_f64(oob_k if oob_k is not None else [10000.0] * g).ravel()  # GridForce.cpp:52 default
        ip = _host_f64(inv_power).ravel() if inv_power is not None else None
        pa = np.ascontiguousarray(particles, dtype=np.int32) if particles is not None else None
        handles = (C.c_void_p * g)(*[gr._h for gr in self.grids])
        self._h = C.c_void_p()
        _check(load_library().gfb_kernel_create(device._h, g, handles, self.n_atoms, _ptr(sc), _ptr(pa), _ptr(ip), _ptr(ok),
                                                C.byref(self._h)))

    def set_energy_slots(self, slots, n_slots):
        """Particle groups: slots[ia] in [0, n_slots); energies then come back as [R, n_slots]."""
        if slots is None:
            _check(load_library().gfb_kernel_set_energy_slots(self._h, None, 1))
            self.n_slots = 1
            return
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        if sl.size != self.n_atoms:
            raise GridForceB200Error(f"{sl.size} slots for {self.n_atoms} atoms")
        _check(load_library().gfb_kernel_set_energy_slots(self._h, _ptr(sl), int(n_slots)))
        self.n_slots = int(n_slots)

    def uses_lines_kernel(self):
        """True when launches of this state go to gf_eval_lines_kernel (see gfb_kernel_eval_path)."""
        return self.eval_path() == 1

    def eval_path(self):
        """gfb_kernel_eval_path: 1 gf_eval_lines_kernel, 2 gf_eval_lines_f64_kernel, 3 gf_eval_bspline_kernel, 4 gf_eval_bspline_f64_kernel, 5 / 6 the MIXED / DOUBLE record kernels with the tricubic arithmetic, 0 general."""
        return int(load_library().gfb_kernel_eval_path(self._h))

    def set_launch_overlap(self, enable=True):
        """PDL for execute_device launches (see gfb_kernel_set_launch_overlap: positions must not come from the kernel
        launched just before)."""
        _check(load_library().gfb_kernel_set_launch_overlap(self._h, 1 if enable else 0))

    def set_resident(self, enable=True, idle_us=0):
        """gfb_kernel_set_resident: one-ligand-per-step host calls are served by a block that stays on the GPU."""
        _check(load_library().gfb_kernel_set_resident(self._h, 1 if enable else 0, int(idle_us)))

    def resident_stop(self):
        _check(load_library().gfb_kernel_resident_stop(self._h))

    def resident_launches(self):
        return int(load_library().gfb_kernel_resident_launches(self._h))

    def resident_timeline(self):
        """Microseconds of the last resident step on the GPU: evaluation, result packets issued."""
        out = np.zeros(2)
        _check(load_library().gfb_kernel_resident_timeline(self._h, _ptr(out)))
        return out

    def update_parameters(self, scaling=None, inv_power=None):
        sc = _host_f64(scaling).reshape(self.n_grids, self.n_atoms) if scaling is not None else None
        ip = _host_f64(inv_power).ravel() if inv_power is not None else None
        _check(load_library().gfb_kernel_update_parameters(self._h, _ptr(sc), _ptr(ip)))

    def execute_host(self, pos, forces=None, force_mode=FORCE_F64_STORE, want_forces=True, want_grid_energies=False,
                     energies_out=None):
        """pos: [R, P, 3] (or [P, 3]) float64. Returns (energies[R], forces[R,P,3] or None, grid_energies or None).
        force_mode FORCE_F32_STORE returns/needs float32 forces; want_forces=False is an energy-only evaluation."""
        pos = np.asarray(pos)
        if pos.dtype != np.float64 or not pos.flags.c_contiguous:
            pos = _host_f64(pos)
        if pos.ndim == 2:
            pos = pos.reshape(1, *pos.shape)
        r, p, _ = pos.shape
        ne = r * self.n_slots
        en = energies_out if energies_out is not None else np.empty(ne, dtype=np.float64)
        ge = np.empty((ne, self.n_grids), dtype=np.float64) if want_grid_energies else None
        fdtype = np.float32 if force_mode == FORCE_F32_STORE else np.float64
        if forces is None and want_forces:
            forces = np.zeros((r, p, 3), dtype=fdtype)
        if forces is not None and (forces.dtype != fdtype or not forces.flags.c_contiguous):
            raise GridForceB200Error(f"forces must be a C-contiguous {np.dtype(fdtype).name} array for force_mode {force_mode}")
        _check(load_library().gfb_kernel_execute_host(self._h, r, p, _ptr(pos), _ptr(en), _ptr(ge), _ptr(forces), force_mode))
        return en, forces, ge

    def request_atom_energies(self, enable=True):
        """GridForce::getParticleAtomEnergies: keep each evaluated atom's energy of the following execute_host calls."""
        _check(load_library().gfb_kernel_request_atom_energies(self._h, 1 if enable else 0))

    def atom_energies(self, n_replicas):
        """[R, A] per-atom energies (summed over the kernel's grids) of the last execute_host call."""
        out = np.empty((n_replicas, self.n_atoms), dtype=np.float64)
        _check(load_library().gfb_kernel_get_atom_energies(self._h, _ptr(out), out.size))
        return out

    def execute_device_gather(self, comm, gather_offset, n_replicas, n_particles, d_pos, d_energies, d_forces=None,
                              force_mode=FORCE_FIXED_ADD, force_stride=0, stream=0, d_energies_clear=None):
        """execute_device whose last block also stores the energies into every rank's gathered array (fused gather)."""
        _check(load_library().gfb_kernel_execute_device_gather(self._h, n_replicas, n_particles, _ptr(d_pos), _ptr(d_energies),
                                                               _ptr(d_forces), force_mode, force_stride, _ptr(d_energies_clear),
                                                               comm._h, gather_offset, _ptr(stream or None)))

    def execute_device(self, n_replicas, n_particles, d_pos, d_energies=None, d_grid_energies=None, d_forces=None,
                       force_mode=FORCE_FIXED_ADD, force_stride=0, d_order=None, stream=0, d_energies_clear=None):
        """All d_* are integer device addresses (e.g. torch_tensor.data_ptr()); stream is a cudaStream_t value."""
        _check(load_library().gfb_kernel_execute_device(self._h, n_replicas, n_particles, _ptr(d_pos), _ptr(d_energies),
                                                        _ptr(d_grid_energies), _ptr(d_forces), force_mode, force_stride,
                                                        _ptr(d_order), _ptr(d_energies_clear), _ptr(stream or None)))

    def sort_atoms(self, n_replicas, n_particles, d_pos, d_order, stream=0):
        _check(load_library().gfb_kernel_sort_atoms(self._h, n_replicas, n_particles, _ptr(d_pos), _ptr(d_order),
                                                    _ptr(stream or None)))

    def classify_host(self, pos, grid_index=0):
        pos = _host_f64(pos)
        if pos.ndim == 2:
            pos = pos.reshape(1, *pos.shape)
        r, p, _ = pos.shape
        out = np.empty(r * self.n_atoms, dtype=CLASS_DTYPE)
        _check(load_library().gfb_kernel_classify_host(self._h, grid_index, r, p, _ptr(pos), _ptr(out)))
        return out

    def close(self):
        if self._h:
            load_library().gfb_kernel_destroy(self._h)
            self._h = C.c_void_p()


class DeviceArrayView:
    """Zero-copy description of device memory owned by the library (e.g. the gathered energies of Comm.gather_wait) in
    the CUDA array interface, so that torch.as_tensor(view, device=...) / cupy.asarray(view) can read it in place."""

    def __init__(self, ptr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def host_register(array):
    """Page-locks a numpy array's memory (gfb_host_register) so that execute_host DMAs from/into it directly."""
    _check(load_library().gfb_host_register(_ptr(array), array.nbytes))


def host_unregister(array):
    _check(load_library().gfb_host_unregister(_ptr(array)))


class Graph:
    """gfb_graph: launches captured on a stream between Graph.begin() and Graph.end(), replayed with launch()."""

    def __init__(self, device, handle):
        self.device, self._h = device, handle

    @staticmethod
    def begin(device, stream=0):
        _check(load_library().gfb_graph_begin(device._h, _ptr(stream or None)))

    @classmethod
    def end(cls, device, stream=0):
        h = C.c_void_p()
        _check(load_library().gfb_graph_end(device._h, _ptr(stream or None), C.byref(h)))
        return cls(device, h)

    def launch(self, stream=0):
        _check(load_library().gfb_graph_launch(self._h, _ptr(stream or None)))

    def close(self):
        if self._h:
            load_library().gfb_graph_destroy(self._h)
            self._h = C.c_void_p()


class Comm:
    """gfb_comm: this process's end of a replica-sharded run with one process per GPU. `unique_id` comes from
    Comm.unique_id() on rank 0 and reaches the other ranks through the launcher; None = no NCCL communicator."""

    def __init__(self, device, world_size, rank, unique_id=None):
        self.device, self.world_size, self.rank = device, world_size, rank
        self._h = C.c_void_p()
        uid = (C.c_ubyte * COMM_ID_BYTES).from_buffer_copy(bytes(unique_id)) if unique_id is not None else None
        _check(load_library().gfb_comm_create(device._h, world_size, rank, uid, C.byref(self._h)))

    @staticmethod
    def unique_id():
        buf = (C.c_ubyte * COMM_ID_BYTES)()
        _check(load_library().gfb_comm_unique_id(buf))
        return bytes(buf)

    def all_gather(self, d_send, d_recv, count, stream=0):
        _check(load_library().gfb_comm_all_gather(self._h, _ptr(d_send), _ptr(d_recv), count, _ptr(stream or None)))

    def gather_alloc(self, count_total):
        """Allocates this rank's gathered array; returns the 64-byte IPC handle the other ranks attach with."""
        buf = (C.c_ubyte * IPC_HANDLE_BYTES)()
        _check(load_library().gfb_comm_gather_alloc(self._h, count_total, buf))
        return bytes(buf)

    def gather_attach(self, handles):
        """handles: every rank's gather_alloc() result, in rank order."""
        blob = b"".join(bytes(h) for h in handles)
        if len(blob) != self.world_size * IPC_HANDLE_BYTES:
            raise GridForceB200Error("gather_attach needs one 64-byte handle per rank")
        _check(load_library().gfb_comm_gather_attach(self._h, (C.c_ubyte * len(blob)).from_buffer_copy(blob)))

    def gather_push(self, d_energies, count, gather_offset, stream=0):
        """Stand-alone producer of the peer-store gather (same protocol as the fused tail, as its own small kernel)."""
        _check(load_library().gfb_comm_gather_push(self._h, _ptr(d_energies), count, gather_offset, _ptr(stream or None)))

    def gather(self, d_energies, count, gather_offset, d_out, stream=0):
        """The whole gather as one flag-in-data kernel: this rank's `count` values out to every rank, all count_total in to d_out."""
        _check(load_library().gfb_comm_gather(self._h, _ptr(d_energies), count, gather_offset, _ptr(d_out), _ptr(stream or None)))

    def gather_wait(self, d_out, stream=0):
        """Enqueues the wait for the oldest unconsumed fused gather; the complete array is copied to device address d_out."""
        _check(load_library().gfb_comm_gather_wait(self._h, _ptr(d_out), _ptr(stream or None)))

    def gather_status(self):
        _check(load_library().gfb_comm_gather_status(self._h))

    def rendezvous(self, stream=0, hold=False):
        """Device-side rendezvous of all ranks on `stream`; hold=True keeps the kernel waiting until rendezvous_release()."""
        _check(load_library().gfb_comm_rendezvous(self._h, 1 if hold else 0, _ptr(stream or None)))

    def rendezvous_release(self):
        _check(load_library().gfb_comm_rendezvous_release(self._h))

    def close(self):
        if self._h:
            load_library().gfb_comm_destroy(self._h)
            self._h = C.c_void_p()


class Multi:
    """gfb_multi: a replica-sharded run over several GPUs driven by this one process."""

    def __init__(self, ordinals):
        ords = (C.c_int * len(ordinals))(*[int(o) for o in ordinals])
        self._h = C.c_void_p()
        _check(load_library().gfb_multi_create(len(ordinals), ords, C.byref(self._h)))
        self.n_devices = len(ordinals)
        self.n_atoms = 0
        self.n_replicas = 0

    def add_grid(self, counts, spacing, origin, values, precision=PRECISION_MIXED, layout=LAYOUT_AUTO):
        v = _host_f64(values).ravel()
        rc = load_library().gfb_multi_add_grid(self._h, (C.c_int * 3)(*[int(c) for c in counts]), (C.c_double * 3)(*spacing),
                                               (C.c_double * 3)(*origin), _ptr(v), v.size, precision, layout)
        if rc < 0:
            _check(rc)
        return rc

    def build(self, scaling, inv_power=None, oob_k=None):
        sc = _host_f64(scaling)
        if sc.ndim == 1:
            sc = sc.reshape(1, -1)
        g = sc.shape[0]
        ok = _host_f64(oob_k if oob_k is not None else [10000.0] * g).ravel()
        ip = _host_f64(inv_power).ravel() if inv_power is not None else None
        _check(load_library().gfb_multi_build(self._h, sc.shape[1], _ptr(sc), _ptr(ip), _ptr(ok)))
        self.n_atoms = sc.shape[1]

    def execute_host(self, pos, want_forces=True, force_mode=FORCE_F64_STORE, forces=None, energies_out=None):
        pos = _host_f64(pos)
        r = pos.shape[0]
        en = energies_out if energies_out is not None else np.empty(r, dtype=np.float64)
        fdtype = np.float32 if force_mode == FORCE_F32_STORE else np.float64
        if forces is None and want_forces:
            forces = np.zeros(pos.shape, dtype=fdtype)
        _check(load_library().gfb_multi_execute_host(self._h, r, _ptr(pos), _ptr(en), _ptr(forces), force_mode))
        return en, forces

    def upload(self, pos):
        pos = _host_f64(pos)
        _check(load_library().gfb_multi_upload(self._h, pos.shape[0], _ptr(pos)))
        self.n_replicas = pos.shape[0]

    def step(self, gather=0):
        """gather: 0 none, 1 ncclAllGather, 2 peer-store gather fused into the evaluation kernel, 3 peer-store push kernel, 4 the
        one-kernel flag-in-data gather (gfb_comm_gather's kernel)."""
        _check(load_library().gfb_multi_step(self._h, gather))

    def download(self, from_device=0, want_forces=True):
        en = np.empty(self.n_replicas, dtype=np.float64)
        f = np.empty((self.n_replicas, self.n_atoms, 3), dtype=np.float64) if want_forces else None
        _check(load_library().gfb_multi_download(self._h, from_device, _ptr(en), _ptr(f)))
        return en, f

    def close(self):
        if self._h:
            load_library().gfb_multi_destroy(self._h)
            self._h = C.c_void_p()
