"""Python face of the plugin path, shaped like the reference's SWIG module `gridforceplugin` plus the sliver of
`openmm` a GridForce script touches (System, Context, Platform.getPlatformByName, State getters).

    import openmmgridforce_b200.gridforceplugin as gfp
    force = gfp.GridForce(); force.addGridCounts(nx, ny, nz); force.addGridSpacing(dx, dy, dz)
    for v in values: force.addGridValue(v)            # or force.setGridValues(values)
    for s in factors: force.addScalingFactor(s)
    system = gfp.System(); [system.addParticle(m) for m in masses]; system.addForce(force)
    context = gfp.Context(system, gfp.Platform.getPlatformByName("B200"))
    context.setPositions(xyz_nm)
    state = context.getState(getEnergy=True, getForces=True)
    state.getPotentialEnergy(), state.getForces()

(cf. python/tests/test_grid_force.py:40-64, 147-159 of the reference.) Everything is executed by the C++ plugin
(libOpenMMGridForceB200.so: registerPlatforms/registerKernelFactories -> B200Platform -> GridForceImpl ->
B200CalcGridForceKernel -> libgridforce_b200.so); this module only marshals arguments through ctypes because neither SWIG
nor OpenMM exists in the build image. C++ OpenMMException surfaces as RuntimeError, as in gridforceplugin.i:49-59.
"""
import ctypes as C
import os

import numpy as np

InvPowerMode_NONE, InvPowerMode_RUNTIME, InvPowerMode_STORED = 0, 1, 2

_LIB = None


def plugin_library_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libOpenMMGridForceB200.so")


def _lib():
    global _LIB
    if _LIB is None:
        path = plugin_library_path()
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `make plugin`")
        lib = C.CDLL(path)
        lib.b200_plugin_last_error.restype = C.c_char_p
        lib.b200_plugin_create.restype = C.c_void_p
        lib.b200_plugin_create.argtypes = [C.c_int]
        lib.b200_plugin_destroy.argtypes = [C.c_void_p]
        lib.b200_plugin_set_property.argtypes = [C.c_char_p, C.c_char_p]
        lib.b200_plugin_add_grid.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p,
                                             C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
        lib.b200_plugin_apply_inv_power.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_double, C.c_void_p]
        lib.b200_plugin_add_nonbonded.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.b200_plugin_set_auto.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_double, C.c_void_p, C.c_int,
                                             C.c_void_p, C.c_int]
        lib.b200_plugin_get_force_data.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p]
        lib.b200_plugin_add_particle_group.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.b200_plugin_finalize.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_int]
        lib.b200_plugin_get_property.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
        lib.b200_plugin_get_default_property.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        lib.b200_plugin_set_particles.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        lib.b200_plugin_atom_energies.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        lib.b200_plugin_batch_evaluate_buffers.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                                           C.c_void_p, C.c_int]
        lib.b200_plugin_pin_buffer.argtypes = [C.c_void_p, C.c_longlong, C.c_int]
        lib.b200_plugin_execute.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.b200_plugin_time_execute.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.b200_plugin_update_scaling.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        lib.b200_plugin_group_energies.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        lib.b200_plugin_batch_evaluate.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _LIB = lib
    return _LIB


def _check(rc):
    if rc != 0:
        raise RuntimeError(_lib().b200_plugin_last_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Platform:
    """openmm.Platform look-alike: name lookup after the plugin's registerPlatforms()/registerKernelFactories()."""

    def __init__(self, name):
        self._name = name

    @staticmethod
    def getPlatformByName(name):
        _check(_lib().b200_plugin_register())
        if name != "B200":
            raise RuntimeError(f'There is no registered Platform called "{name}"')
        return Platform(name)

    def getName(self):
        return self._name

    def setPropertyDefaultValue(self, prop, value):
        """Platform::setPropertyDefaultValue (base-class method): "DeviceIndex", "Precision"; unknown names raise."""
        _check(_lib().b200_plugin_set_property(prop.encode(), str(value).encode()))

    def getPropertyDefaultValue(self, prop):
        buf = C.create_string_buffer(256)
        _check(_lib().b200_plugin_get_default_property(prop.encode(), buf, 256))
        return buf.value.decode()

    def getPropertyValue(self, context, prop):
        """Platform::getPropertyValue(context, name): the value this Context uses (its own property, else the default)."""
        buf = C.create_string_buffer(256)
        _check(_lib().b200_plugin_get_property(context._h, prop.encode(), buf, 256))
        return buf.value.decode()


class GridForce:
    """Same method names/meaning as the reference's GridForce for the evaluation path (openmmapi/include/GridForce.h)."""

    def __init__(self):
        self._counts, self._spacing, self._vals, self._scaling = [], [], [], []
        self._origin = (0.0, 0.0, 0.0)
        self._inv_power, self._inv_power_mode = 0.0, InvPowerMode_NONE
        self._oob_k, self._interp, self._group = 10000.0, 0, 0
        self._ligand_atoms, self._groups, self._particles = [], [], []
        self._context, self._index = None, None
        self._auto_scaling, self._scaling_property = False, ""
        self._auto_generate, self._grid_type, self._grid_cap = False, "", 41840.0
        self._receptor_atoms, self._receptor_positions = [], np.zeros((0, 3))

    def addGridCounts(self, nx, ny, nz):
        self._counts += [int(nx), int(ny), int(nz)]

    def addGridSpacing(self, dx, dy, dz):
        self._spacing += [float(dx), float(dy), float(dz)]

    def addGridValue(self, val):
        self._vals.append(float(val))

    def setGridValues(self, vals):
        self._vals = np.ascontiguousarray(vals, dtype=np.float64).ravel()

    def loadFromFile(self, filename):
        """V3 OMGRID file (reference GridForce::loadFromFile): geometry, origin, values and inv-power state."""
        from . import capi
        try:
            h, vals = capi.read_grid_file(filename)
        except capi.GridForceB200Error as e:
            raise RuntimeError(str(e))
        self._counts, self._spacing, self._origin = list(h["counts"]), list(h["spacing"]), tuple(h["origin"])
        self._vals = vals.ravel()
        self._inv_power, self._inv_power_mode = h["inv_power"], h["inv_power_mode"]

    def saveToFile(self, filename):
        from . import capi
        if len(self._counts) != 3 or len(self._spacing) != 3:
            raise RuntimeError("GridForce: Grid dimensions must be set before saving")
        capi.write_grid_file(filename, self._counts, self._spacing, self._origin, np.asarray(self._vals, dtype=np.float64),
                             inv_power=self._inv_power, inv_power_mode=self._inv_power_mode)

    def addScalingFactor(self, val):
        self._scaling.append(float(val))

    def setScalingFactors(self, vals):
        self._scaling = [float(v) for v in vals]

    def setGridOrigin(self, x, y, z):
        self._origin = (float(x), float(y), float(z))

    def getGridOrigin(self):
        return self._origin

    def setInvPowerMode(self, mode, inv_power):
        if mode != InvPowerMode_NONE and inv_power == 0.0:
            raise RuntimeError("GridForce: inv_power must be non-zero when mode != NONE")
        if mode == InvPowerMode_NONE and inv_power != 0.0:
            raise RuntimeError("GridForce: inv_power must be 0 when mode == NONE")
        if len(self._vals):      # conflicting mode changes once a grid is loaded (reference GridForce.cpp:199-211)
            if self._inv_power_mode == InvPowerMode_STORED and mode == InvPowerMode_RUNTIME:
                raise RuntimeError("GridForce: Cannot set RUNTIME mode on grid that already has STORED transformation. "
                                   "This would apply transformation twice!")
            if self._inv_power_mode == InvPowerMode_RUNTIME and mode == InvPowerMode_STORED:
                raise RuntimeError("GridForce: Cannot set STORED mode on untransformed grid loaded with RUNTIME mode. "
                                   "Call applyInvPowerTransformation() first.")
        self._inv_power_mode, self._inv_power = mode, float(inv_power)

    def getInvPower(self):
        return self._inv_power

    def getInvPowerMode(self):
        return self._inv_power_mode

    def getGridValues(self):
        if self._context is not None and self._auto_generate and len(self._vals) == 0:
            self._vals = self._pull(0)          # the kernel copies a generated grid back into the force
        return np.asarray(self._vals, dtype=np.float64)

    def applyInvPowerTransformation(self):
        """RUNTIME mode: G -> sign(G)|G|^(1/n) once (on the GPU), then the mode is STORED (gridforceplugin.i:181)."""
        vals = np.array(self._vals, dtype=np.float64, order="C", copy=True).ravel()
        mode = C.c_int(-1)
        _check(_lib().b200_plugin_apply_inv_power(_p(vals), vals.size, int(self._inv_power_mode), float(self._inv_power),
                                                  C.byref(mode)))
        self._vals, self._inv_power_mode = vals, mode.value

    def setOutOfBoundsRestraint(self, k):
        self._oob_k = float(k)

    def getOutOfBoundsRestraint(self):
        return self._oob_k

    def setInterpolationMethod(self, method):
        if method < 0 or method > 3:
            raise RuntimeError("GridForce: Invalid interpolation method. Must be 0 (trilinear), 1 (cubic B-spline), 2 (tricubic), or 3 (quintic Hermite)")
        self._interp = int(method)

    # ---- inputs derived from the System's NonbondedForce at Context creation (reference GridForce.h:171-198, 335-342,
    #      523-573); the grid is generated on the GPU ------------------------------------------------------------------
    def setAutoCalculateScalingFactors(self, enable):
        self._auto_scaling = bool(enable)

    def setScalingProperty(self, prop):
        self._scaling_property = str(prop)

    def setAutoGenerateGrid(self, enable):
        self._auto_generate = bool(enable)

    def setGridType(self, grid_type):
        self._grid_type = str(grid_type)

    def setGridCap(self, u_max):
        self._grid_cap = float(u_max)

    def setReceptorAtoms(self, atoms):
        self._receptor_atoms = [int(a) for a in atoms]

    def setReceptorPositions(self, positions):
        self._receptor_positions = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)

    def setReceptorPositionsFromLists(self, x, y, z):
        if not (len(x) == len(y) == len(z)):
            raise RuntimeError("GridForce: x, y, z arrays must have the same size")
        self.setReceptorPositions(np.stack([x, y, z], axis=1))

    def _pull(self, which):
        """Grid values / scaling factors as the kernel left them in the C++ force (generated or auto-calculated)."""
        n = C.c_longlong(0)
        _check(_lib().b200_plugin_get_force_data(self._context._h, self._index, which, None, 0, C.byref(n)))
        out = np.zeros(n.value)
        _check(_lib().b200_plugin_get_force_data(self._context._h, self._index, which, _p(out), out.size, C.byref(n)))
        return out

    def getScalingFactors(self):
        return self._pull(1) if self._context is not None else np.asarray(self._scaling, dtype=np.float64)

    def setLigandAtoms(self, atoms):
        self._ligand_atoms = [int(a) for a in atoms]

    def setParticles(self, particles):
        """Only these particles feel the grid (GridForce.h:433-440); empty = all."""
        self._particles = [int(p) for p in particles]

    def getParticles(self):
        return list(self._particles)

    def setForceGroup(self, group):
        self._group = int(group)

    def getForceGroup(self):
        return self._group

    def addParticleGroup(self, name, particle_indices, scaling_factors=()):
        self._groups.append((name, [int(i) for i in particle_indices], [float(s) for s in scaling_factors]))
        return len(self._groups) - 1

    def getNumParticleGroups(self):
        return len(self._groups)

    def getParticleGroupEnergies(self, context):
        out = np.zeros(max(1, len(self._groups)))
        n = C.c_int(0)
        _check(_lib().b200_plugin_group_energies(context._h, self._index, _p(out), out.size, C.byref(n)))
        return list(out[:n.value])

    def getParticleAtomEnergies(self, context):
        """Per-atom energies of the last evaluation, in the order particles were added to the groups (empty without groups)."""
        cap = max(1, sum(len(g[1]) for g in self._groups))
        out = np.zeros(cap)
        n = C.c_int(0)
        _check(_lib().b200_plugin_atom_energies(context._h, self._index, _p(out), out.size, C.byref(n)))
        return out[:n.value].copy()

    def updateParametersInContext(self, context):
        sc = np.ascontiguousarray(self._scaling, dtype=np.float64)
        _check(_lib().b200_plugin_update_scaling(context._h, self._index, _p(sc), sc.size))


class NonbondedForce:
    """Parameter container (charge e, sigma nm, epsilon kJ/mol per particle): what the auto-derived GridForce inputs read."""

    def __init__(self):
        self._params = []

    def addParticle(self, charge, sigma, epsilon):
        self._params.append((float(charge), float(sigma), float(epsilon)))
        return len(self._params) - 1

    def getNumParticles(self):
        return len(self._params)


class System:
    def __init__(self):
        self._masses, self._forces = [], []

    def addParticle(self, mass):
        self._masses.append(float(mass))
        return len(self._masses) - 1

    def getNumParticles(self):
        return len(self._masses)

    def addForce(self, force):
        self._forces.append(force)
        return len(self._forces) - 1

    def getNumForces(self):
        return len(self._forces)


class State:
    def __init__(self, energy, forces):
        self._e, self._f = energy, forces

    def getPotentialEnergy(self):
        return self._e

    def getForces(self, asNumpy=True):
        return self._f


class Context:
    """Context(system, platform[, properties]): creating it runs GridForceImpl::initialize -> kernel initialize() for every
    GridForce. `properties` are the platform-specific properties of OpenMM's Context constructor ({"Precision": "double"});
    they apply to this Context only and win over the platform defaults."""

    def __init__(self, system, platform, properties=None):
        lib = _lib()
        self._n = system.getNumParticles()
        self._h = C.c_void_p(lib.b200_plugin_create(self._n))
        self._pos = None
        grid_forces = [f for f in system._forces if isinstance(f, GridForce)]
        for nb in (f for f in system._forces if isinstance(f, NonbondedForce)):
            prm = np.ascontiguousarray(nb._params, dtype=np.float64).reshape(-1, 3)
            q, sg, ep = (np.ascontiguousarray(prm[:, k]) for k in range(3))
            _check(lib.b200_plugin_add_nonbonded(self._h, _p(q), _p(sg), _p(ep), prm.shape[0]))
        for index, f in enumerate(grid_forces):
            counts = np.asarray(f._counts, dtype=np.int32)
            spacing = np.asarray(f._spacing, dtype=np.float64)
            if counts.size != 3 or spacing.size != 3:
                raise RuntimeError("GridForce: grid counts and spacing must each be given exactly once")
            vals = np.ascontiguousarray(f._vals, dtype=np.float64)
            sc = np.ascontiguousarray(f._scaling, dtype=np.float64)
            la = np.ascontiguousarray(f._ligand_atoms, dtype=np.int32) if f._ligand_atoms else None
            og = np.asarray(f._origin, dtype=np.float64)
            _check(lib.b200_plugin_add_grid(self._h, _p(counts), _p(spacing), _p(og), _p(vals), vals.size, _p(sc), sc.size, _p(la),
                                            0 if la is None else la.size, f._inv_power, f._oob_k, f._interp, f._group))
            if f._auto_scaling or f._auto_generate:
                ra = np.ascontiguousarray(f._receptor_atoms, dtype=np.int32) if f._receptor_atoms else None
                rp = np.ascontiguousarray(f._receptor_positions, dtype=np.float64)
                _check(lib.b200_plugin_set_auto(self._h, index, int(f._auto_scaling), f._scaling_property.encode(),
                                                int(f._auto_generate), f._grid_type.encode(), f._grid_cap, _p(ra),
                                                0 if ra is None else ra.size, _p(rp), rp.shape[0]))
            for name, idx, scl in f._groups:
                ia = np.ascontiguousarray(idx, dtype=np.int32)
                sa = np.ascontiguousarray(scl, dtype=np.float64) if scl else None
                _check(lib.b200_plugin_add_particle_group(self._h, index, name.encode(), _p(ia), _p(sa), ia.size))
            if f._particles:
                pa = np.ascontiguousarray(f._particles, dtype=np.int32)
                _check(lib.b200_plugin_set_particles(self._h, index, _p(pa), pa.size))
            f._context, f._index = self, index
        props = dict(properties or {})
        names = (C.c_char_p * max(1, len(props)))(*[k.encode() for k in props])
        values = (C.c_char_p * max(1, len(props)))(*[str(v).encode() for v in props.values()])
        _check(lib.b200_plugin_finalize(self._h, platform.getName().encode(), names, values, len(props)))

    def setPositions(self, positions):
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        if pos.shape != (self._n, 3):
            raise RuntimeError("Called setPositions() on a Context with the wrong number of positions")
        self._pos = pos

    def getState(self, getEnergy=False, getForces=False, groups=-1):
        if self._pos is None:
            raise RuntimeError("Particle positions have not been set")
        e = C.c_double(0.0)
        f = np.zeros((self._n, 3))
        _check(_lib().b200_plugin_execute(self._h, _p(self._pos), C.c_int(groups), C.byref(e), _p(f)))
        return State(e.value, f)

    def timeEvaluations(self, reps, groups=-1):
        """`reps` energy+force evaluations issued back to back from C++ (what an integrator's step loop does):
        returns (seconds per evaluation, last energy)."""
        if self._pos is None:
            raise RuntimeError("Particle positions have not been set")
        secs, e = C.c_double(0.0), C.c_double(0.0)
        _check(_lib().b200_plugin_time_execute(self._h, _p(self._pos), C.c_int(groups), int(reps), C.byref(secs), C.byref(e)))
        return secs.value, e.value

    def evaluateBatch(self, positions, precision="mixed", want_forces=True):
        """GridForceBatch over this System's GridForces: positions [R, A, 3] -> (energies [R], forces [R, A, 3] | None)."""
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        r = pos.shape[0]
        en = np.zeros(r)
        f = np.zeros_like(pos) if want_forces else None
        _check(_lib().b200_plugin_batch_evaluate(self._h, precision.encode(), _p(pos), r, _p(en), _p(f)))
        return en, f

    def evaluateBatchBuffers(self, positions, energies, forces=None, precision="mixed", devices=None):
        """GridForceBatch's pointer overloads on caller-owned numpy buffers (no copies): positions [R, A, 3] float64,
        energies [R] float64, forces [R, A, 3] float64 or float32 (None = energy only); devices: GPU ordinals."""
        r = positions.shape[0]
        dv = np.ascontiguousarray(devices, dtype=np.int32) if devices is not None else None
        f32 = forces is not None and forces.dtype == np.float32
        _check(_lib().b200_plugin_batch_evaluate_buffers(self._h, precision.encode(), _p(dv), 0 if dv is None else dv.size, _p(positions),
                                                         r, _p(energies), _p(forces), 1 if f32 else 0))

    def __del__(self):
        try:
            if self._h:
                _lib().b200_plugin_destroy(self._h)
                self._h = None
        except Exception:
            pass


class GridForceBatch:
    """Python face of GridForcePlugin::GridForceBatch (plugin/platform/GridForceBatch.h; SWIG: python/gridforceplugin_b200.i):
    R replicas x A atoms x G GridForces per call, one launch per GPU.

        batch = GridForceBatch()                      # device 0, mixed precision; GridForceBatch([0, 1, 2, 3]) shards replicas
        for f in (ele, ljr, lja): batch.addForce(f)
        energies = batch.evaluate(pos, R)                            # energy-only evaluation
        batch.evaluateWithForces(pos, R, energies, forces)           # caller-owned numpy buffers, no copies
        batch.evaluateWithForcesF32(pos, R, energies, forces_f32)
    """

    def __init__(self, devices=0, precision="mixed"):
        self._devices = [int(devices)] if np.isscalar(devices) else [int(d) for d in devices]
        self._precision = precision
        self._h = None
        self._forces = []

    def addForce(self, force):
        if self._h is not None:
            raise RuntimeError("GridForceBatch: add every force before the first evaluation")
        self._forces.append(force)
        return len(self._forces) - 1

    def getNumForces(self):
        return len(self._forces)

    def getNumAtoms(self):
        return len(self._forces[0]._scaling) if self._forces else 0

    def getNumDevices(self):
        return len(self._devices)

    def _handle(self):
        if self._h is None:
            lib = _lib()
            _check(lib.b200_plugin_register())
            self._h = C.c_void_p(lib.b200_plugin_create(self.getNumAtoms()))
            for f in self._forces:
                counts = np.asarray(f._counts, dtype=np.int32)
                spacing = np.asarray(f._spacing, dtype=np.float64)
                vals = np.ascontiguousarray(f._vals, dtype=np.float64)
                sc = np.ascontiguousarray(f._scaling, dtype=np.float64)
                og = np.asarray(f._origin, dtype=np.float64)
                _check(lib.b200_plugin_add_grid(self._h, _p(counts), _p(spacing), _p(og), _p(vals), vals.size, _p(sc), sc.size, None, 0,
                                                f._inv_power, f._oob_k, f._interp, f._group))
        return self._h

    def _run(self, positions, n_replicas, energies, forces):
        pos = np.asarray(positions)
        if pos.dtype != np.float64 or not pos.flags.c_contiguous or pos.size != n_replicas * self.getNumAtoms() * 3:
            raise RuntimeError("GridForceBatch: positions must hold numReplicas * numAtoms * 3 C-contiguous float64 values")
        dv = np.ascontiguousarray(self._devices, dtype=np.int32)
        f32 = forces is not None and forces.dtype == np.float32
        _check(_lib().b200_plugin_batch_evaluate_buffers(self._handle(), self._precision.encode(), _p(dv), dv.size, _p(pos), n_replicas,
                                                         _p(energies), _p(forces), 1 if f32 else 0))

    def evaluate(self, positions, n_replicas, energies=None):
        en = energies if energies is not None else np.empty(n_replicas)
        self._run(positions, n_replicas, en, None)
        return en

    def evaluateWithForces(self, positions, n_replicas, energies, forces):
        if forces.dtype != np.float64:
            raise RuntimeError("GridForceBatch.evaluateWithForces: forces must be float64 (evaluateWithForcesF32 takes float32)")
        self._run(positions, n_replicas, energies, forces)

    def evaluateWithForcesF32(self, positions, n_replicas, energies, forces):
        if forces.dtype != np.float32:
            raise RuntimeError("GridForceBatch.evaluateWithForcesF32: forces must be float32")
        self._run(positions, n_replicas, energies, forces)

    @staticmethod
    def pinBuffer(array):
        pin_buffer(array, True)

    @staticmethod
    def unpinBuffer(array):
        pin_buffer(array, False)

    def close(self):
        if self._h is not None:
            _lib().b200_plugin_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pin_buffer(array, pin=True):
    """GridForceBatch::pinBuffer / unpinBuffer on a numpy array: page-lock it once so that batched calls DMA directly."""
    _check(_lib().b200_plugin_pin_buffer(_p(array), array.nbytes, 1 if pin else 0))
