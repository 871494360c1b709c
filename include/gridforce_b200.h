/* gridforce_b200.h — C ABI of libgridforce_b200.so, the B200 (sm_100a) GridForce evaluation path.
 *
 * This is the drop-in boundary. Everything above it (the OpenMM platform plugin in
 * openmmgridforce_b200/plugin, a SWIG/cgo/ctypes binding, bench.py) talks to the device only through
 * these entry points: plain pointers and sizes, opaque handles, no C++/torch/OpenMM types.
 * Every function returns 0 on success or a negative gfb_status; gfb_last_error() gives the text
 * (thread-local). There is NO CPU fallback: without a usable CUDA device every compute call fails.
 *
 * Each entry point names the reference interface it stands in for (paths relative to the reference
 * tree jimtufts/openmmgridforce).
 *
 * Units follow OpenMM: nm, kJ/mol, kJ/mol/nm. Grid values are x-major with z fastest,
 * idx = (ix*ny + iy)*nz + iz (openmmapi/include/GridData.h:96-98).
 */
#ifndef GRIDFORCE_B200_H_
#define GRIDFORCE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GFB_API __attribute__((visibility("default")))
#else
#define GFB_API
#endif

#define GFB_VERSION 200          /* 0.2.0 */
#define GFB_MAX_GRIDS 8          /* grids fused into one launch (a System rarely has more than ele/LJr/LJa) */
#define GFB_MAX_PEERS 16         /* GPUs of one replica-sharded run (one NVSwitch box has 8) */
#define GFB_COMM_ID_BYTES 128    /* ncclUniqueId */
#define GFB_IPC_HANDLE_BYTES 64  /* cudaIpcMemHandle_t */

typedef enum {
    GFB_OK = 0,
    GFB_ERR_INVALID = -1,        /* bad argument (message says which) */
    GFB_ERR_CUDA = -2,           /* CUDA runtime/driver error, incl. "no device" */
    GFB_ERR_UNSUPPORTED = -3,    /* feature of the reference outside this path (interp method 1-3, tiled mode...) */
    GFB_ERR_NOMEM = -4
} gfb_status;

/* Arithmetic mode — OpenMM's "mixed"/"double" precision property (CudaPlatform "Precision").
 *   MIXED : grid values stored FP32, index/fraction math FP64 (bit-exact cell index), interpolation FP32,
 *           energy and force accumulation FP64 / 64-bit fixed point.
 *   DOUBLE: values stored FP64, everything FP64. */
typedef enum { GFB_PRECISION_MIXED = 0, GFB_PRECISION_DOUBLE = 1 } gfb_precision;

/* Device layout of a grid. Every stencil read is made of aligned 32-byte loads (LDG.E.256); a layout trades
 * copies of the data for fewer loads per stencil. The gather rate of an SM is about one load-lane per clock
 * whatever its width, and L2 holds whole 128-byte lines, so the best layout is the most replicated one whose
 * TOUCHED footprint still fits L2 (DESIGN.md §3 has the measurements).
 *   ROWS   x-major rows cut into overlapping 32-byte chunks (8 floats advancing by 7 / 4 doubles advancing by 3):
 *          any z-pair lies inside one chunk -> 4 loads per stencil, 1.14x (1.33x) the raw grid.
 *   PAIRS  MIXED only: 32-byte entry = 4 floats (advancing by 3) of row iy and of row iy+1 -> 2 loads, 2.67x.
 *   CELLS  the 8 corners of every cell packed -> 1 load (2 in DOUBLE), 8x.
 *   AUTO   CELLS while that copy is at most 1/16 of the GPU's memory (11 GB on B200), else ROWS.
 * The layout also fixes the INTERPOLATION METHOD (GridForce::setInterpolationMethod, openmmapi/include/GridForce.h:296):
 * the four above are trilinear (method 0, ReferenceGridForceKernels.cpp:1016-1084);
 *   BSPLINE cubic B-spline (method 1, :727-795): the index clamping of the 4x4x4 stencil is baked into a padded copy, and
 *          for every cell and padded x-plane a the 4x4 (y,z) windows of planes a and a+1 are stored back to back as one
 *          128-byte record (256 bytes in DOUBLE), so a stencil is 2 records = 2 full lines read with 8 (16) aligned
 *          32-byte loads of exactly its 64 values; 32x the raw grid. Never chosen by AUTO.
 *   POINTS tricubic Hermite (method 2, :796-893): the x-major points as the API hands them (FP32 in MIXED, FP64 in
 *          DOUBLE; 0.5x / 1x the raw grid) followed by one zero-filled x-slab. The reference forms this method's 32
 *          neighbour reads by FLAT index without clamping, so in the last y/z cells they land in the next row / slab
 *          (reproduced) and in the last x layer past the end of its vector (undefined there; zeros here). The
 *          arithmetic is FP64 in both precisions. Never chosen by AUTO.
 *   HERMITE tricubic Hermite on records: the BSPLINE record format (two 4x4 (y,z) windows of adjacent x-planes per cell
 *          and plane, 128 bytes in MIXED, 256 in DOUBLE; 32x the raw grid) filled by FLAT index instead of clamped indices, so that the
 *          32 neighbours the method reads are inside the stencil's two full lines, with the reference's next-row /
 *          next-slab values at the upper y/z edges and zeros past the array. Same results as POINTS; 2 lines per
 *          stencil instead of 32 scalar loads. Never chosen by AUTO.
 *   BSPLINE_POINTS cubic B-spline (method 1) on the raw points (FP32 / FP64, 0.5x / 1x the raw grid), indices clamped
 *          when they are formed, 64 scalar loads per stencil in the general kernel: the memory-lean counterpart of
 *          BSPLINE for grids whose records (32x) do not fit — the platform falls back to it, and to POINTS for method 2,
 *          when the record copy is refused with GFB_ERR_NOMEM. Same results as BSPLINE. Never chosen by AUTO. */
typedef enum {
    GFB_LAYOUT_AUTO = 0, GFB_LAYOUT_CELLS = 1, GFB_LAYOUT_ROWS = 2, GFB_LAYOUT_PAIRS = 3, GFB_LAYOUT_BSPLINE = 4,
    GFB_LAYOUT_POINTS = 5, GFB_LAYOUT_HERMITE = 6, GFB_LAYOUT_BSPLINE_POINTS = 7
} gfb_layout;

/* How execute writes forces.
 *   GFB_FORCE_F64_STORE : double [n_replicas][n_particles][3], plain stores (entries of particles this
 *                         kernel does not touch are left alone) — ReferencePlatform::PlatformData::forces
 *                         layout (ReferenceGridForceKernels.cpp:139-142).
 *   GFB_FORCE_F64_ADD   : same layout, atomically accumulated (forceData[i] -= ..., :1082).
 *   GFB_FORCE_FIXED_ADD : OpenMM CUDA long-force buffer: unsigned 64-bit, component-planar
 *                         [3][padded_atoms], value = (long long)(f * 2^32), atomically accumulated
 *                         (platforms/cuda/src/kernels/gridForce.cu:487-499). padded_atoms is
 *                         n_replicas*n_particles rounded up to `force_stride` given at execute.
 *   GFB_FORCE_F32_STORE : float [n_replicas][n_particles][3], plain stores. MIXED precision forms the gradient in
 *                         FP32 anyway (the FP64 modes widen it only to store it), so this halves the force bytes
 *                         that leave the GPU — what the host path's PCIe transfer is bound by. No reference
 *                         counterpart (the reference CUDA platform's forces are FP32-accurate fixed point).
 * forces == NULL anywhere means energy only (CalcGridForceKernel::execute with includeForces == false,
 * GridForceBatch::evaluate): the record kernels then skip the gradient arithmetic and the force read-modify-write. */
typedef enum { GFB_FORCE_F64_STORE = 0, GFB_FORCE_F64_ADD = 1, GFB_FORCE_FIXED_ADD = 2, GFB_FORCE_F32_STORE = 3 } gfb_force_mode;

typedef struct gfb_device gfb_device;   /* one GPU: ordinal, default stream, staging buffers */
typedef struct gfb_grid gfb_grid;       /* one grid, repacked cell-major and resident in HBM */
typedef struct gfb_kernel gfb_kernel;   /* state of a CalcGridForceKernel after initialize() */
typedef struct gfb_graph gfb_graph;     /* a captured sequence of launches (CUDA graph) */
typedef struct gfb_comm gfb_comm;       /* this process's end of a replica-sharded multi-GPU run (one rank = one GPU) */
typedef struct gfb_multi gfb_multi;     /* a replica-sharded multi-GPU run driven by ONE process */

typedef struct {
    char name[128];
    int cc_major, cc_minor;
    int sm_count;
    int l2_bytes;
    size_t total_mem_bytes;
} gfb_device_props;

/* Per-atom classification record (debug/parity): the bit-exact contract of
 * ReferenceGridForceKernels.cpp:690-696 (inside test) and :708-710 (cell index). */
typedef struct {
    int32_t inside;      /* 1 if 0 <= p-origin <= spacing*(counts-1) on all axes (upper face inclusive) */
    int32_t cell[3];     /* (ix,iy,iz), or -1,-1,-1 when the restraint branch is taken */
} gfb_class;

/* Library plumbing; no reference counterpart. gfb_last_error() is the text the plugin puts into its OpenMMException — the
 * reference reports every failure that way (e.g. openmmapi/src/GridForce.cpp:193, 306;
 * platforms/cuda/src/CudaGridForceKernels.cpp:387-403). */
GFB_API int gfb_version(void);
GFB_API const char* gfb_last_error(void);
GFB_API int gfb_device_count(int* count);

/* Opens GPU `ordinal` (cudaSetDevice + a non-blocking stream). Fails with GFB_ERR_CUDA when there is no
 * device or it is not compute capability 10.x — the library carries sm_100a code only.
 * Stands in for CudaPlatform context creation (cu.setAsCurrent(), CudaGridForceKernels.cpp:70). */
GFB_API int gfb_device_open(int ordinal, gfb_device** out);
GFB_API int gfb_device_close(gfb_device* dev);
GFB_API int gfb_device_get_props(gfb_device* dev, gfb_device_props* props);
GFB_API int gfb_device_synchronize(gfb_device* dev);

/* Uploads one grid and repacks it on the device into `layout` (gfb_layout above).
 * Replaces GridForce::getGridParameters (openmmapi/src/GridForce.cpp:355-363) + the CUDA platform's
 * float upload (CudaGridForceKernels.cpp:482-486). `vals` is a HOST pointer to counts[0]*counts[1]*counts[2]
 * doubles. counts >= 2 on every axis. */
GFB_API int gfb_grid_create(gfb_device* dev, const int counts[3], const double spacing[3], const double origin[3],
                            const double* vals, size_t n_vals, int precision, int layout, gfb_grid** out);
/* Same, from a DEVICE pointer to doubles (x-major); stream-ordered on the device's stream. What the reference CUDA platform
 * does after its own GPU grid generation (platforms/cuda/src/CudaGridForceKernels.cpp:482-486 uploads; :520-600 generates in place). */
GFB_API int gfb_grid_create_from_device(gfb_device* dev, const int counts[3], const double spacing[3],
                                        const double origin[3], const double* d_vals, size_t n_vals,
                                        int precision, int layout, gfb_grid** out);
/* ---- V3 "OMGRID" grid files: the reference's binary format (GridForce::loadFromFile / saveToFile,
 * openmmapi/src/GridForce.cpp:495-799; GridData::saveToFile, openmmapi/src/GridData.cpp:181-267). Bulk ingest: a
 * 16.8 M-point grid costs 16.8 M addGridValue() calls through SWIG in the reference's own tests
 * (python/tests/test_grid_force.py:58-59); here a file goes disk -> pinned staging -> HBM -> on-device repack. */
typedef struct {
    int counts[3];
    double spacing[3];
    double origin[3];
    int grid_type;                  /* 0 none, 1 charge, 2 ljr, 3 lja */
    double inv_power;
    int inv_power_mode;             /* InvPowerMode: 0 NONE, 1 RUNTIME, 2 STORED */
    unsigned int deriv_count;       /* 0, or 27 when the file carries derivative grids (function values come first) */
    unsigned long long data_offset; /* 128 */
} gfb_gridfile_header;

GFB_API int gfb_gridfile_read_header(const char* path, gfb_gridfile_header* header);
/* Reads the nx*ny*nz function values into a HOST buffer (n_vals must equal the header's point count). */
GFB_API int gfb_gridfile_read_values(const char* path, double* vals, size_t n_vals);
/* Writes a V3 file byte-identical to the reference's: with_trailer = 0 -> GridForce::saveToFile (header + values),
 * with_trailer = 1 -> GridData::saveToFile (adds i32 0 and the origin again). deriv_count is written as 0. */
GFB_API int gfb_gridfile_write(const char* path, const gfb_gridfile_header* header, const double* vals, size_t n_vals,
                               int with_trailer);
/* File -> device grid in one call, streamed through pinned staging in 32 MB pieces (no full host copy). Replaces
 * GridForce::loadFromFile (openmmapi/src/GridForce.cpp:495-692) followed by the kernel's upload of the values. */
GFB_API int gfb_grid_create_from_file(gfb_device* dev, const char* path, int precision, int layout, gfb_grid** out,
                                      gfb_gridfile_header* header_out);

/* Grid generation from receptor atoms — ReferenceCalcGridForceKernel::generateGrid
 * (platforms/reference/src/ReferenceGridForceKernels.cpp:465-544; the auto-generate switch of GridForce,
 * openmmapi/include/GridForce.h:342). grid_type: 1 charge, 2 ljr, 3 lja. Host inputs: pos [n_atoms][3] nm and the
 * NonbondedForce parameters (charge e, sigma nm, epsilon kJ/mol) of the same atoms; grid_cap = GridForce::getGridCap().
 * Outputs (either may be NULL): vals_out, host nx*ny*nz doubles (x-major, z fastest); grid_out, the grid already
 * repacked on the device in `precision`/`layout`, ready for gfb_kernel_create. FP64 on the GPU; matches the reference
 * to ~1e-13 relative (different libm for pow/tanh), tests assert 1e-10. */
GFB_API int gfb_grid_generate(gfb_device* dev, const int counts[3], const double spacing[3], const double origin[3],
                              int grid_type, int n_atoms, const double* pos, const double* charges, const double* sigmas,
                              const double* epsilons, double grid_cap, double* vals_out, int precision, int layout,
                              gfb_grid** grid_out);

/* GridForce::applyInvPowerTransformation (openmmapi/src/GridForce.cpp:221-272; CachedGridData::transformValues,
 * openmmapi/src/CachedGridData.cpp:50-57): G -> sign(G) * |G|^(1/inv_power) for every non-zero value, in place, computed
 * on the GPU in FP64. `vals` is a HOST pointer (uploaded, transformed, downloaded) or, with vals_on_device != 0, a DEVICE
 * pointer (stream-ordered on the device's stream, then synchronised). inv_power must be non-zero. CUDA's pow differs
 * from libm's by at most 2 ulp; tests assert 1e-14 relative against the reference's own method. */
GFB_API int gfb_inv_power_transform(gfb_device* dev, double* vals, size_t n_vals, double inv_power, int vals_on_device);

GFB_API int gfb_grid_destroy(gfb_grid* grid);
/* Frees the grid's own packed-cell copy and keeps the handle (geometry) alive. For a grid that is read ONLY through the
 * interleaved records of kernel states already created from it (2-4 grids of one geometry, no inv-power: the record
 * kernels never touch the per-grid arrays), this returns 8x the raw grid per grid (3 x 192^3: 638 MB of 1.5 GB).
 * Afterwards the grid cannot be used to create further kernel states, and a state that needs the per-grid arrays
 * (one grid, inv-power > 0, more than 4 grids) fails with GFB_ERR_INVALID instead of evaluating. gfb_multi_build does
 * this for the grids it owns. No reference counterpart. */
GFB_API int gfb_grid_release_cells(gfb_grid* grid);
GFB_API size_t gfb_grid_device_bytes(const gfb_grid* grid);
GFB_API int gfb_grid_layout(const gfb_grid* grid);   /* the layout actually chosen (resolves AUTO) */

/* Builds the evaluation state for n_grids GridForces that act on the same atoms — what
 * CalcGridForceKernel::initialize(System, GridForce) captures (ReferenceGridForceKernels.cpp:147-160),
 * for several forces at once so that one launch evaluates all of them (ele + LJr + LJa).
 *   n_atoms    atoms evaluated per replica (the reference's g_scaling_factors.size(), quirk Q6)
 *   scaling    host [n_grids][n_atoms] (GridForce::addScalingFactor order)
 *   particles  host [n_atoms] particle index of each atom inside a replica, or NULL for identity
 *              (GridForce::setLigandAtoms / setParticles). Forces are written at the PARTICLE index
 *              (what the CUDA platform does, gridForce.cu:497; the Reference platform writes at the
 *              ordinal, :1082 — identical when particles == NULL).
 *   inv_power  host [n_grids] or NULL (all 0 = off). > 0: v <- pow(v, n) with the chain rule (:1057-1080).
 *   oob_k      host [n_grids] out-of-grid restraint constants (GridForce::getOutOfBoundsRestraint).
 * All grids must have been created on `dev` with the same precision and layout. */
GFB_API int gfb_kernel_create(gfb_device* dev, int n_grids, gfb_grid* const* grids, int n_atoms,
                              const double* scaling, const int* particles, const double* inv_power,
                              const double* oob_k, gfb_kernel** out);
GFB_API int gfb_kernel_destroy(gfb_kernel* k);
/* CalcGridForceKernel::copyParametersToContext (ReferenceGridForceKernels.cpp:1123-1127): new scaling
 * factors [n_grids][n_atoms] and inv_power [n_grids] (NULL = keep). */
GFB_API int gfb_kernel_update_parameters(gfb_kernel* k, const double* scaling, const double* inv_power);

/* Which evaluation kernel a launch of this state uses: 1 = gf_eval_lines_kernel (MIXED, packed cells, one geometry,
 * 1-4 grids, no inv-power: the 2-4 grid case reads one 128-byte record per atom), 2 = gf_eval_lines_f64_kernel (DOUBLE,
 * same conditions, 2-4 grids, one 256-byte record per atom), 3 = gf_eval_bspline_kernel (MIXED B-spline records),
 * 4 = gf_eval_bspline_f64_kernel (DOUBLE B-spline records), 5 / 6 = gf_eval_bspline_kernel<.., METHOD 2> /
 * gf_eval_bspline_f64_kernel<.., METHOD 2> (MIXED / DOUBLE tricubic Hermite on HERMITE records), 0 = the general gf_eval_kernel. Introspection for tests and bench.py; no reference counterpart. */
GFB_API int gfb_kernel_eval_path(const gfb_kernel* k);

/* Particle groups (GridForce::addParticleGroup / getParticleGroupEnergies, openmmapi/include/GridForce.h:433-508;
 * CUDA platform: flattened groups + particle->group map, CudaGridForceKernels.cpp:607-675, 985-1005): gives every
 * evaluated atom an energy slot. Afterwards every energy array of execute has n_replicas * n_slots entries, entry
 * [r * n_slots + s] being the energy of the atoms of replica r whose slot is s (and [.. ][n_grids] for grid_energies).
 * slots: host [n_atoms] with values in [0, n_slots); NULL restores the default (one slot). Atoms of a group should be
 * contiguous in the atom list (one atomic per run of equal slot per warp). */
GFB_API int gfb_kernel_set_energy_slots(gfb_kernel* k, const int* slots, int n_slots);

/* Device path only: with enable != 0 every gfb_kernel_execute_device launch of this state is made with programmatic
 * stream serialization (PDL): its blocks may start, fetch positions and grid records while the tail of the PREVIOUS
 * kernel on the same stream is still running, and wait for that kernel to complete before their first write. The
 * caller promises that d_pos is not written by the kernel launched immediately before on that stream (true for
 * back-to-back evaluations of resident replicas; NOT true right after an integrator kernel that moves the atoms), and,
 * for the ADD force modes, that this kernel's predecessor only ACCUMULATES into d_forces (the force atomics of an
 * overlapped launch are issued before it waits; they commute with the predecessor's).
 * Small launches (a batch of up to ~6 tiles of 64 atoms per resident block: 8,192 ligand replicas, say) of a plain state
 * (no particle map, evaluation order or energy slots; ADD or no forces; no per-grid or per-atom energies) then run the
 * tile-striding instantiation of gf_eval_lines_kernel: a grid that is resident all at once, whose blocks park their
 * energy sums in shared memory and wait for the previous launch once, at their end — back-to-back
 * launches of that size flow into one another instead of paying ~3 us of ramp and tail each.
 * Default off. No reference counterpart. */
GFB_API int gfb_kernel_set_launch_overlap(gfb_kernel* k, int enable);

/* Resident evaluator for the one-ligand-per-MD-step case (BASELINE configs[1]; the call B200CalcGridForceKernel::execute
 * makes once per step, in place of ReferenceCalcGridForceKernel::execute, ReferenceGridForceKernels.cpp:646-1121).
 * With enable != 0, gfb_kernel_execute_host calls with n_replicas == 1 are served by ONE block that stays on the GPU and
 * is driven through page-locked memory both sides address (positions in, forces and energies out, every double a
 * 16-byte packet that carries the step number in both halves, so the data is its own arrival flag): no kernel launch,
 * no stream synchronisation, no fence and no separate command word per step. The block exits by itself after idle_us microseconds without
 * a step (<= 0: 100 000) and is launched again by the next one, so implicit device-wide synchronisations elsewhere in the
 * process (cudaFree, cudaDeviceSynchronize) wait at most that long; it is also stopped by
 * gfb_kernel_update_parameters, gfb_kernel_set_energy_slots, gfb_kernel_resident_stop and gfb_kernel_destroy.
 * Qualifies: trilinear packed cells (GFB_LAYOUT_CELLS, or the interleaved records after gfb_grid_release_cells), 1..224
 * evaluated atoms with distinct particle indices, no energy slots, F64 force modes, no per-atom energies, and
 * CUDA_LAUNCH_BLOCKING unset; GFB_ERR_UNSUPPORTED otherwise (calls that do not qualify later on — several replicas, FP32
 * forces — take the normal path). Results are those of the general kernel's arithmetic (same device functions).
 * gfb_kernel_resident_launches: how many times the block has been launched (1 for an uninterrupted MD run).
 * No reference counterpart. Default off. */
GFB_API int gfb_kernel_set_resident(gfb_kernel* k, int enable, long long idle_us);
GFB_API int gfb_kernel_resident_stop(gfb_kernel* k);
GFB_API long long gfb_kernel_resident_launches(const gfb_kernel* k);
/* Where the block spent the last resident step, from %globaltimer stamps it leaves in the control block, microseconds:
 * us[0] positions seen -> all grids evaluated, us[1] -> force and energy packets issued. No reference counterpart. */
GFB_API int gfb_kernel_resident_timeline(const gfb_kernel* k, double us[2]);

/* CalcGridForceKernel::execute for host-resident data (Reference-platform style), batched over replicas.
 *   pos       host [n_replicas][n_particles][3] doubles (std::vector<Vec3> layout)
 *   energies  host out [n_replicas] (sum over grids) or NULL   ([n_replicas][n_slots] with energy slots)
 *   grid_energies host out [n_replicas][n_grids] or NULL
 *   forces    host [n_replicas][n_particles][3]; force_mode STORE overwrites the evaluated particles'
 *             entries with the total grid force, ADD adds to what is there. NULL = energy only.
 * Synchronous: returns after the results are in the host buffers. How the data moves:
 *   - one replica of <= 4096 particles (a ligand per MD step): ONE launch on host-mapped pinned staging — positions read
 *     and forces/energy stored over PCIe by the kernel itself, no copy-engine transfers;
 *   - batches: positions uploaded by the copy engine in up to 8 chunks on their own stream; with STORE and no particle
 *     indirection the kernels store the forces straight into the caller's pinned buffer (pageable buffers: into pinned
 *     staging, then memcpy), otherwise chunked D2H copies on a third stream. */
GFB_API int gfb_kernel_execute_host(gfb_kernel* k, int n_replicas, int n_particles, const double* pos,
                                    double* energies, double* grid_energies, void* forces, int force_mode);

/* Page-locks a caller-owned host buffer (cudaHostRegister, portable + mapped) so that gfb_kernel_execute_host DMAs
 * from/into it directly instead of staging through its own pinned buffer — what GridForceBatch does with the caller's
 * position/force arrays. Registering twice is not an error. No reference counterpart. */
GFB_API int gfb_host_register(void* ptr, size_t bytes);
GFB_API int gfb_host_unregister(void* ptr);

/* Per-atom energies (GridForce::getParticleAtomEnergies, openmmapi/include/GridForce.h:508; reference CUDA platform:
 * atomEnergyBuffer, platforms/cuda/src/kernels/gridForce.cu:502-504). After request(enable != 0) every
 * gfb_kernel_execute_host call also keeps each evaluated atom's energy (summed over the kernel's grids) on the device;
 * get() copies the last call's [n_replicas][n_atoms] values (atom-list order) to the host. */
GFB_API int gfb_kernel_request_atom_energies(gfb_kernel* k, int enable);
GFB_API int gfb_kernel_get_atom_energies(gfb_kernel* k, double* out, size_t n);

/* CalcGridForceKernel::execute (openmmapi/include/GridForceKernels.h:63-71) for device-resident data, CUDA-platform
 * style — what CudaCalcGridForceKernel::execute does with cu.getPosq() / cu.getLongForceBuffer()
 * (platforms/cuda/src/CudaGridForceKernels.cpp:787-879, launch at :975-978): enqueues ONE kernel on
 * `stream` (a cudaStream_t; NULL = the device's own stream) and returns without synchronising.
 *   d_pos          device [n_replicas][n_particles][3] doubles
 *   d_energies     device [n_replicas] doubles, ACCUMULATED into (caller zeroes), or NULL
 *   d_grid_energies device [n_replicas][n_grids], accumulated, or NULL
 *   d_forces       device buffer in the layout `force_mode` names, or NULL
 *   force_stride   FIXED_ADD only: padded atom count (>= n_replicas*n_particles) = plane stride
 *   d_order        device [n_replicas*n_atoms] evaluation order (from gfb_kernel_sort_atoms) or NULL
 *   d_energies_clear device [n_replicas] buffer zero-filled by this launch, or NULL: the accumulator of the NEXT
 *                  step when the caller alternates two energy buffers, which saves a memset launch per step */
GFB_API int gfb_kernel_execute_device(gfb_kernel* k, int n_replicas, int n_particles, const double* d_pos,
                                      double* d_energies, double* d_grid_energies, void* d_forces,
                                      int force_mode, long long force_stride, const int* d_order,
                                      double* d_energies_clear, void* stream);

/* Order of the atoms by the brick of grid cells they sit in (grid 0; bricks of 4^3 cells numbered along a Morton
 * curve), so neighbouring lanes read neighbouring lines. Positions move less than a cell per MD step, so the order is
 * reused for many steps. Own counting sort (histogram, scan, scatter): three small kernels, no library. No reference counterpart (BASELINE.json north_star item 2).
 * d_order: device out [n_replicas*n_atoms], a permutation of the flattened [replica][atom] list. Stream-ordered. */
GFB_API int gfb_kernel_sort_atoms(gfb_kernel* k, int n_replicas, int n_particles, const double* d_pos,
                                  int* d_order, void* stream);

/* Runs only the classification stage on the device (same device function the evaluation uses) for grid
 * `grid_index`: host out cls [n_replicas*n_atoms]. Used by the bit-exact index parity tests against
 * platforms/reference/src/ReferenceGridForceKernels.cpp:687-715 (origin shift, inclusive inside test, index and fraction). */
GFB_API int gfb_kernel_classify_host(gfb_kernel* k, int grid_index, int n_replicas, int n_particles,
                                     const double* pos, gfb_class* cls);

/* Converts an OpenMM-style fixed-point force buffer (value = (long long)(f * 2^32), component-planar;
 * platforms/cuda/src/kernels/gridForce.cu:487-499) to doubles [n][3] on the device (stream-ordered). */
GFB_API int gfb_forces_fixed_to_f64(gfb_device* dev, const void* d_fixed, long long force_stride, long long n,
                                    double* d_out, void* stream);

/* Replica-sharded runs: the only cross-GPU traffic is the gather of per-replica energies. This puts
 * `bytes` from local device memory into the same offset of every peer's buffer with the copy engines over NVLink/NVSwitch
 * (cudaMemcpyAsync on peer-mapped addresses: no SM is taken from the evaluation kernel running next to it, unlike an
 * NCCL all-gather's CTAs). peer_dst: n_peers base addresses valid in THIS process (e.g. torch symmetric memory's
 * buffer_ptrs, or cudaIpcOpenMemHandle results; the entry of this rank itself may be included: a local copy);
 * first_peer staggers the order so that ranks do not all write to the same peer at once. Stream-ordered on `stream`;
 * cross-rank completion is the caller's barrier. No reference counterpart (the reference is single-GPU). */
GFB_API int gfb_peer_put(gfb_device* dev, const void* d_src, void* const* peer_dst, int n_peers, size_t dst_offset,
                         size_t bytes, int first_peer, void* stream);

/* ---- CUDA graphs ------------------------------------------------------------------------------------------------
 * Launch-bound loops (a few-microsecond evaluation per step: one ligand, or one rank's shard of a strong-scaled batch)
 * are captured once and replayed: begin() puts `stream` into capture mode, every gfb_kernel_execute_device /
 * gfb_comm_* call made on that stream until end() is recorded instead of run (programmatic-dependent-launch edges
 * included), and launch() replays the whole sequence with one driver call. No reference counterpart. */
GFB_API int gfb_graph_begin(gfb_device* dev, void* stream);
GFB_API int gfb_graph_end(gfb_device* dev, void* stream, gfb_graph** out);
GFB_API int gfb_graph_launch(gfb_graph* g, void* stream);
GFB_API int gfb_graph_destroy(gfb_graph* g);

/* ---- Replica-sharded multi-GPU runs (replaces the sequential replica loop of the reference's
 * example/sampler.py:130-164) -----------------------------------------------------------------------------------------
 * Replicas are independent, so GPU g of N evaluates replicas [g*R/N, (g+1)*R/N) against its own copy of the grids and
 * nothing is exchanged on the force path. The one collective is the gather of per-replica energies. Two deployments:
 *
 * (a) one process per GPU (torchrun/mpirun as the launcher): gfb_comm. Rank 0 calls gfb_comm_unique_id and hands the
 *     128 bytes to every rank by whatever the launcher offers; every rank calls gfb_comm_create (ncclCommInitRank).
 *     NCCL is loaded with dlopen("libnccl.so.2") at that moment; the library itself does not link it.
 *       gfb_comm_all_gather        ncclAllGather of `count` doubles per rank on `stream`.
 *       fused gather               gfb_comm_gather_alloc allocates this rank's gathered array ([2][count_total] doubles,
 *                                  double-buffered) + arrival flags and returns its cudaIpc handle; the launcher
 *                                  all-gathers the handles; gfb_comm_gather_attach maps every peer's array.
 *                                  gfb_kernel_execute_device_gather is then an evaluation launch whose LAST block
 *                                  copies the launch's energies into every rank's gathered array at `gather_offset`
 *                                  (plain stores over NVLink/NVSwitch peer mappings) and raises its arrival flag on
 *                                  every rank: compute and collective in one kernel, no NCCL launch, no extra kernel on
 *                                  the producing side. gfb_comm_gather_wait enqueues a small kernel that waits for
 *                                  all ranks' flags of that gather and copies the array out. Rule: every gather
 *                                  launch is followed by a gather_wait on the same stream before the next one.
 * (b) one process for all GPUs: gfb_multi (ncclCommInitAll, peer access enabled between the devices, one host thread
 *     per device on the host path). This is what GridForceBatch(devices) uses.
 */
GFB_API int gfb_comm_unique_id(unsigned char id[GFB_COMM_ID_BYTES]);
GFB_API int gfb_comm_create(gfb_device* dev, int world_size, int rank, const unsigned char id[GFB_COMM_ID_BYTES], gfb_comm** out);
GFB_API int gfb_comm_destroy(gfb_comm* c);
GFB_API int gfb_comm_all_gather(gfb_comm* c, const double* d_send, double* d_recv, size_t count, void* stream);
GFB_API int gfb_comm_gather_alloc(gfb_comm* c, size_t count_total, unsigned char handle_out[GFB_IPC_HANDLE_BYTES]);
GFB_API int gfb_comm_gather_attach(gfb_comm* c, const unsigned char* handles /* [world_size][GFB_IPC_HANDLE_BYTES], rank order */);
/* As gfb_kernel_execute_device (no per-grid energies, no evaluation order) + the fused gather of d_energies
 * (n_replicas * n_slots doubles) into every rank's gathered array at element gather_offset. No reference counterpart: the
 * reference evaluates replicas one Context after the other on one device (example/sampler.py:153-164). */
GFB_API int gfb_kernel_execute_device_gather(gfb_kernel* k, int n_replicas, int n_particles, const double* d_pos,
                                             double* d_energies, void* d_forces, int force_mode, long long force_stride,
                                             double* d_energies_clear, gfb_comm* c, size_t gather_offset, void* stream);
/* The producer side as a kernel of its own: copies `count` doubles from d_energies into every rank's gathered array at
 * gather_offset and raises the arrival flags, stream-ordered after whatever produced d_energies. Same protocol as the
 * fused tail (one more small launch, nothing added to the evaluation kernel); pair it with gfb_comm_gather_wait. No reference
 * counterpart (the gather of SURVEY.md 8(e); the reference collects energies in a Python list, example/sampler.py:153-164). */
GFB_API int gfb_comm_gather_push(gfb_comm* c, const double* d_energies, size_t count, size_t gather_offset, void* stream);
/* The whole gather as ONE kernel, flag-in-data: every double travels to every rank as a 16-byte packet that carries the
 * gather's sequence number in both 8-byte halves, so the data is its own arrival flag — no fence, no flag round, no
 * second launch (NCCL's LL protocol, over this library's peer mappings). The kernel stores this rank's `count` values into
 * every rank's packet array at gather_offset, then polls this rank's packet array until all count_total values of this
 * gather have arrived and writes them to d_out. Every rank calls it once per gather, in the same order; capturable.
 * No reference counterpart (the "final gather of per-replica energies" of BASELINE.json's north_star). */
GFB_API int gfb_comm_gather(gfb_comm* c, const double* d_energies, size_t count, size_t gather_offset, double* d_out, void* stream);
/* Waits (on `stream`, device side) until every rank's slice of the oldest gather not yet consumed has arrived, then
 * copies the complete [count_total] array into d_out (device memory of the caller). Gather sequence numbers live on the
 * device, so a launch/wait sequence captured with gfb_graph_* can be replayed. A peer that does not arrive within ~20 s
 * raises a flag that gfb_comm_gather_status reports after the stream has been synchronised (the wait kernel never
 * spins forever). */
GFB_API int gfb_comm_gather_wait(gfb_comm* c, double* d_out, void* stream);
GFB_API int gfb_comm_gather_status(gfb_comm* c);   /* GFB_OK, or GFB_ERR_CUDA after a timed-out wait or rendezvous */
/* Device-side rendezvous of all ranks (needs gfb_comm_gather_attach): enqueues a one-block kernel on `stream` that
 * publishes this rank's arrival to every peer over the peer mappings and waits for every peer's; work enqueued behind it
 * starts within an NVLink round trip on all ranks — much tighter than a host barrier, whose ranks leave tens of
 * microseconds apart (that skew is otherwise charged to the first collective that follows, i.e. to the energy gather of
 * a few-hundred-microsecond window). hold != 0: the kernel first waits for gfb_comm_rendezvous_release(), which the host
 * calls after it has enqueued the work that follows, so no rank leaves the rendezvous with an empty stream.
 * Not capturable into a graph when held. No reference counterpart (MPI_Barrier of a host-driven multi-rank run). */
GFB_API int gfb_comm_rendezvous(gfb_comm* c, int hold, void* stream);
GFB_API int gfb_comm_rendezvous_release(gfb_comm* c);

/* One process, n_devices GPUs: what replaces the sequential loop over replica Contexts of the reference's sampler
 * (example/sampler.py:130-164). add_grid uploads and repacks the grid on every device; build creates one
 * evaluation state per device (arguments as gfb_kernel_create, identity particles).
 *   gfb_multi_execute_host   pos host [n_replicas][n_atoms][3] -> energies host [n_replicas], forces host (layout of
 *                            force_mode, STORE modes; NULL = energy only). Replicas are block-partitioned over the
 *                            devices; one host thread per device drives that device's gfb_kernel_execute_host on its
 *                            slice of the caller's arrays, so the per-replica energies land where they belong with no
 *                            device-side collective at all.
 *   gfb_multi_upload         block-partitions and uploads positions once (device-resident shards; forces are kept in
 *                            OpenMM's fixed-point buffer per device).
 *   gfb_multi_step           one evaluation of every shard (one launch per device); gather: 0 none, 1 ncclAllGather of
 *                            the energies on every device, 2 gather over peer memory fused into the evaluation kernel,
 *                            3 the same peer stores as a small kernel behind it (gfb_comm_gather_push's), 4 the one-kernel
 *                            flag-in-data gather (gfb_comm_gather's; the fastest).
 *   gfb_multi_download       energies [n_replicas] as gathered on device `from_device` (after a gathering step; any
 *                            device holds all of them) and, when forces != NULL, the shards' forces as double [R][A][3]. */
GFB_API int gfb_multi_create(int n_devices, const int* ordinals, gfb_multi** out);
GFB_API int gfb_multi_destroy(gfb_multi* m);
GFB_API int gfb_multi_num_devices(const gfb_multi* m);
GFB_API int gfb_multi_add_grid(gfb_multi* m, const int counts[3], const double spacing[3], const double origin[3],
                               const double* vals, size_t n_vals, int precision, int layout);
GFB_API int gfb_multi_build(gfb_multi* m, int n_atoms, const double* scaling, const double* inv_power, const double* oob_k);
GFB_API int gfb_multi_execute_host(gfb_multi* m, int n_replicas, const double* pos, double* energies, void* forces, int force_mode);
GFB_API int gfb_multi_upload(gfb_multi* m, int n_replicas, const double* pos);
GFB_API int gfb_multi_step(gfb_multi* m, int gather);
GFB_API int gfb_multi_download(gfb_multi* m, int from_device, double* energies, double* forces);

/* Number of kernels this library has launched on any device since load (bench.py's gpu_launches). */
GFB_API unsigned long long gfb_launch_count(void);

/* Microbenchmark used for the roofline denominator: random 32-byte-sector gather over `bytes` of device
 * memory (n_loads loads per launch, `reps` launches, CUDA-event timed). Returns GB/s through *gbs. No reference counterpart. */
GFB_API int gfb_bench_sector_gather(gfb_device* dev, size_t bytes, long long n_loads, int reps, double* gbs);
/* Pinned-host <-> device copy bandwidth of this process on this GPU's link: gbs[0] H2D alone, gbs[1] D2H alone, gbs[2]
 * both directions at once (sum). bench.py prints it beside the end-to-end figure (every rank measures at the same time,
 * so it is the share of the host's PCIe/memory path each GPU gets). */
GFB_API int gfb_bench_host_copy(gfb_device* dev, size_t bytes, int reps, double gbs[3]);

#ifdef __cplusplus
}
#endif
#endif /* GRIDFORCE_B200_H_ */
