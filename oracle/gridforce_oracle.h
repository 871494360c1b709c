/* TEST INFRASTRUCTURE — the parity oracle ("port"): a plain-C restatement of the reference's
 * Reference-platform GridForce evaluation. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product (openmmgridforce_b200/) never does.
 *
 * Parity is PINNED: tests/test_oracle.py checks this restatement bit-for-bit against the reference's
 * own unmodified kernel (oracle/_ref/liboracle_ref.so, built by oracle/Makefile from /root/reference)
 * and against the golden vectors under tests/golden/ that the same _ref build generated.
 *
 * Follows /root/reference/platforms/reference/src/ReferenceGridForceKernels.cpp:
 *   :646-696   loop header, origin shift, inclusive inside test
 *   :706-715   cell index and fraction (true FP64 division, truncation)
 *   :1022-1084 trilinear value (z -> y -> x), gradient, inv-power chain rule, accumulate
 *   :727-795   cubic B-spline branch (interpolation method 1): clamped 4x4x4 stencil
 *   :796-893   tricubic Hermite branch (interpolation method 2): finite-difference derivatives, flat-index neighbours
 *   :1093-1117 out-of-grid harmonic restraint (also taken by inside atoms with scale == 0)
 */
#ifndef GRIDFORCE_ORACLE_H_
#define GRIDFORCE_ORACLE_H_
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int counts[3];        /* nx, ny, nz grid points                       (GridForce::addGridCounts) */
    double spacing[3];    /* nm                                           (GridForce::addGridSpacing) */
    double origin[3];     /* nm                                           (GridForce::setGridOrigin) */
    const double* vals;   /* nx*ny*nz, x-major, z fastest                 (GridData.h:96-98) */
    double inv_power;     /* 0 = off; > 0: v <- pow(v, n) with chain rule (:1057-1059, :1076-1080) */
    double oob_k;         /* kJ/mol/nm^2, out-of-grid restraint           (GridForce.cpp:52 default 1e4) */
    int interp_method;    /* 0 trilinear (:1016-1084), 1 cubic B-spline (:727-795), 2 tricubic Hermite (:796-893) (GridForce::setInterpolationMethod) */
} gfo_grid;

/* Per-atom classification record, for the bit-exact index tests. cell = -1 when the atom took the
 * restraint branch. `inside` is the geometric test only (:690-696), before the scale != 0 test. */
typedef struct {
    int inside;
    int cell[3];
} gfo_class;

/* One GridForce, one Context: the loop at :682-1118.
 *   scaling[n_scaling]    loop bound is n_scaling, not the particle count (quirk Q6)
 *   ligand_atoms          NULL => particle index = ia; else pos is read at ligand_atoms[ia] while the force
 *                         is written at ia (quirk Q1, :684 vs :1082)
 *   pos                   [n_particles][3] nm
 *   forces                [>= n_scaling][3], ACCUMULATED into with -= (caller zeroes); may be NULL
 *   cls                   [n_scaling] or NULL
 * Returns the energy (sequential FP64 sum in atom order, as the reference). */
double gfo_execute(const gfo_grid* grid, const double* scaling, int n_scaling, const int* ligand_atoms,
                   const double* pos, double* forces, gfo_class* cls);

/* Batched form used for the CPU baseline and the batched parity tests: R replicas x A atoms, G grids
 * (the reference evaluates this as R Contexts x G GridForces, example/sampler.py:130-164).
 *   scaling   [G][A];   pos [R][A][3];   forces (out, zeroed here) [R][A][3];
 *   energies  (out) [R][G] per-grid energies; the replica's energy is the sum over G in grid order.
 *   n_threads >= 1: replicas are split into contiguous chunks, one pthread each (the reference itself
 *   is single-threaded; this is the "all host cores" arm). */
void gfo_execute_batched(const gfo_grid* grids, int n_grids, const double* scaling, int n_replicas, int n_atoms,
                         const double* pos, double* forces, double* energies, int n_threads);

/* Grid generation from receptor atoms (ReferenceGridForceKernels.cpp:465-544): for every grid point the sum over
 * atoms of   charge: 138.935456*q/r   ljr: sqrt(eps)*(2 sigma)^6/r^12   lja: -2 sqrt(eps)*(2 sigma)^3/r^6
 * with r clamped to >= 1e-6 nm, then capped with U*tanh(v/U) (U = grid cap, GridForce.cpp:52 default 41840).
 * grid_type: 1 charge, 2 ljr, 3 lja (the V3 file codes). pos [n_atoms][3] nm. out: nx*ny*nz, x-major, z fastest.
 * n_threads >= 1 splits the x planes over pthreads (the reference is single-threaded). */
void gfo_generate_grid(const int counts[3], const double spacing[3], const double origin[3], int grid_type, int n_atoms,
                       const double* pos, const double* charges, const double* sigmas, const double* epsilons,
                       double grid_cap, double* out, int n_threads);

/* GridForce::applyInvPowerTransformation (openmmapi/src/GridForce.cpp:262-268), the RUNTIME inv-power mode's one-off
 * transformation of the stored values, in place. */
void gfo_inv_power_transform(double* vals, size_t n, double inv_power);

#ifdef __cplusplus
}
#endif
#endif
