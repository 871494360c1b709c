// Plain-data launch parameters shared by the host launcher (gf_capi.cu) and the kernels (gf_kernels.cuh).
#ifndef GF_PARAMS_H_
#define GF_PARAMS_H_

#include <stdint.h>

#include "gridforce_b200.h"

namespace gfb {

// One grid as the kernel sees it. `cells` points at the device copy in the grid's layout
// (see gfb_layout in gridforce_b200.h and the repack kernels in gf_kernels.cuh):
//   CELLS  cell (ix,iy,iz) -> 8 corners at ((ix*nc[1] + iy)*nc[2] + iz), nc = counts - 1
//   ROWS   row (ix,iy) -> row_chunks 32-byte chunks; chunk j holds z = j*(W-1) .. j*(W-1)+W-1
//   PAIRS  (ix,iy<ny-1) -> row_chunks 32-byte entries; entry j = {row iy: z=3j..3j+3, row iy+1: same}
struct GridView {
    const void* cells;
    const double* scaling;   // [n_atoms] this grid's row of scaling factors (FP64 in both precisions)
    double origin[3];
    double spacing[3];
    double inv_spacing[3];   // fl(1/spacing): fast-path quotient, re-divided exactly near integers
    double hcorner[3];       // spacing*(counts-1), computed on the host exactly as the reference does
    int nc[3];               // cells per axis = counts - 1
    int row_chunks;          // ROWS / PAIRS: 32-byte units per row
    int cell_stride;         // CELLS: elements between consecutive cells (8, or 8*k when k grids are interleaved)
    int pad_;
    double inv_power;        // 0 = off
    double oob_k;
};

// Peer tables of the fused energy gather, in DEVICE memory of each rank (built once by gfb_comm_gather_attach /
// gfb_multi_create). peer_data[r] / peer_flags[r] are addresses valid on THIS device that point into rank r's memory
// (cudaIpcOpenMemHandle mappings, or plain peer pointers inside one process; r == my_rank: local memory).
constexpr int kMaxPeers = 16;
struct GatherTable {
    double* peer_data[kMaxPeers];                // rank r's gathered array: [2][count_total] doubles
    unsigned long long* peer_flags[kMaxPeers];   // rank r's flags: [2][kMaxPeers]; entry [parity][s] = last seq published by rank s
    long long count_total;
    int n_peers, my_rank;
    unsigned int ticket;                         // blocks of the current launch that have finished (reset by the last one)
    unsigned int timed_out;                      // set by gf_gather_wait_kernel when a peer's flag did not arrive in time
    // Sequence numbers live on the DEVICE so that a captured graph can be replayed: nothing about "which gather is this"
    // is baked into a launch's arguments. Every rank issues and waits for gathers in the same order.
    unsigned long long issued;                   // gathers this rank has published (gather_tail)
    unsigned long long waited;                   // gathers this rank has consumed (gf_gather_wait_kernel)
    unsigned int wait_ticket;                    // blocks of the current wait kernel that have finished
    unsigned int copy_ticket;                    // copier blocks of the current gather that have finished their slice
    unsigned long long rendezvous;               // device-side rendezvous of all ranks completed so far (gf_rendezvous_kernel);
                                                 // its arrival flags are a third row of every rank's flag array
    // Flag-in-data gather (gf_gather_ll_kernel): every rank's memory also holds [2][count_total] 16-byte packets
    // {lo32, seq32, hi32, seq32} at byte offset ll_offset from peer_data[r]; a packet is valid for gather number seq once
    // both of its 8-byte halves carry seq.
    long long ll_offset;
    unsigned long long ll_seq;                   // flag-in-data gathers completed by this rank
    unsigned int ll_ticket;                      // blocks of the current flag-in-data gather that have finished
};

struct EvalParams {
    GridView grid[GFB_MAX_GRIDS];
    int n_grids;
    int n_atoms;             // atoms evaluated per replica
    int n_particles;         // particles per replica (stride of pos/forces)
    int n_replicas;
    long long total;         // n_replicas * n_atoms = threads doing work
    const double* pos;       // [n_replicas][n_particles][3]
    const int* particles;    // [n_atoms] or null
    const int* order;        // [total] or null
    const int* slots;        // [n_atoms] energy slot of each atom inside a replica (particle groups), or null
    int n_slots;             // energy slots per replica (1 without groups)
    int energy_store;        // single replica, single block: the block's energy is STORED to *energies (which may then be
                             // host-mapped memory) instead of accumulated with an atomic
    double* energies;        // [n_replicas] or null, accumulated
    double* grid_energies;   // [n_replicas][n_grids] or null, accumulated
    double* energies_clear;  // [n_replicas] or null: zero-filled by this launch (next step's accumulator)
    void* forces;            // layout per force mode, or null
    long long force_stride;  // FIXED_ADD plane stride
    // gf_eval_lines_kernel only (gf_eval_lines.cuh)
    const void* lines;       // 2-4 grids of one geometry: one record per cell = 4 slots of 8 packed corners (128 B MIXED,
                             // 256 B DOUBLE)
    unsigned div_magic;      // floor(2^32 / n_atoms) (saturated): t / n_atoms = umulhi(t, div_magic) or that + 1
    unsigned pdl;            // launch with programmatic stream serialization (gfb_kernel_set_launch_overlap)
    unsigned ahead_blocks;   // gf_eval_lines_kernel: blocks resident at once (SMs x blocks per SM); 0 = no position prefetch
    unsigned defer;          // host -> launcher: this launch MAY run the tile-striding variant of gf_eval_lines_kernel (launch
                             // overlap, ADD / no forces, no per-atom or per-grid energies); launcher -> kernel: it does
    unsigned persist_blocks; // with defer, host -> launcher: the device's SM count (the launcher sizes the resident grid)
    double near_int[3];      // 1.8e-15 * cells per axis: fractions this close to 0 or 1 take the exact division
    double* atom_energies;   // [n_replicas][n_atoms] or null: each evaluated atom's energy, summed over the grids, stored
                             // (GridForce::getParticleAtomEnergies)
    // Fused energy gather (replica-sharded multi-GPU runs, gf_eval_lines_kernel only): the LAST block of the launch to
    // finish copies this launch's n_replicas*n_slots energies into every peer's gathered array at gather_offset and
    // then publishes its gather sequence number in its slot of every peer's flag array. nullptr = off.
    GatherTable* gather;
    long long gather_offset; // first element of this rank's slice in the gathered array
    int force_mode;          // gfb_force_mode for the kernels that switch on it at run time (general, B-spline)
};

struct ClassifyParams {
    GridView grid;
    int n_atoms, n_particles;
    long long total;
    const double* pos;
    const int* particles;
    gfb_class* out;
    int exact_div;
    double near_int[3];      // as EvalParams::near_int (gf_classify_lines_kernel)
};

}  // namespace gfb
#endif
