// Batched multi-replica entry point — the addition to the reference's API surface (wrapped by the same SWIG layer,
// python/gridforceplugin_b200.i): R independent replicas (poses) of the same A atoms are evaluated against G GridForces
// in ONE kernel launch per GPU, returning one energy per replica. It replaces the reference's idiom of one Context per
// replica stepped in a Python loop (example/sampler.py:130-164) for the grid term.
//
//   GridForceBatch batch;                       // device 0, mixed precision
//   batch.addForce(ele); batch.addForce(ljr); batch.addForce(lja);      // GridForce objects, scaling factors set
//   std::vector<double> e = batch.evaluate(positions, R);               // positions: [R][A][3] nm, flat; energy only
//   batch.evaluateWithForces(positions, R, energies, forces);           // forces: [R][A][3] kJ/mol/nm, flat
//
//   GridForceBatch multi(std::vector<int>{0, 1, 2, 3, 4, 5, 6, 7});     // replicas block-partitioned over 8 GPUs
//
// Pointer overloads take caller-owned buffers (numpy arrays through the SWIG buffer typemaps): no std::vector is built
// or zero-filled, and buffers page-locked once with pinBuffer() are DMA'd directly. evaluate() is an energy-only
// evaluation on the device as well (no gradient arithmetic, no force traffic; the sampler's Monte-Carlo use,
// example/sampler.py:186-226). All forces must carry the same number of scaling factors (A). Each replica's energy is
// the sum over the forces, and each force applies its own out-of-grid restraint, exactly as G separate GridForces in one
// System would.
#ifndef B200_GRIDFORCE_BATCH_H_
#define B200_GRIDFORCE_BATCH_H_

#include <memory>
#include <string>
#include <vector>

#include "GridForce.h"

struct gfb_kernel;
struct gfb_device;
struct gfb_multi;

namespace GridForcePlugin {

struct SharedGrid;

class GridForceBatch {
public:
    explicit GridForceBatch(int deviceIndex = 0, const std::string& precision = "mixed");
    explicit GridForceBatch(const std::vector<int>& deviceIndices, const std::string& precision = "mixed");
    ~GridForceBatch();
    int addForce(const GridForce& force);            // returns the force's index; invalidates a built kernel
    int getNumForces() const { return (int) forces.size(); }
    int getNumAtoms() const;                         // scaling factors per force (cached once the batch is built)
    int getNumDevices() const { return (int) devices.size(); }
    std::vector<double> evaluate(const std::vector<double>& positions, int numReplicas);
    void evaluateWithForces(const std::vector<double>& positions, int numReplicas, std::vector<double>& energies,
                            std::vector<double>& forcesOut);
    // Caller-owned buffers: positions [R][A][3], energies [R], forces [R][A][3] (NULL = energy only).
    void evaluate(const double* positions, int numReplicas, double* energies);
    void evaluateWithForces(const double* positions, int numReplicas, double* energies, double* forcesOut);
    void evaluateWithForcesF32(const double* positions, int numReplicas, double* energies, float* forcesOut);
    // Page-lock a caller buffer once (cudaHostRegister) so that every later call DMAs from/into it directly.
    static void pinBuffer(void* ptr, size_t bytes);
    static void unpinBuffer(void* ptr);
    // [R][G] per-force energies of the last single-device call (empty on several devices).
    std::vector<double> getLastGridEnergies() const { return lastGridEnergies; }

private:
    GridForceBatch(const GridForceBatch&);
    GridForceBatch& operator=(const GridForceBatch&);
    void setPrecision(const std::string& name);
    void build();
    void release();
    void run(const double* positions, size_t nPositions, int numReplicas, double* energies, void* forcesOut, int forceMode,
             bool wantGridEnergies);
    std::vector<int> devices;
    int precision;
    std::vector<const GridForce*> forces;
    std::vector<std::shared_ptr<SharedGrid> > grids;
    gfb_device* dev;
    gfb_kernel* kernel;      // one device
    gfb_multi* multi;        // several devices
    int numAtoms;            // cached by build()
    bool built;
    std::vector<double> lastGridEnergies;
};

}  // namespace GridForcePlugin
#endif
