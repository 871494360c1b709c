// Batched multi-replica entry point — the addition to the reference's API surface (wrapped by the same SWIG layer,
// python/gridforceplugin_b200.i): R independent replicas (poses) of the same A atoms are evaluated against G GridForces
// in ONE kernel launch, returning one energy per replica. It replaces the reference's idiom of one Context per replica
// stepped in a Python loop (example/sampler.py:130-164) for the grid term.
//
//   GridForceBatch batch;                       // device 0, mixed precision
//   batch.addForce(ele); batch.addForce(ljr); batch.addForce(lja);      // GridForce objects, scaling factors set
//   std::vector<double> e = batch.evaluate(positions, R);               // positions: [R][A][3] nm, flat
//   batch.evaluateWithForces(positions, R, energies, forces);           // forces: [R][A][3] kJ/mol/nm, flat
//
// All forces must carry the same number of scaling factors (A). Each replica's energy is the sum over the forces,
// and each force applies its own out-of-grid restraint, exactly as G separate GridForces in one System would.
#ifndef B200_GRIDFORCE_BATCH_H_
#define B200_GRIDFORCE_BATCH_H_

#include <memory>
#include <string>
#include <vector>

#include "GridForce.h"

struct gfb_kernel;
struct gfb_device;

namespace GridForcePlugin {

struct SharedGrid;

class GridForceBatch {
public:
    explicit GridForceBatch(int deviceIndex = 0, const std::string& precision = "mixed");
    ~GridForceBatch();
    int addForce(const GridForce& force);            // returns the force's index; invalidates a built kernel
    int getNumForces() const { return (int) forces.size(); }
    int getNumAtoms() const;
    std::vector<double> evaluate(const std::vector<double>& positions, int numReplicas);
    void evaluateWithForces(const std::vector<double>& positions, int numReplicas, std::vector<double>& energies,
                            std::vector<double>& forcesOut);
    std::vector<double> getLastGridEnergies() const { return lastGridEnergies; }   // [R][G] of the last call

private:
    GridForceBatch(const GridForceBatch&);
    GridForceBatch& operator=(const GridForceBatch&);
    void build();
    void run(const std::vector<double>& positions, int numReplicas, std::vector<double>& energies, double* forcesOut);
    int deviceIndex, precision;
    std::vector<const GridForce*> forces;
    std::vector<std::shared_ptr<SharedGrid> > grids;
    gfb_device* dev;
    gfb_kernel* kernel;
    std::vector<double> lastGridEnergies;
};

}  // namespace GridForcePlugin
#endif
