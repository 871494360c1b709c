/* SWIG interface additions for the B200 platform — meant to be %include'd at the end of the reference's
 * python/gridforceplugin.i (after its GridForce declaration, :158-282), so that the existing module `gridforceplugin`
 * keeps every declaration it has and gains the batched multi-replica entry point. It uses only typemaps the reference
 * module already instantiates (vectord, vectori: gridforceplugin.i:16-28) and its %exception block (:49-59), which maps
 * OpenMMException to RuntimeError.
 *
 * Cannot be built in this repository's container (no SWIG, no OpenMM); the same C++ classes are exercised from Python by
 * tests/test_plugin.py through openmmgridforce_b200/gridforceplugin.py (ctypes).
 *
 * Nothing needs wrapping for the platform itself: OpenMM loads lib/plugins/libOpenMMGridForceB200.so, calls
 * registerPlatforms()/registerKernelFactories(), and scripts select it with Platform.getPlatformByName("B200").
 */
%{
#include "GridForceBatch.h"
%}

namespace GridForcePlugin {

class GridForceBatch {
public:
    GridForceBatch(int deviceIndex = 0, const std::string& precision = "mixed");
    ~GridForceBatch();

    int addForce(const GridForce& force);
    int getNumForces() const;
    int getNumAtoms() const;

    /* positions: flat [numReplicas][numAtoms][3] in nm -> one energy (kJ/mol) per replica */
    std::vector<double> evaluate(const std::vector<double>& positions, int numReplicas);

    /* Python: energies, forces = batch.evaluateWithForces(positions, numReplicas) */
    %apply std::vector<double>& OUTPUT { std::vector<double>& energies };
    %apply std::vector<double>& OUTPUT { std::vector<double>& forcesOut };
    void evaluateWithForces(const std::vector<double>& positions, int numReplicas,
                            std::vector<double>& energies, std::vector<double>& forcesOut);
    %clear std::vector<double>& energies;
    %clear std::vector<double>& forcesOut;

    std::vector<double> getLastGridEnergies() const;
};

}  // namespace GridForcePlugin
