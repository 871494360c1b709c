// ForceImpl glue: creates the platform kernel by name and forwards to it
// (reference openmmapi/src/GridForceImpl.cpp:55-86).
#ifndef B200_GRIDFORCE_IMPL_H_
#define B200_GRIDFORCE_IMPL_H_

#include <map>
#include <string>
#include <vector>

#include "GridForce.h"
#include "openmm/Kernel.h"
#include "openmm/internal/ForceImpl.h"

namespace GridForcePlugin {

class GridForceImpl : public OpenMM::ForceImpl {
public:
    explicit GridForceImpl(const GridForce& owner) : owner(owner) {}
    void initialize(OpenMM::ContextImpl& context);
    const GridForce& getOwner() const { return owner; }
    void updateContextState(OpenMM::ContextImpl&, bool&) {}
    double calcForcesAndEnergy(OpenMM::ContextImpl& context, bool includeForces, bool includeEnergy, int groups);
    std::map<std::string, double> getDefaultParameters() { return std::map<std::string, double>(); }
    std::vector<std::string> getKernelNames();
    void updateParametersInContext(OpenMM::ContextImpl& context);
    std::vector<double> getParticleGroupEnergies();
    std::vector<double> getParticleAtomEnergies();

private:
    const GridForce& owner;
    OpenMM::Kernel kernel;
};

}  // namespace GridForcePlugin
#endif
