// The kernel interface a platform implements — same name ("CalcGridForce") and the same five virtuals as the
// reference's openmmapi/include/GridForceKernels.h:46-92. This is the C++ side of the drop-in boundary.
#ifndef B200_GRIDFORCE_KERNELS_H_
#define B200_GRIDFORCE_KERNELS_H_

#include <string>
#include <vector>

#include "GridForce.h"
#include "openmm/KernelImpl.h"
#include "openmm/Platform.h"
#include "openmm/System.h"

namespace GridForcePlugin {

class CalcGridForceKernel : public OpenMM::KernelImpl {
public:
    static std::string Name() { return "CalcGridForce"; }
    CalcGridForceKernel(std::string name, const OpenMM::Platform& platform) : OpenMM::KernelImpl(name, platform) {}
    virtual void initialize(const OpenMM::System& system, const GridForce& force) = 0;
    virtual double execute(OpenMM::ContextImpl& context, bool includeForces, bool includeEnergy) = 0;
    virtual void copyParametersToContext(OpenMM::ContextImpl& context, const GridForce& force) = 0;
    virtual std::vector<double> getParticleGroupEnergies() = 0;
    virtual std::vector<double> getParticleAtomEnergies() = 0;
};

}  // namespace GridForcePlugin
#endif
