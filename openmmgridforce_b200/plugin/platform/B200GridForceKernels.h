// CalcGridForceKernel for the B200 platform. Host logic only: every number comes from libgridforce_b200.so through
// the C ABI in include/gridforce_b200.h. There is no CPU path: if the library cannot open an sm_100 GPU, initialize()
// throws OpenMMException.
#ifndef B200_GRIDFORCE_KERNELS_IMPL_H_
#define B200_GRIDFORCE_KERNELS_IMPL_H_

#include <memory>
#include <string>
#include <vector>

#include "GridForceKernels.h"
#include "gridforce_b200.h"

namespace GridForcePlugin {

// Device grids are shared between Contexts of one process (the sampler pattern: many Contexts over one System,
// example/sampler.py:130-151): keyed by device, precision, geometry and a hash of the values, held by weak_ptr —
// the role of the reference CUDA platform's grid cache (CudaGridForceKernels.cpp:29-40, 468-509).
struct SharedGrid {
    gfb_grid* handle;
    explicit SharedGrid(gfb_grid* h) : handle(h) {}
    ~SharedGrid() { gfb_grid_destroy(handle); }
};

gfb_device* b200Device(int ordinal);    // process-wide, opened on first use; throws OpenMMException
// GridForce::getInterpolationMethod -> device layout: 0 trilinear (AUTO), 1 cubic B-spline (BSPLINE records), 2 tricubic Hermite (HERMITE records); 3 throws.
int b200LayoutForMethod(int interpolationMethod, int precision, const char* who);
std::shared_ptr<SharedGrid> b200AcquireGrid(gfb_device* dev, int ordinal, int precision, int layout, const std::vector<int>& counts,
                                            const std::vector<double>& spacing, const double origin[3],
                                            const std::vector<double>& vals);

struct B200StepFusion;   // the GridForces of one Context that can be evaluated by ONE launch per step (B200GridForceKernels.cpp)

class B200CalcGridForceKernel : public CalcGridForceKernel {
public:
    B200CalcGridForceKernel(std::string name, const OpenMM::Platform& platform, int deviceIndex, int precision,
                            OpenMM::ContextImpl* owner = 0)
        : CalcGridForceKernel(name, platform), deviceIndex(deviceIndex), precision(precision), dev(0), numParticles(0), owner(owner) {}
    ~B200CalcGridForceKernel();
    void initialize(const OpenMM::System& system, const GridForce& force);
    double execute(OpenMM::ContextImpl& context, bool includeForces, bool includeEnergy);
    void copyParametersToContext(OpenMM::ContextImpl& context, const GridForce& force);
    std::vector<double> getParticleGroupEnergies();
    std::vector<double> getParticleAtomEnergies();
    // Platform property "ResidentKernel": evaluation states that qualify (gfb_kernel_set_resident) keep a block on the GPU.
    void setResident(bool on, long long idleMicroseconds) {
        resident = on;
        residentIdleUs = idleMicroseconds;
    }

private:
    void applyResident(gfb_kernel* k) const;   // after gfb_kernel_create; states that do not qualify stay on the launch path
    bool resident = false;
    long long residentIdleUs = 100000;
    void release();
    void build(const GridForce& force);
    const OpenMM::System* system = 0;       // for the auto-derived inputs (NonbondedForce parameters)
    // ---- step fusion -------------------------------------------------------------------------------------------------
    friend struct B200StepFusion;
    double executeFused(OpenMM::ContextImpl& context, const double* pos, double* frc, bool& done);
    B200StepFusion* fusion = 0;
    std::string fusionKey;                  // kernels with equal, non-empty keys can share a launch
    int forceGroup = 0;
    std::vector<double> scalingCopy;        // this force's scaling factors (a fused state concatenates the members')
    std::vector<int> ligandCopy;
    double invPowerCopy = 0.0, oobKCopy = 0.0;
    int deviceIndex, precision;
    gfb_device* dev;
    std::shared_ptr<SharedGrid> grid;
    std::vector<gfb_kernel*> kernels;       // the one evaluation state of this force
    int numGroups = 0;
    size_t numGroupAtoms = 0;               // atoms over all particle groups (length of getParticleAtomEnergies())
    bool evaluatedOnce = false;
    std::vector<double> lastGroupEnergies;
    int numParticles;
    OpenMM::ContextImpl* owner;             // the Context this kernel belongs to (fusion registry key), or null
    bool groupMode = false;
};

}  // namespace GridForcePlugin
#endif
