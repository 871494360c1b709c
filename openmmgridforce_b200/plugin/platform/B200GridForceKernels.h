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
// GridForce::getInterpolationMethod -> device layout: 0 trilinear (AUTO), 1 cubic B-spline (BSPLINE tiles); 2 and 3 throw.
int b200LayoutForMethod(int interpolationMethod, const char* who);
std::shared_ptr<SharedGrid> b200AcquireGrid(gfb_device* dev, int ordinal, int precision, int layout, const std::vector<int>& counts,
                                            const std::vector<double>& spacing, const double origin[3],
                                            const std::vector<double>& vals);

class B200CalcGridForceKernel : public CalcGridForceKernel {
public:
    B200CalcGridForceKernel(std::string name, const OpenMM::Platform& platform, int deviceIndex, int precision)
        : CalcGridForceKernel(name, platform), deviceIndex(deviceIndex), precision(precision), dev(0), numParticles(0) {}
    ~B200CalcGridForceKernel();
    void initialize(const OpenMM::System& system, const GridForce& force);
    double execute(OpenMM::ContextImpl& context, bool includeForces, bool includeEnergy);
    void copyParametersToContext(OpenMM::ContextImpl& context, const GridForce& force);
    std::vector<double> getParticleGroupEnergies();
    std::vector<double> getParticleAtomEnergies();

private:
    void release();
    void build(const GridForce& force);
    const OpenMM::System* system = 0;       // for the auto-derived inputs (NonbondedForce parameters)
    int deviceIndex, precision;
    gfb_device* dev;
    std::shared_ptr<SharedGrid> grid;
    std::vector<gfb_kernel*> kernels;       // the one evaluation state of this force
    int numGroups = 0;
    std::vector<double> lastGroupEnergies;
    int numParticles;
    bool groupMode = false;
};

}  // namespace GridForcePlugin
#endif
