#ifndef B200_GRIDFORCE_KERNEL_FACTORY_H_
#define B200_GRIDFORCE_KERNEL_FACTORY_H_

#include <string>

#include "openmm/KernelFactory.h"

namespace GridForcePlugin {

// Same signature as every platform factory of the reference (platforms/cuda/include/CudaGridForceKernelFactory.h:12-15).
class B200GridForceKernelFactory : public OpenMM::KernelFactory {
public:
    OpenMM::KernelImpl* createKernelImpl(std::string name, const OpenMM::Platform& platform, OpenMM::ContextImpl& context) const;
};

}  // namespace GridForcePlugin

// The two symbols OpenMM's plugin loader calls after dlopen (reference: ReferenceGridForceKernelFactory.cpp:44-60,
// CudaGridForceKernelFactory.cpp:17-29), plus an explicit entry for static linking.
extern "C" void registerPlatforms();
extern "C" void registerKernelFactories();
extern "C" void registerB200GridForceKernelFactories();

#endif
