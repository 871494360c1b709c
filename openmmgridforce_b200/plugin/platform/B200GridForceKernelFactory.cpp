#include "B200GridForceKernelFactory.h"

#include <cstdlib>

#include "B200GridForceKernels.h"
#include "B200Platform.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/windowsExport.h"

using namespace OpenMM;

namespace GridForcePlugin {

B200Platform::B200Platform() : ReferencePlatform("B200") {
    defaults[DeviceIndex()] = "0";
    defaults[Precision()] = "mixed";
}

const std::string& B200Platform::getPropertyDefaultValue(const std::string& property) const {
    std::map<std::string, std::string>::const_iterator it = defaults.find(property);
    if (it == defaults.end()) throw OpenMMException("B200 platform: unknown property '" + property + "'");
    return it->second;
}

KernelImpl* B200GridForceKernelFactory::createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const {
    if (name != CalcGridForceKernel::Name())
        throw OpenMMException("Tried to create kernel with illegal kernel name '" + name + "'");
    const B200Platform& b200 = dynamic_cast<const B200Platform&>(platform);
    const std::string& prec = b200.getPropertyDefaultValue(B200Platform::Precision());
    if (prec != "mixed" && prec != "double")
        throw OpenMMException("B200 platform: Precision must be 'mixed' or 'double', got '" + prec + "'");
    const int device = atoi(b200.getPropertyDefaultValue(B200Platform::DeviceIndex()).c_str());
    return new B200CalcGridForceKernel(name, platform, device, prec == "double" ? GFB_PRECISION_DOUBLE : GFB_PRECISION_MIXED, &context);
}

}  // namespace GridForcePlugin

using namespace GridForcePlugin;

extern "C" OPENMM_EXPORT void registerPlatforms() {
    for (int i = 0; i < Platform::getNumPlatforms(); i++)
        if (Platform::getPlatform(i).getName() == "B200") return;      // idempotent
    Platform::registerPlatform(new B200Platform());
}

extern "C" OPENMM_EXPORT void registerKernelFactories() {
    for (int i = 0; i < Platform::getNumPlatforms(); i++) {
        Platform& platform = Platform::getPlatform(i);
        if (dynamic_cast<B200Platform*>(&platform) != 0)
            platform.registerKernelFactory(CalcGridForceKernel::Name(), new B200GridForceKernelFactory());
    }
}

extern "C" OPENMM_EXPORT void registerB200GridForceKernelFactories() {
    registerPlatforms();
    registerKernelFactories();
}
