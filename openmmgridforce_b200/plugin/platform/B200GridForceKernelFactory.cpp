#include "B200GridForceKernelFactory.h"

#include <cstdlib>

#include "B200GridForceKernels.h"
#include "B200Platform.h"
#include "openmm/Context.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/windowsExport.h"

using namespace OpenMM;

namespace GridForcePlugin {

B200Platform::B200Platform() {
    platformProperties.push_back(DeviceIndex());
    platformProperties.push_back(Precision());
    setPropertyDefaultValue(DeviceIndex(), "0");
    platformProperties.push_back(ResidentKernel());
    platformProperties.push_back(ResidentIdleMicroseconds());
    setPropertyDefaultValue(Precision(), "mixed");
    setPropertyDefaultValue(ResidentKernel(), "false");
    setPropertyDefaultValue(ResidentIdleMicroseconds(), "100000");
}

void B200Platform::contextCreated(ContextImpl& context, const std::map<std::string, std::string>& properties) const {
    ReferencePlatform::contextCreated(context, properties);      // the host arrays every Reference kernel works on
    std::lock_guard<std::mutex> g(lock);
    contextProperties[&context] = properties;
}

void B200Platform::contextDestroyed(ContextImpl& context) const {
    {
        std::lock_guard<std::mutex> g(lock);
        contextProperties.erase(&context);
    }
    ReferencePlatform::contextDestroyed(context);
}

const std::string& B200Platform::propertyFor(const ContextImpl& context, const std::string& property) const {
    std::lock_guard<std::mutex> g(lock);
    std::map<const ContextImpl*, std::map<std::string, std::string> >::const_iterator c = contextProperties.find(&context);
    if (c != contextProperties.end()) {
        std::map<std::string, std::string>::const_iterator it = c->second.find(property);
        if (it != c->second.end()) return it->second;
    }
    return getPropertyDefaultValue(property);       // throws "Illegal property name" for an unknown one
}

const std::string& B200Platform::getPropertyValue(const Context& context, const std::string& property) const {
    return propertyFor(getContextImpl(context), property);
}

KernelImpl* B200GridForceKernelFactory::createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const {
    if (name != CalcGridForceKernel::Name())
        throw OpenMMException("Tried to create kernel with illegal kernel name '" + name + "'");
    const B200Platform& b200 = dynamic_cast<const B200Platform&>(platform);
    const std::string& prec = b200.propertyFor(context, B200Platform::Precision());
    if (prec != "mixed" && prec != "double")
        throw OpenMMException("B200 platform: Precision must be 'mixed' or 'double', got '" + prec + "'");
    const int device = atoi(b200.propertyFor(context, B200Platform::DeviceIndex()).c_str());
    const std::string& res = b200.propertyFor(context, B200Platform::ResidentKernel());
    if (res != "true" && res != "false")
        throw OpenMMException("B200 platform: ResidentKernel must be 'true' or 'false', got '" + res + "'");
    const long long idle = atoll(b200.propertyFor(context, B200Platform::ResidentIdleMicroseconds()).c_str());
    B200CalcGridForceKernel* k = new B200CalcGridForceKernel(name, platform, device, prec == "double" ? GFB_PRECISION_DOUBLE : GFB_PRECISION_MIXED, &context);
    k->setResident(res == "true", idle);
    return k;
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
