// The "B200" OpenMM platform: host-side state is the Reference platform's (std::vector<Vec3> positions/forces,
// ReferencePlatform::PlatformData — what OpenMM hands a Reference kernel, ReferenceGridForceKernels.cpp:134-142), so every
// stock kernel of ReferencePlatform (integrators, bonded forces) keeps working, while "CalcGridForce" is served by
// the sm_100a kernels behind libgridforce_b200.so. Scripts select it by name: Platform.getPlatformByName("B200").
//
// Properties, registered the OpenMM way (names pushed into Platform::platformProperties, defaults through the
// base class's non-virtual setPropertyDefaultValue — the reference CUDA platform's pattern, and what
// platforms/reference/src/ReferenceGridForceKernelFactory.cpp:44-72 relies on for the Reference platform):
//   "DeviceIndex"  GPU ordinal, default "0"
//   "Precision"    "mixed" (default) or "double"
//   "ResidentKernel"  "false" (default) or "true": one-ligand-per-step evaluations are served by a block that stays on the
//                  GPU between steps (gfb_kernel_set_resident: no launch, no synchronise per step)
//   "ResidentIdleMicroseconds"  how long that block waits for the next step before it leaves the GPU, default "100000"
// Per-Context values given to the Context constructor arrive in contextCreated() and win over the defaults;
// getPropertyValue(context, name) reports what a Context uses.
#ifndef B200_PLATFORM_H_
#define B200_PLATFORM_H_

#include <map>
#include <mutex>
#include <string>

#include "openmm/reference/ReferencePlatform.h"

namespace GridForcePlugin {

class B200Platform : public OpenMM::ReferencePlatform {
public:
    B200Platform();
    const std::string& getName() const {
        static const std::string name = "B200";
        return name;
    }
    double getSpeed() const { return 200.0; }
    bool supportsDoublePrecision() const { return true; }
    const std::string& getPropertyValue(const OpenMM::Context& context, const std::string& property) const;
    void contextCreated(OpenMM::ContextImpl& context, const std::map<std::string, std::string>& properties) const;
    void contextDestroyed(OpenMM::ContextImpl& context) const;
    // The value a kernel of `context` uses: the Context's own property if it gave one, else the platform default.
    const std::string& propertyFor(const OpenMM::ContextImpl& context, const std::string& property) const;
    static const std::string& DeviceIndex() {
        static const std::string key = "DeviceIndex";
        return key;
    }
    static const std::string& Precision() {
        static const std::string key = "Precision";
        return key;
    }
    static const std::string& ResidentKernel() {
        static const std::string key = "ResidentKernel";
        return key;
    }
    static const std::string& ResidentIdleMicroseconds() {
        static const std::string key = "ResidentIdleMicroseconds";
        return key;
    }

private:
    mutable std::mutex lock;
    mutable std::map<const OpenMM::ContextImpl*, std::map<std::string, std::string> > contextProperties;
};

}  // namespace GridForcePlugin
#endif
