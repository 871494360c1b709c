// The "B200" OpenMM platform: host-side state is the Reference platform's (std::vector<Vec3> positions/forces,
// ReferencePlatform::PlatformData — what OpenMM hands a Reference kernel, ReferenceGridForceKernels.cpp:134-142), so every
// stock kernel of ReferencePlatform (integrators, bonded forces) keeps working, while "CalcGridForce" is served by
// the sm_100a kernels behind libgridforce_b200.so. Scripts select it by name: Platform.getPlatformByName("B200").
// Platform properties: "DeviceIndex" (default "0"), "Precision" ("mixed" default, or "double").
#ifndef B200_PLATFORM_H_
#define B200_PLATFORM_H_

#include <map>
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
    void setPropertyDefaultValue(const std::string& property, const std::string& value) { defaults[property] = value; }
    const std::string& getPropertyDefaultValue(const std::string& property) const;
    static const std::string& DeviceIndex() {
        static const std::string key = "DeviceIndex";
        return key;
    }
    static const std::string& Precision() {
        static const std::string key = "Precision";
        return key;
    }

private:
    std::map<std::string, std::string> defaults;
};

}  // namespace GridForcePlugin
#endif
