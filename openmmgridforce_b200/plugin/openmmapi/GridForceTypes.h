#ifndef B200_GRIDFORCE_TYPES_H_
#define B200_GRIDFORCE_TYPES_H_
// Same enumerators and values as the reference's openmmapi/include/GridForceTypes.h:10-32.
namespace GridForcePlugin {
enum class InvPowerMode { NONE = 0, RUNTIME = 1, STORED = 2 };
}
#endif
