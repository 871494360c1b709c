// C entry points that drive the PLUGIN path end to end the way OpenMM would — registerPlatforms() /
// registerKernelFactories() -> Platform::getPlatformByName("B200") -> System + GridForce -> Context -> GridForceImpl ->
// B200CalcGridForceKernel::execute — so that tests/test_plugin.py can compare it with the oracle from Python (there is
// no SWIG or OpenMM in the build container; with them, python/gridforceplugin_b200.i exposes the same classes).
// Mirrors oracle/ref_driver.cpp's call shape on purpose: same inputs, two implementations.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "B200GridForceKernelFactory.h"
#include "B200Platform.h"
#include "GridForce.h"
#include "GridForceBatch.h"
#include "openmm/Context.h"
#include "openmm/NonbondedForce.h"
#include "openmm/Platform.h"
#include "openmm/System.h"
#include "openmm/internal/windowsExport.h"

using namespace OpenMM;
using namespace GridForcePlugin;

namespace {
struct Handle {
    System system;
    std::vector<GridForce*> forces;   // owned by system
    Context* context;
    GridForceBatch* batch;          // std::vector entry points
    GridForceBatch* batchBuffers;   // caller-buffer entry points (possibly several devices)
    std::vector<int> batchDevices;
    std::string batchPrecision;
    int numParticles;
    Handle() : context(0), batch(0), batchBuffers(0), numParticles(0) {}
    ~Handle() {
        delete context;
        delete batch;
        delete batchBuffers;
    }
};
thread_local std::string lastError;
}  // namespace

#define GUARD(...)                    \
    try {                             \
        __VA_ARGS__;                  \
        return 0;                     \
    } catch (std::exception & e) {    \
        lastError = e.what();         \
        return 1;                     \
    }

extern "C" {

OPENMM_EXPORT const char* b200_plugin_last_error() { return lastError.c_str(); }

// 1 if a platform called "B200" is registered and serves the "CalcGridForce" kernel name.
OPENMM_EXPORT int b200_plugin_register() {
    GUARD({
        registerPlatforms();
        registerKernelFactories();
        Platform& p = Platform::getPlatformByName("B200");
        if (!p.supportsKernels(std::vector<std::string>(1, "CalcGridForce"))) throw OpenMMException("B200 platform lacks CalcGridForce");
    })
}

OPENMM_EXPORT int b200_plugin_set_property(const char* name, const char* value) {
    GUARD({
        registerB200GridForceKernelFactories();
        // through the BASE class, as a script's platform.setPropertyDefaultValue(...) does (the method is not virtual in
        // OpenMM): works because the B200 platform registers its property names, throws for an unknown name
        Platform& base = Platform::getPlatformByName("B200");
        base.setPropertyDefaultValue(name, value);
    })
}

OPENMM_EXPORT int b200_plugin_get_default_property(const char* name, char* out, int capacity) {
    GUARD({
        registerB200GridForceKernelFactories();
        const std::string& v = Platform::getPlatformByName("B200").getPropertyDefaultValue(name);
        snprintf(out, capacity, "%s", v.c_str());
    })
}

OPENMM_EXPORT void* b200_plugin_create(int numParticles) {
    Handle* h = new Handle();
    h->numParticles = numParticles;
    for (int i = 0; i < numParticles; i++) h->system.addParticle(1.0);
    return h;
}

OPENMM_EXPORT int b200_plugin_add_grid(void* handle, const int* counts, const double* spacing, const double* origin, const double* vals,
                                       long long nVals, const double* scaling, int nScaling, const int* ligandAtoms, int nLigandAtoms,
                                       double invPower, double oobK, int interpolationMethod, int forceGroup) {
    Handle* h = static_cast<Handle*>(handle);
    GUARD({
        GridForce* f = new GridForce();
        h->system.addForce(f);
        h->forces.push_back(f);
        f->addGridCounts(counts[0], counts[1], counts[2]);
        f->addGridSpacing(spacing[0], spacing[1], spacing[2]);
        f->setGridOrigin(origin[0], origin[1], origin[2]);
        f->setGridValues(std::vector<double>(vals, vals + nVals));
        for (int i = 0; i < nScaling; i++) f->addScalingFactor(scaling[i]);
        if (ligandAtoms && nLigandAtoms > 0) f->setLigandAtoms(std::vector<int>(ligandAtoms, ligandAtoms + nLigandAtoms));
        if (invPower != 0.0) f->setInvPowerMode(InvPowerMode::STORED, invPower);
        f->setOutOfBoundsRestraint(oobK);
        f->setInterpolationMethod(interpolationMethod);
        f->setForceGroup(forceGroup);
    })
}

// GridForce::setInvPowerMode(RUNTIME, n) + applyInvPowerTransformation() on `vals` (in place); *modeAfter = the mode the
// force is left in. Errors (wrong mode, empty grid, n == 0) surface exactly as the C++ class raises them.
OPENMM_EXPORT int b200_plugin_apply_inv_power(double* vals, long long nVals, int mode, double invPower, int* modeAfter) {
    GUARD({
        GridForce f;
        f.setGridValues(std::vector<double>(vals, vals + nVals));
        f.setInvPowerMode(static_cast<InvPowerMode>(mode), invPower);
        f.applyInvPowerTransformation();
        const std::vector<double>& v = f.getGridValues();
        memcpy(vals, v.data(), v.size() * sizeof(double));
        *modeAfter = static_cast<int>(f.getInvPowerMode());
    })
}

// System.addForce(NonbondedForce): the parameter source of the auto-derived inputs (charge e, sigma nm, epsilon kJ/mol).
OPENMM_EXPORT int b200_plugin_add_nonbonded(void* handle, const double* charges, const double* sigmas, const double* epsilons, int n) {
    Handle* h = static_cast<Handle*>(handle);
    GUARD({
        NonbondedForce* nb = new NonbondedForce();
        for (int i = 0; i < n; i++) nb->addParticle(charges[i], sigmas[i], epsilons[i]);
        h->system.addForce(nb);
    })
}

// GridForce::setAutoCalculateScalingFactors / setScalingProperty / setAutoGenerateGrid / setGridType / setGridCap /
// setReceptorAtoms / setReceptorPositions on force `force` (reference GridForce.h:171-198, 335-342, 523-573).
OPENMM_EXPORT int b200_plugin_set_auto(void* handle, int force, int autoScaling, const char* scalingProperty, int autoGenerate,
                                       const char* gridType, double gridCap, const int* receptorAtoms, int nReceptorAtoms,
                                       const double* receptorPositions, int nReceptorPositions) {
    Handle* h = static_cast<Handle*>(handle);
    GUARD({
        GridForce* f = h->forces.at(force);
        f->setAutoCalculateScalingFactors(autoScaling != 0);
        f->setScalingProperty(scalingProperty ? scalingProperty : "");
        f->setAutoGenerateGrid(autoGenerate != 0);
        f->setGridType(gridType ? gridType : "");
        f->setGridCap(gridCap);
        if (receptorAtoms && nReceptorAtoms > 0) f->setReceptorAtoms(std::vector<int>(receptorAtoms, receptorAtoms + nReceptorAtoms));
        std::vector<Vec3> pos(nReceptorPositions);
        for (int i = 0; i < nReceptorPositions; i++)
            pos[i] = Vec3(receptorPositions[3 * i], receptorPositions[3 * i + 1], receptorPositions[3 * i + 2]);
        f->setReceptorPositions(pos);
    })
}

// What the kernel wrote back into the force at Context creation: which = 0 grid values, 1 scaling factors.
OPENMM_EXPORT int b200_plugin_get_force_data(void* handle, int force, int which, double* out, long long capacity, long long* count) {
    Handle* h = static_cast<Handle*>(handle);
    GUARD({
        std::vector<int> c;
        std::vector<double> sp, v, sc;
        h->forces.at(force)->getGridParameters(c, sp, v, sc);
        const std::vector<double>& src = which == 0 ? v : sc;
        *count = (long long) src.size();
        for (long long i = 0; i < (long long) src.size() && i < capacity; i++) out[i] = src[i];
    })
}

OPENMM_EXPORT int b200_plugin_add_particle_group(void* handle, int force, const char* name, const int* particles, const double* scaling, int n) {
    Handle* h = static_cast<Handle*>(handle);
    GUARD({
        h->forces.at(force)->addParticleGroup(name, std::vector<int>(particles, particles + n),
                                              scaling ? std::vector<double>(scaling, scaling + n) : std::vector<double>());
    })
}

// Context creation on the platform named `platform` ("B200"): GridForceImpl::initialize -> kernel initialize().
// properties: n (name, value) pairs, the Context constructor's platform-specific properties (OpenMM: Context(system,
// integrator, platform, properties)); they win over the platform's defaults for this Context only.
OPENMM_EXPORT int b200_plugin_finalize(void* handle, const char* platform, const char* const* names, const char* const* values, int n) {
    Handle* h = static_cast<Handle*>(handle);
    GUARD({
        registerB200GridForceKernelFactories();
        std::map<std::string, std::string> props;
        for (int i = 0; i < n; i++) props[names[i]] = values[i];
        h->context = new Context(h->system, Platform::getPlatformByName(platform), props);
    })
}

// Platform::getPropertyValue(context, name): what this Context actually uses.
OPENMM_EXPORT int b200_plugin_get_property(void* handle, const char* name, char* out, int capacity) {
    Handle* h = static_cast<Handle*>(handle);
    GUARD({
        if (!h->context) throw OpenMMException("finalize first");
        const std::string& v = h->context->getPlatform().getPropertyValue(*h->context, name);
        snprintf(out, capacity, "%s", v.c_str());
    })
}

OPENMM_EXPORT int b200_plugin_set_particles(void* handle, int force, const int* particles, int n) {
    Handle* h = static_cast<Handle*>(handle);
    GUARD({ h->forces.at(force)->setParticles(std::vector<int>(particles, particles + n)); })
}

OPENMM_EXPORT int b200_plugin_atom_energies(void* handle, int force, double* out, int capacity, int* count) {
    Handle* h = static_cast<Handle*>(handle);
    GUARD({
        std::vector<double> e = h->forces.at(force)->getParticleAtomEnergies(*h->context);
        *count = (int) e.size();
        for (int i = 0; i < (int) e.size() && i < capacity; i++) out[i] = e[i];
    })
}

OPENMM_EXPORT int b200_plugin_execute(void* handle, const double* positions, int groups, double* energy, double* forces) {
    Handle* h = static_cast<Handle*>(handle);
    GUARD({
        if (!h->context) throw OpenMMException("finalize first");
        std::vector<Vec3> pos(h->numParticles);
        for (int i = 0; i < h->numParticles; i++) pos[i] = Vec3(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]);
        h->context->setPositions(pos);
        *energy = h->context->computeForcesAndEnergy(true, true, groups);
        if (forces) memcpy(forces, &h->context->getForces()[0], sizeof(double) * 3 * h->numParticles);
    })
}

// `reps` back-to-back evaluations from C++, as an integrator's step loop issues them (no marshalling inside the loop):
// seconds per call through *seconds, last energy through *energy. Used by bench.py's configs[1] figure.
OPENMM_EXPORT int b200_plugin_time_execute(void* handle, const double* positions, int groups, int reps, double* seconds, double* energy) {
    Handle* h = static_cast<Handle*>(handle);
    GUARD({
        if (!h->context) throw OpenMMException("finalize first");
        std::vector<Vec3> pos(h->numParticles);
        for (int i = 0; i < h->numParticles; i++) pos[i] = Vec3(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]);
        h->context->setPositions(pos);
        double e = 0.0;
        const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < reps; r++) e = h->context->computeForcesAndEnergy(true, true, groups);
        *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / (reps > 0 ? reps : 1);
        *energy = e;
    })
}

OPENMM_EXPORT int b200_plugin_update_scaling(void* handle, int force, const double* scaling, int n) {
    Handle* h = static_cast<Handle*>(handle);
    GUARD({
        h->forces.at(force)->setScalingFactors(std::vector<double>(scaling, scaling + n));
        h->forces.at(force)->updateParametersInContext(*h->context);
    })
}

OPENMM_EXPORT int b200_plugin_group_energies(void* handle, int force, double* out, int capacity, int* count) {
    Handle* h = static_cast<Handle*>(handle);
    GUARD({
        std::vector<double> e = h->forces.at(force)->getParticleGroupEnergies(*h->context);
        *count = (int) e.size();
        for (int i = 0; i < (int) e.size() && i < capacity; i++) out[i] = e[i];
    })
}

// Batched entry point over the forces added so far (GridForceBatch): positions [R][A][3] -> energies [R], forces.
// The std::vector overloads (what SWIG's vectord typemaps call).
OPENMM_EXPORT int b200_plugin_batch_evaluate(void* handle, const char* precision, const double* positions, int numReplicas,
                                             double* energies, double* forces) {
    Handle* h = static_cast<Handle*>(handle);
    GUARD({
        if (!h->batch) {
            h->batch = new GridForceBatch(0, precision);
            for (size_t i = 0; i < h->forces.size(); i++) h->batch->addForce(*h->forces[i]);
        }
        const size_t n = (size_t) numReplicas * h->batch->getNumAtoms() * 3;
        std::vector<double> pos(positions, positions + n), e, f;
        if (forces) {
            h->batch->evaluateWithForces(pos, numReplicas, e, f);
            memcpy(forces, f.data(), sizeof(double) * n);
        } else {
            e = h->batch->evaluate(pos, numReplicas);
        }
        memcpy(energies, e.data(), sizeof(double) * numReplicas);
    })
}

// The pointer overloads on caller-owned buffers (numpy arrays through the SWIG buffer typemaps), optionally over several
// devices: no copy, no vector. forcesF32 != 0: `forces` is float [R][A][3]. devices == NULL: device 0.
OPENMM_EXPORT int b200_plugin_batch_evaluate_buffers(void* handle, const char* precision, const int* devices, int nDevices,
                                                     const double* positions, int numReplicas, double* energies, void* forces,
                                                     int forcesF32) {
    Handle* h = static_cast<Handle*>(handle);
    GUARD({
        const std::vector<int> devs = devices && nDevices > 0 ? std::vector<int>(devices, devices + nDevices) : std::vector<int>(1, 0);
        if (h->batchBuffers && (h->batchDevices != devs || h->batchPrecision != precision)) {
            delete h->batchBuffers;
            h->batchBuffers = 0;
        }
        if (!h->batchBuffers) {
            h->batchBuffers = devs.size() > 1 ? new GridForceBatch(devs, precision) : new GridForceBatch(devs[0], precision);
            h->batchDevices = devs;
            h->batchPrecision = precision;
            for (size_t i = 0; i < h->forces.size(); i++) h->batchBuffers->addForce(*h->forces[i]);
        }
        if (!forces) h->batchBuffers->evaluate(positions, numReplicas, energies);
        else if (forcesF32) h->batchBuffers->evaluateWithForcesF32(positions, numReplicas, energies, static_cast<float*>(forces));
        else h->batchBuffers->evaluateWithForces(positions, numReplicas, energies, static_cast<double*>(forces));
    })
}

OPENMM_EXPORT int b200_plugin_pin_buffer(void* ptr, long long bytes, int pin) {
    GUARD({
        if (pin) GridForceBatch::pinBuffer(ptr, (size_t) bytes);
        else GridForceBatch::unpinBuffer(ptr);
    })
}

OPENMM_EXPORT void b200_plugin_destroy(void* handle) { delete static_cast<Handle*>(handle); }

}  // extern "C"
