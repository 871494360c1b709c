// TEST INFRASTRUCTURE — not product code. Nothing under openmmgridforce_b200/ may link this.
//
// C-ABI wrapper around the reference's own, unmodified Reference-platform kernel
// (ReferenceCalcGridForceKernel, /root/reference/platforms/reference/src/
// ReferenceGridForceKernels.cpp:147-160 initialize, :646-1121 execute), driven the way
// OpenMM drives it: System -> GridForce -> Context(ReferencePlatform) -> GridForceImpl
// (openmmapi/src/GridForceImpl.cpp:55-68) -> CalcGridForceKernel::execute.
// The reference sources are compiled where they lie (see Makefile); only OpenMM itself
// is replaced by third_party/openmm_shim.
#include <cstring>
#include <iostream>
#include <mutex>
#include <sstream>
#include <vector>

#include "GridData.h"
#include "GridForce.h"
#include "GridForceKernels.h"
#include "openmm/Context.h"
#include "openmm/NonbondedForce.h"
#include "openmm/OpenMMException.h"
#include "openmm/Platform.h"
#include "openmm/System.h"
#include "openmm/reference/ReferencePlatform.h"

extern "C" void registerKernelFactories();  // ReferenceGridForceKernelFactory.cpp:47

using namespace OpenMM;
using GridForcePlugin::GridForce;
using GridForcePlugin::InvPowerMode;

namespace {

// The reference prints debug text from execute() on its first calls (:662-704); swallow it. std::cout's buffer is
// process-wide and bench.py --impl reference drives one Context per thread, so the swap is reference-counted under a
// mutex: the first caller installs a discarding buffer that lives for the whole process, the last one restores.
struct DiscardBuf : std::streambuf {
    int_type overflow(int_type c) { return traits_type::not_eof(c); }
    std::streamsize xsputn(const char*, std::streamsize n) { return n; }
};
struct CoutSilencer {
    static std::mutex& lock() { static std::mutex m; return m; }
    static int& users() { static int n = 0; return n; }
    static std::streambuf*& saved() { static std::streambuf* b = 0; return b; }
    CoutSilencer() {
        static DiscardBuf* discard = new DiscardBuf();   // never destroyed: threads may still be printing at exit
        std::lock_guard<std::mutex> g(lock());
        if (users()++ == 0) saved() = std::cout.rdbuf(discard);
    }
    ~CoutSilencer() {
        std::lock_guard<std::mutex> g(lock());
        if (--users() == 0) std::cout.rdbuf(saved());
    }
};

Platform& referencePlatform() {
    static ReferencePlatform* platform = 0;
    if (!platform) {
        platform = new ReferencePlatform();
        Platform::registerPlatform(platform);
        registerKernelFactories();
    }
    return *platform;
}

struct Handle {
    System system;
    std::vector<GridForce*> forces;  // owned by system
    Context* context;
    int numParticles;
    Handle() : context(0), numParticles(0) {}
    ~Handle() { delete context; }
};

thread_local std::string lastError;

}  // namespace

extern "C" {

const char* oracle_ref_last_error() { return lastError.c_str(); }

// One System holding `numParticles` particles; forces are added with oracle_ref_add_grid,
// then oracle_ref_finalize creates the Context (-> kernel initialize()).
void* oracle_ref_create(int numParticles) {
    Handle* h = new Handle();
    h->numParticles = numParticles;
    for (int i = 0; i < numParticles; i++) h->system.addParticle(1.0);
    return h;
}

// Adds one GridForce the way python/tests/test_grid_force.py:40-64 does (addGridCounts,
// addGridSpacing, values, scaling factors). ligandAtoms may be NULL (identity mapping).
int oracle_ref_add_grid(void* handle, const int* counts, const double* spacing, const double* origin,
                        const double* vals, long long nVals, const double* scaling, int nScaling,
                        const int* ligandAtoms, int nLigandAtoms, double invPower, double oobK,
                        int interpolationMethod, int forceGroup) {
    Handle* h = static_cast<Handle*>(handle);
    try {
        GridForce* f = new GridForce();
        f->addGridCounts(counts[0], counts[1], counts[2]);
        f->addGridSpacing(spacing[0], spacing[1], spacing[2]);
        f->setGridOrigin(origin[0], origin[1], origin[2]);
        f->setGridValues(std::vector<double>(vals, vals + nVals));
        f->setScalingFactors(std::vector<double>(scaling, scaling + nScaling));
        if (ligandAtoms && nLigandAtoms > 0)
            f->setLigandAtoms(std::vector<int>(ligandAtoms, ligandAtoms + nLigandAtoms));
        if (invPower != 0.0) f->setInvPowerMode(InvPowerMode::STORED, invPower);
        f->setOutOfBoundsRestraint(oobK);
        f->setInterpolationMethod(interpolationMethod);
        f->setForceGroup(forceGroup);
        h->system.addForce(f);
        h->forces.push_back(f);
        return 0;
    } catch (std::exception& e) {
        lastError = e.what();
        return 1;
    }
}

int oracle_ref_finalize(void* handle) {
    Handle* h = static_cast<Handle*>(handle);
    try {
        CoutSilencer quiet;
        h->context = new Context(h->system, referencePlatform());
        return 0;
    } catch (std::exception& e) {
        lastError = e.what();
        return 1;
    }
}

// positions: [numParticles][3] nm. forces (out, may be NULL): [numParticles][3] kJ/mol/nm,
// zeroed then accumulated by every GridForce whose group bit is set, exactly as
// forceData[] is in the reference (:1082, :1116). Returns the summed energy through *energy.
int oracle_ref_execute(void* handle, const double* positions, int groups, double* energy, double* forces) {
    Handle* h = static_cast<Handle*>(handle);
    try {
        std::vector<Vec3> pos(h->numParticles);
        for (int i = 0; i < h->numParticles; i++)
            pos[i] = Vec3(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]);
        h->context->setPositions(pos);
        CoutSilencer quiet;
        *energy = h->context->computeForcesAndEnergy(true, true, groups);
        if (forces) {
            const std::vector<Vec3>& f = h->context->getForces();
            for (int i = 0; i < h->numParticles; i++) {
                forces[3 * i] = f[i][0];
                forces[3 * i + 1] = f[i][1];
                forces[3 * i + 2] = f[i][2];
            }
        }
        return 0;
    } catch (std::exception& e) {
        lastError = e.what();
        return 1;
    }
}

// Timing entry point: `reps` back-to-back evaluations on the positions already
// resident in the Context (no marshalling inside the loop). Returns the last energy.
int oracle_ref_execute_repeat(void* handle, const double* positions, int reps, double* energy) {
    Handle* h = static_cast<Handle*>(handle);
    try {
        std::vector<Vec3> pos(h->numParticles);
        for (int i = 0; i < h->numParticles; i++)
            pos[i] = Vec3(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]);
        h->context->setPositions(pos);
        CoutSilencer quiet;
        double e = 0.0;
        for (int r = 0; r < reps; r++) e = h->context->computeForcesAndEnergy(true, true, 0xFFFFFFFF);
        *energy = e;
        return 0;
    } catch (std::exception& e) {
        lastError = e.what();
        return 1;
    }
}

void oracle_ref_destroy(void* handle) { delete static_cast<Handle*>(handle); }

// ---- V3 grid files through the reference's own reader/writers (for tests/test_gridfile.py) -----------------------
// mode 0: GridForce::saveToFile (openmmapi/src/GridForce.cpp:694-799); mode 1: GridData::saveToFile (GridData.cpp:181-267).
int oracle_ref_save_file(int mode, const char* path, const int* counts, const double* spacing, const double* origin,
                         const double* vals, long long nVals, const char* gridType, double invPower, int invPowerMode) {
    try {
        std::vector<double> v(vals, vals + nVals);
        if (mode == 0) {
            GridForce f;
            f.addGridCounts(counts[0], counts[1], counts[2]);
            f.addGridSpacing(spacing[0], spacing[1], spacing[2]);
            f.setGridOrigin(origin[0], origin[1], origin[2]);
            f.setGridValues(v);
            f.setGridType(gridType);
            if (invPowerMode != 0) f.setInvPowerMode(static_cast<InvPowerMode>(invPowerMode), invPower);
            f.saveToFile(path);
        } else {
            GridForcePlugin::GridData d(counts[0], counts[1], counts[2], spacing[0], spacing[1], spacing[2]);
            d.setOrigin(origin[0], origin[1], origin[2]);
            d.setValues(v);
            d.setGridType(gridType);
            d.setInvPower(invPower);
            d.setInvPowerMode(static_cast<InvPowerMode>(invPowerMode));
            d.saveToFile(path);
        }
        return 0;
    } catch (std::exception& e) {
        lastError = e.what();
        return 1;
    }
}

// Grid generation through the reference's own auto-generate path: a System with a NonbondedForce carrying the
// receptor parameters and a GridForce with setAutoGenerateGrid(true); creating the Context runs
// ReferenceCalcGridForceKernel::initialize -> generateGrid (:213-278, :465-544), which copies the values back into the
// GridForce (:272). gridType: "charge" | "ljr" | "lja".
int oracle_ref_generate_grid(const int* counts, const double* spacing, const double* origin, const char* gridType, int nAtoms,
                             const double* pos, const double* charges, const double* sigmas, const double* epsilons,
                             double gridCap, double* out) {
    try {
        System system;
        NonbondedForce* nb = new NonbondedForce();
        std::vector<double> xs(nAtoms), ys(nAtoms), zs(nAtoms);
        for (int i = 0; i < nAtoms; i++) {
            system.addParticle(1.0);
            nb->addParticle(charges[i], sigmas[i], epsilons[i]);
            xs[i] = pos[3 * i];
            ys[i] = pos[3 * i + 1];
            zs[i] = pos[3 * i + 2];
        }
        system.addForce(nb);
        GridForce* f = new GridForce();
        system.addForce(f);
        f->addGridCounts(counts[0], counts[1], counts[2]);
        f->addGridSpacing(spacing[0], spacing[1], spacing[2]);
        f->setGridOrigin(origin[0], origin[1], origin[2]);
        f->setGridCap(gridCap);
        f->setAutoGenerateGrid(true);
        f->setGridType(gridType);
        f->setReceptorPositionsFromArrays(xs, ys, zs);
        {
            CoutSilencer quiet;
            Context context(system, referencePlatform());
        }
        const std::vector<double>& v = f->getGridValues();
        if (v.size() != (size_t) counts[0] * counts[1] * counts[2]) throw OpenMMException("generation produced no values");
        memcpy(out, v.data(), v.size() * sizeof(double));
        return 0;
    } catch (std::exception& e) {
        lastError = e.what();
        return 1;
    }
}

// GridForce::loadFromFile (GridForce.cpp:495-692) -> what the kernel would then pull with getGridParameters.
int oracle_ref_load_file(const char* path, int* counts, double* spacing, double* origin, double* vals, long long capacity,
                         double* invPower, int* invPowerMode) {
    try {
        GridForce f;
        f.loadFromFile(path);
        std::vector<int> c;
        std::vector<double> sp, v, sc;
        f.getGridParameters(c, sp, v, sc);
        if ((long long) v.size() > capacity) throw OpenMMException("buffer too small");
        for (int k = 0; k < 3; k++) {
            counts[k] = c[k];
            spacing[k] = sp[k];
        }
        f.getGridOrigin(origin[0], origin[1], origin[2]);
        memcpy(vals, v.data(), v.size() * sizeof(double));
        *invPower = f.getInvPower();
        *invPowerMode = static_cast<int>(f.getInvPowerMode());
        return 0;
    } catch (std::exception& e) {
        lastError = e.what();
        return 1;
    }
}

// GridForce::applyInvPowerTransformation (GridForce.cpp:221-272) on a grid in RUNTIME mode: vals are transformed in
// place; *modeAfter receives the mode the reference leaves the force in (STORED).
int oracle_ref_apply_inv_power(const int* counts, double* vals, long long nVals, double invPower, int* modeAfter) {
    try {
        GridForce f;
        f.addGridCounts(counts[0], counts[1], counts[2]);
        f.addGridSpacing(0.1, 0.1, 0.1);
        f.setGridValues(std::vector<double>(vals, vals + nVals));
        f.setInvPowerMode(InvPowerMode::RUNTIME, invPower);
        f.applyInvPowerTransformation();
        const std::vector<double>& v = f.getGridValues();
        if ((long long) v.size() != nVals) throw OpenMMException("value count changed");
        memcpy(vals, v.data(), v.size() * sizeof(double));
        *modeAfter = static_cast<int>(f.getInvPowerMode());
        return 0;
    } catch (std::exception& e) {
        lastError = e.what();
        return 1;
    }
}

}  // extern "C"
