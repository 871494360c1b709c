#include "B200GridForceKernels.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <sstream>

#include "B200Platform.h"
#include "openmm/NonbondedForce.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"

using namespace OpenMM;

namespace GridForcePlugin {

static void check(int rc, const char* what) {
    if (rc != GFB_OK) throw OpenMMException(std::string("GridForce[B200]: ") + what + ": " + gfb_last_error());
}

// ---- process-wide device and grid registries ---------------------------------------------------------------------
static std::mutex registryMutex;

gfb_device* b200Device(int ordinal) {
    static std::map<int, gfb_device*> devices;
    std::lock_guard<std::mutex> lock(registryMutex);
    std::map<int, gfb_device*>::iterator it = devices.find(ordinal);
    if (it != devices.end()) return it->second;
    gfb_device* dev = 0;
    check(gfb_device_open(ordinal, &dev), "cannot open the GPU (this platform has no CPU fallback)");
    devices[ordinal] = dev;
    return dev;
}

// Word-wise multiply/xor-shift hash (8 bytes per step, several GB/s): grids are 10^7 doubles and this runs at every
// Context creation and copyParametersToContext, so a byte-wise FNV (1 GB/s) would cost more than the upload.
static unsigned long long hashWords(const void* data, size_t bytes, unsigned long long h = 0x9E3779B97F4A7C15ull) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    size_t i = 0;
    for (; i + 8 <= bytes; i += 8) {
        unsigned long long w;
        memcpy(&w, p + i, 8);
        h = (h ^ w) * 0xBF58476D1CE4E5B9ull;
        h ^= h >> 29;
    }
    unsigned long long tail = 0;
    if (i < bytes) memcpy(&tail, p + i, bytes - i);
    h = (h ^ tail ^ (unsigned long long) bytes) * 0x94D049BB133111EBull;
    return h ^ (h >> 32);
}

int b200LayoutForMethod(int interpolationMethod, int precision, const char* who) {
    if (interpolationMethod == 0) return GFB_LAYOUT_AUTO;        // trilinear
    if (interpolationMethod == 1) return GFB_LAYOUT_BSPLINE;     // cubic B-spline
    // tricubic Hermite (finite-difference derivatives): records (two full lines per stencil in MIXED, four in DOUBLE);
    // GFB_LAYOUT_POINTS (the raw points, 32 scalar loads per stencil) stays available through the C ABI for memory-lean use
    (void) precision;
    if (interpolationMethod == 2) return GFB_LAYOUT_HERMITE;
    throw OpenMMException(std::string(who) + ": interpolation method 3 (quintic Hermite) needs the 27 derivative grids and is "
                          "not implemented on this platform; use 0 (trilinear), 1 (cubic B-spline) or 2 (tricubic)");
}

void B200CalcGridForceKernel::applyResident(gfb_kernel* k) const {
    if (!resident) return;
    // GFB_ERR_UNSUPPORTED (B-spline / tricubic layouts, more than 224 atoms, repeated particles): that state keeps the
    // launch-per-step path; the property is a request, not a requirement.
    if (gfb_kernel_set_resident(k, 1, residentIdleUs) == GFB_ERR_CUDA)
        throw OpenMMException(std::string("GridForce[B200]: resident evaluator: ") + gfb_last_error());
}

std::shared_ptr<SharedGrid> b200AcquireGrid(gfb_device* dev, int ordinal, int precision, int layout, const std::vector<int>& counts,
                                            const std::vector<double>& spacing, const double origin[3],
                                            const std::vector<double>& vals) {
    static std::map<std::string, std::weak_ptr<SharedGrid> > cache;
    // Two independent 64-bit hashes of the values (different seeds and a different word order for the second) plus the
    // geometry: a hit needs all 128 bits and the size to agree, so one unlucky collision cannot hand a force another
    // force's grid.
    unsigned long long h1 = hashWords(vals.data(), vals.size() * sizeof(double));
    unsigned long long h2 = hashWords(vals.data(), vals.size() * sizeof(double), 0xD6E8FEB86659FD93ull);
    h1 = hashWords(counts.data(), 3 * sizeof(int), h1);
    h1 = hashWords(spacing.data(), 3 * sizeof(double), h1);
    h1 = hashWords(origin, 3 * sizeof(double), h1);
    h2 = hashWords(origin, 3 * sizeof(double), h2);
    h2 = hashWords(spacing.data(), 3 * sizeof(double), h2);
    h2 = hashWords(counts.data(), 3 * sizeof(int), h2);
    std::ostringstream key;
    key << ordinal << ':' << precision << ':' << layout << ':' << vals.size() << ':' << h1 << ':' << h2;
    std::lock_guard<std::mutex> lock(registryMutex);
    for (std::map<std::string, std::weak_ptr<SharedGrid> >::iterator it = cache.begin(); it != cache.end();) {
        if (it->second.expired()) cache.erase(it++);      // grids whose last Context is gone
        else ++it;
    }
    std::shared_ptr<SharedGrid> hit = cache[key.str()].lock();
    if (hit) return hit;
    gfb_grid* g = 0;
    // Record layouts are 32x the raw grid. When that copy does not fit (GFB_ERR_NOMEM, said before any allocation), or
    // B200_COMPACT_LAYOUTS=1 asks for it, the same method runs on the raw points through the general kernel.
    static const bool compact = [] {
        const char* e = getenv("B200_COMPACT_LAYOUTS");
        return e && e[0] == '1';
    }();
    const int lean = layout == GFB_LAYOUT_BSPLINE ? GFB_LAYOUT_BSPLINE_POINTS : (layout == GFB_LAYOUT_HERMITE ? GFB_LAYOUT_POINTS : layout);
    int rc = (compact && lean != layout) ? GFB_ERR_NOMEM
                                         : gfb_grid_create(dev, counts.data(), spacing.data(), origin, vals.data(), vals.size(), precision, layout, &g);
    if (rc == GFB_ERR_NOMEM && lean != layout)
        rc = gfb_grid_create(dev, counts.data(), spacing.data(), origin, vals.data(), vals.size(), precision, lean, &g);
    check(rc, "grid upload");
    std::shared_ptr<SharedGrid> made(new SharedGrid(g));
    cache[key.str()] = made;
    return made;
}

// ---- step fusion ---------------------------------------------------------------------------------------------------
// An OpenMM System for docking-style MD carries one GridForce per receptor grid (electrostatic, LJ repulsive, LJ
// attractive: python/tests/test_grid_force.py:117-159, example/sampler.py), and ContextImpl::calcForcesAndEnergy calls
// their kernels one after the other. Each call is pure latency (one launch + one synchronize, ~16 us), so three forces
// cost three of them per step. The kernels of one Context that evaluate the same atoms on grids of one geometry
// therefore share a launch: the first member called in an evaluation looks at ContextImpl::getLastForceGroups(),
// evaluates every member whose force group is part of this evaluation in ONE launch (per-grid energies, summed
// forces), and the other members return their cached energy when OpenMM calls them moments later.
// B200_FUSE_FORCES=0 switches it off.
struct B200StepFusion {
    std::mutex lock;
    std::vector<B200CalcGridForceKernel*> members;
    std::map<unsigned, gfb_kernel*> fused;     // by member mask
    // Per member: the energy a fused launch of ANOTHER member computed for it, valid only for the evaluation that
    // produced it — same force-group flags and same positions (stamp = hash of a sample of the position array).
    // A member consumes its own entry when OpenMM calls it; nobody else's entries are touched by that.
    std::vector<double> cachedEnergy;
    std::vector<int> pendingGroups;
    std::vector<unsigned long long> pendingStamp;
    unsigned pendingMask = 0;
    void resize(size_t n) {
        if (cachedEnergy.size() == n) return;
        cachedEnergy.assign(n, 0.0);
        pendingGroups.assign(n, 0);
        pendingStamp.assign(n, 0);
        pendingMask = 0;
    }
    void dropFused() {
        for (std::map<unsigned, gfb_kernel*>::iterator it = fused.begin(); it != fused.end(); ++it) gfb_kernel_destroy(it->second);
        fused.clear();
        pendingMask = 0;
    }
};

static std::map<ContextImpl*, B200StepFusion*>& fusionRegistry() {
    static std::map<ContextImpl*, B200StepFusion*> reg;
    return reg;
}

static bool fusionEnabled() {
    static const bool on = [] {
        const char* e = getenv("B200_FUSE_FORCES");
        return !(e && e[0] == '0');
    }();
    return on;
}

// ---- the kernel ------------------------------------------------------------------------------------------------------
B200CalcGridForceKernel::~B200CalcGridForceKernel() {
    if (fusion) {
        std::lock_guard<std::mutex> reg(registryMutex);
        {
            std::lock_guard<std::mutex> g(fusion->lock);
            fusion->dropFused();
            std::vector<B200CalcGridForceKernel*>& m = fusion->members;
            m.erase(std::remove(m.begin(), m.end(), this), m.end());
        }
        if (fusion->members.empty()) {
            fusionRegistry().erase(owner);
            delete fusion;
        }
        fusion = 0;
    }
    release();
}

void B200CalcGridForceKernel::release() {
    for (size_t i = 0; i < kernels.size(); i++) gfb_kernel_destroy(kernels[i]);
    kernels.clear();
}

// What CalcGridForceKernel::initialize captures (ReferenceGridForceKernels.cpp:147-160), validated the way the CUDA
// platform validates (CudaGridForceKernels.cpp:387-403), and refusing the reference features outside this path.
void B200CalcGridForceKernel::build(const GridForce& force) {
    std::vector<int> counts;
    std::vector<double> spacing, vals, scaling;
    force.getGridParameters(counts, spacing, vals, scaling);
    const int layout = b200LayoutForMethod(force.getInterpolationMethod(), precision, "GridForce[B200]");
    if (force.getTiledMode()) throw OpenMMException("GridForce[B200]: tiled grids are not supported on this platform");
    double origin[3];
    force.getGridOrigin(origin[0], origin[1], origin[2]);
    const double invPower = force.getInvPower(), oobK = force.getOutOfBoundsRestraint();
    dev = b200Device(deviceIndex);

    const NonbondedForce* nonbonded = 0;
    if (system)
        for (int i = 0; i < system->getNumForces() && !nonbonded; i++)
            nonbonded = dynamic_cast<const NonbondedForce*>(&system->getForce(i));

    // Scaling factors from the NonbondedForce (ReferenceGridForceKernels.cpp:163-209; the Reference platform's formulas,
    // SURVEY.md quirk Q9), written back to the force as the reference does (:209).
    if (force.getAutoCalculateScalingFactors() && scaling.empty()) {
        const std::string prop = force.getScalingProperty();
        if (prop.empty()) throw OpenMMException("GridForce: Auto-calculate scaling factors enabled but no scaling property specified");
        if (prop != "charge" && prop != "ljr" && prop != "lja")
            throw OpenMMException("GridForce: Invalid scaling property '" + prop + "'. Must be 'charge', 'ljr', or 'lja'");
        if (!nonbonded) throw OpenMMException("GridForce: Auto-calculate scaling factors requires a NonbondedForce in the system");
        scaling.resize(numParticles);
        for (int i = 0; i < numParticles; i++) {
            double q, sigma, eps;
            nonbonded->getParticleParameters(i, q, sigma, eps);
            if (prop == "charge") scaling[i] = q;
            else scaling[i] = std::sqrt(eps) * std::pow(2.0 * sigma, prop == "ljr" ? 6.0 : 3.0);
        }
        const_cast<GridForce&>(force).setScalingFactors(scaling);
    }

    // Grid from the receptor atoms (ReferenceGridForceKernels.cpp:213-278 + generateGrid :465-544): generated on the GPU,
    // left there ready for evaluation, and copied back into the force (:272) so that saveToFile()/getGridParameters() see it.
    std::shared_ptr<SharedGrid> generated;
    if (force.getAutoGenerateGrid() && vals.empty()) {
        const std::string type = force.getGridType();
        if (type != "charge" && type != "ljr" && type != "lja")
            throw OpenMMException("GridForce: Invalid grid type '" + type + "'. Must be 'charge', 'ljr', or 'lja'");
        if (counts.size() != 3 || spacing.size() != 3)
            throw OpenMMException("GridForce: Grid counts and spacing must be set before auto-generation");
        if (!nonbonded) throw OpenMMException("GridForce: Auto-grid generation requires a NonbondedForce in the system");
        std::vector<int> receptor = force.getReceptorAtoms();
        const std::vector<int>& ligand = force.getLigandAtoms();
        const std::vector<Vec3>& rpos = force.getReceptorPositions();
        if (receptor.empty())
            for (int i = 0; i < numParticles; i++)
                if (std::find(ligand.begin(), ligand.end(), i) == ligand.end()) receptor.push_back(i);
        if (rpos.empty()) throw OpenMMException("GridForce: Receptor positions must be set for auto-grid generation");
        if (rpos.size() < receptor.size()) throw OpenMMException("GridForce: Not enough receptor positions provided");
        const size_t n = receptor.size();
        std::vector<double> q(n), sig(n), eps(n), xyz(3 * n);
        for (size_t i = 0; i < n; i++) {
            nonbonded->getParticleParameters(receptor[i], q[i], sig[i], eps[i]);
            xyz[3 * i] = rpos[i][0];
            xyz[3 * i + 1] = rpos[i][1];
            xyz[3 * i + 2] = rpos[i][2];
        }
        vals.resize((size_t) counts[0] * counts[1] * counts[2]);
        gfb_grid* g = 0;
        check(gfb_grid_generate(dev, counts.data(), spacing.data(), origin, type == "charge" ? 1 : type == "ljr" ? 2 : 3, (int) n,
                                xyz.data(), q.data(), sig.data(), eps.data(), force.getGridCap(), vals.data(), precision, layout, &g),
              "grid generation");
        generated.reset(new SharedGrid(g));
        const_cast<GridForce&>(force).setGridValues(vals);
    }
    if (counts.size() != 3 || spacing.size() != 3)
        throw OpenMMException("GridForce[B200]: grid counts and spacing must each be given exactly once");
    if (vals.size() != (size_t) counts[0] * counts[1] * counts[2])
        throw OpenMMException("GridForce[B200]: number of grid values does not match the grid counts");

    grid = generated ? generated : b200AcquireGrid(dev, deviceIndex, precision, layout, counts, spacing, origin, vals);
    release();
    gfb_grid* handle = grid->handle;
    groupMode = force.getNumParticleGroups() > 0;
    if (groupMode) {
        // Multi-ligand mode (GridForce.h:433-508): all groups are flattened into ONE atom list (particle index +
        // scaling factor + the group's energy slot per atom) and evaluated by one launch; the kernel reduces energies
        // per slot, which is what getParticleGroupEnergies() returns. Same flattening as the reference CUDA platform
        // (CudaGridForceKernels.cpp:607-675), minus its per-step host round trips (:840-853, 987-996).
        std::vector<int> particlesFlat, slots;
        std::vector<double> scalingFlat;
        numGroups = force.getNumParticleGroups();
        for (int g = 0; g < numGroups; g++) {
            const ParticleGroup& grp = force.getParticleGroup(g);
            for (size_t i = 0; i < grp.particleIndices.size(); i++) {
                if (grp.particleIndices[i] < 0 || grp.particleIndices[i] >= numParticles)
                    throw OpenMMException("GridForce[B200]: particle group '" + grp.name + "' has an index outside the System");
                particlesFlat.push_back(grp.particleIndices[i]);
                scalingFlat.push_back(grp.scalingFactors[i]);
                slots.push_back(g);
            }
        }
        gfb_kernel* k = 0;
        check(gfb_kernel_create(dev, 1, &handle, (int) particlesFlat.size(), scalingFlat.data(), particlesFlat.data(), &invPower,
                                &oobK, &k), "kernel setup");
        kernels.push_back(k);
        check(gfb_kernel_set_energy_slots(k, slots.data(), numGroups), "particle group slots");
        check(gfb_kernel_request_atom_energies(k, 1), "per-atom energies");   // getParticleAtomEnergies (GridForce.h:508)
        numGroupAtoms = particlesFlat.size();
        fusionKey.clear();
        return;
    }
    numGroups = 0;
    // Single mode: scaling factor ia belongs to particle ligandAtoms[ia] (identity when no ligand atoms are set).
    // The loop bound is the number of scaling factors, not the particle count (reference quirk Q6).
    std::vector<int> ligand = force.getLigandAtoms();
    if (!ligand.empty()) {
        if (ligand.size() != scaling.size())
            throw OpenMMException("GridForce[B200]: number of ligand atoms differs from the number of scaling factors");
        for (size_t i = 0; i < ligand.size(); i++)
            if (ligand[i] < 0 || ligand[i] >= numParticles)
                throw OpenMMException("GridForce[B200]: ligand atom index outside the System");
    } else if ((int) scaling.size() > numParticles) {
        throw OpenMMException("GridForce[B200]: more scaling factors than particles in the System");
    }
    // Particle filter (GridForce::setParticles, openmmapi/include/GridForce.h:433-440), honoured the way the reference
    // CUDA platform does (CudaGridForceKernels.cpp:122-127, 512-515; gridForce.cu:45-49): only the listed particles are
    // evaluated, each with the scaling factor that belongs to ITS particle index (zero when the force has none for it,
    // the CUDA platform's zero padding, :398-424). With ligand atoms set, the filter keeps the ligand atoms it lists.
    const std::vector<int>& filter = force.getParticles();
    if (!filter.empty()) {
        std::vector<double> byParticle(numParticles, 0.0);
        std::vector<char> has(numParticles, 0);
        for (size_t ia = 0; ia < scaling.size(); ia++) {
            const int particle = ligand.empty() ? (int) ia : ligand[ia];
            byParticle[particle] = scaling[ia];
            has[particle] = 1;
        }
        std::vector<int> keptParticles;
        std::vector<double> keptScaling;
        for (size_t i = 0; i < filter.size(); i++) {
            if (filter[i] < 0 || filter[i] >= numParticles)
                throw OpenMMException("GridForce[B200]: setParticles() index outside the System");
            if (!ligand.empty() && !has[filter[i]]) continue;
            keptParticles.push_back(filter[i]);
            keptScaling.push_back(byParticle[filter[i]]);
        }
        ligand = keptParticles;
        scaling = keptScaling;
        // (a filter that keeps nothing leaves an empty atom list: the force then evaluates nothing)
    }
    const int* particles = ligand.empty() ? 0 : ligand.data();
    gfb_kernel* k = 0;
    check(gfb_kernel_create(dev, 1, &handle, (int) scaling.size(), scaling.data(), particles, &invPower, &oobK, &k), "kernel setup");
    kernels.push_back(k);
    applyResident(k);

    // what a fused launch needs from this member, and which other kernels may share one with it
    scalingCopy = scaling;
    ligandCopy = ligand;
    invPowerCopy = invPower;
    oobKCopy = oobK;
    forceGroup = force.getForceGroup();
    std::ostringstream key;
    key.precision(17);
    key << deviceIndex << '|' << precision << '|' << layout << '|' << counts[0] << 'x' << counts[1] << 'x' << counts[2] << '|'
        << spacing[0] << ',' << spacing[1] << ',' << spacing[2] << '|' << origin[0] << ',' << origin[1] << ',' << origin[2] << '|'
        << scaling.size() << '|' << hashWords(ligand.data(), ligand.size() * sizeof(int));
    fusionKey = key.str();
    if (owner && fusionEnabled()) {
        std::lock_guard<std::mutex> reg(registryMutex);
        if (!fusion) {
            B200StepFusion*& slot = fusionRegistry()[owner];
            if (!slot) slot = new B200StepFusion();
            fusion = slot;
            std::lock_guard<std::mutex> g(fusion->lock);
            if (fusion->members.size() < 32) fusion->members.push_back(this);
            else fusion = 0;
        }
        if (fusion) {   // (re)built parameters invalidate every fused state of this Context
            std::lock_guard<std::mutex> g(fusion->lock);
            fusion->dropFused();
        }
    }
}

void B200CalcGridForceKernel::initialize(const System& system, const GridForce& force) {
    this->system = &system;
    numParticles = system.getNumParticles();
    build(force);
}

// CalcGridForceKernel::execute: positions and forces are the Reference platform's host arrays
// (std::vector<Vec3> = contiguous double[3]); forces are accumulated (forceData[i] -= ..., :1082) and the energy is
// returned, as the Reference kernel does (:1120) — includeForces/includeEnergy only prune work, never change values.
double B200CalcGridForceKernel::execute(ContextImpl& context, bool includeForces, bool includeEnergy) {
    ReferencePlatform::PlatformData* data = reinterpret_cast<ReferencePlatform::PlatformData*>(context.getPlatformData());
    std::vector<Vec3>& pos = *data->positions;
    std::vector<Vec3>& frc = *data->forces;
    static_assert(sizeof(Vec3) == 3 * sizeof(double), "Vec3 must be three packed doubles");
    if (numParticles == 0 || kernels.empty()) return 0.0;
    const double* p = &pos[0][0];
    double* f = includeForces ? &frc[0][0] : 0;
    if (fusion && !groupMode && !fusionKey.empty()) {
        bool done = false;
        const double e = executeFused(context, p, f, done);
        if (done) return e;
    }
    double total = 0.0;
    lastGroupEnergies.assign(groupMode ? numGroups : 1, 0.0);      // one energy per group slot (or the single total)
    check(gfb_kernel_execute_host(kernels[0], 1, numParticles, p, lastGroupEnergies.data(), 0, f, GFB_FORCE_F64_ADD), "execute");
    evaluatedOnce = true;
    for (size_t i = 0; i < lastGroupEnergies.size(); i++) total += lastGroupEnergies[i];
    (void) includeEnergy;
    return total;
}

// See B200StepFusion. Returns with done = false when this call has to run on its own (no partner in this evaluation).
double B200CalcGridForceKernel::executeFused(ContextImpl& context, const double* pos, double* frc, bool& done) {
    B200StepFusion& fu = *fusion;
    std::lock_guard<std::mutex> g(fu.lock);
    const size_t n = fu.members.size();
    size_t me = n;
    for (size_t i = 0; i < n; i++)
        if (fu.members[i] == this) me = i;
    if (me == n || n < 2) return 0.0;
    const int groups = context.getLastForceGroups();
    fu.resize(n);
    // Which evaluation is this? OpenMM calls every ForceImpl once per evaluation, in System order, with the same
    // force-group flags and the same positions: flags + a hash over a sample of the positions identify it.
    const size_t nd = (size_t) numParticles * 3;
    unsigned long long stamp = hashWords(pos, std::min<size_t>(nd, 96) * sizeof(double), (unsigned long long) nd);
    if (nd > 96) {
        stamp = hashWords(pos + nd - 96, 96 * sizeof(double), stamp);
        const size_t stride = std::max<size_t>(1, nd / 64);
        for (size_t i = 0; i < nd; i += stride) stamp = hashWords(pos + i, sizeof(double), stamp);
    }
    if ((fu.pendingMask >> me) & 1u) {
        fu.pendingMask &= ~(1u << me);      // consumed, or stale: either way this member's entry is gone
        if (fu.pendingGroups[me] == groups && fu.pendingStamp[me] == stamp) {   // evaluated moments ago by the member called first
            done = true;
            return fu.cachedEnergy[me];
        }
    }
    // members that OpenMM will call in this evaluation and that can share a launch with this one
    unsigned mask = 0;
    int count = 0;
    for (size_t i = 0; i < n; i++) {
        const B200CalcGridForceKernel* k = fu.members[i];
        if (k->fusionKey == fusionKey && !k->groupMode && k->numParticles == numParticles &&
            ((groups >> k->forceGroup) & 1) != 0 && count < GFB_MAX_GRIDS) {
            mask |= 1u << i;
            count++;
        }
    }
    if (count < 2 || !((mask >> me) & 1u)) return 0.0;
    gfb_kernel*& fk = fu.fused[mask];
    if (!fk) {
        std::vector<gfb_grid*> grids;
        std::vector<double> scalingAll, invPower, oobK;
        for (size_t i = 0; i < n; i++) {
            if (!((mask >> i) & 1u)) continue;
            const B200CalcGridForceKernel* k = fu.members[i];
            grids.push_back(k->grid->handle);
            scalingAll.insert(scalingAll.end(), k->scalingCopy.begin(), k->scalingCopy.end());
            invPower.push_back(k->invPowerCopy);
            oobK.push_back(k->oobKCopy);
        }
        check(gfb_kernel_create(dev, count, grids.data(), (int) scalingCopy.size(), scalingAll.data(),
                                ligandCopy.empty() ? 0 : ligandCopy.data(), invPower.data(), oobK.data(), &fk), "fused kernel setup");
        applyResident(fk);
    }
    double total = 0.0;
    std::vector<double> perGrid(count, 0.0);
    check(gfb_kernel_execute_host(fk, 1, numParticles, pos, &total, perGrid.data(), frc, GFB_FORCE_F64_ADD), "fused execute");
    int slot = 0;
    double mine = 0.0;
    for (size_t i = 0; i < n; i++) {
        if (!((mask >> i) & 1u)) continue;
        if (i == me) mine = perGrid[slot];
        else {                              // only the bits of THIS launch's members change
            fu.cachedEnergy[i] = perGrid[slot];
            fu.pendingGroups[i] = groups;
            fu.pendingStamp[i] = stamp;
            fu.pendingMask |= 1u << i;
        }
        fu.members[i]->lastGroupEnergies.assign(1, perGrid[slot]);
        slot++;
    }
    done = true;
    return mine;
}

// Reference: re-reads grid parameters and inv_power (ReferenceGridForceKernels.cpp:1123-1127).
void B200CalcGridForceKernel::copyParametersToContext(ContextImpl& context, const GridForce& force) {
    build(force);
}

std::vector<double> B200CalcGridForceKernel::getParticleGroupEnergies() {
    return groupMode ? lastGroupEnergies : std::vector<double>();
}

// Per-atom energies of the most recent evaluation, in the order the particles were added to the groups; empty without
// particle groups (openmmapi/include/GridForce.h:500-508; reference CUDA platform CudaGridForceKernels.cpp:1057-1069 —
// its Reference platform returns an empty vector, ReferenceGridForceKernels.cpp:1134-1137). The kernel keeps them on
// the device; they are copied out only when asked for.
std::vector<double> B200CalcGridForceKernel::getParticleAtomEnergies() {
    std::vector<double> out;
    if (!groupMode || kernels.empty() || !evaluatedOnce) return out;
    out.resize(numGroupAtoms);
    check(gfb_kernel_get_atom_energies(kernels[0], out.data(), out.size()), "per-atom energies");
    return out;
}

}  // namespace GridForcePlugin
