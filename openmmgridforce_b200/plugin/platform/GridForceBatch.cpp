#include "GridForceBatch.h"

#include "B200GridForceKernels.h"
#include "openmm/OpenMMException.h"

using OpenMM::OpenMMException;

namespace GridForcePlugin {

void GridForceBatch::setPrecision(const std::string& name) {
    precision = GFB_PRECISION_MIXED;
    if (name == "double") precision = GFB_PRECISION_DOUBLE;
    else if (name != "mixed") throw OpenMMException("GridForceBatch: precision must be 'mixed' or 'double'");
}

GridForceBatch::GridForceBatch(int deviceIndex, const std::string& precisionName)
    : devices(1, deviceIndex), dev(0), kernel(0), multi(0), numAtoms(0), built(false) {
    setPrecision(precisionName);
}

GridForceBatch::GridForceBatch(const std::vector<int>& deviceIndices, const std::string& precisionName)
    : devices(deviceIndices), dev(0), kernel(0), multi(0), numAtoms(0), built(false) {
    setPrecision(precisionName);
    if (devices.empty()) throw OpenMMException("GridForceBatch: the device list is empty");
    if ((int) devices.size() > GFB_MAX_PEERS) throw OpenMMException("GridForceBatch: too many devices");
}

GridForceBatch::~GridForceBatch() { release(); }

void GridForceBatch::release() {
    if (kernel) gfb_kernel_destroy(kernel);
    if (multi) gfb_multi_destroy(multi);
    kernel = 0;
    multi = 0;
    grids.clear();
    built = false;
}

int GridForceBatch::addForce(const GridForce& force) {
    if ((int) forces.size() >= GFB_MAX_GRIDS) throw OpenMMException("GridForceBatch: too many forces in one batch");
    forces.push_back(&force);
    release();
    return (int) forces.size() - 1;
}

int GridForceBatch::getNumAtoms() const {
    if (built) return numAtoms;
    if (forces.empty()) return 0;
    std::vector<int> c;
    std::vector<double> s, v, sc;
    forces[0]->getGridParameters(c, s, v, sc);      // before the first evaluation only: copies the grid to read one size
    return (int) sc.size();
}

void GridForceBatch::build() {
    if (forces.empty()) throw OpenMMException("GridForceBatch: add at least one GridForce before evaluating");
    release();
    const bool several = devices.size() > 1;
    if (several) {
        if (gfb_multi_create((int) devices.size(), devices.data(), &multi) != GFB_OK)
            throw OpenMMException(std::string("GridForceBatch: ") + gfb_last_error());
    } else {
        dev = b200Device(devices[0]);
    }
    std::vector<gfb_grid*> handles;
    std::vector<double> scalingAll, invPower, oobK;
    size_t nAtoms = 0;
    for (size_t g = 0; g < forces.size(); g++) {
        const GridForce& f = *forces[g];
        const int layout = b200LayoutForMethod(f.getInterpolationMethod(), precision, "GridForceBatch");
        if (f.getInterpolationMethod() != forces[0]->getInterpolationMethod())
            throw OpenMMException("GridForceBatch: all forces must use the same interpolation method");
        std::vector<int> counts;
        std::vector<double> spacing, vals, scaling;
        f.getGridParameters(counts, spacing, vals, scaling);
        if (counts.size() != 3 || spacing.size() != 3 || vals.size() != (size_t) counts[0] * counts[1] * counts[2])
            throw OpenMMException("GridForceBatch: force has an incomplete grid definition");
        if (g == 0) nAtoms = scaling.size();
        if (scaling.size() != nAtoms)
            throw OpenMMException("GridForceBatch: all forces must have the same number of scaling factors");
        double origin[3];
        f.getGridOrigin(origin[0], origin[1], origin[2]);
        if (several) {
            if (gfb_multi_add_grid(multi, counts.data(), spacing.data(), origin, vals.data(), vals.size(), precision, layout) < 0)
                throw OpenMMException(std::string("GridForceBatch: ") + gfb_last_error());
        } else {
            grids.push_back(b200AcquireGrid(dev, devices[0], precision, layout, counts, spacing, origin, vals));
            handles.push_back(grids.back()->handle);
        }
        scalingAll.insert(scalingAll.end(), scaling.begin(), scaling.end());
        invPower.push_back(f.getInvPower());
        oobK.push_back(f.getOutOfBoundsRestraint());
    }
    const int rc = several ? gfb_multi_build(multi, (int) nAtoms, scalingAll.data(), invPower.data(), oobK.data())
                           : gfb_kernel_create(dev, (int) handles.size(), handles.data(), (int) nAtoms, scalingAll.data(), 0,
                                               invPower.data(), oobK.data(), &kernel);
    if (rc != GFB_OK) throw OpenMMException(std::string("GridForceBatch: ") + gfb_last_error());
    numAtoms = (int) nAtoms;
    built = true;
}

void GridForceBatch::run(const double* positions, size_t nPositions, int numReplicas, double* energies, void* forcesOut,
                         int forceMode, bool wantGridEnergies) {
    if (!built) build();
    if (numReplicas < 0 || nPositions != (size_t) numReplicas * numAtoms * 3)
        throw OpenMMException("GridForceBatch: positions must hold numReplicas * numAtoms * 3 values");
    if (numReplicas == 0) return;
    if (!positions || !energies) throw OpenMMException("GridForceBatch: NULL positions or energies buffer");
    int rc;
    if (multi) {
        lastGridEnergies.clear();
        rc = gfb_multi_execute_host(multi, numReplicas, positions, energies, forcesOut, forceMode);
    } else {
        double* ge = 0;
        if (wantGridEnergies) {
            lastGridEnergies.resize((size_t) numReplicas * forces.size());
            ge = lastGridEnergies.data();
        }
        rc = gfb_kernel_execute_host(kernel, numReplicas, numAtoms, positions, energies, ge, forcesOut, forceMode);
    }
    if (rc != GFB_OK) throw OpenMMException(std::string("GridForceBatch: ") + gfb_last_error());
}

std::vector<double> GridForceBatch::evaluate(const std::vector<double>& positions, int numReplicas) {
    std::vector<double> energies(numReplicas > 0 ? numReplicas : 0);
    run(positions.data(), positions.size(), numReplicas, energies.data(), 0, GFB_FORCE_F64_STORE, true);
    return energies;
}

void GridForceBatch::evaluateWithForces(const std::vector<double>& positions, int numReplicas, std::vector<double>& energies,
                                        std::vector<double>& forcesOut) {
    energies.resize(numReplicas > 0 ? numReplicas : 0);     // resize, not assign: every entry is overwritten by the kernels
    forcesOut.resize(positions.size());
    run(positions.data(), positions.size(), numReplicas, energies.data(), forcesOut.empty() ? 0 : forcesOut.data(),
        GFB_FORCE_F64_STORE, true);
}

void GridForceBatch::evaluate(const double* positions, int numReplicas, double* energies) {
    if (!built) build();
    run(positions, (size_t) (numReplicas > 0 ? numReplicas : 0) * numAtoms * 3, numReplicas, energies, 0, GFB_FORCE_F64_STORE, false);
}

void GridForceBatch::evaluateWithForces(const double* positions, int numReplicas, double* energies, double* forcesOut) {
    if (!built) build();
    run(positions, (size_t) (numReplicas > 0 ? numReplicas : 0) * numAtoms * 3, numReplicas, energies, forcesOut, GFB_FORCE_F64_STORE, false);
}

void GridForceBatch::evaluateWithForcesF32(const double* positions, int numReplicas, double* energies, float* forcesOut) {
    if (!built) build();
    run(positions, (size_t) (numReplicas > 0 ? numReplicas : 0) * numAtoms * 3, numReplicas, energies, forcesOut, GFB_FORCE_F32_STORE, false);
}

void GridForceBatch::pinBuffer(void* ptr, size_t bytes) {
    if (gfb_host_register(ptr, bytes) != GFB_OK) throw OpenMMException(std::string("GridForceBatch: ") + gfb_last_error());
}

void GridForceBatch::unpinBuffer(void* ptr) {
    if (gfb_host_unregister(ptr) != GFB_OK) throw OpenMMException(std::string("GridForceBatch: ") + gfb_last_error());
}

}  // namespace GridForcePlugin
