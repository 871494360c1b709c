#include "GridForceBatch.h"

#include "B200GridForceKernels.h"
#include "openmm/OpenMMException.h"

using OpenMM::OpenMMException;

namespace GridForcePlugin {

GridForceBatch::GridForceBatch(int deviceIndex, const std::string& precisionName)
    : deviceIndex(deviceIndex), precision(GFB_PRECISION_MIXED), dev(0), kernel(0) {
    if (precisionName == "double") precision = GFB_PRECISION_DOUBLE;
    else if (precisionName != "mixed") throw OpenMMException("GridForceBatch: precision must be 'mixed' or 'double'");
}

GridForceBatch::~GridForceBatch() {
    if (kernel) gfb_kernel_destroy(kernel);
}

int GridForceBatch::addForce(const GridForce& force) {
    if ((int) forces.size() >= GFB_MAX_GRIDS) throw OpenMMException("GridForceBatch: too many forces in one batch");
    forces.push_back(&force);
    if (kernel) {
        gfb_kernel_destroy(kernel);
        kernel = 0;
    }
    return (int) forces.size() - 1;
}

int GridForceBatch::getNumAtoms() const {
    if (forces.empty()) return 0;
    std::vector<int> c;
    std::vector<double> s, v, sc;
    forces[0]->getGridParameters(c, s, v, sc);
    return (int) sc.size();
}

void GridForceBatch::build() {
    if (forces.empty()) throw OpenMMException("GridForceBatch: add at least one GridForce before evaluating");
    dev = b200Device(deviceIndex);
    grids.clear();
    std::vector<gfb_grid*> handles;
    std::vector<double> scalingAll, invPower, oobK;
    size_t nAtoms = 0;
    for (size_t g = 0; g < forces.size(); g++) {
        const GridForce& f = *forces[g];
        const int layout = b200LayoutForMethod(f.getInterpolationMethod(), "GridForceBatch");
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
        grids.push_back(b200AcquireGrid(dev, deviceIndex, precision, layout, counts, spacing, origin, vals));
        handles.push_back(grids.back()->handle);
        scalingAll.insert(scalingAll.end(), scaling.begin(), scaling.end());
        invPower.push_back(f.getInvPower());
        oobK.push_back(f.getOutOfBoundsRestraint());
    }
    if (gfb_kernel_create(dev, (int) handles.size(), handles.data(), (int) nAtoms, scalingAll.data(), 0, invPower.data(),
                          oobK.data(), &kernel) != GFB_OK)
        throw OpenMMException(std::string("GridForceBatch: ") + gfb_last_error());
}

void GridForceBatch::run(const std::vector<double>& positions, int numReplicas, std::vector<double>& energies, double* forcesOut) {
    if (!kernel) build();
    const int nAtoms = getNumAtoms();
    if (numReplicas < 0 || positions.size() != (size_t) numReplicas * nAtoms * 3)
        throw OpenMMException("GridForceBatch: positions must hold numReplicas * numAtoms * 3 values");
    energies.assign(numReplicas, 0.0);
    lastGridEnergies.assign((size_t) numReplicas * forces.size(), 0.0);
    if (numReplicas == 0) return;
    if (gfb_kernel_execute_host(kernel, numReplicas, nAtoms, positions.data(), energies.data(), lastGridEnergies.data(), forcesOut,
                                GFB_FORCE_F64_STORE) != GFB_OK)
        throw OpenMMException(std::string("GridForceBatch: ") + gfb_last_error());
}

std::vector<double> GridForceBatch::evaluate(const std::vector<double>& positions, int numReplicas) {
    std::vector<double> energies;
    run(positions, numReplicas, energies, 0);
    return energies;
}

void GridForceBatch::evaluateWithForces(const std::vector<double>& positions, int numReplicas, std::vector<double>& energies,
                                        std::vector<double>& forcesOut) {
    forcesOut.assign(positions.size(), 0.0);
    run(positions, numReplicas, energies, forcesOut.empty() ? 0 : forcesOut.data());
}

}  // namespace GridForcePlugin
