// See GridForce.h. Behaviour (defaults, validation messages' meaning) follows the reference's
// openmmapi/src/GridForce.cpp for the members that exist here.
#include "GridForce.h"

#include <string>

#include "GridForceKernels.h"
#include "gridforce_b200.h"
#include "internal/GridForceImpl.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"

using OpenMM::OpenMMException;

namespace GridForcePlugin {

// defaults: reference GridForce.cpp:52 (cap 41840 kJ/mol, restraint 10000 kJ/mol/nm^2, trilinear, no inv-power)
GridForce::GridForce()
    : m_origin(3, 0.0), m_invPower(0.0), m_gridCap(41840.0), m_oobK(10000.0), m_invPowerMode(InvPowerMode::NONE),
      m_interpolation(0), m_systemPtr(0) {}

void GridForce::addGridCounts(int nx, int ny, int nz) {
    m_counts.push_back(nx);
    m_counts.push_back(ny);
    m_counts.push_back(nz);
}
void GridForce::addGridSpacing(double dx, double dy, double dz) {
    m_spacing.push_back(dx);
    m_spacing.push_back(dy);
    m_spacing.push_back(dz);
}
void GridForce::addGridValue(double val) { m_vals.push_back(val); }
void GridForce::setGridValues(const std::vector<double>& vals) { m_vals = vals; }
const std::vector<double>& GridForce::getGridValues() const { return m_vals; }
void GridForce::setGridOrigin(double x, double y, double z) {
    m_origin[0] = x;
    m_origin[1] = y;
    m_origin[2] = z;
}
void GridForce::getGridOrigin(double& x, double& y, double& z) const {
    x = m_origin[0];
    y = m_origin[1];
    z = m_origin[2];
}

// V3 files go through the C ABI's reader/writer (gf_gridfile.h), which is byte-compatible with the reference's
// GridForce::saveToFile / loadFromFile (tests/test_gridfile.py).
void GridForce::loadFromFile(const std::string& filename) {
    gfb_gridfile_header h;
    if (gfb_gridfile_read_header(filename.c_str(), &h) != GFB_OK) throw OpenMMException(gfb_last_error());
    std::vector<double> vals((size_t) h.counts[0] * h.counts[1] * h.counts[2]);
    if (gfb_gridfile_read_values(filename.c_str(), vals.data(), vals.size()) != GFB_OK) throw OpenMMException(gfb_last_error());
    m_counts.assign(h.counts, h.counts + 3);
    m_spacing.assign(h.spacing, h.spacing + 3);
    m_origin.assign(h.origin, h.origin + 3);
    m_vals.swap(vals);
    m_invPower = h.inv_power;                                  // restored without transforming (GridForce.cpp:646-662)
    m_invPowerMode = static_cast<InvPowerMode>(h.inv_power_mode);
    static const char* names[] = {"", "charge", "ljr", "lja"};
    m_gridType = h.grid_type >= 1 && h.grid_type <= 3 ? names[h.grid_type] : "";
}

void GridForce::saveToFile(const std::string& filename) const {
    if (m_counts.size() != 3 || m_spacing.size() != 3) throw OpenMMException("GridForce: Grid dimensions must be set before saving");
    gfb_gridfile_header h;
    for (int k = 0; k < 3; k++) {
        h.counts[k] = m_counts[k];
        h.spacing[k] = m_spacing[k];
        h.origin[k] = m_origin[k];
    }
    h.grid_type = m_gridType == "charge" ? 1 : m_gridType == "ljr" ? 2 : m_gridType == "lja" ? 3 : 0;
    h.inv_power = m_invPower;
    h.inv_power_mode = static_cast<int>(m_invPowerMode);
    h.deriv_count = 0;
    h.data_offset = 128;
    if (gfb_gridfile_write(filename.c_str(), &h, m_vals.data(), m_vals.size(), 0) != GFB_OK) throw OpenMMException(gfb_last_error());
}

void GridForce::setReceptorPositionsFromArrays(const std::vector<double>& x, const std::vector<double>& y,
                                               const std::vector<double>& z) {
    if (x.size() != y.size() || y.size() != z.size()) throw OpenMMException("GridForce: x, y, z arrays must have the same size");
    m_receptorPositions.clear();
    m_receptorPositions.reserve(x.size());
    for (size_t i = 0; i < x.size(); i++) m_receptorPositions.push_back(OpenMM::Vec3(x[i], y[i], z[i]));
}

void GridForce::addScalingFactor(double val) { m_scaling.push_back(val); }
void GridForce::setScalingFactor(int index, double val) {
    if (index < 0 || index >= (int) m_scaling.size()) throw OpenMMException("GridForce: scaling factor index out of range");
    m_scaling[index] = val;
}
void GridForce::setScalingFactors(const std::vector<double>& vals) { m_scaling = vals; }

void GridForce::setInvPowerMode(InvPowerMode mode, double inv_power) {
    if (mode != InvPowerMode::NONE && inv_power == 0.0)
        throw OpenMMException("GridForce: inv_power must be non-zero when mode != NONE");
    if (mode == InvPowerMode::NONE && inv_power != 0.0)
        throw OpenMMException("GridForce: inv_power must be 0 when mode == NONE");
    if (!m_vals.empty()) {   // conflicting mode changes once a grid is loaded (reference GridForce.cpp:199-211)
        if (m_invPowerMode == InvPowerMode::STORED && mode == InvPowerMode::RUNTIME)
            throw OpenMMException("GridForce: Cannot set RUNTIME mode on grid that already has STORED transformation. "
                                  "This would apply transformation twice!");
        if (m_invPowerMode == InvPowerMode::RUNTIME && mode == InvPowerMode::STORED)
            throw OpenMMException("GridForce: Cannot set STORED mode on untransformed grid loaded with RUNTIME mode. "
                                  "Call applyInvPowerTransformation() first.");
    }
    m_invPowerMode = mode;
    m_invPower = inv_power;
}

// Reference GridForce.cpp:241-271 (the direct path; this stand-in has no CachedGridData), with the pow loop moved to the GPU.
void GridForce::applyInvPowerTransformation(int deviceIndex) {
    if (m_invPowerMode != InvPowerMode::RUNTIME)
        throw OpenMMException("GridForce: Can only call applyInvPowerTransformation() when mode == RUNTIME. Current mode: " +
                              std::to_string(static_cast<int>(m_invPowerMode)));
    if (m_invPower == 0.0) throw OpenMMException("GridForce: inv_power must be non-zero");
    if (m_vals.empty()) throw OpenMMException("GridForce: No grid values to transform. Load grid first.");
    gfb_device* dev = 0;
    if (gfb_device_open(deviceIndex, &dev) != GFB_OK) throw OpenMMException(gfb_last_error());
    const int rc = gfb_inv_power_transform(dev, m_vals.data(), m_vals.size(), m_invPower, 0);
    const std::string err = rc == GFB_OK ? "" : gfb_last_error();
    gfb_device_close(dev);
    if (rc != GFB_OK) throw OpenMMException(err);
    m_invPowerMode = InvPowerMode::STORED;
}
InvPowerMode GridForce::getInvPowerMode() const { return m_invPowerMode; }
double GridForce::getInvPower() const { return m_invPower; }
void GridForce::setGridCap(double uMax) { m_gridCap = uMax; }
double GridForce::getGridCap() const { return m_gridCap; }
void GridForce::setOutOfBoundsRestraint(double k) { m_oobK = k; }
double GridForce::getOutOfBoundsRestraint() const { return m_oobK; }
void GridForce::setInterpolationMethod(int method) {
    if (method < 0 || method > 3)
        throw OpenMMException("GridForce: Invalid interpolation method. Must be 0 (trilinear), 1 (cubic B-spline), 2 (tricubic), or 3 (quintic Hermite)");
    m_interpolation = method;
}
int GridForce::getInterpolationMethod() const { return m_interpolation; }

void GridForce::setLigandAtoms(const std::vector<int>& atomIndices) { m_ligandAtoms = atomIndices; }
const std::vector<int>& GridForce::getLigandAtoms() const { return m_ligandAtoms; }
void GridForce::setParticles(const std::vector<int>& particles) { m_particles = particles; }
const std::vector<int>& GridForce::getParticles() const { return m_particles; }

int GridForce::addParticleGroup(const std::string& name, const std::vector<int>& particleIndices,
                                const std::vector<double>& scalingFactors) {
    for (size_t i = 0; i < m_groups.size(); i++)
        if (m_groups[i].name == name) throw OpenMMException("Particle group '" + name + "' already exists");
    if (!scalingFactors.empty() && scalingFactors.size() != particleIndices.size())
        throw OpenMMException("Particle group '" + name + "': one scaling factor per particle is required");
    m_groups.push_back(ParticleGroup(name, particleIndices, scalingFactors));
    return (int) m_groups.size() - 1;
}
int GridForce::getNumParticleGroups() const { return (int) m_groups.size(); }
const ParticleGroup& GridForce::getParticleGroup(int index) const {
    if (index < 0 || index >= (int) m_groups.size()) throw OpenMMException("Particle group index out of range");
    return m_groups[index];
}
const ParticleGroup* GridForce::getParticleGroupByName(const std::string& name) const {
    for (size_t i = 0; i < m_groups.size(); i++)
        if (m_groups[i].name == name) return &m_groups[i];
    return 0;
}
void GridForce::removeParticleGroup(int index) {
    if (index < 0 || index >= (int) m_groups.size()) throw OpenMMException("Particle group index out of range");
    m_groups.erase(m_groups.begin() + index);
}
void GridForce::clearParticleGroups() { m_groups.clear(); }

void GridForce::getGridParameters(std::vector<int>& g_counts, std::vector<double>& g_spacing, std::vector<double>& g_vals,
                                  std::vector<double>& g_scaling_factors) const {
    g_counts = m_counts;
    g_spacing = m_spacing;
    g_vals = m_vals;
    g_scaling_factors = m_scaling;
}

OpenMM::ForceImpl* GridForce::createImpl() const { return new GridForceImpl(*this); }

void GridForce::updateParametersInContext(OpenMM::Context& context) {
    dynamic_cast<GridForceImpl&>(getImplInContext(context)).updateParametersInContext(getContextImpl(context));
}
std::vector<double> GridForce::getParticleGroupEnergies(OpenMM::Context& context) const {
    return dynamic_cast<GridForceImpl&>(getImplInContext(context)).getParticleGroupEnergies();
}
std::vector<double> GridForce::getParticleAtomEnergies(OpenMM::Context& context) const {
    return dynamic_cast<GridForceImpl&>(getImplInContext(context)).getParticleAtomEnergies();
}

// ---- GridForceImpl (reference openmmapi/src/GridForceImpl.cpp:55-86) ---------------------------------------------
void GridForceImpl::initialize(OpenMM::ContextImpl& context) {
    const_cast<GridForce&>(owner).setSystemPointer(&context.getSystem());
    kernel = context.getPlatform().createKernel(CalcGridForceKernel::Name(), context);
    kernel.getAs<CalcGridForceKernel>().initialize(context.getSystem(), owner);
}
double GridForceImpl::calcForcesAndEnergy(OpenMM::ContextImpl& context, bool includeForces, bool includeEnergy, int groups) {
    if ((groups & (1 << owner.getForceGroup())) != 0)
        return kernel.getAs<CalcGridForceKernel>().execute(context, includeForces, includeEnergy);
    return 0.0;
}
std::vector<std::string> GridForceImpl::getKernelNames() { return std::vector<std::string>(1, CalcGridForceKernel::Name()); }
void GridForceImpl::updateParametersInContext(OpenMM::ContextImpl& context) {
    kernel.getAs<CalcGridForceKernel>().copyParametersToContext(context, owner);
}
std::vector<double> GridForceImpl::getParticleGroupEnergies() { return kernel.getAs<CalcGridForceKernel>().getParticleGroupEnergies(); }
std::vector<double> GridForceImpl::getParticleAtomEnergies() { return kernel.getAs<CalcGridForceKernel>().getParticleAtomEnergies(); }

}  // namespace GridForcePlugin
