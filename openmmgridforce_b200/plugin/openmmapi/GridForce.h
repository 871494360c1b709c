// In-repo stand-in for the reference's public Force class (openmmapi/include/GridForce.h), restricted to what the
// evaluation path reads. Method names, argument meaning and error behaviour follow the reference so that
// (a) a script written against the reference API runs unchanged for this path, and
// (b) platform/B200GridForceKernels.cpp compiles against either this header or the reference's own.
// Members of the reference API that feed other subsystems (GridData/CachedGridData sharing, tiled streaming,
// derivative grids) are declared only as far as the kernel must be
// able to ask "is this on?" and refuse; they are out of scope of this repository.
#ifndef B200_GRIDFORCE_H_
#define B200_GRIDFORCE_H_

#include <string>
#include <vector>

#include "GridForceTypes.h"
#include "openmm/Context.h"
#include "openmm/Force.h"
#include "openmm/Vec3.h"

namespace GridForcePlugin {

// reference: GridForce.h:56-77
struct ParticleGroup {
    ParticleGroup(const std::string& name, const std::vector<int>& particleIndices,
                  const std::vector<double>& scalingFactors = std::vector<double>())
        : name(name), particleIndices(particleIndices), scalingFactors(scalingFactors) {
        if (this->scalingFactors.empty()) this->scalingFactors.assign(particleIndices.size(), 1.0);
    }
    std::string name;
    std::vector<int> particleIndices;
    std::vector<double> scalingFactors;
};

class GridForce : public OpenMM::Force {
public:
    GridForce();

    // ---- grid definition (reference GridForce.cpp:153-181) --------------------------------------------------
    void addGridCounts(int nx, int ny, int nz);
    void addGridSpacing(double dx, double dy, double dz);   // nm
    void addGridValue(double val);                          // x-major, z fastest
    void setGridValues(const std::vector<double>& vals);    // bulk form of addGridValue
    const std::vector<double>& getGridValues() const;
    void setGridOrigin(double x, double y, double z);
    void getGridOrigin(double& x, double& y, double& z) const;

    // ---- V3 "OMGRID" files (reference GridForce.cpp:495-799): values, geometry, origin, type, inv-power state ----
    void loadFromFile(const std::string& filename);
    void saveToFile(const std::string& filename) const;
    void setGridType(const std::string& type) { m_gridType = type; }        // "", "charge", "ljr", "lja"
    const std::string& getGridType() const { return m_gridType; }

    // ---- per-atom scaling factors ------------------------------------------------------------------------------
    void addScalingFactor(double val);
    void setScalingFactor(int index, double val);
    void setScalingFactors(const std::vector<double>& vals);

    // ---- evaluation options --------------------------------------------------------------------------------------
    void setInvPowerMode(InvPowerMode mode, double inv_power);
    InvPowerMode getInvPowerMode() const;
    // RUNTIME mode: G -> sign(G)|G|^(1/n) once, after which the mode is STORED (reference GridForce.cpp:221-272).
    // Computed on the GPU (gfb_inv_power_transform); `deviceIndex` picks which one.
    void applyInvPowerTransformation(int deviceIndex = 0);
    double getInvPower() const;
    void setGridCap(double uMax);
    double getGridCap() const;
    void setOutOfBoundsRestraint(double k);                 // kJ/mol/nm^2, default 10000
    double getOutOfBoundsRestraint() const;
    void setInterpolationMethod(int method);                // 0 trilinear, 1 cubic B-spline, 2 tricubic run on this platform; 3 throws at Context creation
    int getInterpolationMethod() const;

    // ---- which particles ------------------------------------------------------------------------------------------
    void setLigandAtoms(const std::vector<int>& atomIndices);
    const std::vector<int>& getLigandAtoms() const;
    void setParticles(const std::vector<int>& particles);
    const std::vector<int>& getParticles() const;
    int addParticleGroup(const std::string& name, const std::vector<int>& particleIndices,
                         const std::vector<double>& scalingFactors = std::vector<double>());
    int getNumParticleGroups() const;
    const ParticleGroup& getParticleGroup(int index) const;
    const ParticleGroup* getParticleGroupByName(const std::string& name) const;
    void removeParticleGroup(int index);
    void clearParticleGroups();
    std::vector<double> getParticleGroupEnergies(OpenMM::Context& context) const;
    std::vector<double> getParticleAtomEnergies(OpenMM::Context& context) const;

    // ---- derive inputs from the System's NonbondedForce at Context creation (reference GridForce.h:171-198, 335-342,
    //      523-573; ReferenceGridForceKernels.cpp:163-278). The grid is generated on the GPU (gfb_grid_generate). -----
    void setAutoCalculateScalingFactors(bool enable) { m_autoScaling = enable; }
    bool getAutoCalculateScalingFactors() const { return m_autoScaling; }
    void setScalingProperty(const std::string& property) { m_scalingProperty = property; }   // validated by the kernel
    const std::string& getScalingProperty() const { return m_scalingProperty; }
    void setAutoGenerateGrid(bool enable) { m_autoGenerate = enable; }
    bool getAutoGenerateGrid() const { return m_autoGenerate; }
    void setReceptorAtoms(const std::vector<int>& atomIndices) { m_receptorAtoms = atomIndices; }
    const std::vector<int>& getReceptorAtoms() const { return m_receptorAtoms; }
    void setReceptorPositions(const std::vector<OpenMM::Vec3>& positions) { m_receptorPositions = positions; }
    void setReceptorPositionsFromArrays(const std::vector<double>& x, const std::vector<double>& y, const std::vector<double>& z);
    const std::vector<OpenMM::Vec3>& getReceptorPositions() const { return m_receptorPositions; }

    // ---- switches of subsystems outside this path: always off here; the kernel checks them -------------------------
    bool getTiledMode() const { return false; }
    bool hasDerivatives() const { return false; }

    // ---- what the kernel pulls in initialize() (reference GridForce.cpp:355-363) -------------------------------------
    void getGridParameters(std::vector<int>& g_counts, std::vector<double>& g_spacing, std::vector<double>& g_vals,
                           std::vector<double>& g_scaling_factors) const;
    void updateParametersInContext(OpenMM::Context& context);

    void setSystemPointer(const void* systemPtr) { m_systemPtr = systemPtr; }
    const void* getSystemPointer() const { return m_systemPtr; }

protected:
    OpenMM::ForceImpl* createImpl() const;

private:
    std::vector<int> m_counts;
    std::vector<double> m_spacing, m_vals, m_scaling, m_origin;
    std::vector<int> m_ligandAtoms, m_particles;
    std::vector<ParticleGroup> m_groups;
    double m_invPower, m_gridCap, m_oobK;
    InvPowerMode m_invPowerMode;
    int m_interpolation;
    const void* m_systemPtr;
    std::string m_gridType, m_scalingProperty;
    bool m_autoScaling = false, m_autoGenerate = false;
    std::vector<int> m_receptorAtoms;
    std::vector<OpenMM::Vec3> m_receptorPositions;
};

}  // namespace GridForcePlugin
#endif
