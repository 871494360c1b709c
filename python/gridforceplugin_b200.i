/* SWIG interface additions for the B200 platform — meant to be %include'd at the end of the reference's
 * python/gridforceplugin.i (after its GridForce declaration, :158-282), so that the existing module `gridforceplugin`
 * keeps every declaration it has and gains the batched multi-replica entry point. The std::vector overloads use only
 * typemaps the reference module already instantiates (vectord, vectori: gridforceplugin.i:16-28) and its %exception
 * block (:49-59), which maps OpenMMException to RuntimeError. The buffer overloads take any object with the Python buffer
 * protocol (numpy arrays) through the typemaps below — self-contained, no numpy.i needed.
 *
 * Cannot be built in this repository's container (no SWIG, no OpenMM); the same C++ classes are exercised from Python by
 * tests/test_plugin.py through openmmgridforce_b200/gridforceplugin.py (ctypes).
 *
 * Nothing needs wrapping for the platform itself: OpenMM loads lib/plugins/libOpenMMGridForceB200.so, calls
 * registerPlatforms()/registerKernelFactories(), and scripts select it with Platform.getPlatformByName("B200");
 * platform.setPropertyDefaultValue("Precision", "double") and Context(system, integrator, platform, {"DeviceIndex": "1"})
 * go through OpenMM's own Platform wrapper.
 */
%{
#include "GridForceBatch.h"

/* A C-contiguous buffer of `itemsize`-byte items -> pointer + item count; *view must be released by the caller. */
static int gfb_get_buffer(PyObject* obj, Py_buffer* view, int writable, Py_ssize_t itemsize, const char* what) {
    if (PyObject_GetBuffer(obj, view, (writable ? PyBUF_WRITABLE : 0) | PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) != 0) {
        PyErr_Format(PyExc_TypeError, "%s: expected a C-contiguous %s buffer (e.g. a numpy array)", what, writable ? "writable" : "readable");
        return 0;
    }
    if (view->itemsize != itemsize) {
        PyBuffer_Release(view);
        PyErr_Format(PyExc_TypeError, "%s: items must be %d bytes wide", what, (int) itemsize);
        return 0;
    }
    return 1;
}
%}

/* (const double* positions, int numReplicas): positions buffer of float64; numReplicas stays a separate argument */
%typemap(in) const double* positions (Py_buffer view) {
    if (!gfb_get_buffer($input, &view, 0, sizeof(double), "positions")) SWIG_fail;
    $1 = (double*) view.buf;
}
%typemap(freearg) const double* positions { PyBuffer_Release(&view$argnum); }
%typemap(in) double* energies (Py_buffer view) {
    if (!gfb_get_buffer($input, &view, 1, sizeof(double), "energies")) SWIG_fail;
    $1 = (double*) view.buf;
}
%typemap(freearg) double* energies { PyBuffer_Release(&view$argnum); }
%typemap(in) double* forcesOut (Py_buffer view) {
    if (!gfb_get_buffer($input, &view, 1, sizeof(double), "forcesOut")) SWIG_fail;
    $1 = (double*) view.buf;
}
%typemap(freearg) double* forcesOut { PyBuffer_Release(&view$argnum); }
%typemap(in) float* forcesOut (Py_buffer view) {
    if (!gfb_get_buffer($input, &view, 1, sizeof(float), "forcesOut")) SWIG_fail;
    $1 = (float*) view.buf;
}
%typemap(freearg) float* forcesOut { PyBuffer_Release(&view$argnum); }
/* pinBuffer(buffer): pointer + byte count from one Python object */
%typemap(in) (void* ptr, size_t bytes) (Py_buffer view) {
    if (PyObject_GetBuffer($input, &view, PyBUF_C_CONTIGUOUS) != 0) SWIG_fail;
    $1 = view.buf;
    $2 = (size_t) view.len;
}
%typemap(freearg) (void* ptr, size_t bytes) { PyBuffer_Release(&view$argnum); }
%typemap(in) void* ptr (Py_buffer view) {
    if (PyObject_GetBuffer($input, &view, PyBUF_C_CONTIGUOUS) != 0) SWIG_fail;
    $1 = view.buf;
}
%typemap(freearg) void* ptr { PyBuffer_Release(&view$argnum); }

namespace GridForcePlugin {

class GridForceBatch {
public:
    GridForceBatch(int deviceIndex = 0, const std::string& precision = "mixed");
    GridForceBatch(const std::vector<int>& deviceIndices, const std::string& precision = "mixed");
    ~GridForceBatch();

    int addForce(const GridForce& force);
    int getNumForces() const;
    int getNumAtoms() const;
    int getNumDevices() const;

    /* positions: flat [numReplicas][numAtoms][3] in nm -> one energy (kJ/mol) per replica (energy-only evaluation) */
    std::vector<double> evaluate(const std::vector<double>& positions, int numReplicas);

    /* Python: energies, forces = batch.evaluateWithForces(positions, numReplicas) */
    %apply std::vector<double>& OUTPUT { std::vector<double>& energies };
    %apply std::vector<double>& OUTPUT { std::vector<double>& forcesOut };
    void evaluateWithForces(const std::vector<double>& positions, int numReplicas,
                            std::vector<double>& energies, std::vector<double>& forcesOut);
    %clear std::vector<double>& energies;
    %clear std::vector<double>& forcesOut;

    /* numpy buffers, no copies: batch.evaluateWithForces(pos, R, energies, forces) with
     * pos float64 [R, A, 3], energies float64 [R], forces float64 (or float32 for ...F32) [R, A, 3] */
    void evaluate(const double* positions, int numReplicas, double* energies);
    void evaluateWithForces(const double* positions, int numReplicas, double* energies, double* forcesOut);
    void evaluateWithForcesF32(const double* positions, int numReplicas, double* energies, float* forcesOut);

    /* GridForceBatch.pinBuffer(array): page-lock once, DMA directly afterwards */
    static void pinBuffer(void* ptr, size_t bytes);
    static void unpinBuffer(void* ptr);

    std::vector<double> getLastGridEnergies() const;
};

}  // namespace GridForcePlugin
