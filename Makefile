# Builds everything in-tree:
#   openmmgridforce_b200/lib/libgridforce_b200.so   — CUDA kernels + C ABI (sm_100a only)
#   openmmgridforce_b200/lib/libOpenMMGridForceB200.so — OpenMM platform plugin (see openmmgridforce_b200/plugin)
#   oracle/…                                        — test-only parity oracles (see oracle/Makefile)
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Iinclude -Iopenmmgridforce_b200/csrc
LIBDIR    := openmmgridforce_b200/lib
CSRC      := openmmgridforce_b200/csrc

PLUGIN    := openmmgridforce_b200/plugin
# OPENMM_INCLUDE: point it at a real OpenMM install's include dir to build against OpenMM itself; default = the shim.
OPENMM_INCLUDE ?= third_party/openmm_shim
PLUGIN_SRCS := $(PLUGIN)/openmmapi/GridForce.cpp $(PLUGIN)/platform/B200GridForceKernels.cpp \
               $(PLUGIN)/platform/B200GridForceKernelFactory.cpp $(PLUGIN)/platform/GridForceBatch.cpp $(PLUGIN)/plugin_driver.cpp
PLUGIN_HDRS := $(wildcard $(PLUGIN)/openmmapi/*.h $(PLUGIN)/openmmapi/internal/*.h $(PLUGIN)/platform/*.h)

all: lib plugin oracle

plugin: $(LIBDIR)/libOpenMMGridForceB200.so

$(LIBDIR)/libOpenMMGridForceB200.so: $(PLUGIN_SRCS) $(PLUGIN_HDRS) $(LIBDIR)/libgridforce_b200.so include/gridforce_b200.h
	g++ -std=c++11 -O2 -fPIC -shared -Wall -Wno-unused-parameter -I$(OPENMM_INCLUDE) -Iinclude -I$(PLUGIN)/openmmapi -I$(PLUGIN)/platform \
	    -o $@ $(PLUGIN_SRCS) -L$(LIBDIR) -lgridforce_b200 -Wl,-rpath,'$$ORIGIN' -Wl,-Bsymbolic -lpthread

lib: $(LIBDIR)/libgridforce_b200.so

$(LIBDIR)/libgridforce_b200.so: $(CSRC)/gf_capi.cu $(CSRC)/gf_kernels.cuh $(CSRC)/gf_eval_lines.cuh $(CSRC)/gf_gridfile.h $(CSRC)/gf_params.h include/gridforce_b200.h
	mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/gf_capi.cu

ptxas-info:
	$(NVCC) $(NVFLAGS) -Xptxas -v -cubin -o /tmp/gf_capi.cubin $(CSRC)/gf_capi.cu

oracle:
	$(MAKE) -C oracle all

clean:
	rm -f $(LIBDIR)/*.so
	$(MAKE) -C oracle clean

.PHONY: all lib plugin oracle clean ptxas-info
