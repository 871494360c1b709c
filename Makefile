# Builds everything in-tree:
#   openmmgridforce_b200/lib/libgridforce_b200.so   — CUDA kernels + C ABI (sm_100a only)
#   openmmgridforce_b200/lib/libOpenMMGridForceB200.so — OpenMM platform plugin (see openmmgridforce_b200/plugin)
#   oracle/…                                        — test-only parity oracles (see oracle/Makefile)
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
# EXTRA: additional nvcc flags for A/B builds, e.g. make lib EXTRA=-DGFB_LINES_BLOCK_MULTI=64 OBJDIR=build/obj_b64 LIBOUT=ab/libgf_b64.so (ab/ travels with gpurun; GFB_LIB_PATH selects it)
EXTRA     ?=
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Iinclude -Iopenmmgridforce_b200/csrc $(EXTRA)
LIBDIR    := openmmgridforce_b200/lib
CSRC      := openmmgridforce_b200/csrc

OBJDIR    ?= build/obj
LIBOUT    ?= $(LIBDIR)/libgridforce_b200.so

PLUGIN    := openmmgridforce_b200/plugin
# OPENMM_INCLUDE: point it at a real OpenMM install's include dir to build against OpenMM itself; default = the shim.
OPENMM_INCLUDE ?= third_party/openmm_shim
PLUGIN_SRCS := $(PLUGIN)/openmmapi/GridForce.cpp $(PLUGIN)/platform/B200GridForceKernels.cpp \
               $(PLUGIN)/platform/B200GridForceKernelFactory.cpp $(PLUGIN)/platform/GridForceBatch.cpp $(PLUGIN)/plugin_driver.cpp
PLUGIN_HDRS := $(wildcard $(PLUGIN)/openmmapi/*.h $(PLUGIN)/openmmapi/internal/*.h $(PLUGIN)/platform/*.h third_party/openmm_shim/openmm/*.h third_party/openmm_shim/openmm/*/*.h)

all: lib plugin oracle

plugin: $(LIBDIR)/libOpenMMGridForceB200.so

$(LIBDIR)/libOpenMMGridForceB200.so: $(PLUGIN_SRCS) $(PLUGIN_HDRS) $(LIBDIR)/libgridforce_b200.so include/gridforce_b200.h
	g++ -std=c++11 -O2 -fPIC -shared -Wall -Wno-unused-parameter -I$(OPENMM_INCLUDE) -Iinclude -I$(PLUGIN)/openmmapi -I$(PLUGIN)/platform \
	    -o $@ $(PLUGIN_SRCS) -L$(LIBDIR) -lgridforce_b200 -Wl,-rpath,'$$ORIGIN' -Wl,-Bsymbolic -lpthread

lib: $(LIBOUT)

# One object per translation unit so that `make -j` builds the kernel families in parallel (the single-file build took
# 90 s). Every header of csrc/ is a prerequisite of every object: an edit anywhere rebuilds, a stale .so cannot happen.
CUDA_HDRS := $(wildcard $(CSRC)/*.cuh $(CSRC)/*.h) include/gridforce_b200.h
CUDA_OBJS := $(addprefix $(OBJDIR)/, gf_capi.o gf_grids.o gf_aux.o gf_multi.o gf_resident.o gf_launch_general_f32.o gf_launch_general_f64.o \
               gf_launch_lines_1.o gf_launch_lines_2.o gf_launch_lines_3.o gf_launch_lines_4.o gf_launch_records_f64.o gf_launch_bspline.o)

$(OBJDIR)/gf_launch_lines_%.o: $(CSRC)/gf_launch_lines.cu $(CUDA_HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -DGFB_LINES_NG=$* -c -o $@ $<

$(OBJDIR)/%.o: $(CSRC)/%.cu $(CUDA_HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c -o $@ $<

$(LIBOUT): $(CUDA_OBJS)
	@mkdir -p $(dir $(LIBOUT))
	$(NVCC) $(ARCH) -shared -o $@ $(CUDA_OBJS) -ldl

# The platform sources compiled against the REFERENCE's own openmmapi headers (GridForce.h, GridForceKernels.h,
# internal/GridForceImpl.h) instead of the in-repo stand-ins: proves the plugin is source-compatible with the interface it
# replaces. OpenMM itself is modelled by the shim (real OpenMM is not installable here). Needs REFERENCE_DIR.
REFERENCE_DIR ?= /root/reference
plugin-check-reference:
	@test -d $(REFERENCE_DIR)/openmmapi/include || (echo "no reference tree at $(REFERENCE_DIR)"; exit 1)
	for f in $(PLUGIN)/platform/B200GridForceKernels.cpp $(PLUGIN)/platform/B200GridForceKernelFactory.cpp $(PLUGIN)/platform/GridForceBatch.cpp; do \
	    g++ -std=c++11 -fsyntax-only -Wall -Wno-unused-parameter -I$(REFERENCE_DIR)/openmmapi/include -I$(OPENMM_INCLUDE) -Iinclude \
	        -I$(PLUGIN)/platform $$f || exit 1; done
	@echo "plugin sources compile against $(REFERENCE_DIR)/openmmapi/include"

ptxas-info:
	$(NVCC) $(NVFLAGS) -Xptxas -v -DGFB_LINES_NG=3 -cubin -o /tmp/gf_lines3.cubin $(CSRC)/gf_launch_lines.cu
	$(NVCC) $(NVFLAGS) -Xptxas -v -cubin -o /tmp/gf_lines_f64.cubin $(CSRC)/gf_launch_records_f64.cu

oracle:
	$(MAKE) -C oracle all

clean:
	rm -f $(LIBDIR)/*.so
	rm -rf $(OBJDIR)
	$(MAKE) -C oracle clean

.PHONY: all lib plugin oracle clean ptxas-info plugin-check-reference
