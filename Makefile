# Builds librover_fe.so (C ABI, include/rover_fe.h) for sm_100a.  nvcc cross-compiles without a GPU.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
CSRC := rover_slam_b200/csrc
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Iinclude -I$(CSRC) --expt-relaxed-constexpr
OBJ := $(CSRC)/rover_fe.o $(CSRC)/sp_kernels.o $(CSRC)/lg_kernels.o $(CSRC)/tensormap.o $(CSRC)/weights.o
# `make PROBES=1` adds the tcgen05 hardware probes (csrc/probe_kernels.cu, rfe_debug_probe) that tools/gpu_probe.py drives;
# they are measurement scaffolding and stay out of the product library by default.
ifeq ($(PROBES),1)
OBJ += $(CSRC)/probe_kernels.o
NVCCFLAGS += -DRFE_ENABLE_PROBES
endif
LIB := rover_slam_b200/librover_fe.so

all: $(LIB)

$(CSRC)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/rover_fe.h
	$(NVCC) $(NVCCFLAGS) -Xptxas -v -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; exit 1)

$(CSRC)/%.o: $(CSRC)/%.cc $(wildcard $(CSRC)/*.h)
	$(NVCC) $(NVCCFLAGS) -x cu -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart

clean:
	rm -f $(OBJ) $(LIB) $(CSRC)/*.ptxas.log

# ---- host C++ class surface (SPextractor / SPmatcher / runners) on top of the C ABI -------------------------
HOST := rover_slam_b200/host
HOSTSRC := $(HOST)/src/transform.cpp $(HOST)/src/superpoint_onnx.cc $(HOST)/src/lightglue_onnx.cpp \
           $(HOST)/src/SPextractor.cc $(HOST)/src/SPmatcher_onnx.cc
HOSTFLAGS := -O2 -std=c++14 -fPIC -Wall -Iinclude -I$(HOST)/include -I$(HOST)/shim -DROVER_FE_OPENCV_SHIM -DROVER_FE_STANDALONE
HOSTLIB := rover_slam_b200/librover_slam_frontend.so
HOSTDRV := rover_slam_b200/host_driver
LATDRV := rover_slam_b200/latency_driver

host: $(HOSTLIB) $(HOSTDRV) $(LATDRV)

$(LATDRV): $(HOST)/test/latency_driver.cpp $(HOSTLIB)
	g++ $(HOSTFLAGS) -o $@ $< -Lrover_slam_b200 -lrover_slam_frontend -lrover_fe -Wl,-rpath,'$$ORIGIN'

$(HOSTLIB): $(HOSTSRC) $(wildcard $(HOST)/include/*/*.h) $(LIB)
	g++ $(HOSTFLAGS) -shared -o $@ $(HOSTSRC) -Lrover_slam_b200 -lrover_fe -Wl,-rpath,'$$ORIGIN'

$(HOSTDRV): $(HOST)/test/host_driver.cpp $(HOSTLIB)
	g++ $(HOSTFLAGS) -o $@ $< -Lrover_slam_b200 -lrover_slam_frontend -lrover_fe -Wl,-rpath,'$$ORIGIN'

all: host

# ---- debug build: every mbarrier wait is bounded and traps with a message instead of hanging (common.cuh, RFE_DEBUG_WAIT).
# Select it at run time with ROVER_FE_LIB=rover_slam_b200/librover_fe_dbg.so; used for the first GPU run of a new kernel.
DBGOBJ := $(patsubst $(CSRC)/%.o,$(CSRC)/dbg_%.o,$(OBJ))
$(CSRC)/dbg_%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/rover_fe.h
	$(NVCC) $(NVCCFLAGS) -DRFE_DEBUG_WAIT -c $< -o $@
$(CSRC)/dbg_%.o: $(CSRC)/%.cc $(wildcard $(CSRC)/*.h)
	$(NVCC) $(NVCCFLAGS) -DRFE_DEBUG_WAIT -x cu -c $< -o $@
dbg: $(DBGOBJ)
	$(NVCC) $(ARCH) -shared -o rover_slam_b200/librover_fe_dbg.so $(DBGOBJ) -lcudart

# ---- A/B build WITH the strip kernel's issuer clock reads (conv_strip.cuh, RFE_STRIP_PACE): ROVER_FE_LIB=.../librover_fe_pace.so
$(CSRC)/pace_rover_fe.o: $(CSRC)/rover_fe.cu $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/rover_fe.h
	$(NVCC) $(NVCCFLAGS) -DRFE_STRIP_PACE=1 -c $< -o $@
pace: $(CSRC)/pace_rover_fe.o $(OBJ)
	$(NVCC) $(ARCH) -shared -o rover_slam_b200/librover_fe_pace.so $(CSRC)/pace_rover_fe.o $(filter-out $(CSRC)/rover_fe.o,$(OBJ)) -lcudart

# ---- LINEAR-epilogue phase counters (umma_kernel.cuh, RFE_EPI_PROF): ROVER_FE_LIB=.../librover_fe_epiprof.so, tools/gpu_umma_prof.py
$(CSRC)/epiprof_rover_fe.o: $(CSRC)/rover_fe.cu $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/rover_fe.h
	$(NVCC) $(NVCCFLAGS) -DRFE_EPI_PROF -c $< -o $@
epiprof: $(CSRC)/epiprof_rover_fe.o $(OBJ)
	$(NVCC) $(ARCH) -shared -o rover_slam_b200/librover_fe_epiprof.so $(CSRC)/epiprof_rover_fe.o $(filter-out $(CSRC)/rover_fe.o,$(OBJ)) -lcudart
