# Top-level build: the CUDA library (C ABI), the host CLI that mirrors clumpy's command plugins,
# and the checkers under oracle/.  Everything is built in-tree; built files are git-ignored.
#
#   make lib      clumpy_b200/libclumpy_cuda.so   (nvcc, sm_100a only)
#   make cli      clumpy_b200/clumpy              (g++, links the library)
#   make oracle   oracle/liboracle.so (+ oracle/_ref/ when /root/reference is present)

NVCC     ?= nvcc
CXX      ?= g++
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -Iinclude \
            -Xcompiler -fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function $(EXTRA_NVFLAGS)
CSRC     := clumpy_b200/csrc
CUFILES  := $(CSRC)/context.cu $(CSRC)/simplex.cu $(CSRC)/curl.cu $(CSRC)/splat.cu $(CSRC)/splat_multi.cu $(CSRC)/age_offsets_dev.cu \
            $(CSRC)/composite_group.cu $(CSRC)/composite_exact.cu $(CSRC)/advect.cu $(CSRC)/advect_multi.cu $(CSRC)/producers.cu $(CSRC)/microbench.cu
CUOBJS   := $(CUFILES:.cu=.o) $(CSRC)/age_offsets.o
LIB      := clumpy_b200/libclumpy_cuda.so
HOST     := clumpy_b200/host
CLI      := clumpy_b200/clumpy
CLISRC   := $(HOST)/main_clumpy.cc $(HOST)/npy.cc $(wildcard $(HOST)/commands/*.cc)

all: lib cli oracle

lib: $(LIB)

$(CSRC)/%.o: $(CSRC)/%.cu $(CSRC)/internal.h $(CSRC)/device_utils.cuh $(CSRC)/rings.cuh $(CSRC)/advect_kernels.cuh include/clumpy_cuda.h
	$(NVCC) $(NVFLAGS) -c $< -o $@

# plain host code (no CUDA): g++ so that function multi-versioning (AVX2 clone of the generator) is available
$(CSRC)/%.o: $(CSRC)/%.cc $(CSRC)/internal.h
	$(CXX) -std=c++17 -O3 -fPIC -fvisibility=hidden -Wall -c $< -o $@

$(LIB): $(CUOBJS)
	$(NVCC) $(ARCH) -shared -o $@ $^

cli: $(CLI)

$(CLI): $(CLISRC) $(HOST)/clumpy_command.hh $(HOST)/npy.hh $(LIB)
	$(CXX) -std=c++17 -O2 -ffp-contract=off -Wall -Iinclude -I$(HOST) -o $@ $(CLISRC) -Lclumpy_b200 -lclumpy_cuda \
	  -Wl,-rpath,'$$ORIGIN' -lpthread

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(CUOBJS) $(LIB) $(CLI)
	$(MAKE) -C oracle clean

.PHONY: all lib cli oracle clean
