# Top-level build: libeuler_gpu.so (CUDA, sm_100a), the host C program, the oracle.
#   make            -> everything that can be built here (nvcc cross-compiles without a GPU)
#   make gpu        -> euler_b200/lib/libeuler_gpu.so
#   make host       -> euler_b200/lib/libeuler_host.so + bin/euler-gpu
#   make oracle     -> oracle/liboracle.so (+ oracle/_ref when /root/reference is present)
NVCC     ?= /usr/local/cuda/bin/nvcc
CC       ?= gcc
ARCH     := -gencode arch=compute_100a,code=sm_100a
# -fmad=false: bit-exact parity with the reference's non-contracted fp32/fp64 arithmetic
NVFLAGS  := $(ARCH) -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -Wall \
            -Xptxas -warn-spills --expt-relaxed-constexpr --expt-extended-lambda
CSRC     := euler_b200/csrc
LIBDIR   := euler_b200/lib
OBJDIR   := build/obj
CU_SRCS  := $(CSRC)/api.cu $(CSRC)/grid_kernels.cu $(CSRC)/marker_kernels.cu \
            $(CSRC)/pcg_kernels.cu $(CSRC)/wavefront.cu $(CSRC)/comm.cu
CU_OBJS  := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(CU_SRCS))
HDRS     := $(wildcard $(CSRC)/*.cuh $(CSRC)/*.h) include/euler_gpu.h
HOST     := euler_b200/host

CXX      ?= g++
CUDA_INC ?= $(dir $(NVCC))../include

.PHONY: all gpu host oracle hostops clean
all: gpu host oracle hostops

gpu: $(LIBDIR)/libeuler_gpu.so
$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@
$(LIBDIR)/libeuler_gpu.so: $(CU_OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -cudart static -o $@ $^ -ldl

host: $(LIBDIR)/libeuler_host.so bin/euler-gpu
$(LIBDIR)/libeuler_host.so: $(HOST)/scenario.c $(HOST)/scenario.h $(HOST)/checkpoint.c $(HOST)/checkpoint.h \
                            $(HOST)/rendezvous.c $(HOST)/rendezvous.h $(HOST)/render.c $(HOST)/render.h include/euler_gpu.h
	@mkdir -p $(LIBDIR)
	$(CC) -std=gnu99 -O2 -ffp-contract=off -Wall -Wextra -fPIC -shared -Iinclude $(HOST)/scenario.c $(HOST)/checkpoint.c \
	  $(HOST)/rendezvous.c $(HOST)/render.c -lm -o $@
bin/euler-gpu: $(HOST)/main.c $(HOST)/scenario.c $(HOST)/render.c $(HOST)/checkpoint.c $(HOST)/rendezvous.c \
               $(HOST)/scenario.h $(HOST)/render.h $(HOST)/checkpoint.h $(HOST)/rendezvous.h include/euler_gpu.h
	@mkdir -p bin
	$(CC) -std=gnu99 -O2 -ffp-contract=off -Wall -Wextra -Iinclude $(HOST)/main.c $(HOST)/scenario.c $(HOST)/render.c \
	  $(HOST)/checkpoint.c $(HOST)/rendezvous.c \
	  -ldl -lm -o $@

oracle:
	$(MAKE) -C oracle all

# TEST INFRASTRUCTURE: the row operators of the PCG kernels (csrc/pcg_ops.cuh) compiled for the
# host, same no-contraction arithmetic as the oracle (tests/test_kernel_arith_host.py)
hostops: build/libkernels_host.so
HOSTOPS_SRC := tests/csrc/pcg_ops_host.cpp tests/csrc/interp_host.cpp tests/csrc/grid_ops_host.cpp
build/libkernels_host.so: $(HOSTOPS_SRC) $(CSRC)/pcg_ops.cuh $(CSRC)/pcg_pipe.cuh $(CSRC)/common.cuh $(CSRC)/interp.cuh \
                          $(CSRC)/rng.cuh $(CSRC)/marker_walk.cuh $(CSRC)/grid_ops.cuh
	@mkdir -p build
	$(CXX) -std=c++17 -O2 -ffp-contract=off -Wall -Wextra -Wno-unknown-pragmas -Wno-unused-function -fPIC -shared \
	  -I$(CUDA_INC) -I$(CSRC) $(HOSTOPS_SRC) -o $@

clean:
	rm -rf build bin $(LIBDIR)/*.so
	$(MAKE) -C oracle clean
