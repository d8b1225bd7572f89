# Builds everything in-tree:
#   mmannot_b200/lib/libmmannot_host.so   host front-end (config / GTF / SAM+BAM decode), g++
#   mmannot_b200/lib/libmmannot_b200.so   CUDA hot path + C ABI, nvcc sm_100a
#   mmannot_b200/bin/mmannot_b200         drop-in CLI
#   oracle/_build/liboracle.so            CPU restatement (test infrastructure only)
#   oracle/_ref/*                         the reference itself, compiled where it lies (test infrastructure only)
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
CXXFLAGS  := -O3 -std=c++17 -fPIC -Wall -Wextra -Iinclude
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Iinclude

HOST_SRC  := $(wildcard mmannot_b200/csrc/host/*.cpp)
CLI_SRC   := mmannot_b200/csrc/host/main.cpp mmannot_b200/csrc/host/counter.cpp mmannot_b200/csrc/host/stats_writers.cpp mmannot_b200/csrc/host/bam_device.cpp
HOST_LIB_SRC := $(filter-out $(CLI_SRC),$(HOST_SRC))
HOST_HDR  := $(wildcard mmannot_b200/csrc/host/*.hpp) $(wildcard include/*.h)
CU_SRC    := $(wildcard mmannot_b200/csrc/*.cu)
CU_HDR    := $(wildcard mmannot_b200/csrc/*.cuh) $(wildcard include/*.h)

all: host cuda cli oracle

host: mmannot_b200/lib/libmmannot_host.so
cuda: mmannot_b200/lib/libmmannot_b200.so
cli: mmannot_b200/bin/mmannot_b200
oracle: oracle/_build/liboracle.so ref

mmannot_b200/lib/libmmannot_host.so: $(HOST_LIB_SRC) $(HOST_HDR)
	@mkdir -p mmannot_b200/lib
	$(CXX) $(CXXFLAGS) -shared -o $@ $(HOST_LIB_SRC) -lz -pthread

mmannot_b200/lib/libmmannot_b200.so: $(CU_SRC) $(CU_HDR)
	@mkdir -p mmannot_b200/lib
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CU_SRC)

mmannot_b200/bin/mmannot_b200: $(CLI_SRC) $(HOST_HDR) mmannot_b200/lib/libmmannot_host.so mmannot_b200/lib/libmmannot_b200.so
	@mkdir -p mmannot_b200/bin
	$(CXX) $(CXXFLAGS) -o $@ $(CLI_SRC) -Lmmannot_b200/lib -lmmannot_host -lmmannot_b200 -lz -pthread -Wl,-rpath,'$$ORIGIN/../lib'

oracle/_build/liboracle.so: oracle/oracle.c oracle/oracle.h
	@mkdir -p oracle/_build
	$(CC) -O2 -std=c99 -fPIC -Wall -shared -o $@ oracle/oracle.c -lm

ref:
	bash oracle/build_ref.sh

# tuning builds (not shipped): make variants V="t128b6:-DMMA_FAST_THREADS=128,-DMMA_FAST_BLOCKS_PER_SM=6 ..." ;
# MMANNOT_B200_LIB=mmannot_b200/lib/variants/t128b6.so python bench.py ...
variants: $(CU_SRC) $(CU_HDR)
	@mkdir -p mmannot_b200/lib/variants
	for v in $(V); do name=$${v%%:*}; defs=$$(echo $${v#*:} | tr ',' ' '); $(NVCC) $(NVFLAGS) $$defs -shared -o mmannot_b200/lib/variants/$$name.so mmannot_b200/csrc/mma_api.cu & done; wait

clean:
	rm -rf mmannot_b200/lib mmannot_b200/bin oracle/_build

.PHONY: all host cuda cli oracle ref clean variants
