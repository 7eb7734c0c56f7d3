# Builds the C-ABI library (sm_100a only) and the CPU oracle (test infrastructure).
NVCC      ?= nvcc
CC        := gcc
PKG       := motifs.jl_b200
CSRC      := $(PKG)/csrc
BUILD     ?= build
LIB       ?= $(PKG)/lib/libmotifs_b200.so
NVFLAGS   := -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -Wall
ifdef TCS_PROFILE
NVFLAGS   += -DTCS_PROFILE=1
endif
ifdef EXTRA_DEFS
NVFLAGS   += $(EXTRA_DEFS)
endif
ifdef FZ_CL
NVFLAGS   += -DFZ_CL=$(FZ_CL)
endif
ifdef FZ_PROFILE
NVFLAGS   += -DFZ_PROFILE=1
endif
CU        := $(wildcard $(CSRC)/*.cu)
HDR       := $(wildcard $(CSRC)/*.cuh) include/motifs_b200.h

all: $(LIB) oracle

OBJ       := $(patsubst $(CSRC)/%.cu,$(BUILD)/%.o,$(CU))

# one object per translation unit (make -j compiles them in parallel); NCCL is bound at run time (csrc/comm.cu), so only libdl is linked
$(BUILD)/%.o: $(CSRC)/%.cu $(HDR)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c -o $@ $<

$(LIB): $(OBJ)
	@mkdir -p $(dir $(LIB))
	$(NVCC) $(NVFLAGS) -shared -o $@ $(OBJ) -ldl

oracle: oracle/liboracle.so

oracle/liboracle.so: oracle/scan_oracle.c
	$(CC) -O3 -march=x86-64-v3 -fopenmp -fPIC -shared -Wall -o $@ $< -lm

ptxas-info:
	$(NVCC) $(NVFLAGS) -Xptxas -v -shared -o /tmp/_mb200_info.so $(CU) -ldl

clean:
	rm -rf $(LIB) oracle/liboracle.so build

.PHONY: all oracle clean ptxas-info
