# Makefile -- same targets as the reference's (all, check, clean; Makefile:14-24 there)
# plus the CUDA library.  `make` builds liblbm_b200.so (nvcc, sm_100a) and the C host
# program d2q9-bgk, which keeps the reference's command line and file contract.

EXE=d2q9-bgk
PKG=advanced-hpc-lbm_b200
LIB=$(PKG)/liblbm_b200.so

CC=gcc
CFLAGS=-std=c99 -Wall -O2
NVCC=nvcc
NVCCFLAGS=-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
LIBS=-lm

FINAL_STATE_FILE=./final_state.dat
AV_VELS_FILE=./av_vels.dat
REF_FINAL_STATE_FILE=check/128x128.final_state.dat
REF_AV_VELS_FILE=check/128x128.av_vels.dat

all: $(LIB) $(EXE)

$(LIB): $(PKG)/csrc/lbm_gpu.cu $(PKG)/csrc/lbm_kernels.cuh include/lbm_gpu.h
	$(NVCC) $(NVCCFLAGS) -shared -Iinclude $(PKG)/csrc/lbm_gpu.cu -o $@

$(EXE): $(PKG)/host/d2q9-bgk.c $(PKG)/host/lbm_io.c $(PKG)/host/lbm_io.h include/lbm_gpu.h $(LIB)
	$(CC) $(CFLAGS) -Iinclude -I$(PKG)/host $(PKG)/host/d2q9-bgk.c $(PKG)/host/lbm_io.c \
	    -L$(PKG) -llbm_b200 -Wl,-rpath,'$$ORIGIN/$(PKG)' $(LIBS) -o $@

# golden text files are expanded from the compact fixtures in tests/golden/
check/%.av_vels.dat check/%.final_state.dat:
	python tests/golden/expand_golden.py check

check: $(REF_AV_VELS_FILE) $(REF_FINAL_STATE_FILE)
	python check/check.py --ref-av-vels-file=$(REF_AV_VELS_FILE) --ref-final-state-file=$(REF_FINAL_STATE_FILE) --av-vels-file=$(AV_VELS_FILE) --final-state-file=$(FINAL_STATE_FILE)

oracle:
	python oracle/build_oracle.py

.PHONY: all check clean oracle

clean:
	rm -f $(EXE) $(LIB)
