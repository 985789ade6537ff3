# Build of the B200-native IMSAME hot path.
#   make            host library, GPU library (sm_100a), CLI binaries
#   make oracle     test-only CPU restatement (oracle/_build/liboracle.so)
#   make ref        the unmodified reference compiled from /root/reference/src into oracle/_ref/
NVCC      ?= /usr/local/cuda/bin/nvcc
# the image exports CC=/opt/gcc/bin/gcc, a relocated gcc without libgomp.spec: use the system one
CC        := $(firstword $(wildcard /usr/bin/gcc) gcc)
OMPFLAG   := $(shell echo 'int main(){return 0;}' | $(CC) -fopenmp -x c - -o /dev/null 2>/dev/null && echo -fopenmp)
CFLAGS    := -O3 -fPIC -Wall -D_FILE_OFFSET_BITS=64 -D_LARGEFILE64_SOURCE $(OMPFLAG)
NVFLAGS   := -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
             -Xcompiler -fPIC,-Wall -Xptxas -v
OUT       := imsame_b200/_lib
HOST_SRC  := imsame_b200/host/fasta.c imsame_b200/host/thresholds.c imsame_b200/host/synth.c \
             imsame_b200/host/render.c
GPU_SRC   := imsame_b200/csrc/capi.cu imsame_b200/csrc/nwp_launch.cu
GPU_OBJ   := $(OUT)/capi.o $(OUT)/nwp_launch.o
GPU_HDR   := $(wildcard imsame_b200/csrc/*.cuh) $(wildcard imsame_b200/csrc/*.inc) $(wildcard imsame_b200/csrc/*.h) include/imsame_gpu.h imsame_b200/host/imsame_host.h

all: $(OUT)/libimsame_host.so $(OUT)/libimsame_gpu.so bin/IMSAME bin/IMSAME_allvsall bin/revComp bin/all_vs_all_metagenomes_IMSAME.sh

$(OUT)/libimsame_host.so: $(HOST_SRC) imsame_b200/host/imsame_host.h include/imsame_gpu.h
	@mkdir -p $(OUT)
	$(CC) $(CFLAGS) -shared $(HOST_SRC) -lm -o $@

$(OUT)/%.o: imsame_b200/csrc/%.cu $(GPU_HDR)
	@mkdir -p $(OUT)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OUT)/ptxas_$*.log || (cat $(OUT)/ptxas_$*.log; false)
	@grep -E "error|warning|spill" $(OUT)/ptxas_$*.log | grep -v "0 bytes spill\|pragma unroll" || true

$(OUT)/libimsame_gpu.so: $(GPU_OBJ) imsame_b200/host/thresholds.c
	$(CC) $(CFLAGS) -c imsame_b200/host/thresholds.c -o $(OUT)/thresholds.o
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -shared $(GPU_OBJ) $(OUT)/thresholds.o -ldl -o $@
	@cat $(OUT)/ptxas_capi.log $(OUT)/ptxas_nwp_launch.log > $(OUT)/ptxas.log

JOB_SRC   := imsame_b200/host/imsame_job.c imsame_b200/host/imsame_job.h

bin/IMSAME: imsame_b200/host/imsame_main.c $(JOB_SRC) $(OUT)/libimsame_host.so $(OUT)/libimsame_gpu.so
	@mkdir -p bin
	$(CC) $(CFLAGS) -fPIE imsame_b200/host/imsame_main.c imsame_b200/host/imsame_job.c -L$(OUT) -limsame_host -limsame_gpu \
	    -Wl,-rpath,'$$ORIGIN/../$(OUT)' -lpthread -lm -o $@

bin/IMSAME_allvsall: imsame_b200/host/imsame_allvsall_main.c $(JOB_SRC) $(OUT)/libimsame_host.so $(OUT)/libimsame_gpu.so
	@mkdir -p bin
	$(CC) $(CFLAGS) -fPIE imsame_b200/host/imsame_allvsall_main.c imsame_b200/host/imsame_job.c -L$(OUT) -limsame_host -limsame_gpu \
	    -Wl,-rpath,'$$ORIGIN/../$(OUT)' -lpthread -lm -o $@

bin/revComp: imsame_b200/host/revcomp_main.c $(OUT)/libimsame_host.so
	@mkdir -p bin
	$(CC) $(CFLAGS) -fPIE imsame_b200/host/revcomp_main.c -L$(OUT) -limsame_host -Wl,-rpath,'$$ORIGIN/../$(OUT)' -lm -o $@

tools: tools/int_peak
tools/int_peak: tools/int_peak.cu
	$(NVCC) -O3 -gencode arch=compute_100a,code=sm_100a -o $@ tools/int_peak.cu

bin/all_vs_all_metagenomes_IMSAME.sh: scripts/all_vs_all_metagenomes_IMSAME.sh
	@mkdir -p bin
	cp scripts/all_vs_all_metagenomes_IMSAME.sh $@ && chmod +x $@

oracle: oracle/_build/liboracle.so
oracle/_build/liboracle.so: oracle/imsame_oracle.c oracle/imsame_sampled.c oracle/imsame_oracle.h
	@mkdir -p oracle/_build
	$(CC) -O2 -fPIC -Wall $(OMPFLAG) -shared oracle/imsame_oracle.c oracle/imsame_sampled.c -lm -o $@

ref:
	sh oracle/build_ref.sh

clean:
	rm -rf $(OUT) bin oracle/_build

.PHONY: all oracle ref clean tools
