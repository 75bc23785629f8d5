# Build of the B200-native DIAMOND hot path.  Everything is compiled in-tree (the .so travels to the GPU box).
#   make lib      -> diamond_b200/libdmnd_b200.so      product: host pipeline (C++) + sm_100a CUDA kernels, C ABI
#   make cli      -> diamond_b200/bin/dmnd-b200        product CLI (links the .so)
#   make oracle   -> oracle/_build/libdmnd_oracle.so   TEST ONLY: same host pipeline over the CPU restatement (oracle/dmnd_oracle.c)
#                    oracle/_build/dmnd-oracle-cli
#   make ref      -> oracle/_ref/diamond               the unmodified reference, compiled from /root/reference (if present)
NVCC ?= /usr/local/cuda/bin/nvcc
CXX  := g++
HOST := diamond_b200/csrc/host
CUDA := diamond_b200/csrc/cuda
OBJ  := build/obj
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v
CXXFLAGS := -O2 -std=c++17 -ffp-contract=off -fPIC -pthread -Wall -Wextra

HOST_SRC := pipeline.cpp chaining.cpp scoring.cpp
CUDA_SRC := ctx.cu swipe.cu seed.cu mask.cu chain.cu comm.cu fs.cu
HOST_OBJ := $(patsubst %.cpp,$(OBJ)/host/%.o,$(HOST_SRC))
CUDA_OBJ := $(patsubst %.cu,$(OBJ)/cuda/%.o,$(CUDA_SRC))

all: lib cli oracle
lib: diamond_b200/libdmnd_b200.so
cli: diamond_b200/bin/dmnd-b200
oracle: oracle/_build/libdmnd_oracle.so oracle/_build/dmnd-oracle-cli

$(OBJ)/host/%.o: $(HOST)/%.cpp $(wildcard $(HOST)/*.h) $(wildcard $(HOST)/*.inc) include/dmnd_b200.h
	@mkdir -p $(dir $@)
	$(CXX) $(CXXFLAGS) -c $< -o $@
$(OBJ)/cuda/%.o: $(CUDA)/%.cu $(CUDA)/ctx.cuh $(CUDA)/dev_params.h $(CUDA)/mask_kernels.cuh $(CUDA)/gf_kernels.cuh $(CUDA)/seed_kernels.cuh $(CUDA)/swipe16.cuh $(CUDA)/chain_kernels.cuh $(CUDA)/fs_kernels.cuh include/dmnd_b200.h $(HOST)/motif_table.h
	@mkdir -p $(dir $@)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJ)/cuda/$*.ptxas.log || (cat $(OBJ)/cuda/$*.ptxas.log; false)

diamond_b200/libdmnd_b200.so: $(HOST_OBJ) $(CUDA_OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $^ -lcudart_static -lpthread -ldl -lrt

diamond_b200/bin/dmnd-b200: $(HOST)/cli.cpp $(HOST)/range_cover.h diamond_b200/libdmnd_b200.so
	@mkdir -p $(dir $@)
	$(CXX) $(CXXFLAGS) $< -o $@ -Ldiamond_b200 -ldmnd_b200 -lz -Wl,-rpath,'$$ORIGIN/..'

$(OBJ)/oracle/dmnd_oracle.o: oracle/dmnd_oracle.c include/dmnd_b200.h $(HOST)/motif_table.h
	@mkdir -p $(dir $@)
	gcc -O2 -ffp-contract=off -fPIC -Wall -Wextra -c $< -o $@
oracle/_build/libdmnd_oracle.so: $(HOST_OBJ) $(OBJ)/oracle/dmnd_oracle.o
	@mkdir -p $(dir $@)
	$(CXX) -shared -pthread -o $@ $^ -lm
oracle/_build/dmnd-oracle-cli: $(HOST)/cli.cpp $(HOST)/range_cover.h oracle/_build/libdmnd_oracle.so
	$(CXX) $(CXXFLAGS) $< -o $@ -Loracle/_build -ldmnd_oracle -lz -Wl,-rpath,'$$ORIGIN'

ref:
	$(MAKE) -C oracle/ref_build -j$$(nproc)

clean:
	rm -rf build diamond_b200/libdmnd_b200.so diamond_b200/bin oracle/_build
.PHONY: all lib cli oracle ref clean
