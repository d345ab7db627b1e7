# Developer convenience; __graft_entry__.build() is the contract entry point and does the same compile.
NVCC ?= /usr/local/cuda/bin/nvcc
FLAGS = -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared
SRC = hipgp_b200/csrc
lib: $(SRC)/libhipgp_b200.so
$(SRC)/libhipgp_b200.so: $(SRC)/*.cu $(SRC)/*.cuh $(SRC)/*.inl $(SRC)/*.h include/hipgp_b200.h
	$(NVCC) $(FLAGS) $(SRC)/plan.cu -o $@
dev:
	$(NVCC) $(FLAGS) -DHIPGP_DEV_SMALL $(SRC)/plan.cu -o $(SRC)/libhipgp_b200.so
emu:
	python -c "import sys; sys.path.insert(0,'tests'); import emu_build; print(emu_build.build())"
