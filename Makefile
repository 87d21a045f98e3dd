# Builds the product: imd_b200/libimd_b200.so (hand-written sm_100a kernels + C ABI + C host helpers).
# The oracle/ directory has its own Makefile (test infrastructure).
NVCC     ?= nvcc
CC       ?= gcc
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr
CFLAGS   := -O2 -fPIC -Wall -std=c99 -D_POSIX_C_SOURCE=200809L
CU       := $(wildcard imd_b200/csrc/*.cu)
HC       := $(wildcard imd_b200/host/*.c)
OBJ      := $(CU:imd_b200/csrc/%.cu=build/%.o) build/forces_cubic.o build/forces_eeam.o build/forces_cubic_eeam.o $(HC:imd_b200/host/%.c=build/host_%.o)
LIB      := imd_b200/libimd_b200.so

all: $(LIB)

build/%.o: imd_b200/csrc/%.cu imd_b200/csrc/internal.cuh include/imd_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)

# the force kernels once more for the cubic table interpolations (4point / spline), see forces.cu
build/forces_cubic.o: imd_b200/csrc/forces.cu imd_b200/csrc/internal.cuh include/imd_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -DIMDB_CUBIC=1 -c $< -o $@ 2> build/forces_cubic.ptxas.log || (cat build/forces_cubic.ptxas.log; false)

# ... and for the extended-EAM terms (EEAM builds of the reference)
build/forces_eeam.o: imd_b200/csrc/forces.cu imd_b200/csrc/internal.cuh include/imd_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -DIMDB_EEAM=1 -c $< -o $@ 2> build/forces_eeam.ptxas.log || (cat build/forces_eeam.ptxas.log; false)

build/forces_cubic_eeam.o: imd_b200/csrc/forces.cu imd_b200/csrc/internal.cuh include/imd_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -DIMDB_CUBIC=1 -DIMDB_EEAM=1 -c $< -o $@ 2> build/forces_cubic_eeam.ptxas.log || (cat build/forces_cubic_eeam.ptxas.log; false)

build/host_%.o: imd_b200/host/%.c include/imd_b200.h
	@mkdir -p build
	$(CC) $(CFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIB)

.PHONY: all oracle clean
