# Build the C-ABI shared library (CUDA, sm_100a) and the oracle (test infrastructure).
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -Xcompiler -fopenmp -Xcompiler -Wall -Xptxas -v --expt-relaxed-constexpr
CSRC      := $(wildcard parelag_b200/csrc/*.cu)
HSRC      := $(wildcard parelag_b200/src/*.cpp)
OBJ       := $(patsubst %.cu,build/%.o,$(CSRC)) $(patsubst %.cpp,build/%.o,$(HSRC))
LIB       := parelag_b200/lib/libparelag_b200.so

all: $(LIB) oracle

$(LIB): $(OBJ)
	@mkdir -p $(dir $@)
	$(NVCC) -shared $(ARCH) -o $@ $(OBJ) -lcudart -ldl -lgomp

build/%.o: %.cu $(wildcard parelag_b200/csrc/*.cuh) $(wildcard include/*.h)
	@mkdir -p $(dir $@)
	$(NVCC) $(NVFLAGS) -Iinclude -c $< -o $@ 2> build/$(notdir $<).ptxas.log || (cat build/$(notdir $<).ptxas.log; false)

build/%.o: %.cpp $(wildcard parelag_b200/src/*.hpp) $(wildcard include/*.h)
	@mkdir -p $(dir $@)
	g++ -O3 -std=c++17 -fopenmp -fPIC -Wall -Wno-misleading-indentation -Iinclude -Iparelag_b200/src -c $< -o $@

oracle: oracle/libsolve_oracle.so

oracle/libsolve_oracle.so: oracle/solve_oracle.c
	gcc -O3 -march=x86-64-v3 -fopenmp -fPIC -shared -o $@ $< -lm

clean:
	rm -rf build $(LIB) oracle/*.so

.PHONY: all oracle clean
