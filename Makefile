# Builds the CUDA library, the C++ driver above its C-ABI and the CPU oracle without Python
# (python __graft_entry__.py does the same through ndpp_b200/build.py and oracle/pyoracle.py).
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
CSRC      := ndpp_b200/csrc
LIB       := $(CSRC)/libndppgpu.so
TOOL      := tools/ndpp_calc_scatt
# -fmad=false: the reference is compiled without FMA contraction and its closed forms cancel catastrophically;
# contracting a*b+c on the device would move the results outside the parity tolerance (csrc/legendre.cuh)
NVCCFLAGS := -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared -ccbin /usr/bin/g++

all: $(LIB) $(TOOL) oracle

$(LIB): $(wildcard $(CSRC)/*.cu $(CSRC)/*.cuh $(CSRC)/*.inc) include/ndppgpu.h
	$(NVCC) $(NVCCFLAGS) $(NDPP_NVCC_EXTRA) -o $@ $(CSRC)/ndppgpu.cu

$(TOOL): tools/ndpp_calc_scatt.cpp include/ndpp_host.hpp include/ndpp_library.hpp include/ndppgpu.h $(LIB)
	$(CXX) -O2 -std=c++17 -Wall -Wextra -I include $< -o $@ -L $(CSRC) -lndppgpu '-Wl,-rpath,$$ORIGIN/../ndpp_b200/csrc'

oracle:
	$(MAKE) -C oracle

# regenerate the fused closed-form Legendre integrals from the reference text in csrc/legendre.cuh and check them on the host
fused:
	python scripts/gen_legendre_fused.py && python scripts/gen_legendre_fused.py --check

clean:
	rm -f $(LIB) $(TOOL)
	$(MAKE) -C oracle clean

.PHONY: all oracle fused clean
