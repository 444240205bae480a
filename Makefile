# Builds everything in-tree:
#   saamge_b200/lib/libsaamge_b200.so   CUDA kernels + C ABI (include/saamge_b200.h), sm_100a
#   saamge_b200/lib/libsaamge_host.so   C++ host mirror of the reference API + driver API
#   oracle/liboracle.so                 CPU oracle (test infrastructure only)
CXX := g++
NVCC ?= /usr/local/cuda/bin/nvcc
CUDA_HOME ?= /usr/local/cuda
CXXFLAGS = -O2 -g -fPIC -fopenmp -std=c++14 -Wall -Wno-unknown-pragmas
NVFLAGS = -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a \
          -Xcompiler -fPIC -Xcompiler -fopenmp -Xptxas -v
METIS_A = $(CUDA_HOME)/targets/x86_64-linux/lib/libmetis_static.a

LIBDIR = saamge_b200/lib
HOST_SRC = $(wildcard saamge_b200/host/*.cpp)
HOST_HDR = $(wildcard saamge_b200/host/*.hpp) $(wildcard include/*.h)
CU_SRC = $(wildcard saamge_b200/csrc/*.cu)
CU_HDR = $(wildcard saamge_b200/csrc/*.cuh) $(wildcard include/*.h)
ORC_SRC = $(wildcard oracle/*.cpp)
ORC_HDR = $(wildcard oracle/*.hpp)

all: gpu host oracle
gpu: $(LIBDIR)/libsaamge_b200.so
host: $(LIBDIR)/libsaamge_host.so
oracle: oracle/liboracle.so

OBJDIR = build/obj
CU_OBJ = $(patsubst saamge_b200/csrc/%.cu,$(OBJDIR)/%.o,$(CU_SRC))

# one object per translation unit (kernels are file-local: no relocatable device code), so a
# change to one .cu recompiles that file only; `make -j` builds them side by side
$(OBJDIR)/%.o: saamge_b200/csrc/%.cu $(CU_HDR)
	@mkdir -p $(OBJDIR) $(LIBDIR)
	$(NVCC) $(NVFLAGS) -c -o $@ $< -Iinclude 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; false)

$(LIBDIR)/libsaamge_b200.so: $(CU_OBJ)
	@mkdir -p $(LIBDIR)
	$(NVCC) -shared -o $@ $(CU_OBJ) -lcudart
	@cat $(OBJDIR)/*.ptxas.log > $(LIBDIR)/ptxas.log

$(LIBDIR)/libsaamge_host.so: $(HOST_SRC) $(HOST_HDR) $(LIBDIR)/libsaamge_b200.so
	@mkdir -p $(LIBDIR)
	$(CXX) $(CXXFLAGS) -shared -o $@ $(HOST_SRC) -Iinclude $(METIS_A) \
	    -L$(LIBDIR) -lsaamge_b200 -Wl,-rpath,'$$ORIGIN' -lm

oracle/liboracle.so: $(ORC_SRC) $(ORC_HDR) $(HOST_HDR) $(LIBDIR)/libsaamge_host.so
	$(CXX) $(CXXFLAGS) -shared -o $@ $(ORC_SRC) -Iinclude \
	    -L$(LIBDIR) -lsaamge_host -Wl,-rpath,'$$ORIGIN/../$(LIBDIR)' -ldl -lm

clean:
	rm -rf $(LIBDIR)/*.so $(LIBDIR)/ptxas.log oracle/liboracle.so $(OBJDIR)

.PHONY: all gpu host oracle clean
