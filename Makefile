# Build libtaper_b200.so (hand-written sm_100a CUDA + C ABI + C++ host layer), in-tree.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function \
             -fmad=false --expt-relaxed-constexpr -Iinclude
CSRC      := taper_b200/csrc
OBJDIR    := build/obj
CU_SRCS   := $(wildcard $(CSRC)/*.cu)
CPP_SRCS  := $(wildcard $(CSRC)/host/*.cpp)
OBJS      := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(CU_SRCS)) $(patsubst $(CSRC)/host/%.cpp,$(OBJDIR)/host_%.o,$(CPP_SRCS))
LIB       := taper_b200/libtaper_b200.so

all: $(LIB) examples

$(OBJDIR)/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh $(CSRC)/tc_ptx.cuh include/taper_b200.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(OBJDIR)/host_%.o: $(CSRC)/host/%.cpp $(wildcard $(CSRC)/host/*.hpp) include/taper_b200.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -x cu -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -ldl

EXAMPLES := build/train_mnist build/train_mnist_cnn build/xor
examples: $(EXAMPLES)

build/%: examples/%.cpp $(LIB) $(wildcard $(CSRC)/host/*.hpp) include/taper_b200.h
	@mkdir -p build
	g++ -O2 -std=c++17 -Iinclude -I$(CSRC)/host $< -o $@ -Ltaper_b200 -ltaper_b200 -Wl,-rpath,'$$ORIGIN/../taper_b200'

clean:
	rm -rf build $(LIB)

.PHONY: all clean examples
