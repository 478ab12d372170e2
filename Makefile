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

all: $(LIB)

$(OBJDIR)/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh include/taper_b200.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(OBJDIR)/host_%.o: $(CSRC)/host/%.cpp $(wildcard $(CSRC)/host/*.hpp) include/taper_b200.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -x cu -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -ldl

examples: $(LIB)
	@mkdir -p build
	g++ -O2 -std=c++17 -Iinclude -I$(CSRC)/host examples/train_mnist.cpp -o build/train_mnist -Ltaper_b200 -ltaper_b200 -Wl,-rpath,'$$ORIGIN/../taper_b200'
	g++ -O2 -std=c++17 -Iinclude -I$(CSRC)/host examples/train_mnist_cnn.cpp -o build/train_mnist_cnn -Ltaper_b200 -ltaper_b200 -Wl,-rpath,'$$ORIGIN/../taper_b200'
	g++ -O2 -std=c++17 -Iinclude -I$(CSRC)/host examples/xor.cpp -o build/xor -Ltaper_b200 -ltaper_b200 -Wl,-rpath,'$$ORIGIN/../taper_b200'

clean:
	rm -rf build $(LIB)

.PHONY: all clean examples
