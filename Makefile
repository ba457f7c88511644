# Builds libcapr_b200.so (hand-written sm_100a CUDA behind the C ABI of include/capr_b200.h) and the
# C-ABI symbol check.  `python -c "import __graft_entry__ as g; g.build()"` drives the same recipe.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Iinclude -Icapreolus_b200/csrc
SRCDIR    := capreolus_b200/csrc
BUILDDIR  := build/obj
SRCS      := $(wildcard $(SRCDIR)/*.cu)
OBJS      := $(patsubst $(SRCDIR)/%.cu,$(BUILDDIR)/%.o,$(SRCS))
LIB       := capreolus_b200/libcapr_b200.so

all: $(LIB)

$(BUILDDIR)/%.o: $(SRCDIR)/%.cu $(wildcard $(SRCDIR)/*.cuh) include/capr_b200.h
	@mkdir -p $(BUILDDIR)
	$(NVCC) $(NVCCFLAGS) $(EXTRA) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) 

clean:
	rm -rf build $(LIB)

.PHONY: all clean
