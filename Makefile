# Builds the hand-written sm_100a CUDA behind the C ABI of include/capr_b200.h:
#   capreolus_b200/libcapr_b200.so      the PRODUCT library (what capreolus_b200/_lib.py loads)
#   capreolus_b200/libcapr_b200_dbg.so  the same sources with -DCAPR_DEBUG_BUILD (profiling switches, capr_gemm_test) plus the
#                                       micro-benchmarks under csrc/bench/ (capr_debug_*); used by tests / scripts / the
#                                       L2-gather roofline probe of bench.py, never by a scoring call
# `python -c "import __graft_entry__ as g; g.build()"` drives the same recipe.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Iinclude -Icapreolus_b200/csrc
SRCDIR    := capreolus_b200/csrc
BUILDDIR  := build/obj
DBGDIR    := build/obj_dbg
SRCS      := $(wildcard $(SRCDIR)/*.cu)
BENCHSRCS := $(wildcard $(SRCDIR)/bench/*.cu)
HDRS      := $(wildcard $(SRCDIR)/*.cuh) include/capr_b200.h
OBJS      := $(patsubst $(SRCDIR)/%.cu,$(BUILDDIR)/%.o,$(SRCS))
DBGOBJS   := $(patsubst $(SRCDIR)/%.cu,$(DBGDIR)/%.o,$(SRCS)) $(patsubst $(SRCDIR)/bench/%.cu,$(DBGDIR)/bench_%.o,$(BENCHSRCS))
LIB       := capreolus_b200/libcapr_b200.so
DBGLIB    := capreolus_b200/libcapr_b200_dbg.so

all: $(LIB) $(DBGLIB)

$(BUILDDIR)/%.o: $(SRCDIR)/%.cu $(HDRS)
	@mkdir -p $(BUILDDIR)
	$(NVCC) $(NVCCFLAGS) $(EXTRA) -c $< -o $@

$(DBGDIR)/%.o: $(SRCDIR)/%.cu $(HDRS)
	@mkdir -p $(DBGDIR)
	$(NVCC) $(NVCCFLAGS) -DCAPR_DEBUG_BUILD $(EXTRA) -c $< -o $@

$(DBGDIR)/bench_%.o: $(SRCDIR)/bench/%.cu $(HDRS)
	@mkdir -p $(DBGDIR)
	$(NVCC) $(NVCCFLAGS) -DCAPR_DEBUG_BUILD $(EXTRA) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS)

$(DBGLIB): $(DBGOBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(DBGOBJS)

clean:
	rm -rf build $(LIB) $(DBGLIB)

.PHONY: all clean
