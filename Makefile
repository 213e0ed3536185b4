# Builds every native artefact in-tree (the .so files travel to the GPU box with
# the snapshot; they are git-ignored).
#   make            -> product library + synthetic-store generator + oracle
#   make product    -> oarfish_b200/lib/liboarfish_em.so, liboarsynth.so
#   make oracle     -> oracle/liboarfish_oracle.so   (test infrastructure only)
NVCC      ?= nvcc
# the image exports CC=/opt/gcc/bin/gcc, a wrapper that cannot find libgomp.spec; use the system gcc
HOSTCC    := $(shell [ -x /usr/bin/gcc ] && echo /usr/bin/gcc || echo gcc)
HOSTCXX   := $(shell [ -x /usr/bin/g++ ] && echo /usr/bin/g++ || echo g++)
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wextra -Xptxas -v --expt-relaxed-constexpr
CFLAGS    := -O3 -std=gnu11 -fPIC -Wall -Wextra -fopenmp

CSRC      := oarfish_b200/csrc
LIBDIR    := oarfish_b200/lib
CU_SRCS   := $(wildcard $(CSRC)/*.cu)
CU_HDRS   := $(wildcard $(CSRC)/*.cuh) include/oarfish_em.h

all: product oracle

product: $(LIBDIR)/liboarfish_em.so $(LIBDIR)/liboarsynth.so $(LIBDIR)/host_mirror_test

# one object per translation unit, so that `make -j` compiles them side by side
OBJDIR    := $(LIBDIR)/obj
CU_OBJS   := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(CU_SRCS))

$(OBJDIR)/%.o: $(CSRC)/%.cu $(CU_HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) $(EXTRA_NVFLAGS) -c -o $@ $< 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; exit 1)

$(LIBDIR)/liboarfish_em.so: $(CU_OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(CU_OBJS)
	@cat $(OBJDIR)/*.ptxas.log > $(LIBDIR)/ptxas.log

$(LIBDIR)/liboarsynth.so: $(CSRC)/synth.c
	@mkdir -p $(LIBDIR)
	$(HOSTCC) $(CFLAGS) -shared -o $@ $< -lm

# C++ host mirror of the reference interface (include/oarfish_em.hpp) + its test driver
$(LIBDIR)/host_mirror_test: tests/cpp/host_mirror_test.cpp include/oarfish_em.hpp include/oarfish_em.h $(LIBDIR)/liboarfish_em.so
	$(HOSTCXX) -O2 -std=c++17 -Wall -Wextra -Iinclude -o $@ $< -L$(LIBDIR) -loarfish_em -pthread -Wl,-rpath,'$$ORIGIN'

oracle: oracle/liboarfish_oracle.so

oracle/liboarfish_oracle.so: oracle/em_oracle.c oracle/em_par_port.c oracle/coverage_oracle.c oracle/filter_oracle.c
	$(HOSTCC) $(CFLAGS) -shared -o $@ $^ -lm

clean:
	rm -rf $(OBJDIR) $(LIBDIR)/*.so $(LIBDIR)/ptxas.log $(LIBDIR)/host_mirror_test oracle/*.so

.PHONY: all product oracle clean
