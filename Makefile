# Build recipe of the C-ABI library for callers that do not use Python (same command as
# soc_project_stereo_matching_b200/build.py).  Needs nvcc (CUDA 12.x); cross-compiles for sm_100a without a GPU.
NVCC    ?= nvcc
LIBDIR  := soc_project_stereo_matching_b200/lib
LIB     := $(LIBDIR)/libsgm_b200.so
CSRC    := soc_project_stereo_matching_b200/csrc
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC \
           -Xcompiler -fvisibility=default --fmad=false -Iinclude

.PHONY: lib example clean
lib: $(LIB)

$(LIB): $(CSRC)/sgm_b200.cu $(wildcard $(CSRC)/*.cuh $(CSRC)/*.h include/*.h)
	mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -o $@ $(CSRC)/sgm_b200.cu

# plain C caller: the reference's call sequence (main.c:72,83) plus the board's frame loop
example: $(LIB) examples/frame_loop.c
	$(CC) -std=gnu11 -Wall -Iinclude -o examples/frame_loop examples/frame_loop.c -L$(LIBDIR) -lsgm_b200 -Wl,-rpath,$(abspath $(LIBDIR)) -lm

clean:
	rm -f $(LIB) examples/frame_loop
