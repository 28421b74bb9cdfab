# Build of the B200-native seq-align hot path.
#
#   make            libseqalign_b200.so (CUDA engine + C-ABI + seq-align C API)
#                   and libalign.a (same objects, the reference's library name)
#   make emu        tests/emu/libseqalign_emu.so: the same sources compiled by
#                   g++ against the lane emulator (test infrastructure for the
#                   GPU-less container; never shipped)
#   make oracle     CPU checkers under oracle/ (test infrastructure)
#   make tools      bin/needleman_wunsch, bin/smith_waterman (when present)
NVCC ?= nvcc
CC ?= gcc
CXX ?= g++
ARCH = -gencode arch=compute_100a,code=sm_100a
NVFLAGS = $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall -Iinclude -Iseq-align_b200/csrc
CFLAGS = -std=gnu99 -O2 -Wall -Wextra -fPIC -Iinclude -Iseq-align_b200/host

PKG = seq-align_b200
LIBDIR = $(PKG)/lib
HOST_SRCS = $(PKG)/host/sa_scoring.c $(PKG)/host/sa_alignment.c $(PKG)/host/sa_nw.c $(PKG)/host/sa_sw.c $(PKG)/host/sa_multi.c $(PKG)/host/sa_cli.c $(PKG)/host/sa_cmdline.c
HOST_OBJS = $(HOST_SRCS:.c=.o)
CU_DEPS = $(wildcard $(PKG)/csrc/*.cuh $(PKG)/csrc/*.h include/*.h)

all: $(LIBDIR)/libseqalign_b200.so $(LIBDIR)/libalign.a tools

$(PKG)/host/%.o: $(PKG)/host/%.c $(wildcard include/*.h $(PKG)/host/*.h)
	$(CC) $(CFLAGS) -c $< -o $@

$(PKG)/csrc/sa_engine.o: $(PKG)/csrc/sa_engine.cu $(CU_DEPS)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(PKG)/csrc/sa_decode.o: $(PKG)/csrc/sa_decode.cu $(PKG)/csrc/sa_platform.h include/seqalign_b200.h
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIBDIR)/libseqalign_b200.so: $(PKG)/csrc/sa_engine.o $(PKG)/csrc/sa_decode.o $(HOST_OBJS)
	mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $^ -lpthread -lz

$(LIBDIR)/libalign.a: $(PKG)/csrc/sa_engine.o $(PKG)/csrc/sa_decode.o $(HOST_OBJS)
	mkdir -p $(LIBDIR)
	ar rcs $@ $^

EMU = tests/emu
emu: $(EMU)/libseqalign_emu.so
$(EMU)/libseqalign_emu.so: $(PKG)/csrc/sa_engine.cu $(PKG)/csrc/sa_decode.cu $(CU_DEPS) $(EMU)/cuda_emu.cpp $(EMU)/cuda_emu.h $(HOST_SRCS)
	$(CXX) -O1 -g -std=c++17 -fPIC -fsanitize=alignment -fsanitize-undefined-trap-on-error -DSA_EMU -Iinclude -I$(PKG)/csrc -I$(EMU) -c -x c++ $(PKG)/csrc/sa_engine.cu -o $(EMU)/sa_engine_emu.o
	$(CXX) -O1 -g -std=c++17 -fPIC -fsanitize=alignment -fsanitize-undefined-trap-on-error -DSA_EMU -Iinclude -I$(PKG)/csrc -I$(EMU) -c -x c++ $(PKG)/csrc/sa_decode.cu -o $(EMU)/sa_decode_emu.o
	$(CXX) -O1 -g -std=c++17 -fPIC -I$(EMU) -c $(EMU)/cuda_emu.cpp -o $(EMU)/cuda_emu.o
	for f in $(HOST_SRCS); do $(CC) $(CFLAGS) -c $$f -o $(EMU)/`basename $$f .c`_emu.o || exit 1; done
	$(CXX) -shared -o $@ $(EMU)/sa_engine_emu.o $(EMU)/sa_decode_emu.o $(EMU)/cuda_emu.o $(EMU)/sa_scoring_emu.o $(EMU)/sa_alignment_emu.o $(EMU)/sa_nw_emu.o $(EMU)/sa_sw_emu.o $(EMU)/sa_multi_emu.o $(EMU)/sa_cli_emu.o $(EMU)/sa_cmdline_emu.o -lpthread -lz

# batching command-line tools (same flags / stdout as the reference's bin/*)
TOOLS = bin/needleman_wunsch bin/smith_waterman bin/lcs
TOOL_COMMON =
TOOL_DEPS = $(wildcard $(PKG)/tools/*.h $(PKG)/host/*.h include/*.h) $(LIBDIR)/libseqalign_b200.so
TOOL_LINK = -L$(LIBDIR) -lseqalign_b200 -Wl,-rpath,'$$ORIGIN/../$(LIBDIR)' -lz
tools: $(TOOLS)
bin/needleman_wunsch: $(PKG)/tools/nw_main.c $(TOOL_COMMON) $(TOOL_DEPS)
	mkdir -p bin
	$(CC) $(CFLAGS) -I$(PKG)/tools $< -o $@ $(TOOL_LINK)
bin/smith_waterman: $(PKG)/tools/sw_main.c $(TOOL_COMMON) $(TOOL_DEPS)
	mkdir -p bin
	$(CC) $(CFLAGS) -I$(PKG)/tools $< -o $@ $(TOOL_LINK)
bin/lcs: $(PKG)/tools/lcs_main.c $(TOOL_DEPS)
	mkdir -p bin
	$(CC) $(CFLAGS) $< -o $@ $(TOOL_LINK)

# the same tool sources against the lane emulator (CPU-side tests of the CLI logic)
emu-tools: $(EMU)/libseqalign_emu.so
	mkdir -p $(EMU)/bin
	$(CC) $(CFLAGS) -I$(PKG)/tools $(PKG)/tools/nw_main.c -o $(EMU)/bin/needleman_wunsch -L$(EMU) -lseqalign_emu -Wl,-rpath,'$$ORIGIN/..' -lz
	$(CC) $(CFLAGS) -I$(PKG)/tools $(PKG)/tools/sw_main.c -o $(EMU)/bin/smith_waterman -L$(EMU) -lseqalign_emu -Wl,-rpath,'$$ORIGIN/..' -lz
	$(CC) $(CFLAGS) $(PKG)/tools/lcs_main.c -o $(EMU)/bin/lcs -L$(EMU) -lseqalign_emu -Wl,-rpath,'$$ORIGIN/..' -lz

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(PKG)/host/*.o $(PKG)/csrc/*.o $(LIBDIR)/*.so $(LIBDIR)/*.a $(EMU)/*.o $(EMU)/*.so

.PHONY: all emu oracle clean tools emu-tools
