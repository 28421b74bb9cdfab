#!/bin/bash
# experiment builds: the library with a row checkpoint every 32 / 16 rows instead of 64 (wide-pair traceback)
# -> seq-align_b200/lib_ck32/, lib_ck16/ (SEQALIGN_LIB selects one); the default build is untouched
set -e
cd "$(dirname "$0")/.."
make -j8 > /dev/null
for r in 32 16; do
  mkdir -p seq-align_b200/lib_ck$r
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall -Iinclude -Iseq-align_b200/csrc \
       -DSA_CK_ROWS=$r -c seq-align_b200/csrc/sa_engine.cu -o seq-align_b200/lib_ck$r/sa_engine.o &
done
wait
for r in 32 16; do
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o seq-align_b200/lib_ck$r/libseqalign_b200.so seq-align_b200/lib_ck$r/sa_engine.o \
       seq-align_b200/csrc/sa_decode.o seq-align_b200/host/*.o -lpthread -lz
done
ls -la seq-align_b200/lib_ck*/libseqalign_b200.so
