#!/bin/bash
# The host-callable part of the C ABI (codec, Fano, nhash, unpack, channel symbols, the reference's own unit tests linked
# against the library, file formats) under AddressSanitizer + UndefinedBehaviorSanitizer: builds the library's host code with
# -fsanitize=address,undefined into a scratch directory and runs the CPU tests of the ABI against it.  No GPU needed.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${1:-/tmp/wspr_b200_san}
mkdir -p "$OUT/obj"
cd "$ROOT/rtlsdr_wsprd_b200/csrc"
for f in wspr_kernels wspr_decode wspr_frontend wspr_abi; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -fmad=false \
       -Xcompiler -fPIC,-fvisibility=default,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer -c $f.cu -o "$OUT/obj/$f.o"
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libwsprd_b200.so" "$OUT"/obj/*.o -cudart shared \
     -Xcompiler -fsanitize=address,-fsanitize=undefined
cd "$ROOT"
WSPR_B200_LIB="$OUT/libwsprd_b200.so" LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
  ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
  python -m pytest tests/test_abi_cpu.py tests/test_host_formats.py -q
