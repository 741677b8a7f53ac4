#!/bin/bash
# The kernels and the scheduler under AddressSanitizer + UndefinedBehaviorSanitizer: the emulated build (see README.md) is
# compiled with -fsanitize=address,undefined and the GPU parity tests are run against it -- out-of-bounds accesses to device
# buffers, shared arrays (static storage with red zones) and dynamic shared memory, signed overflows, bad shifts ... stop the
# run at the first report.  (object-size / pointer-overflow are off: shared-window addresses are offsets from one anchor
# symbol by construction.)  Usage: tools/cuda_emu/sanitize.sh [outdir] [pytest -k expression]
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
OUT=${1:-/tmp/wspr_b200_emu_san}
SEL=${2:-"reference_fixture or golden_option_variants or weak_signals_exercise or drifting_and_edge or degenerate or short_capture or persistent_hashtable_option or fano_kernel or sync_and_demodulate_abi or subtract_signal2_abi or subtract_signal_abi or stage_spectrogram or frontend_against_reference_golden or frontend_ragged or streaming_frontend or one_shot_batch_entry or quick_and_normal or four_passes or hashtable_batch"}
cd "$ROOT"
python - "$OUT" <<'PY'
import os, subprocess, sys
sys.path.insert(0, os.path.join(os.getcwd(), "tools", "cuda_emu"))
import build
out = sys.argv[1]
san = ["-fsanitize=address", "-fsanitize=undefined", "-fno-omit-frame-pointer", "-fno-sanitize=alignment,object-size,pointer-overflow"]
lib = build.build(out, opt="-O1", defs=san)
dst = os.path.join(out, "rtlsdr_wsprd_b200", "csrc")
objs = [os.path.join(dst, n + ".o") for n in build.SOURCES] + [os.path.join(dst, "emu_runtime.o")]
subprocess.run(["g++", "-shared", "-fsanitize=address", "-fsanitize=undefined", "-o", lib] + objs + ["-lpthread"], check=True)
PY
WSPR_B200_LIB="$OUT/libwsprd_b200_emu.so" LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
  ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -s -q -p no:cacheprovider -k "$SEL"
