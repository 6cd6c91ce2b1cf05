#!/usr/bin/env bash
# Memory-safety check without a GPU tool (compute-sanitizer is closed on this pool): the kernel sources compiled for the CPU
# fiber emulator (tests/emu, g++ -DSYG_EMU) with AddressSanitizer, driven by the emulator legs of the C-ABI parity tests.
# Device buffers are numpy allocations and shared memory is a per-block heap block, so an out-of-range global or shared
# index aborts the run with an ASan report.  Usage: bash scripts/emu_asan.sh [pytest args]
set -euo pipefail
cd "$(dirname "$0")/.."
export SYG_EMU_ASAN=1
python tests/emu/build_emu.py
LD_PRELOAD="$(gcc -print-file-name=libasan.so)" ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
    python -m pytest tests/test_parity_cabi.py -m "not gpu" -x -q "$@"
