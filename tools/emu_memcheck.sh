#!/usr/bin/env bash
# memcheck without a GPU: the emulated kernel sources (tests/emu) under AddressSanitizer — every global / shared / local access
# of every kernel the emulation tests run (default kernels and the opt-in revisions) is bounds-checked against the host
# buffers that stand in for device memory. Needs the system g++ (the image's toolchain g++ ships no libasan).
# usage: tools/emu_memcheck.sh [pytest args, default: the three emulation test files]
cd "$(dirname "$0")/.." || exit 1
ASAN=$(/usr/bin/g++ -print-file-name=libasan.so)
[ -f "$ASAN" ] || { echo "no libasan for /usr/bin/g++"; exit 2; }
ARGS=("$@"); [ ${#ARGS[@]} -eq 0 ] && ARGS=(tests/test_emu_integrate.py tests/test_emu_engine.py tests/test_emu_stream.py)
# -s: AddressSanitizer reports go to stderr and the process aborts; pytest's capture would swallow them
LD_PRELOAD="$ASAN" ASAN_OPTIONS=detect_leaks=0 VH_EMU_LIB=libvh_emu_asan.so python -m pytest "${ARGS[@]}" -x -q -s
