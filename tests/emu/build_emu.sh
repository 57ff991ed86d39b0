#!/bin/sh
# TEST INFRASTRUCTURE ONLY: builds tests/emu/libdpc_b200_emu.so (CPU emulation of the kernels).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
CSRC="$HERE/../../differentiable-point-clouds_b200/csrc"
g++ -std=c++20 -O1 -g -fPIC -shared -pthread -ffp-contract=off -DDPC_EMU -DDPC_EXPERIMENTS -I"$HERE" -I"$CSRC" \
    -x c++ "$CSRC/dpc_capi.cu" "$HERE/cuda_emu.cpp" -o "$HERE/libdpc_b200_emu.so"
echo built "$HERE/libdpc_b200_emu.so"
