#!/bin/bash
# developer helper: tools/build_variant.sh NAME "EXTRA NVCC FLAGS" -> build_variants/NAME.so (select with SLATE_B200_LIB)
set -e
cd "$(dirname "$0")/.."
mkdir -p build_variants
SQB_NVCC_EXTRA="$2" python -m slate_quantum_b200.build --force > /dev/null
cp slate_quantum_b200/libslate_b200.so build_variants/$1.so
grep -A3 "tridiag_panel_kernelILi8ELi1" slate_quantum_b200/csrc/build/ptxas.log | grep -i "registers\|spill" | head -3
