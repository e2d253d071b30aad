#!/bin/bash
# Builds treatppmx_b200/libtppmx.so for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libtppmx.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC,-O2 -shared ${TPPMX_NVCC_EXTRA:-} \
  -I"${HERE}/../../include" -o "${OUT}" "${HERE}/tppmx.cu"
echo "built ${OUT}"
