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
# optional debug variant with device-side index assertions (tests: TPPMX_LIB=.../libtppmx_check.so)
if [ "${TPPMX_BUILD_CHECK:-0}" = "1" ]; then
  "${NVCC}" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DTPPMX_CHECK \
    -Xcompiler -fPIC,-O2 -shared -I"${HERE}/../../include" -o "${HERE}/../libtppmx_check.so" "${HERE}/tppmx.cu"
  echo "built ${HERE}/../libtppmx_check.so"
fi
