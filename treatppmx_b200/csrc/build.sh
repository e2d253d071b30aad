#!/bin/bash
# Builds treatppmx_b200/libtppmx.so for sm_100a (B200).  nvcc cross-compiles without a GPU.
# Four translation units compiled in parallel (the fast sweep kernel's instantiations per LPU
# dominate compile time), then one link.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 -I"${HERE}/../../include")
UNITS=(tppmx sweep_fast_lpu8 sweep_fast_lpu16 sweep_fast_lpu32 sweep_big multi)

build_variant() {  # $1 = output .so, $2 = object dir, rest = extra flags
  local out="$1" odir="$2"; shift 2
  mkdir -p "${odir}"
  local pids=()
  for u in "${UNITS[@]}"; do
    "${NVCC}" "${FLAGS[@]}" "$@" ${TPPMX_NVCC_EXTRA:-} -c "${HERE}/${u}.cu" -o "${odir}/${u}.o" &
    pids+=($!)
  done
  for p in "${pids[@]}"; do wait "$p"; done
  local objs=()
  for u in "${UNITS[@]}"; do objs+=("${odir}/${u}.o"); done
  "${NVCC}" -gencode arch=compute_100a,code=sm_100a -shared -o "${out}" "${objs[@]}" -ldl
  echo "built ${out}"
}

# experiment builds: TPPMX_VARIANT=name [TPPMX_NVCC_EXTRA="-DX=1 ..."] -> libtppmx_name.so (A/B runs
# select it with TPPMX_LIB=...); the default build is untouched
if [ -n "${TPPMX_VARIANT:-}" ]; then
  build_variant "${HERE}/../libtppmx_${TPPMX_VARIANT}.so" "${HERE}/build/obj_${TPPMX_VARIANT}"
  exit 0
fi
build_variant "${HERE}/../libtppmx.so" "${HERE}/build/obj"
# optional debug variant with device-side index assertions (tests: TPPMX_LIB=.../libtppmx_check.so)
if [ "${TPPMX_BUILD_CHECK:-0}" = "1" ]; then
  build_variant "${HERE}/../libtppmx_check.so" "${HERE}/build/obj_check" -DTPPMX_CHECK
fi
