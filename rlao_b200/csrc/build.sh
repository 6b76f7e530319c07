#!/usr/bin/env bash
# Builds libaoenv_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
OUT=../libaoenv_b200.so
SRCS="api.cu atm.cu step.cu vk.cu gemm.cu wfs.cu wfs_fused.cu pyr.cu ctrl.cu dm.cu"
[ -f psf.cu ] && SRCS="$SRCS psf.cu"
[ -f gemm_tc.cu ] && SRCS="$SRCS gemm_tc.cu"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
TMP="$OUT.tmp.$$"
set +e
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
      -Xcompiler -fPIC,-O2,-Wall -Xptxas -v --shared -o "$TMP" $SRCS > build.log 2>&1
rc=$?
set -e
grep -E "error|warning|spill|Used" build.log || true
if [ $rc -ne 0 ] || [ ! -f "$TMP" ]; then
  rm -f "$TMP"
  echo "BUILD FAILED (nvcc exit $rc); $OUT left untouched" >&2
  exit 1
fi
mv -f "$TMP" "$OUT"
echo "built $OUT"
