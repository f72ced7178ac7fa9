#!/usr/bin/env bash
# Builds libmmf_b200.so (sm_100a only) in-tree.  Usage: ./build.sh [extra nvcc flags]
set -euo pipefail
cd "$(dirname "$0")"
PKG="multi-modal-misinformation-detection-with-explanation-generation_b200"
SRC="$PKG/csrc"
OUT="$PKG/libmmf_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall "$@")
mkdir -p build
pids=()
for f in api cosine fusion vault_build vault_stream vault_mma exchange shard; do
  if [ ! -f "build/$f.o" ] || [ "$SRC/$f.cu" -nt "build/$f.o" ] || [ -n "$(find "$SRC" include -name '*.cuh' -newer "build/$f.o" -o -name '*.h' -newer "build/$f.o")" ]; then
    "$NVCC" "${FLAGS[@]}" -c "$SRC/$f.cu" -o "build/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" -shared -o "$OUT" build/api.o build/cosine.o build/fusion.o build/vault_build.o build/vault_stream.o build/vault_mma.o build/exchange.o build/shard.o -ldl
echo "built $OUT"
# torch-free C-ABI self test (tools/cabi_selftest.cu): the quick look on a GPU box
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/cabi_selftest tools/cabi_selftest.cu \
  -L"$PKG" -lmmf_b200 -Xlinker -rpath -Xlinker "\$ORIGIN/../$PKG"
echo "built tools/cabi_selftest"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -o tools/hbm_stride_micro tools/hbm_stride_micro.cu
echo "built tools/hbm_stride_micro"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_stream_micro tools/tma_stream_micro.cu
echo "built tools/tma_stream_micro"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/shard_selftest tools/shard_selftest.cu \
  -L"$PKG" -lmmf_b200 -lpthread -Xlinker -rpath -Xlinker "\$ORIGIN/../$PKG"
echo "built tools/shard_selftest"
