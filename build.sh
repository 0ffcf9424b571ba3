#!/bin/bash
# Build the in-tree CUDA shared library for sm_100a (cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
PKG="$HERE/self-supervised-image-enhancement-network-training-with-low-light-images-only_b200"
SRC="$PKG/csrc"
OUT="$PKG/libsshslie_b200.so"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -DSSHSLIE_BUILD"
mkdir -p "$PKG/build"
pids=()
for f in engine conv_simt conv_umma conv_pipe elementwise attention attention_tc loss fft_loss adam; do
  if [ ! -f "$PKG/build/$f.o" ] || [ "$SRC/$f.cu" -nt "$PKG/build/$f.o" ] || [ -n "$(find "$SRC" -name '*.h' -newer "$PKG/build/$f.o" -o -name '*.cuh' -newer "$PKG/build/$f.o")" ] || [ "$HERE/include/sshslie_b200.h" -nt "$PKG/build/$f.o" ]; then
    $NVCC $FLAGS $EXTRA -c "$SRC/$f.cu" -o "$PKG/build/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -Wno-deprecated-gpu-targets -shared -o "$OUT" "$PKG"/build/*.o -cudart static
echo "built $OUT"
