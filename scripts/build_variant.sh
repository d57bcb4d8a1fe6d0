#!/bin/bash
# Builds a variant of the library with extra compile flags: gnn_fpga_b200/libgnnseg_<name>.so
# (select it with GNNSEG_LIB=gnn_fpga_b200/libgnnseg_<name>.so).  For A/B runs of kernel variants.
# usage: scripts/build_variant.sh <name> <extra nvcc flags...>
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/gnn_fpga_b200/csrc
OUT=$ROOT/build/$NAME
mkdir -p "$OUT"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-O3,-Wall -I$ROOT/include $*"
pids=()
for f in gnnseg_forward gnnseg_node_tc gnnseg_backward gnnseg_graph gnnseg_segments gnnseg_abi gnnseg_fused; do
    [ -f "$SRC/$f.cu" ] || continue
    $NVCC $FLAGS -c "$SRC/$f.cu" -o "$OUT/$f.o" &
    pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
objs=$(ls "$OUT"/*.o)
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$ROOT/gnn_fpga_b200/libgnnseg_$NAME.so" $objs "$SRC/gnnseg_host.o" "$SRC/gnnseg_npz.o" "$SRC/gnnseg_store.o" -lpthread -lgomp
echo "built gnn_fpga_b200/libgnnseg_$NAME.so"
