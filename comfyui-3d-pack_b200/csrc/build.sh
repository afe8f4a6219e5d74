#!/bin/bash
# Builds libgs_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH --expt-relaxed-constexpr ${GS_NVCC_EXTRA}"
mkdir -p build
pids=()
# preprocess forward: NO FMA contraction (bit-exact integer outputs vs the oracle)
$NVCC $COMMON -fmad=false -c gs_preprocess.cu -o build/gs_preprocess.o & pids+=($!)
$NVCC $COMMON -fmad=false -c dr_raster.cu -o build/dr_raster.o & pids+=($!)
for f in gs_sort gs_binning gs_render gs_composite gs_backward gs_knn gs_train gs_densify gs_loss gs_capi dr_ops dr_capi ngp_ops ngp_mlp ngp_capi; do
  $NVCC $COMMON -c $f.cu -o build/$f.o & pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared $ARCH -o ../gs_b200/libgs_b200.so build/*.o -lcudart
echo "built $(realpath ../gs_b200/libgs_b200.so)"
