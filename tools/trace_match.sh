#!/bin/bash
# Trace variant of libgadm.so: clock64 accounting in the matcher's UMMA and epilogue warps (-DGADM_MATCH_TRACE).
#   tools/trace_match.sh build   (in the build container: nvcc -> tools/_trace/libgadm.so, travels with gpurun)
#   tools/trace_match.sh run     (on the GPU box: runs one launch of each mode with the trace library swapped in;
#                                 the box's copy of the repo is scratch, the product library here is untouched)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
T=$ROOT/tools/_trace
if [ "$1" = "build" ]; then
  rm -rf $T; mkdir -p $T/build
  cd $ROOT/geometric-aware-dense-matching_b200/csrc
  for f in gadm_api match_sm100 prep knn3d knn_feat gather; do
    nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -Xcompiler -fPIC \
      --expt-relaxed-constexpr -DGADM_MATCH_TRACE -c $f.cu -o $T/build/$f.o &
  done; wait
  nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o $T/libgadm.so $T/build/*.o -cudart static
  rm -rf $T/build
else
  cd $ROOT
  cp geometric-aware-dense-matching_b200/libgadm.so /tmp/libgadm_product.so
  cp $T/libgadm.so geometric-aware-dense-matching_b200/libgadm.so
  python tools/prof_step.py match 1 || true
  cp /tmp/libgadm_product.so geometric-aware-dense-matching_b200/libgadm.so
fi
