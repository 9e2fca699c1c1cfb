#!/bin/bash
# builds a trace variant of libgadm.so (clock64 stamps in the matcher epilogue) and runs one launch of each mode
set -e
cd geometric-aware-dense-matching_b200/csrc
mkdir -p /tmp/tr && for f in gadm_api match_sm100 prep knn3d knn_feat gather; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr -DGADM_MATCH_TRACE -c $f.cu -o /tmp/tr/$f.o; done
cp ../libgadm.so /tmp/libgadm_orig.so
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libgadm.so /tmp/tr/*.o -cudart static
cd ../..
python tools/prof_step.py match 1
cp /tmp/libgadm_orig.so geometric-aware-dense-matching_b200/libgadm.so
