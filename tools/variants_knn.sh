#!/bin/bash
# timing experiments: libgadm.so variants (tools/_variants/libgadm_<V>.so) run through tools/bench_knn.py on the GPU box
ROOT=$(cd "$(dirname "$0")/.." && pwd); cd $ROOT
cp geometric-aware-dense-matching_b200/libgadm.so /tmp/libgadm_product.so
TAG=product python tools/bench_knn.py 2>&1 | tail -1
for v in $VARS; do
  cp tools/_variants/libgadm_$v.so geometric-aware-dense-matching_b200/libgadm.so
  TAG=$v python tools/bench_knn.py 2>&1 | tail -1
done
cp /tmp/libgadm_product.so geometric-aware-dense-matching_b200/libgadm.so
