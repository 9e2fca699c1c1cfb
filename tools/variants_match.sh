#!/bin/bash
# Timing experiments (NOT product builds; results of the variants are wrong by construction): libgadm.so variants with
# one epilogue ingredient removed, to see what bounds the matcher.   build (here) / run (GPU box)
#   VARS="BASE NOEPI NOSTASH NOLDS NOMUFU NOMAX LDONLY" tools/variants_match.sh build      (-DGADM_DBG_<name>)
#   NOEPI   no epilogue at all: the TMA -> UMMA -> commit pipeline ceiling (match_kernel, pair, frag, TMEM-A kernels)
#   NOSTASH / NOLDS / NOMUFU / NOMAX / LDONLY: fragment-layout kernel without stash stores / constant loads /
#   exponentials / max tree + stash / everything but the TMEM loads.
#   Select the kernel under test at run time with GADM_MATCH_FRAG / GADM_MATCH_TA / GADM_MATCH_ALT / GADM_MATCH_PAIR /
#   GADM_MATCH_RT (see match_launch_t); results: profiles/SUMMARY_r1.md.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
T=$ROOT/tools/_variants
VARS="${VARS:-BASE}"
if [ "$1" = "build" ]; then
  rm -rf $T; mkdir -p $T
  cd $ROOT/geometric-aware-dense-matching_b200/csrc
  for v in $VARS; do
    nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -Xcompiler -fPIC \
      --expt-relaxed-constexpr -DGADM_DBG_$v -c match_sm100.cu -o $T/match_$v.o &
  done; wait
  for v in $VARS; do
    nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o $T/libgadm_$v.so $T/match_$v.o \
      build/gadm_api.o build/prep.o build/knn3d.o build/knn_feat.o build/gather.o -cudart static
  done
  rm -f $T/*.o
else
  cd $ROOT
  cp geometric-aware-dense-matching_b200/libgadm.so /tmp/libgadm_product.so
  for v in $VARS; do
    cp $T/libgadm_$v.so geometric-aware-dense-matching_b200/libgadm.so
    echo "== $v"; python tools/bench_match.py 2>&1 | tail -2
  done
  cp /tmp/libgadm_product.so geometric-aware-dense-matching_b200/libgadm.so
fi
