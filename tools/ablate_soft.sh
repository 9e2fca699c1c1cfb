#!/bin/bash
# Timing-only ablation builds of the SOFT matcher (results are WRONG by construction for the GADM_DBG_* flags; never
# shipped): what each part of the epilogue costs.  Builds tools/_stats/libgadm_<name>.so per flag set; time them with
# tools/_stats/time_soft.py.
#   GADM_DBG_NOEPI    no epilogue at all (TMA -> UMMA -> commit only)
#   GADM_DBG_NOSTASH  no predicated stash stores          GADM_DBG_NOMAX  no max tree / stash / running maximum
#   GADM_DBG_NOXYZ    no LDS of the coordinate planes     GADM_DBG_NOSUMS no FFMA2 coordinate sums
#   GADM_DBG_NOEXP    no MUFU.EX2
#   GADM_SOFT_POLY=n  (a product parameter, results stay valid) n of 4 exponential pairs on the FMA pipe
set -e
cd "$(dirname "$0")/../geometric-aware-dense-matching_b200/csrc"
mkdir -p ../../tools/_stats /tmp/abl
NV="nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -Xcompiler -fPIC --expt-relaxed-constexpr -I."
build() {  # name flags...
  name=$1; shift
  $NV "$@" -c match_sm100.cu -o /tmp/abl/match_$name.o
  nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o ../../tools/_stats/libgadm_$name.so \
    build/gadm_api.o /tmp/abl/match_$name.o build/circle_sm100.o build/prep.o build/knn3d.o build/knn_feat.o build/knn_feat_tc.o build/gather.o -cudart static
  echo built $name
}
if [ $# -gt 0 ]; then
  for spec in "$@"; do   # name:flag,flag
    name=${spec%%:*}; flags=${spec#*:}; build $name $(echo $flags | tr ',' ' ')
  done
  exit 0
fi
build base
build nostash -DGADM_DBG_NOSTASH
build nomax -DGADM_DBG_NOMAX
build noxyz -DGADM_DBG_NOXYZ
build nosums -DGADM_DBG_NOSUMS -DGADM_DBG_NOXYZ
build noexp -DGADM_DBG_NOEXP
build onlyexp -DGADM_DBG_NOMAX -DGADM_DBG_NOSUMS -DGADM_DBG_NOXYZ
build noepi -DGADM_DBG_NOEPI
