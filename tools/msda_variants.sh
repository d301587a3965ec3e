#!/bin/bash
# A/B builds of the MSDeformAttn tuning knobs (csrc/msda.cu: ASIS_MSDA_FWD_MINB, ASIS_MSDA_FWD_UNROLL,
# ASIS_MSDA_GATHER_U, ASIS_MSDA_GATHER_WIDE): one complete library per variant under tools/variants/ (git-ignored, travels to
# the GPU box), timed with  ASIS_LIB=tools/variants/libasis_<tag>.so python tools/msda_ab.py
#   usage: tools/msda_variants.sh build        (here, no GPU)
#          tools/msda_variants.sh run          (on the GPU box) -> gpurun_out/msda_variants.jsonl
set -eu
cd "$(dirname "$0")/.."
V=tools/variants
C=adaptersis_b200/csrc
variants=("fwd7u2:-DASIS_MSDA_FWD_MINB=7" "fwd8u2:-DASIS_MSDA_FWD_MINB=8" "fwd6u1:-DASIS_MSDA_FWD_MINB=6 -DASIS_MSDA_FWD_UNROLL=1"
          "gatb6:-DASIS_MSDA_GATHER_MINB=6" "gatb8u2:-DASIS_MSDA_GATHER_MINB=8 -DASIS_MSDA_GATHER_U=2" "loc5:-DASIS_MSDA_LOCAW_MINB=5" "loc6:-DASIS_MSDA_LOCAW_MINB=6")
if [ "${1:-build}" = build ]; then
  mkdir -p $V
  python -m adaptersis_b200.build
  others=$(ls $C/build/*.o | grep -v msda.o)
  for v in "${variants[@]}"; do
    tag=${v%%:*}; flags=${v#*:}
    nvcc -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC -Xptxas -v -gencode arch=compute_100a,code=sm_100a $flags \
      -c $C/msda.cu -o $V/msda_$tag.o 2> $V/msda_$tag.ptxas.log
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $V/libasis_$tag.so $V/msda_$tag.o $others -lcudart_static -ldl -lpthread -lrt
    echo "built $tag ($flags): spills $(grep -c 'bytes spill stores' $V/msda_$tag.ptxas.log) entries, $(grep -v ' 0 bytes spill stores' $V/msda_$tag.ptxas.log | grep -c 'spill stores') with spills"
  done
else
  mkdir -p gpurun_out
  : > gpurun_out/msda_variants.jsonl
  python tools/msda_ab.py base >> gpurun_out/msda_variants.jsonl
  for v in "${variants[@]}"; do
    tag=${v%%:*}
    ASIS_LIB=$PWD/$V/libasis_$tag.so python tools/msda_ab.py $tag >> gpurun_out/msda_variants.jsonl
  done
  cat gpurun_out/msda_variants.jsonl
fi
