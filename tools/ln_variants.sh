#!/bin/bash
# A/B builds of the LayerNorm backward knobs (csrc/norm.cu: ASIS_LN_BWD_ROWS, ASIS_LN_BWD_MINB), as tools/msda_variants.sh:
#   tools/ln_variants.sh build   (here)      tools/ln_variants.sh run   (GPU box) -> gpurun_out/ln_variants.jsonl
set -eu
cd "$(dirname "$0")/.."
V=tools/variants
C=adaptersis_b200/csrc
variants=("r2b4:-DASIS_LN_BWD_ROWS=2 -DASIS_LN_BWD_MINB=4" "r2b5:-DASIS_LN_BWD_ROWS=2 -DASIS_LN_BWD_MINB=5" "r2b6:-DASIS_LN_BWD_ROWS=2 -DASIS_LN_BWD_MINB=6"
          "r4b4:-DASIS_LN_BWD_ROWS=4 -DASIS_LN_BWD_MINB=4" "r3b4:-DASIS_LN_BWD_ROWS=3 -DASIS_LN_BWD_MINB=4" "r1b8:-DASIS_LN_BWD_ROWS=1 -DASIS_LN_BWD_MINB=8")
if [ "${1:-build}" = build ]; then
  mkdir -p $V
  python -m adaptersis_b200.build
  others=$(ls $C/build/*.o | grep -v norm.o)
  for v in "${variants[@]}"; do
    tag=${v%%:*}; flags=${v#*:}
    nvcc -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC -Xptxas -v -gencode arch=compute_100a,code=sm_100a $flags \
      -c $C/norm.cu -o $V/norm_$tag.o 2> $V/norm_$tag.ptxas.log
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $V/libasis_$tag.so $V/norm_$tag.o $others -lcudart_static -ldl -lpthread -lrt
    echo "built $tag ($flags): $(awk '/Compiling entry function/{n=$7} /spill stores/{sp=$0} /Used [0-9]+ registers/{print n, $5, sp}' $V/norm_$tag.ptxas.log | grep 'ln_bwd2_kernelI13__nv_bfloat16fE' | sed 's/.*EEv[^ ]* //')"
  done
else
  mkdir -p gpurun_out
  : > gpurun_out/ln_variants.jsonl
  python tools/ln_ab.py base >> gpurun_out/ln_variants.jsonl
  for v in "${variants[@]}"; do
    tag=${v%%:*}
    ASIS_LIB=$PWD/$V/libasis_$tag.so python tools/ln_ab.py $tag >> gpurun_out/ln_variants.jsonl
  done
  cat gpurun_out/ln_variants.jsonl
fi
