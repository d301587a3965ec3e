#!/bin/bash
# ncu --set full captures of the step's top kernels (one GPU), summarised ON THE BOX: the .ncu-rep files of a full set are
# 10-30 MB each and gpurun brings back at most 64 MiB, so only the markdown summaries and the stall hot spots travel.
#   usage: tools/profile_round.sh <tag> ["capture names"]      ->  gpurun_out/<tag>_ncu_*.md, gpurun_out/<tag>_hot_*.txt
set -u
tag=${1:-r2}
out=gpurun_out
tmp=/tmp/ncu_$tag
mkdir -p $tmp $out
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {   # name, kernel regex, skip, count, one_kernel argument
  python tools/one_kernel.py $5 > /dev/null 2> $tmp/$1.err || { echo "FAIL (without ncu) $5"; tail -3 $tmp/$1.err; return; }
  $NCU -k regex:$2 -s $3 -c $4 -o $tmp/$1 python tools/one_kernel.py $5 > $tmp/$1.log 2>&1 || { echo "FAIL (ncu) $1"; tail -3 $tmp/$1.log; return; }
  python tools/ncu_md.py "$tag: $1 ($5)" $out/${tag}_ncu_$1.md $tmp/$1.ncu-rep
  python tools/ncu_hot.py $tmp/$1.ncu-rep "" 30 > $out/${tag}_hot_$1.txt 2>&1
  echo "ok $1"
}
cd "$(dirname "$0")/.."
which=${2:-"gemm_gelu_grad gemm_proj attn conv msda"}
for w in $which; do
  case $w in
    gemm_gelu_grad) cap gemm_gelu_grad gemm_tc_pair 2 1 gemm_gelu_grad ;;
    gemm_proj) cap gemm_proj gemm_tc_pair 2 1 gemm_proj ;;
    attn) cap attn attn_ 3 4 attn_bwd ;;
    attn_fwd) cap attn_fwd attn_fwd 2 1 attn ;;
    conv) cap conv gemm_tc_pair 3 3 conv ;;
    msda) cap msda msda_ 5 5 msda ;;
  esac
done
ls -la $out | grep $tag
