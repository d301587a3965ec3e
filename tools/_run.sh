timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for k in gemm_gelu attn_fwd attn_bwd; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|attn_" --launch-skip 2 -c 2 -o gpurun_out/r62_$k python tools/one_kernel.py $k > gpurun_out/r62_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
timeout 500 python bench.py > gpurun_out/r62_bench.json 2> gpurun_out/r62_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r62_bench.err | cut -c1-600
cut -c1-200 gpurun_out/r62_bench.json
