timeout 300 python tools/torch_profile_step.py > gpurun_out/r61_prof.log 2>&1; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r61_step_launches.csv python bench.py --steps 1 --warmup 3 --only-timed > gpurun_out/r61_ncu_bench.log 2>&1; echo rc=$?
