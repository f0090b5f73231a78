#!/bin/bash
# round 2, GPU call 4: the new bench.py on BASELINE config 4 (1 GPU), launch list, full capture of the dominant kernel
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 1 --warmup 3 > gpurun_out/bench_cfg4_n1.json 2> gpurun_out/bench_cfg4_n1.err
tail -c 1500 gpurun_out/bench_cfg4_n1.json; tail -3 gpurun_out/bench_cfg4_n1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_cfg4_ref.json 2> gpurun_out/bench_cfg4_ref.err
tail -c 800 gpurun_out/bench_cfg4_ref.json
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-secondary > gpurun_out/r02_c4_ncu_launches.log 2>&1
tail -2 gpurun_out/r02_c4_ncu_launches.log | cut -c1-200
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_fixed_point -c 1 -o gpurun_out/prof_r02_cfg4 -f \
  python scripts/run_config.py --config 4 --max-alphas 1 > gpurun_out/r02_c4_ncu_full.log 2>&1
tail -2 gpurun_out/r02_c4_ncu_full.log | cut -c1-200
