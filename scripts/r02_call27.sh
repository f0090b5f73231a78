#!/bin/bash
# round 2, GPU call 27 (final build): whole GPU suite, the default bench line, ncu launch list of a step,
# ncu --set full of the fixed-point kernel on config 4 (first alpha)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_c27_pytest.txt 2>&1
tail -4 gpurun_out/r02_c27_pytest.txt
timeout 1500 python bench.py > gpurun_out/r02_c27_bench_n1.json 2> gpurun_out/r02_c27_bench_n1.err
python -c "
import json
l=json.loads(open('gpurun_out/r02_c27_bench_n1.json').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step'], l['e2e']['ms_per_step'], l['roofline']['frac'], l['roofline']['avg_pass_us'], l['roofline']['fused_pass']['avg_launch_us'], l['clocks'], {k: v.get('ok') for k, v in l.items() if k.startswith('parity')})"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c27_launches.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-secondary > gpurun_out/r02_c27_ncu.log 2>&1
tail -1 gpurun_out/r02_c27_ncu.log | cut -c1-200
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_fixed_point -c 1 -o gpurun_out/prof_r02_cfg4_final -f \
  python scripts/run_config.py --config 4 --max-alphas 1 > gpurun_out/r02_c27_ncu_full.log 2>&1
tail -1 gpurun_out/r02_c27_ncu_full.log | cut -c1-200
ls -la gpurun_out/prof_r02_cfg4_final.ncu-rep
