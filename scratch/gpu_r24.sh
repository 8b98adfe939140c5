#!/bin/bash
# one GPU session: parity subset, both benches, ncu launch list and a full capture of k_exp_var2
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mix_and_vmnv or decryption or group_ops" 2>&1 | tail -3
python bench.py --workload verify-mix --no-cpu --trace gpurun_out/r24_trace_mix.json > gpurun_out/r24_bench_mix.log 2> gpurun_out/r24_bench_mix.err
tail -c 400 gpurun_out/r24_bench_mix.log
python bench.py --no-cpu --phases > gpurun_out/r24_bench.log 2> gpurun_out/r24_bench.err
tail -c 330 gpurun_out/r24_bench.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r24_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r24_ncu_bench.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_exp_var2 -c 1 -o gpurun_out/r24_var2 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r24_ncu_var2.log 2>&1
ls -la gpurun_out/r24*
