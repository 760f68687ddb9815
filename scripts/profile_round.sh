set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-other-precision"
timeout 280 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_v9.json 2> gpurun_out/bench_r01_v9.err || exit 1
timeout 200 $B > gpurun_out/b2.json 2>&1 || exit 1
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv -k regex:"logmel|emotion|dual_stream|ema" --log-file gpurun_out/launches_r01_v9_bf16.csv $B > gpurun_out/ncu_l.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"logmel_power" -s 3 -c 1 -f -o gpurun_out/k1_v9 $B > gpurun_out/ncu_k1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"dual_stream_tc" -s 3 -c 1 -f -o gpurun_out/tc_v9 $B > gpurun_out/ncu_tc.log 2>&1
timeout 400 python scripts/bench_configs.py > gpurun_out/bench_configs_r01_v9_1gpu.jsonl 2> gpurun_out/bc.err
ls -la gpurun_out/ | tail -12
