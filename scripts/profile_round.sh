# One profiling pass of a round on a 1-GPU box (run under gpurun): bench line, launch list of the step's kernels, one
# `ncu --set full` capture of each hot kernel.  Every ncu command follows a plain run of the same command line.
#   gpurun --timeout 900 -- 'bash scripts/profile_round.sh r02'
set -x
R=${1:-r02}
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-other-precision"
timeout 280 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err || exit 1
timeout 200 $B > gpurun_out/plain.log 2>&1 || exit 1
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv -k regex:"logmel|emotion|dual_stream|ema_scan" --log-file gpurun_out/launches_${R}.csv $B > gpurun_out/ncu_l.log 2>&1
timeout 120 python scripts/k1_check.py 512 > gpurun_out/k1_check_${R}.txt 2>&1 || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"logmel_power" -s 6 -c 1 -f -o gpurun_out/k1_${R} python scripts/k1_check.py 512 > gpurun_out/ncu_k1.log 2>&1
timeout 120 python scripts/step_timing.py > gpurun_out/step_timing_${R}.txt 2>&1 || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"dual_stream_tc" -s 8 -c 1 -f -o gpurun_out/tc_${R} python scripts/step_timing.py > gpurun_out/ncu_tc.log 2>&1
ls -la gpurun_out/ | tail -12
