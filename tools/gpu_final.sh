set -x
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_final.log 2>&1; tail -3 gpurun_out/gpu_tests_final.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cat gpurun_out/bench_final.json
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_launch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01w.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches_w.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lu_(refactor|sweep)_wide" -s 9 -c 3 -o gpurun_out/wide_bench_r01 -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_wide.log 2>&1; tail -2 gpurun_out/ncu_full_wide.log
