set -x
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_w9.log 2>&1; tail -4 gpurun_out/gpu_tests_w9.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_w9.json 2> gpurun_out/bench_w9.err; cat gpurun_out/bench_w9.json; tail -2 gpurun_out/bench_w9.err
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_launch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01w.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches_w.log 2>&1
tail -2 gpurun_out/ncu_launches_w.log
