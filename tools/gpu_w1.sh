set -x
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_w1.log 2>&1; tail -5 gpurun_out/gpu_tests_w1.log
python tools/tune.py --workload c3 --batch 10000 --iters 4 --cfg "ws:;ws:WF=128;ws:WF=96;ws:WF=64;ws:WS=32;ws:WIDE=0" > gpurun_out/tune_w1.log 2>&1; cat gpurun_out/tune_w1.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_w1.json 2> gpurun_out/bench_w1.err; cat gpurun_out/bench_w1.json; tail -3 gpurun_out/bench_w1.err
