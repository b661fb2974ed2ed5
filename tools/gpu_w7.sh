set -x
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_w7.log 2>&1; tail -5 gpurun_out/gpu_tests_w7.log
python tools/tune.py --workload c3 --batch 10000 --iters 4 --cfg "ws:WS=8;ws:WIDE=0" > gpurun_out/tune_w7.log 2>&1; cat gpurun_out/tune_w7.log
