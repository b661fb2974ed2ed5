set -x
python -m pytest tests -m gpu -x -q -k "lu" > gpurun_out/gpu_tests_w8.log 2>&1; tail -3 gpurun_out/gpu_tests_w8.log
python tools/tune.py --workload c3 --batch 10000 --iters 4 --cfg "ws:WS=8" > gpurun_out/tune_w8.log 2>&1; cat gpurun_out/tune_w8.log
