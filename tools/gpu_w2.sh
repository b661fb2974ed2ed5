set -x
python tools/tune.py --workload c3 --batch 10000 --iters 4 --cfg "ws:WS=8;ws:WS=16;ws:WS=4;ws:WS=8,WF=96;ws:WS=8,WF=64" > gpurun_out/tune_w2.log 2>&1; cat gpurun_out/tune_w2.log
python tools/prof_lu.py --iters 3 > gpurun_out/prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lu_refactor_wide -s 1 -c 1 -o gpurun_out/rfw_v2 python tools/prof_lu.py --iters 2 > gpurun_out/ncu_rfw_v2.log 2>&1; cat gpurun_out/prof_plain.log; tail -3 gpurun_out/ncu_rfw_v2.log
