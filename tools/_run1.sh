cd /root/repo
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "rowlane_refactor_geometries" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02c_n1_c3.json 2> gpurun_out/bench_r02c_n1_c3.err; tail -c 600 gpurun_out/bench_r02c_n1_c3.json
python bench.py --batch 1250 --steps 20 --warmup 3 --no-secondary --no-cpu > gpurun_out/bench_r02c_n1_c3_b1250.json 2> gpurun_out/bench_r02c_n1_c3_b1250.err
python bench.py --workload c4 --steps 5 --warmup 3 --no-secondary > gpurun_out/bench_r02c_n1_c4.json 2> gpurun_out/bench_r02c_n1_c4.err
python bench.py --workload c4 --batch 1496 --steps 10 --warmup 3 --no-secondary --no-cpu > gpurun_out/bench_r02c_n1_c4_b1496.json 2> gpurun_out/bench_r02c_n1_c4_b1496.err
python bench.py --workload c2 --steps 20 --warmup 3 --no-secondary > gpurun_out/bench_r02c_n1_c2.json 2> gpurun_out/bench_r02c_n1_c2.err
python tools/latency.py > gpurun_out/latency_r02c.txt 2>&1
