cd /root/repo
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/bench_r02b_n1_c4.json 2> gpurun_out/bench_r02b_n1_c4.err
tail -c 1500 gpurun_out/bench_r02b_n1_c4.json
