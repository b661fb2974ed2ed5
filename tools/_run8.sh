cd /root/repo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --no-secondary > gpurun_out/bench_r02c_n8_c3.json 2> gpurun_out/bench_r02c_n8_c3.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload c4 --steps 10 --warmup 3 --no-secondary > gpurun_out/bench_r02c_n8_c4.json 2> gpurun_out/bench_r02c_n8_c4.err
