cd /root/repo
ncu --set full --clock-control none --import-source on -k regex:rowlane -s 1 -c 1 -o gpurun_out/rl_w4_c3_3552 -f python tools/prof_lu.py --workload c3 --batch 3552 --iters 2 --check 0 > gpurun_out/ncu_rl_w4.log 2>&1
ncu --set full --clock-control none -k regex:rowlane -s 1 -c 1 -o gpurun_out/rl_w8_c3_8 -f python tools/prof_lu.py --workload c3 --batch 8 --iters 2 --check 0 > gpurun_out/ncu_rl_w8.log 2>&1
ncu --set full --clock-control none -k regex:rowlane -s 1 -c 1 -o gpurun_out/rl_w1_c4 -f python tools/prof_lu.py --workload c4 --batch 11926 --iters 2 --check 0 > gpurun_out/ncu_rl_w1_c4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02c_b1250.csv python bench.py --batch 1250 --steps 2 --warmup 1 --no-secondary --no-cpu > gpurun_out/ncu_launch_b1250.log 2>&1
