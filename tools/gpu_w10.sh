set -x
for wl in c4 c2; do
  timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_${wl}_wide.json 2> gpurun_out/bench_${wl}_wide.err; python -c "import json;d=json.load(open('gpurun_out/bench_${wl}_wide.json'));print('$wl wide',d['value'],d['config']['kernel_ms'])"
  CSP3_WIDE=0 timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_${wl}_v3.json 2> gpurun_out/bench_${wl}_v3.err; python -c "import json;d=json.load(open('gpurun_out/bench_${wl}_v3.json'));print('$wl v3',d['value'],d['config']['kernel_ms'])"
done
