#!/usr/bin/env bash
# Round-2 GPU session V (1 GPU): does the RK4 kernel hold its rate in a sustained run? 300 and 3000 timed steps of the
# default workload with SM clock / power / throttle-reason samples every 100 ms beside them.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=timestamp,clocks.sm,power.draw,power.limit,temperature.gpu,clocks_event_reasons.active --format=csv -lms 100 > gpurun_out/r2v_smi.csv 2>&1 &
SMI=$!
for k in 300 3000; do
  python bench.py --steps $k --warmup 20 --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/r2v_bench_$k.json 2> gpurun_out/r2v_bench_$k.err; echo "steps $k rc $?"
  python -c "import json;d=json.loads(open('gpurun_out/r2v_bench_$k.json').read().strip().splitlines()[-1]);print(d['steps'],d['ms_per_step'],d['clocks'])"
done
kill $SMI
wc -l gpurun_out/r2v_smi.csv
