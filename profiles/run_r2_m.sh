#!/usr/bin/env bash
# Round-2 GPU session M (1 GPU): last check of the final tree: extended-physics tests, smoke, the default bench line
# (now with the prim2048x64_ext sub-line) and the reference arm.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -k "extended or primitive or step_host" -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2m_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2m_pytest.log
tail -4 gpurun_out/r2m_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1; echo "smoke rc $?"
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2m_bench_default.json 2> gpurun_out/r2m_bench_default.err; echo "default bench rc $?"; tail -3 gpurun_out/r2m_bench_default.err
python -c "
import json
d=json.loads(open('gpurun_out/r2m_bench_default.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['fp32_pipe'])
for o in d['other_configs']+[d['strong_scaling']]: print(o.get('name'), o.get('ms_per_step'), o.get('roofline',{}).get('frac'), o.get('kernel'), o.get('error'))"
