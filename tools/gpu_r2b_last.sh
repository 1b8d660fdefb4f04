#!/bin/bash
# last pass of the session: full GPU suite + bench C3 / C2 with the final code
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02b_last_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -9 gpurun_out/r02b_last_pytest_gpu.log
timeout 900 python bench.py --config c3 --steps 3 --warmup 3 > gpurun_out/r02b_last_bench_c3_n1.json 2> gpurun_out/r02b_last_bench_c3_n1.err; echo "bench c3 rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02b_last_bench_c2_n1.json 2> gpurun_out/r02b_last_bench_c2_n1.err; echo "bench c2 rc=$?"
for f in c2 c3; do python - <<PY
import json
d=json.loads(open('gpurun_out/r02b_last_bench_${f}_n1.json').read().strip().splitlines()[-1])
print('${f}', round(d['ms_per_step'],2), 'ms', '%.3g' % d['value'], 'frac', round(d['roofline']['frac'],3), 'whole', round(d['roofline']['whole_step']['frac'],3), 'parity', d['parity_checked'], 'e2e %.3g' % d['e2e']['value'], d['clocks']['reasons'], d['stage_ms'])
PY
done
