#!/bin/bash
# Round 2, second half, final one-GPU pass: full GPU test suite, bench C2 (default) / C3 / C1, reference arm, launch lists
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02b_final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02b_final_pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02b_final_bench_c2_n1.json 2> gpurun_out/r02b_final_bench_c2_n1.err; echo "bench c2 rc=$?"
timeout 900 python bench.py --config c3 --steps 3 --warmup 3 > gpurun_out/r02b_final_bench_c3_n1.json 2> gpurun_out/r02b_final_bench_c3_n1.err; echo "bench c3 rc=$?"
timeout 900 python bench.py --config c1 --steps 5 --warmup 3 > gpurun_out/r02b_final_bench_c1_n1.json 2> gpurun_out/r02b_final_bench_c1_n1.err; echo "bench c1 rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/r02b_final_ref_c2.json 2> gpurun_out/r02b_final_ref_c2.err; echo "ref c2 rc=$?"
bash tools/gpu_launchlist_cfg.sh c2 100000000 r02b_final_c2 "--sort-mode 3" > gpurun_out/r02b_final_c2_summary.txt 2>&1
bash tools/gpu_launchlist_cfg.sh c3 100000000 r02b_final_c3_100m > gpurun_out/r02b_final_c3_100m_summary.txt 2>&1
for f in c2 c3 c1; do python - <<PY
import json
d=json.loads(open('gpurun_out/r02b_final_bench_${f}_n1.json').read().strip().splitlines()[-1])
print('${f}', round(d['ms_per_step'],2), 'ms', '%.3g' % d['value'], 'frac', round(d['roofline']['frac'],3), 'whole', round(d['roofline']['whole_step']['frac'],3), 'parity', d['parity_checked'], 'e2e %.3g' % d['e2e']['value'], 'payload', round(d['payload'].get('ms_per_step',0),2), 'cpu %.3g' % d['cpu_baseline']['value'], d['clocks']['reasons'], d['sort'])
PY
done
