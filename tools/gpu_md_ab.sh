#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batches.py tests/test_gpu_sharded.py tests/test_gpu_multi_process_one.py -q -k "MIN_DISTANCE or min_distance" 2>&1 | tail -3
python bench.py --config c4 --points 125000000 --steps 2 --warmup 1 --no-e2e --no-parity --no-cpu-baseline --no-payload 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms']['ms_sample'])"
bash tools/gpu_launchlist_cfg.sh c4 125000000 r02_c4_125m_v4 | grep -E "md_|launches,"
