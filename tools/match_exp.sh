#!/usr/bin/env bash
# Local helper: one GPU call = matcher parity tests + the matching leg of the bench; prints pairs/s.
#   tools/match_exp.sh <tag>
TAG=$1
tools/gpu_retry.sh gpurun_out/${TAG}_call.log --timeout 420 -- "timeout 200 python -m pytest tests/test_match_gpu.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/${TAG}_tests.txt; timeout 200 python bench.py --path match --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_match.json 2> gpurun_out/${TAG}_match.err; cat gpurun_out/${TAG}_tests.txt; python -c \"
import json; d=json.load(open('gpurun_out/${TAG}_match.json')); print('VALUE', d['value'], 'E2E', d['e2e']['value'], 'COMPAT', d['per_pair_compat']['value'])\"; tail -3 gpurun_out/${TAG}_match.err"
grep -E "passed|failed|VALUE|charged|left" gpurun_out/${TAG}_call.log
