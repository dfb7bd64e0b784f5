#!/bin/bash
# Round 2, GPU call 34 (1 GPU): the staging-copy pool at its new default size (12 of 16 host threads): host-buffer tests, launch stress (host ring),
# pageable rate, bench headline.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  echo "== host-buffer tests"
  timeout 300 python -m pytest tests -m gpu -q -x -k "host or dropin or cpp" 2>&1 | tail -3
  echo "== launch stress (includes the host ring on pageable arrays)"
  timeout 200 tests/cpp/sched_stress 16 2>&1 | tail -12
  echo "== pageable rate"
  timeout 200 python tools/pageable_sweep.py --threads default,default
  echo "== bench (headline only)"
  timeout 300 python bench.py --headline-only --no-cpu > gpurun_out/r02af_bench.json 2> gpurun_out/r02af_bench.err; echo "rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/r02af_bench.json'));e=d['e2e']
print('value',d['value'],'e2e',e['value'],'pageable',e['pageable']['value'],'registered',e['pageable_registered']['value'],'ceiling',e['copy_only_ceiling']['value'],'lat',e['single_poly_latency_us'])"
} 2>&1 | tee gpurun_out/r02af_call34.log
