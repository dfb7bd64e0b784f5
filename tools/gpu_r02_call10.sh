#!/bin/bash
# Round 2, GPU call 10 (1 GPU): the three-stream ring host pipeline -- parity of the host-buffer tests, chunk x ring sweep, bench e2e.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02j
{
  echo "== host-buffer tests"
  timeout 900 python -m pytest tests -m gpu -x -q -k "host or dropin or round2 or reference_programs" 2>&1 | tail -5
  echo "== e2e sweep (tools/e2e_sweep.py)"
  timeout 900 python tools/e2e_sweep.py 2:4 2:8 4:4 4:8 4:16 8:4 8:8 16:4 16:8 1:16 2>&1 | tee ${O}_e2e_sweep.txt
  echo "== bench (N=1, headline only)"
  timeout 900 python bench.py --headline-only --no-cpu > ${O}_bench.json 2> ${O}_bench.err; echo "rc=$?"; python -c "
import json;d=json.load(open('${O}_bench.json'));print(json.dumps(d['e2e'],indent=1)[:1800])"; tail -3 ${O}_bench.err
} 2>&1 | tee ${O}_call10.log
