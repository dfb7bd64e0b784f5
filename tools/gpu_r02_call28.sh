#!/bin/bash
# Round 2, GPU call 28 (1 GPU): the tree as it ships after call 27 (N^-1 fold on for the N = 2^15 cluster kernel only): whole GPU suite,
# smoke, ncu --set full of the cluster kernels, bench of both arms with the driver's command lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02ab
{
  echo "== GPU suite"
  s=$(date +%s); timeout 1200 python -m pytest tests -m gpu -q -x > ${O}_pytest_gpu.log 2>&1; echo "rc=$? wall=$(( $(date +%s) - s )) s"; tail -4 ${O}_pytest_gpu.log
  echo "== smoke"
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
  echo "== cluster kernels, N = 2^15 (tree)"
  for i in 1 2; do timeout 300 python tools/kbench.py --bits 64 --degree 32768 --nmoduli 2 --batch 256 2>&1 | tail -1; done
  timeout 300 python tools/kbench.py --bits 64 --degree 32768 --nmoduli 4 --batch 512 2>&1 | tail -1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_cluster -s 4 -c 2 -f -o /tmp/n15 python tools/kbench.py --bits 64 --degree 32768 --nmoduli 2 --batch 256 --iters 2 > /tmp/n15.log 2>&1
  echo "ncu n15 rc=$?"
  python tools/ncu_summary.py /tmp/n15.ncu-rep > ${O}_ncu_n15.txt 2>&1
  python tools/ncu_stalls.py /tmp/n15.ncu-rep >> ${O}_ncu_n15.txt 2>&1
  grep -h -E "^==|time_duration|registers_per_thread|issue_active|inst_executed.sum" ${O}_ncu_n15.txt | cut -c1-150
  echo "== reference arm (driver's command line)"
  s=$(date +%s); timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > ${O}_ref.json 2>/dev/null; echo "rc=$? wall=$(( $(date +%s) - s )) s"; cut -c1-300 ${O}_ref.json
  echo "== bench (N=1)"
  s=$(date +%s); timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > ${O}_bench.json 2> ${O}_bench.err; echo "rc=$? wall=$(( $(date +%s) - s )) s"; cut -c1-300 ${O}_bench.json; tail -2 ${O}_bench.err
  python - <<'EOF'
import json
d = json.load(open('gpurun_out/r02ab_bench.json'))
r = d['roofline']
print('value', d['value'], 'fwd_ms', r['fwd_ms_per_launch'], 'inv_ms', r['inv_ms_per_launch'], 'frac', r['frac'], 'e2e', d['e2e']['value'])
for k, c in d['configs'].items():
    print(k, {x: c[x] for x in c if x.endswith('_ms') or x == 'checked_vs_oracle'})
EOF
} 2>&1 | tee ${O}_call28.log
