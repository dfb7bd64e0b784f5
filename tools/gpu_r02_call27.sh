#!/bin/bash
# Round 2, GPU call 27 (1 GPU): N^-1 folded into the inverse twiddles (ntt_plan.h plan_fold; 64-bit words) against the old form
# (-DNFLGPU_FOLD=0: one extra Shoup multiplication per butterfly of the last inverse stage), per size; whole GPU suite and the bench
# on the tree (fold on).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02aa
{
  echo "== GPU suite (tree: fold on)"
  s=$(date +%s); timeout 1200 python -m pytest tests -m gpu -q -x > ${O}_pytest_gpu.log 2>&1; echo "rc=$? wall=$(( $(date +%s) - s )) s"; tail -4 ${O}_pytest_gpu.log
  echo "== fold vs nofold (forward kernels are identical in both builds)"
  kb() { for v in nofold$1 fold$1 nofold$1 fold$1; do timeout 300 python tools/kbench.py $2 --lib build/variants/$v/libnflgpu.so 2>&1 | tail -1; done; }
  echo "# C2 u64 N=1024 M=4 batch=4096";   kb 10 "--bits 64 --degree 1024 --nmoduli 4 --batch 4096"
  echo "# u64 N=4096 M=4 batch=1024";      kb 12 "--bits 64 --degree 4096 --nmoduli 4 --batch 1024"
  echo "# C5 u64 N=8192 M=6 batch=2048";   kb 13 "--bits 64 --degree 8192 --nmoduli 6 --batch 2048"
  echo "# C3 u64 N=16384 M=8 batch=1024";  kb 14 "--bits 64 --degree 16384 --nmoduli 8 --batch 1024"
  echo "# u64 N=32768 M=2 batch=256 (cluster)"; kb 15 "--bits 64 --degree 32768 --nmoduli 2 --batch 256"
  echo "== bench (N=1, tree)"
  s=$(date +%s); timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > ${O}_bench.json 2> ${O}_bench.err; echo "rc=$? wall=$(( $(date +%s) - s )) s"; cut -c1-300 ${O}_bench.json; tail -2 ${O}_bench.err
  python - <<'EOF'
import json
d = json.load(open('gpurun_out/r02aa_bench.json'))
r = d['roofline']
print('value', d['value'], 'fwd_ms', r['fwd_ms_per_launch'], 'inv_ms', r['inv_ms_per_launch'], 'e2e', d['e2e']['value'])
for k, c in d['configs'].items():
    print(k, {x: c[x] for x in c if x.endswith('_ms') or x == 'checked_vs_oracle'})
EOF
} 2>&1 | tee ${O}_call27.log
