#!/bin/bash
# One GPU-box visit of the development loop: GPU suite, all-config kernel table, bench line, ncu launch list + full capture.
# Results land in gpurun_out/ (copy what should be judged into profiles/).
cd "$(dirname "$0")/.."
TAG=${1:-r01e}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python tools/kbench_all.py > gpurun_out/${TAG}_kbench_all.txt 2>&1
cat gpurun_out/${TAG}_kbench_all.txt
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt_ -s 4 -c 2 -f -o gpurun_out/${TAG}_prof_ntt \
    python tools/kbench.py --iters 2 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out
