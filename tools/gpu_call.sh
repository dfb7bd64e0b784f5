cd /root/repo
timeout 600 python -m pytest tests/test_gaussian.py -m gpu -x -q 2>&1 | tail -8
timeout 120 python tools/gauss_bench.py 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/gauss_launches.csv python tools/gauss_bench.py > gpurun_out/gauss_ncu.log 2>&1
grep -v "^==" gpurun_out/gauss_launches.csv | awk -F'","' '{print $5, $NF}' | tail -8
