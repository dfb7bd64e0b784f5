cd /root/repo
N="nfllib_b200/libnflgpu.so"
tools/gpu_variants.sh "--bits 64 --degree 1024 --nmoduli 4 --batch 4096" $N build/variants/old10/libnflgpu.so $N build/variants/old10/libnflgpu.so
tools/gpu_variants.sh "--bits 64 --degree 8192 --nmoduli 6 --batch 2048" $N build/variants/old13/libnflgpu.so
tools/gpu_variants.sh "--bits 64 --degree 16384 --nmoduli 8 --batch 1024" $N build/variants/old14/libnflgpu.so
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01e_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r01e_pytest_gpu.log; tail -5 gpurun_out/r01e_pytest_gpu.log
