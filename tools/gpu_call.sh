cd /root/repo
N="nfllib_b200/libnflgpu.so"
tools/gpu_variants.sh "--bits 64 --degree 1024 --nmoduli 4 --batch 4096" $N build/variants/nopf10/libnflgpu.so $N build/variants/nopf10/libnflgpu.so
tools/gpu_variants.sh "--bits 64 --degree 8192 --nmoduli 6 --batch 2048" $N build/variants/nopf13/libnflgpu.so
tools/gpu_variants.sh "--bits 64 --degree 16384 --nmoduli 8 --batch 1024" $N build/variants/nopf14/libnflgpu.so
tools/gpu_variants.sh "--bits 32 --degree 4096 --nmoduli 14 --batch 2048" $N build/variants/nopf12_32/libnflgpu.so
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
