#!/bin/bash
# First GPU call of the next round: the items DESIGN.md section 8.1 lists as CPU-checked only.
#   before calling:  tools/variants.sh 13 lazy2_13 "-DNFLGPU_LAZY64=2";  tools/variants.sh 14 lazy2_14 "-DNFLGPU_LAZY64=2"
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  echo "== reference demo programs against the drop-in header"
  for b in nfllib_demo_main_op1024_60_uint32_t nfllib_demo_main_func1024_60_uint32_t nfllib_demo_main_op8192_124_uint64_t \
           nfllib_demo_main_func8192_124_uint64_t ntt_multi; do
    timeout 600 tests/cpp/_ref/$b > gpurun_out/demo_$b.log 2>&1; echo "$b rc=$?"; tail -3 gpurun_out/demo_$b.log
  done
  echo "== memcheck over the Gaussian sampler tests"
  timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gaussian.py -m gpu -x -q 2>&1 | tail -6
  echo "== select-form top-bit reduction on the 32-coefficient forward kernels"
  N=nfllib_b200/libnflgpu.so
  [ -f build/variants/lazy2_13/libnflgpu.so ] && tools/gpu_variants.sh "--bits 64 --degree 8192 --nmoduli 6 --batch 2048" $N build/variants/lazy2_13/libnflgpu.so
  [ -f build/variants/lazy2_14/libnflgpu.so ] && tools/gpu_variants.sh "--bits 64 --degree 16384 --nmoduli 8 --batch 1024" $N build/variants/lazy2_14/libnflgpu.so
} 2>&1 | tee gpurun_out/next_round_first_call.log
