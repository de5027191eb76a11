#!/bin/bash
# Runs the GPU parity tests group by group in separate processes (a trapped kernel poisons its CUDA
# context), each under its own timeout; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, timeout, pytest args...
  local name=$1 to=$2; shift 2
  timeout "$to" python -m pytest -q -m gpu -p no:cacheprovider "$@" > "gpurun_out/$name.log" 2>&1
  echo "== $name: exit $?"; tail -n 25 "gpurun_out/$name.log"
}
run conv 300 tests/test_gpu_kernels.py -k conv_gemm
run gn 300 tests/test_gpu_kernels.py -k 'test_groupnorm'
run gnfused 300 tests/test_gpu_kernels.py -k single_pass
run attn 300 tests/test_gpu_kernels.py -k attention
run sampler 300 tests/test_gpu_kernels.py -k sampler_step
run unet 600 tests/test_gpu_unet.py -s
