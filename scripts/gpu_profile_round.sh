#!/bin/bash
# Run ON THE GPU BOX (gpurun): collects the profile evidence of one build into gpurun_out/ under the tag $1 (e.g. r2).
#   launch list of one denoising step (time + DRAM bytes per launch), `ncu --set full` captures of the dominant
#   kernels with source, and the headline bench lines.  Post-process here with scripts/summarize_profiles.py.
tag=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
# 1. every launch of one step over one 1024-row chunk (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/${tag}_launches_1024rows.csv python bench.py --steps 1 --warmup 3 --batch 512 --max-rows 1024 --no-cpu-baseline \
    > gpurun_out/${tag}_ncu_bench.log 2>&1
echo "== launch list: $(wc -l < gpurun_out/${tag}_launches_1024rows.csv) lines"
# 2. full captures (one launch each, with source)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_gemm|groupnorm_kernel|attention' \
    -o gpurun_out/${tag}_kernels python scripts/prof_kernels.py --once > gpurun_out/${tag}_ncu_kernels.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -c 3 \
    -o gpurun_out/${tag}_kernels_smallk python scripts/prof_kernels.py --small --once > gpurun_out/${tag}_ncu_smallk.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sampler_step -c 1 \
    -o gpurun_out/${tag}_sampler_step python scripts/prof_kernels.py --sampler --once > gpurun_out/${tag}_ncu_sampler.log 2>&1
ls -la gpurun_out/${tag}_*.ncu-rep
# 3. isolated CUDA-event timings of the same launches (not under the profiler)
timeout 300 python scripts/prof_kernels.py > gpurun_out/${tag}_kernel_timings.txt 2>&1
timeout 120 python scripts/prof_kernels.py --sampler >> gpurun_out/${tag}_kernel_timings.txt 2>&1
tail -n 20 gpurun_out/${tag}_kernel_timings.txt
# 4. the training slice's backward kernels at the configs[4] batch (128): time + DRAM bytes per launch
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_backward_launches.csv python scripts/prof_kernels.py --backward --once > gpurun_out/${tag}_ncu_backward.log 2>&1
echo "== backward launch list: $(wc -l < gpurun_out/${tag}_backward_launches.csv) lines"
