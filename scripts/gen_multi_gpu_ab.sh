#!/bin/bash
# gen_multi_gpu_ab.sh <n_gpus> [total] [batch]: static slices vs the dynamic batch queue of v_diffusion_b200.generate on one box
# (CIFAR-10 cond UNet, random init, CFG w=1, 100-step DDIM); writes gpurun_out/generate_{static,dynamic}_<n>gpu.json
N=${1:-8}; TOTAL=${2:-24576}; BS=${3:-128}
mkdir -p gpurun_out /tmp/vdt_cfg
python - <<'PY'
import json
m = json.load(open("tests/golden/merged_configs.json"))["cifar10_cond"]
json.dump({"data": {"name": "cifar10"}, "model": m["model"], "diffusion": m["diffusion"], "conditional": m["conditional"]},
          open("/tmp/vdt_cfg/cfg.json", "w"))
json.dump({}, open("/tmp/vdt_cfg/defaults.json", "w"))
PY
for s in static dynamic; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    -m v_diffusion_b200.generate --config-path /tmp/vdt_cfg/cfg.json --default-config-path /tmp/vdt_cfg/defaults.json \
    --use-ddim --sample-timesteps 100 --w-guide 1.0 --batch-size $BS --total-size $TOTAL --schedule $s --warmup-batches 1 \
    2> gpurun_out/generate_${s}.err > /dev/null
  grep '"schedule"' gpurun_out/generate_${s}.err | tail -1 | tee gpurun_out/generate_${s}_${N}gpu.json
done
