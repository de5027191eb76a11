#!/bin/bash
# Run ON THE GPU BOX: the whole evidence set of the final build of a round (tag $1) into gpurun_out/.
tag=${1:-r2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "== smoke exit $?"
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/${tag}_gputest.log 2>&1; echo "== gpu tests exit $?"; tail -n 3 gpurun_out/${tag}_gputest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_batch4096.json 2> gpurun_out/${tag}_bench.err; echo "== bench exit $?"
timeout 900 python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "== reference arm exit $?"
for wl in celeba cifar10_uncond mnist28; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_${wl}.json 2> gpurun_out/${tag}_bench_${wl}.err; echo "== bench $wl exit $?"
done
# the composed training step: parity cases one by one, then three steps at batch 128 with wall-clock per step
for spec in "graph_parity cifar_cond 4 fp16" "graph_parity small 3 bf16" "train_steps small 8 fp16" "dropout small 4 fp16" "grad_golden small 0 fp16" "grad_golden cifar 0 fp16" "autograd_step small 4 fp16"; do
  timeout 600 python -m tests.train_step_worker $spec > gpurun_out/${tag}_train_$(echo $spec | tr " " _).log 2>&1; echo "== train_step_worker $spec exit $?"
done
timeout 300 python scripts/quick_train_step.py 128 > gpurun_out/${tag}_quick_train_step_b128.log 2>&1; echo "== quick_train_step exit $?"
bash scripts/gpu_profile_round.sh $tag
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("value"), (d.get("e2e") or {}).get("value"), (d.get("clocks") or {}).get("sm_mhz"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "unreadable", e)
PY
