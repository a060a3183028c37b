#!/bin/bash
# the driver's scaling run in small: bench.py under torchrun at N = 2, 4 and all GPUs of the box, default flags
TAG=${1:-r02s}
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for N in 2 4 $NG; do
  [ $N -le $NG ] || continue
  ( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err
  python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/bench_${N}gpu_$TAG.json").read().strip().splitlines()[-1])
    g=l.get("gather") or {}
    print("N", l["n_gpus"], "value %.4g ms %.4f e2e %.4g" % (l["value"], l["ms_per_step"], l["e2e"]["value"]), "gather fused %.4f nccl %.4f equal %s" % (g.get("fused_ms_per_step",0), g.get("nccl_all_gather_ms_per_step") or 0, g.get("fused_equals_nccl_bit_for_bit")), l["clocks"]["sm_mhz"], l["clocks"]["samples"])
    for k,v in (l.get("strong") or {}).items():
        print("  strong", k, "fused %.4f local %.4f nccl %.4f equal %s" % (v["ms_per_step"], v["ms_per_step_local_rows_only"], v["ms_per_step_nccl_all_gather_after_kernel"] or 0, v["fused_equals_nccl_gather_bit_for_bit"]))
except Exception as e:
    print("N $N failed", e)
PY
  tail -2 gpurun_out/bench_${N}gpu_$TAG.err | cut -c1-300
done
