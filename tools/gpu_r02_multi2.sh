#!/bin/bash
# multi-GPU call: bench (no CPU legs) with the gather / strong legs, with and without padded gather rows; sharded check
TAG=${1:-r02n}
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for PAD in 1 0; do
( time WLSQM_GATHER_PAD=$PAD timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 2951$PAD bench.py --gpus $NG --steps 20 --warmup 5 ) > gpurun_out/bench_${NG}gpu_pad${PAD}_$TAG.json 2> gpurun_out/bench_${NG}gpu_pad${PAD}_$TAG.err
python - <<PY
import json
l=json.loads(open("gpurun_out/bench_${NG}gpu_pad${PAD}_$TAG.json").read().strip().splitlines()[-1])
g=l.get("gather") or {}
print("pad $PAD N", l["n_gpus"], "value %.4g ms %.4f" % (l["value"], l["ms_per_step"]), "gather fused %.4f nccl %.4f equal %s" % (g.get("fused_ms_per_step",0), g.get("nccl_all_gather_ms_per_step",0), g.get("fused_equals_nccl_bit_for_bit")))
for k,v in (l.get("strong") or {}).items():
    print("  strong", k, "fused %.4f local %.4f nccl %.4f equal %s err %s" % (v["ms_per_step"], v["ms_per_step_local_rows_only"], v["ms_per_step_nccl_all_gather_after_kernel"], v["fused_equals_nccl_gather_bit_for_bit"], v["fused_gather_error"]))
PY
tail -3 gpurun_out/bench_${NG}gpu_pad${PAD}_$TAG.err | cut -c1-300
done
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 benchmarks/multi_gpu_check.py ) > gpurun_out/multi_gpu_check_${NG}gpu_$TAG.json 2>&1; tail -3 gpurun_out/multi_gpu_check_${NG}gpu_$TAG.json | cut -c1-900
