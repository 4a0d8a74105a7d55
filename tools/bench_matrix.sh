#!/bin/bash
# Per-shape bench lines at N GPUs of one box (north_star: "throughput of each named shape at 1, 2, 4 and 8 GPUs").
#   tools/bench_matrix.sh N [workloads...]     -> gpurun_out/r2_bench_<workload>_<N>gpu.json
N=${1:-1}; shift
WL=${@:-"c3 c4"}
mkdir -p gpurun_out
for w in $WL; do
  steps=30; [ "$w" = "c5" ] && steps=2
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --workload $w --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${w}_${N}gpu.json 2> gpurun_out/r2_bench_${w}_${N}gpu.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --workload $w --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${w}_${N}gpu.json 2> gpurun_out/r2_bench_${w}_${N}gpu.err
  fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_${w}_${N}gpu.json").read().strip().splitlines()[-1])
    print("$w", "N=$N", "value %.4g %s" % (d["value"], d["unit"]), "ms/step %.4g" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"],
          "replicas_identical", d.get("replicas_bit_identical"))
except Exception as e:
    print("$w N=$N FAILED", e)
PY
done
