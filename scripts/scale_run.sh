#!/bin/bash
# The driver's SCALE sequence on one 8-GPU box: reference arm once, then bench.py at N = 1, 2, 4, 8 back to back (N > 1 under
# torchrun), plus the in-process engine's strong scaling (cuspmm_mgpu_*) with the peer-store gather.  Outputs under gpurun_out/.
set -u
OUT=${1:-gpurun_out}
STEPS=${STEPS:-10}
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/scale_ref.json 2> $OUT/scale_ref.err
for n in 1 2 4 8; do
  if [ $n = 1 ]; then
    timeout 400 python bench.py --gpus 1 --steps $STEPS --warmup 3 --no-extras > $OUT/scale_n1.json 2> $OUT/scale_n1.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps $STEPS --warmup 3 > $OUT/scale_n$n.json 2> $OUT/scale_n$n.err
  fi
  echo "bench N=$n rc=$?"
done
timeout 300 python scripts/mgpu_strong.py --workload large_25605 --gather > $OUT/mgpu_strong_25605.jsonl 2> $OUT/mgpu_strong.err
timeout 300 python scripts/mgpu_strong.py --workload large_20000 --gather > $OUT/mgpu_strong_20000.jsonl 2>> $OUT/mgpu_strong.err
python - <<PY
import json
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open("$OUT/scale_n%d.json" % n).read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        print(n, "value %.0f GF" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e ms", e.get("ms_per_step"), "parity", (d.get("parity") or {}).get("ok"),
              "x cusparse", (d.get("cusparse") or {}).get("speedup_vs_cusparse"), "kernel", d["roofline"].get("kernel"))
    except Exception as ex:
        print(n, "failed", ex)
PY
cat $OUT/mgpu_strong_25605.jsonl $OUT/mgpu_strong_20000.jsonl
