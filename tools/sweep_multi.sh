#!/bin/bash
# A/B sweep of the routed multi-GPU path under torchrun: NGPU ranks, optionally a 1/DIV slice of the C4 workload
# (CLUMPY_BENCH_ROWS_DIV: NGPU ranks then see the per-rank working set of an NGPU*DIV-rank run).  Usage:
#   tools/sweep_multi.sh NGPU DIV STEPS "VAR=VAL ..." "VAR=VAL ..." ...      (one quoted environment per variant; "" = defaults)
ngpu=$1; div=$2; steps=$3; shift 3
i=0
for variant in "$@"; do
  i=$((i+1))
  out=gpurun_out/sweep_mg${ngpu}_div${div}_$i
  env CLUMPY_BENCH_ROWS_DIV=$div CLUMPY_BENCH_NO_CLOCKS=1 $variant python -m torch.distributed.run --nnodes=1 --nproc-per-node $ngpu \
      --master-addr 127.0.0.1 --master-port $((29520+i)) bench.py --gpus $ngpu --steps $steps --warmup 5 > $out.json 2> $out.err
  python - "$out.json" "$variant" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    k = d["roofline"]["per_rank_kernel_ms"]
    print(f'{sys.argv[2]!r:60s} ms/step {d["ms_per_step"]:.5f}  route {max(k["route_per_frame"]):.4f} drain {max(k["drain_incl_wait_per_frame"]):.4f} composite {max(k["composite_per_frame"]):.4f}  e2e {d["e2e"]["seconds"]:.4f}s')
except Exception as e:
    print(f'{sys.argv[2]!r}: failed ({e})')
PY
done
