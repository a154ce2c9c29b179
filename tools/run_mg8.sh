#!/bin/bash
# The multi-GPU lines on an 8-GPU box (through gpurun --gpus 8): parity over NVLink at 8 ranks, then bench.py at
# N = 4 (20 steps) and N = 8 (20 and 2000 steps), results under gpurun_out/.
run() { # N K port
  CLUMPY_BENCH_VERBOSE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus $1 --steps $2 --warmup 5 > gpurun_out/bench_mg$1_$2.json 2> gpurun_out/bench_mg$1_$2.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_mg$1_$2.json')); k=d['roofline']['per_rank_kernel_ms']
print('N=$1 K=$2 ms/step', round(d['ms_per_step'],5), 'steady', round(d['steady_ms_per_step'],5), 'e2e', round(d['e2e']['seconds'],4))
print(' route', [round(x,4) for x in k['route_per_frame']]); print(' drain', [round(x,4) for x in k['drain_incl_wait_per_frame']]); print(' comp ', [round(x,4) for x in k['composite_per_frame']])
" || tail -5 gpurun_out/bench_mg$1_$2.err
}
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/check_multigpu.py > gpurun_out/check_mg8.log 2>&1; tail -9 gpurun_out/check_mg8.log | cut -c1-150
run 4 20 29512
run 8 20 29513
run 8 2000 29514
grep "bench rank 0" gpurun_out/bench_mg8_20.err | grep "e2e +" | tail -5
