#!/bin/bash
# bench.py at N = 2 and N = 4 on a 4-GPU box (through gpurun --gpus 4): the 2000-step lines and the 2-GPU 20-step line.
run() { # N K port
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus $1 --steps $2 --warmup 5 > gpurun_out/bench_mg$1_$2.json 2> gpurun_out/bench_mg$1_$2.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_mg$1_$2.json'))
print('N=$1 K=$2 ms/step', round(d['ms_per_step'],5), 'steady', round(d['steady_ms_per_step'],5), 'e2e', round(d['e2e']['seconds'],4))" || tail -5 gpurun_out/bench_mg$1_$2.err
}
run 2 2000 29521
run 4 2000 29522
run 2 20 29523
