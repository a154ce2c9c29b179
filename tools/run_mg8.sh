python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/check_multigpu.py > gpurun_out/check_mg8.log 2>&1; tail -9 gpurun_out/check_mg8.log | cut -c1-150
for K in 20 2000; do
  CLUMPY_BENCH_VERBOSE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps $K --warmup 5 > gpurun_out/bench_mg8_$K.json 2> gpurun_out/bench_mg8_$K.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_mg8_$K.json')); k=d['roofline']['per_rank_kernel_ms']
print('K=$K ms/step', d['ms_per_step'], 'steady', d['steady_ms_per_step'], 'e2e', d['e2e']['seconds'])
print(' route', [round(x,4) for x in k['route_per_frame']]); print(' drain', [round(x,4) for x in k['drain_incl_wait_per_frame']]); print(' comp ', [round(x,4) for x in k['composite_per_frame']])
" || tail -5 gpurun_out/bench_mg8_$K.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --workload c3 --steps 20 --warmup 5 > gpurun_out/bench_c3_mg8.json 2> gpurun_out/bench_c3_mg8.err; cat gpurun_out/bench_c3_mg8.json | cut -c1-900; tail -3 gpurun_out/bench_c3_mg8.err
