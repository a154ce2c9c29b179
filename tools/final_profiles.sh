#!/bin/bash
# End-of-round evidence on ONE GPU (run on the GPU box through gpurun; everything lands in gpurun_out/ with the tag):
# the bench lines, the ncu launch list of the bench command, ncu --set full of the steady-state kernels (single GPU
# and one rank of 8 in loopback), the e2e trace, the C3 bench and the simplex kernel under ncu.
tag=${1:-final}
o=gpurun_out
tools/gpu_check.sh $tag
python bench.py --steps 70 --warmup 5 --no-cpu-baseline > $o/bench_${tag}_70.json 2> $o/bench_${tag}_70.err &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_${tag}.csv \
      python bench.py --steps 70 --warmup 5 --no-cpu-baseline > $o/ncu_launches_${tag}.log 2>&1
python tools/profile_target.py > $o/profile_target_${tag}.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'advect_fused|composite_group' -s 302 -c 2 -f \
      -o $o/steady_${tag} python tools/profile_target.py > $o/ncu_steady_${tag}.log 2>&1
python tools/e2e_trace.py > $o/e2e_trace_${tag}.txt 2>&1; tail -4 $o/e2e_trace_${tag}.txt
python bench.py --workload c3 --steps 20 --warmup 5 > $o/bench_c3_${tag}.json 2> $o/bench_c3_${tag}.err; cut -c1-1200 $o/bench_c3_${tag}.json; tail -2 $o/bench_c3_${tag}.err
ncu --set full --clock-control none --import-source on -k regex:'simplex_octaves|curl_pair' -s 4 -c 2 -f -o $o/c3_${tag} \
    python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $o/ncu_c3_${tag}.log 2>&1
export CLUMPY_ROUTE_LOOPBACK=1 ROUTED_SINGLE_WARM=1200
python tools/routed_single.py 8 64 > $o/routed_single_${tag}.log 2>&1; tail -1 $o/routed_single_${tag}.log
ncu --set full --clock-control none --import-source on -k regex:'advect_route|drain_group|composite_group' -s 465 -c 3 -f -o $o/route_${tag} \
    python tools/routed_single.py 8 64 > $o/ncu_route_${tag}.log 2>&1; tail -2 $o/ncu_route_${tag}.log
ls -la $o | tail -20
