#!/bin/bash
# The standard single-GPU check of a change (run on the GPU box through gpurun): the gpu tests, the driver's 20-step
# bench window and the default 2000-step bench; results under gpurun_out/ with the given tag.
tag=${1:-check}
python -m pytest tests -m gpu -q > gpurun_out/pytest_$tag.log 2>&1; tail -12 gpurun_out/pytest_$tag.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${tag}_20.json 2> gpurun_out/bench_${tag}_20.err; tail -2 gpurun_out/bench_${tag}_20.err
python bench.py --steps 2000 --no-cpu-baseline > gpurun_out/bench_${tag}_2000.json 2> gpurun_out/bench_${tag}_2000.err; tail -2 gpurun_out/bench_${tag}_2000.err
python - <<PY
import json
for f in ("gpurun_out/bench_${tag}_20.json", "gpurun_out/bench_${tag}_2000.json"):
    try:
        d = json.load(open(f)); r = d["roofline"]
    except Exception as e:
        print(f, "unreadable:", e); continue
    print(f, "ms/step", round(d["ms_per_step"], 4), "steady", round(d["steady_ms_per_step"], 4), "frac", round(r["frac"], 3), "steady_frac", round(r["steady_frac"], 3),
          "e2e", f'{d["e2e"]["value"]:.3e}', round(d["e2e"]["seconds"], 4), "parity", d.get("parity") and (d["parity"]["within1"], d["parity"]["max"]))
    print("   ", {k: (round(v["launch_ms"], 4), round(v["launch_ms_alone"], 4), round(v["frac"], 3)) for k, v in r["stages"].items()})
PY
