#!/usr/bin/env python3
"""The atomic leg of the splat roofline on ONE B200: updates per second of the reductions the
advect kernel issues -- red.global.add.noftz.f16x2 (half cells, the C4 path) and red.global.add.u32 -- for
16 777 216 updates per launch (one per C4 point) on 67 MB / 134 MB regions (one half / u32 counter image at
8192 x 4096), uniform random, tile-ordered ("local") and on a hot spot.  One JSON object per line.

    python tools/atomic_peak.py [--quick]

bench.py calls the same library entry (clumpy_cuda_atomic_peak) for the pattern and size of its workload."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import clumpy_b200  # noqa: E402

OPS = {0: "red.global.add.noftz.f16x2", 1: "red.global.add.u32"}
PATTERNS = {0: "uniform random", 1: "hot spot", 2: "local (tile order)"}


def main():
    quick = "--quick" in sys.argv
    n = 1 << (22 if quick else 24)
    with clumpy_b200.Context(0) as ctx:
        for op in (0, 1):
            for region_mb in (67, 134):
                region = region_mb * 1000 * 1000 // 256 * 256
                for pattern, hot in ((0, 1), (2, 1), (1, 1), (1, 32), (1, 4096)):
                    ms = ctx.atomic_peak(op, pattern, region, n, hot_cells=hot, iters=3 if quick else 10)
                    print(json.dumps({"op": OPS[op], "pattern": PATTERNS[pattern] + (f" ({hot} cells)" if pattern == 1 else ""),
                                      "region_MB": region_mb, "updates_per_launch": n, "ms_per_launch": ms,
                                      "G_updates_per_s": n / (ms * 1e-3) / 1e9 if ms > 0 else None}), flush=True)


if __name__ == "__main__":
    main()
