#!/bin/bash
# usage: tools/bench_kernels.sh [label]  -- short bench run, prints ms/iteration and the per-kernel live timings
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --residual-iters 0 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$1', 'ms/it', round(d['ms_per_step'], 4), 'kernels', d['kernels_per_iteration'], 'clk', d['clocks']['sm_mhz'])
for r in [d['roofline']] + d['roofline_kernels']:
    print('   %-70s %7.1f us  %.3f' % (r['kernel'][:70], r['ms_per_launch'] * 1e3, r['frac']))
"
