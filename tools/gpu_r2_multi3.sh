#!/bin/bash
# Round 2, very last build: the same bench.py command at N GPUs (N = 1: plain python, N > 1: torchrun), 30 iterations after 5 warm-up.
N=${1:-1}
mkdir -p gpurun_out; out=gpurun_out/r2c_multi_$N.txt; : > $out
if [ "$N" = "1" ]; then T="python"; else T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
timeout 900 $T bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_r2c_${N}gpu.json 2> gpurun_out/bench_r2c_${N}gpu.err
echo "bench $N gpus rc=$?" >> $out
python - $N >> $out 2>&1 <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open(f'gpurun_out/bench_r2c_{n}gpu.json').read().strip().splitlines()[-1])
print('strong K=30', 'n_gpus', d['n_gpus'], 'value', round(d['value'], 1), 'ms/it', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1),
      'sustained', d.get('sustained') and round(d['sustained']['value'], 1), 'weak', d.get('weak_scaling') and round(d['weak_scaling']['value'], 1),
      'batch/gpu', d['config']['batch_per_gpu'])
if d.get('other_configs'):
    print('   others', {k: (round(v['ms_per_step'], 4), round(v['value'], 1)) for k, v in d['other_configs'].items()})
PY
cat $out
