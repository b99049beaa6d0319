#!/bin/bash
# Round 2, final build, multi-GPU (gpurun --gpus N): ONE bench.py run under torchrun (strong headline + weak + C4/C5 shares in the same line).
N=${1:-2}
mkdir -p gpurun_out; out=gpurun_out/r2b_multi_$N.txt; : > $out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $T bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2b_${N}gpu.json 2> gpurun_out/bench_r2b_${N}gpu.err
echo "bench $N gpus rc=$?" >> $out
python - $N >> $out 2>&1 <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open(f'gpurun_out/bench_r2b_{n}gpu.json').read().strip().splitlines()[-1])
print('strong K=30', 'n_gpus', d['n_gpus'], 'value', round(d['value'], 1), 'ms/it', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1),
      'sustained', d.get('sustained') and round(d['sustained']['value'], 1), 'weak', d.get('weak_scaling') and round(d['weak_scaling']['value'], 1),
      'batch/gpu', d['config']['batch_per_gpu'])
if d.get('other_configs'):
    print('   others', {k: (round(v['ms_per_step'], 4), round(v['value'], 1)) for k, v in d['other_configs'].items()})
PY
tail -3 gpurun_out/bench_r2b_${N}gpu.err >> $out
cat $out
