#!/bin/bash
# Round 2, multi-GPU call (gpurun --gpus N): bench.py under torchrun on N GPUs (strong headline + weak in the same run), the
# reference arm under the same launcher, and on GPU 0 the parity suite.
N=${1:-2}
mkdir -p gpurun_out; out=gpurun_out/r2_multi_$N.txt; : > $out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $T bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err
echo "bench $N gpus rc=$?" >> $out
timeout 600 $T bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline --no-extras --residual-iters 0 > gpurun_out/bench_r2_${N}gpu_k100.json 2>> gpurun_out/bench_r2_${N}gpu.err
echo "bench $N gpus K=100 rc=$?" >> $out
timeout 600 $T bench.py --gpus $N --scaling weak --steps 30 --warmup 5 --no-cpu-baseline --no-extras --residual-iters 0 > gpurun_out/bench_r2_${N}gpu_weak.json 2>> gpurun_out/bench_r2_${N}gpu.err
echo "bench $N gpus weak rc=$?" >> $out
timeout 300 $T bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_r2_${N}gpu_reference.json 2>> gpurun_out/bench_r2_${N}gpu.err
echo "reference arm rc=$? lines=$(wc -l < gpurun_out/bench_r2_${N}gpu_reference.json)" >> $out
python - $N >> $out 2>&1 <<'PY'
import json, sys
n = sys.argv[1]
for tag in ("", "_k100", "_weak"):
    try:
        d = json.loads(open(f'gpurun_out/bench_r2_{n}gpu{tag}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(tag, 'unreadable', e); continue
    print(tag or 'strong K=30', 'scaling', d['scaling'], 'n_gpus', d['n_gpus'], 'value', round(d['value'], 1), 'ms/it', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1),
          'sustained', d.get('sustained') and round(d['sustained']['value'], 1), 'weak', d.get('weak_scaling') and round(d['weak_scaling']['value'], 1),
          'batch/gpu', d['config']['batch_per_gpu'])
    if d.get('other_configs'):
        print('   others', {k: (round(v['ms_per_step'], 4), round(v['value'], 1)) for k, v in d['other_configs'].items()})
PY
if [ "$2" = "tests" ]; then
    rm -f gpurun_out/parity_measured.jsonl
    CUDA_VISIBLE_DEVICES=0 timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/tests_r2_multi.log 2>&1
    echo "tests rc=$?  $(tail -1 gpurun_out/tests_r2_multi.log)" >> $out
    grep -E "^FAILED|^E  " gpurun_out/tests_r2_multi.log | cut -c1-300 | head -30 >> $out
fi
tail -5 gpurun_out/bench_r2_${N}gpu.err >> $out
cat $out
