mkdir -p gpurun_out; out=gpurun_out/ab100.txt; : > $out
for rep in 1 2 3; do for p in 0 2; do
HELMNET_PDL=$p python bench.py --steps 100 --warmup 5 --no-cpu-baseline --residual-iters 0 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('pdl=$p steps=100 ms/it', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'],1), 'clk', d['clocks'])
" >> $out 2>&1
done; done
for p in 0 2; do
HELMNET_PDL=$p python bench.py --steps 400 --warmup 5 --no-cpu-baseline --residual-iters 0 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('pdl=$p steps=400 ms/it', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'],1), 'clk', d['clocks'])
" >> $out 2>&1
done
cat $out
