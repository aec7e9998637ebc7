G=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
if [ "$2" = full ]; then
  timeout 900 $TR --master-port 29531 bench.py --gpus $G --steps 5 --warmup 3 > gpurun_out/r3_bench_gl_2p20x135_n${G}_ring.json 2> gpurun_out/r3_n${G}_ring.err
else
  timeout 600 $TR --master-port 29532 bench.py --gpus $G --steps 5 --warmup 3 --no-e2e --no-library-multi --no-cpu-baseline --no-fri --no-latency > gpurun_out/r3_bench_gl_2p20x135_n${G}_ring.json 2> gpurun_out/r3_n${G}_ring.err
fi
python - gpurun_out/r3_bench_gl_2p20x135_n${G}_ring.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','stage_ms')}, d['config']['exchange'])
    print('e2e', (d.get('e2e') or {}).get('ms_per_step'), 'lib', {k:(v.get('ms_per_step') if isinstance(v,dict) else v) for k,v in (d.get('library_multi_device') or {}).items() if k not in ('how','parity','cap0')})
except Exception as e:
    print('no line:', e)
PY
grep -h "parity check\|PARITY\|Error\|error" gpurun_out/r3_n${G}_ring.err | tail -6
