TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1"
port=29610
for L in $2; do
for ex in ${3:-fused ring}; do
  port=$((port + 1))
  timeout 600 $TR --master-port $port bench.py --gpus $1 --steps 5 --warmup 3 --log-n $L --exchange $ex --no-e2e --no-library-multi --no-cpu-baseline --no-fri --no-latency >> gpurun_out/r3_ring_n$1.jsonl 2>> gpurun_out/r3_ring_n$1.err
done
done
python - <<'PY'
import json,sys,glob
for f in glob.glob('gpurun_out/r3_ring_n*.jsonl'):
    for l in open(f):
        try: d=json.loads(l)
        except Exception: continue
        print(f, d['n_gpus'], d['config']['log_rows'], d['config']['exchange'], '%.2f ms'%d['ms_per_step'], {k:round(v,2) for k,v in d['stage_ms'].items()})
PY
grep -h "parity\|PARITY\|Error\|error\|Traceback" gpurun_out/r3_ring_n$1.err | tail -8
