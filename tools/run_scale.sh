#!/bin/bash
# Multi-GPU evidence kept under profiles/: config 2 on G GPUs (rank-per-GPU and the library's multi-device pass), on 8 GPUs
# also the north-star size 2^23 x 135, on 2 GPUs also the multi-device parity tests over the real device lists.
# Usage: tools/run_scale.sh <gpus> <out-prefix>
G=${1:-8}
P=${2:-gpurun_out/r3}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus $G --steps 5 --warmup 3 --no-cpu-baseline --no-fri --no-latency > ${P}_bench_gl_2p20x135_n$G.json 2> ${P}_n${G}_a.err
if [ "$G" = 8 ]; then
  timeout 900 $TR --master-port 29522 bench.py --gpus $G --steps 3 --warmup 3 --log-n 23 --no-e2e --no-library-multi --no-cpu-baseline --no-fri --no-latency > ${P}_bench_gl_2p23x135_n$G.json 2> ${P}_n${G}_b.err
fi
if [ "$G" = 2 ] || [ "$G" = 4 ]; then
  timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "multi_device or shard" > ${P}_pytest_multi_device_n$G.log 2>&1
  tail -3 ${P}_pytest_multi_device_n$G.log
fi
for f in ${P}_bench_*_n$G.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','stage_ms')})
    print('e2e', (d.get('e2e') or {}).get('ms_per_step'), 'lib', {k:(v.get('ms_per_step') if isinstance(v,dict) else v) for k,v in (d.get('library_multi_device') or {}).items() if k!='how'})
except Exception as e:
    print('no line:', e)
PY
done
grep -h "parity check\|PARITY\|Error\|error" ${P}_n${G}_*.err | tail -8
