#!/bin/bash
# 8-GPU measurements kept under profiles/: config 2 (with the library's multi-device pass), the north-star size 2^23 x 135,
# config 5's 2^24 x 135 and config 3 (BabyBear 2^22 x 256).  Usage: tools/run_n8.sh <gpus> <out-prefix>
G=${1:-8}
P=${2:-gpurun_out/r2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus $G --steps 5 --warmup 3 > ${P}_bench_gl_2p20x135_n$G.json 2> ${P}_n${G}_a.err
timeout 900 $TR --master-port 29522 bench.py --gpus $G --steps 3 --warmup 3 --log-n 23 --no-e2e --no-library-multi > ${P}_bench_gl_2p23x135_n$G.json 2> ${P}_n${G}_b.err
timeout 900 $TR --master-port 29523 bench.py --gpus $G --steps 3 --warmup 3 --log-n 24 --no-e2e --no-library-multi > ${P}_bench_gl_2p24x135_n$G.json 2> ${P}_n${G}_c.err
timeout 900 $TR --master-port 29524 bench.py --gpus $G --steps 3 --warmup 3 --field bb --no-e2e --no-library-multi > ${P}_bench_bb_2p22x256_n$G.json 2> ${P}_n${G}_d.err
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
