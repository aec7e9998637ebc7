#!/bin/bash
# How the coefficients of the column-sharded inverse NTT travel (P2G_EXCHANGE=fused|push|ce, p2g_api.cu exchange_mode),
# measured device-resident on G GPUs.  Usage: tools/run_exchange_modes.sh <gpus> <out.jsonl> [log_n ...]
G=${1:-8}
OUT=${2:-gpurun_out/r2_exchange_modes_n$G.jsonl}
shift 2
SIZES=${@:-20}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
: > $OUT
port=29610
for L in $SIZES; do
  for mode in fused push ce; do
    port=$((port + 1))
    P2G_EXCHANGE=$mode timeout 600 $TR --master-port $port bench.py --gpus $G --steps 5 --warmup 3 --log-n $L \
        --no-e2e --no-library-multi --no-cpu-baseline --no-fri --no-latency >> $OUT 2> ${OUT%.jsonl}_${mode}_$L.err
  done
done
python - "$OUT" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    try:
        d = json.loads(l)
    except Exception:
        continue
    print(d["config"]["log_rows"], d["config"].get("exchange_transport"), "%.2f ms" % d["ms_per_step"], {k: round(v, 2) for k, v in d["stage_ms"].items()})
PY
