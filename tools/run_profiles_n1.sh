#!/bin/bash
# Round-2 single-GPU evidence kept under profiles/: bench lines (both fields), the size sweep, the ncu launch list of the
# bench command and full ncu captures of the NTT kernels (each ncu run only after the same command exited 0 without it).
P=${1:-gpurun_out/r2}
timeout 900 python bench.py --steps 5 --warmup 3 > ${P}_bench_gl_2p20x135.json 2> ${P}_bench_gl.err
timeout 900 python bench.py --field bb --steps 3 --warmup 3 > ${P}_bench_bb_2p22x256.json 2> ${P}_bench_bb.err
timeout 900 python tools/sweep.py 16 23 > ${P}_sweep_gl_135cols.jsonl 2> ${P}_sweep.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-fri --no-latency --no-parity-check > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${P}_launch_list_bench_gl_2p20x135.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-fri --no-latency --no-parity-check > ${P}_ncu_list.log 2>&1
timeout 200 python tools/quick_bench.py gl 20 135 1 > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:ntt_ -c 4 -o ${P}_ntt_gl_final -f python tools/quick_bench.py gl 20 135 1 > ${P}_ncu_gl.log 2>&1
timeout 200 python tools/quick_bench.py bb 20 256 1 > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:ntt_ -c 4 -o ${P}_ntt_bb_final -f python tools/quick_bench.py bb 20 256 1 > ${P}_ncu_bb.log 2>&1
tail -c 600 ${P}_bench_gl_2p20x135.json; echo; tail -c 300 ${P}_bench_bb_2p22x256.json; echo; cat ${P}_sweep_gl_135cols.jsonl | cut -c1-160
