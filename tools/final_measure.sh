#!/bin/bash
# Round-end measurement on one B200 (run through gpurun): bench line, reference arm, ncu launch list of the bench
# command, one `ncu --set full` capture of every kernel of a 30-slot forward.  Outputs under gpurun_out/<tag>_*.
tag=${1:-r02b}
o=gpurun_out
python bench.py > $o/${tag}_bench_1gpu.json 2> $o/${tag}_bench_1gpu.err; echo "bench rc=$?"; tail -c 600 $o/${tag}_bench_1gpu.json
python bench.py --impl reference --steps 4 --warmup 1 > $o/${tag}_bench_ref.json 2> $o/${tag}_bench_ref.err; echo "ref rc=$?"; tail -c 400 $o/${tag}_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 17 --warmup 3 --no-cpu-baseline --no-latency > $o/${tag}_ncu_l.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on --launch-skip 20 --launch-count 20 -f -o $o/${tag}_full \
    python tools/profile_run.py nrx_large 30 2 > $o/${tag}_ncu_f.log 2>&1; echo "ncu full rc=$?"
ncu -i $o/${tag}_full.ncu-rep --page raw --csv > $o/${tag}_full_raw.csv 2>/dev/null; ls -la $o/${tag}_*
