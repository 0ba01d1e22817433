set -x
mkdir -p gpurun_out
for g in breakout amidar space_invaders; do
timeout 300 python bench.py --wrapped --game $g --steps 100 --warmup 10 > gpurun_out/bench21_wrapped_$g.log 2>&1; tail -1 gpurun_out/bench21_wrapped_$g.log | cut -c1-330
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches21w.csv python bench.py --wrapped --steps 20 --warmup 3 > gpurun_out/launches21w.log 2>&1
grep -c . gpurun_out/launches21w.csv
