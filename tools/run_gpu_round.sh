set -x
mkdir -p gpurun_out
B="--steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches20.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches20.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:area_tile -s 6 -c 1 -o gpurun_out/prof_v20_area_brk python bench.py $B > gpurun_out/ncu20a.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:step_kernel -s 6 -c 1 -o gpurun_out/prof_v20_step_brk python bench.py $B > gpurun_out/ncu20b.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:render_kernel -s 6 -c 1 -o gpurun_out/prof_v20_rgba_brk python bench.py --obs rgba $B > gpurun_out/ncu20c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 6 -c 1 -o gpurun_out/prof_v20_rgb_amidar python bench.py --game amidar --obs rgb $B > gpurun_out/ncu20d.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:area_tile -s 6 -c 1 -o gpurun_out/prof_v20_area_amidar python bench.py --game amidar $B > gpurun_out/ncu20e.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:area_tile -s 6 -c 1 -o gpurun_out/prof_v20_area_si python bench.py --game space_invaders $B > gpurun_out/ncu20f.log 2>&1
for f in area_brk step_brk rgba_brk rgb_amidar area_amidar area_si; do python tools/ncu_summary.py gpurun_out/prof_v20_$f.ncu-rep > gpurun_out/ncu_v20_$f.txt 2>&1; done
ls -la gpurun_out
rm -f gpurun_out/prof_v20_step_brk.ncu-rep gpurun_out/prof_v20_area_si.ncu-rep gpurun_out/prof_v20_area_amidar.ncu-rep
timeout 600 python bench.py > gpurun_out/bench20_default.log 2>&1; tail -1 gpurun_out/bench20_default.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench20_reference.log 2>&1
