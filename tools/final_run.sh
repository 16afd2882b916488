python -m pytest tests -m gpu -x -q > gpurun_out/t_r1f.log 2>&1; tail -2 gpurun_out/t_r1f.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1f.log 2>&1; tail -1 gpurun_out/smoke_r1f.log
python bench.py > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; tail -c 300 gpurun_out/bench_r1f.err
python bench.py --impl reference --steps 100 > gpurun_out/bench_r1f_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/ncu_r1f_1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_strip -s 6 -c 1 -o gpurun_out/prof_r1f_cfg3 python bench.py --instances 16384 --size 256 --window 64 --rule B368/S245 --fused-reductions --pool-mib 1024 --steps 4 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/ncu_r1f_2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_tiled -s 4 -c 1 -o gpurun_out/prof_r1f_tiled2 python tools/biggrid.py 65536 > gpurun_out/ncu_r1f_3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_stream -s 6 -c 1 -o gpurun_out/prof_r1f_cfg2 python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/ncu_r1f_4.log 2>&1
ls -la gpurun_out/*r1f*
