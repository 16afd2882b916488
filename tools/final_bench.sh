python bench.py > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; tail -c 200 gpurun_out/bench_r1g.err; cut -c1-300 gpurun_out/bench_r1g.json
