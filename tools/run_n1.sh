# N = 1 run: GPU tests, launch list of the bench step, default bench line
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02g_gpu_tests.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02g_launches_bench.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu --no-api > gpurun_out/r02g_bench_under_ncu.log 2>&1
python bench.py > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err
true
