# N = 1 run: the resample tests again, launch list of the bench step (bench.py --steps 3 --warmup 1)
python -m pytest tests/test_gpu_resample.py -x -q 2>&1 | tail -5 > gpurun_out/r02j_gpu_tests.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused_features|delta_fixed|cmvn|expand_tiles" -c 24 --csv --log-file gpurun_out/r02j_launches_bench.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu --no-api > gpurun_out/r02j_bench_under_ncu.log 2>&1
true
