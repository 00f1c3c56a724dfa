# N = 1 run: GPU tests, host-API profile, default bench line, parity report
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02d_gpu_tests.log
python tools/profile_api.py > gpurun_out/r02d_profile_api.txt 2>&1
python bench.py > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err
python tools/parity_report.py > gpurun_out/r02d_parity_report.txt 2>&1
true
