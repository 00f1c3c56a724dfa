# N = 1 run: the whole GPU test-suite and smoke()
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02k_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02k_smoke.log 2>&1
true
