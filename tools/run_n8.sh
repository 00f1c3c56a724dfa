# N = 8 run (gpurun --gpus 8): the default bench line (collection inside the step) and two variants
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29611 bench.py --gpus 8 > gpurun_out/r02f_bench_n8.json 2> gpurun_out/r02f_bench_n8.err
$TR --master-port 29612 bench.py --gpus 8 --gather-base-chunks 0 --steps 6 --no-e2e --no-api --no-cpu > gpurun_out/r02f_bench_n8_b0.json 2> gpurun_out/r02f_bench_n8_b0.err
$TR --master-port 29613 bench.py --gpus 8 --gather-base-chunks 4 --steps 6 --no-e2e --no-api --no-cpu > gpurun_out/r02f_bench_n8_b4.json 2> gpurun_out/r02f_bench_n8_b4.err
true
