# N = 8 run (gpurun --gpus 8): the default bench line (collection inside the step, method chosen by measurement)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29611 bench.py --gpus 8 --no-e2e --no-api --no-cpu > gpurun_out/r02i_bench_n8.json 2> gpurun_out/r02i_bench_n8.err
true
