# N = 4 run (gpurun --gpus 4): the default bench line without the host legs (collection inside the
# step, transport and base-row chunks chosen by the trial steps)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
$TR --master-port 29641 bench.py --gpus 4 --steps 6 --no-e2e --no-api --no-cpu > gpurun_out/r02i_bench_n4.json 2> gpurun_out/r02i_bench_n4.err
true
