# N = 8 run (gpurun --gpus 8): the default bench line and the two 8-GPU configurations of BASELINE.json
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29611 bench.py --gpus 8 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
$TR --master-port 29612 bench.py --gpus 8 --config 3 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_n8.json 2> gpurun_out/r02_bench_cfg3_n8.err
$TR --master-port 29613 bench.py --gpus 8 --config 4 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg4_n8.json 2> gpurun_out/r02_bench_cfg4_n8.err
$TR --master-port 29614 bench.py --gpus 8 --gather nccl --no-e2e --no-api --no-cpu > gpurun_out/r02_bench_n8_nccl.json 2> gpurun_out/r02_bench_n8_nccl.err
$TR --master-port 29615 bench.py --gpus 8 --gather-ctas 148 --no-e2e --no-api --no-cpu > gpurun_out/r02_bench_n8_ctas148.json 2> gpurun_out/r02_bench_n8_ctas148.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_n8box.json 2>/dev/null
nvidia-smi topo -m > gpurun_out/r02_n8_topo.txt 2>&1
true
