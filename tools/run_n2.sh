# N = 2 validation run (gpurun --gpus 2): GPU tests incl. the multi-process one, benches
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_n2_tests.log
python tools/bench_configs.py --utts 2000 --only 8k > gpurun_out/r02_cfg_bench6.log 2>&1
$TR --master-port 29511 bench.py --gpus 2 > gpurun_out/r02_bench_n2_p2p.json 2> gpurun_out/r02_bench_n2_p2p.err
$TR --master-port 29512 bench.py --gpus 2 --gather nccl --no-e2e --no-api --no-cpu > gpurun_out/r02_bench_n2_nccl.json 2> gpurun_out/r02_bench_n2_nccl.err
rm -f gpurun_out/r02_bench_n2_ctas.json
for c in 74 148 592; do $TR --master-port 2952$((c%10)) bench.py --gpus 2 --gather-ctas $c --steps 5 --no-e2e --no-api --no-cpu >> gpurun_out/r02_bench_n2_ctas.json 2>> gpurun_out/r02_bench_n2_ctas.err; done
$TR --master-port 29513 bench.py --gpus 2 --config 3 --steps 4 --warmup 3 --no-cpu --no-api > gpurun_out/r02_bench_cfg3_n2.json 2> gpurun_out/r02_bench_cfg3_n2.err
$TR --master-port 29514 bench.py --gpus 2 --config 4 --steps 4 --warmup 3 --no-cpu --no-api > gpurun_out/r02_bench_cfg4_n2.json 2> gpurun_out/r02_bench_cfg4_n2.err
true
