# N = 2 validation run (gpurun --gpus 2): multi-process GPU test, --gather auto with and without the
# NVLink load of N = 8 (--gather-fanout 7: every push is delivered seven times to the peer)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_distributed.py -x -q 2>&1 | tail -15 > gpurun_out/r02h_n2_tests.log
show='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); g=d.get("gather") or {}
        print(sys.argv[1], "%.4e"%d["value"], "ms/step %.2f"%d["ms_per_step"], "feat %.2f"%d["roofline"]["kernel_ms"], "how", g.get("how"), "alone %.2f"%g.get("alone_ms",0), "base_chunks", g.get("chunks_as_base_rows"), g.get("rows_of_all_ranks_match_checksums"), (d.get("check") or {}).get("ok"), g.get("calibration_ms_per_step"))'
$TR --master-port 29901 bench.py --gpus 2 --steps 6 --no-e2e --no-api --no-cpu 2>gpurun_out/r02h_n2_auto.err | tee gpurun_out/r02h_bench_n2_auto.json | python -c "$show" auto >> gpurun_out/r02h_n2_summary.txt
$TR --master-port 29902 bench.py --gpus 2 --steps 6 --gather-fanout 7 --no-e2e --no-api --no-cpu 2>gpurun_out/r02h_n2_auto_f7.err | tee gpurun_out/r02h_bench_n2_auto_f7.json | python -c "$show" auto_fanout7 >> gpurun_out/r02h_n2_summary.txt
$TR --master-port 29903 bench.py --gpus 2 --steps 6 --gather nccl --gather-base-chunks 5 --no-e2e --no-api --no-cpu 2>gpurun_out/r02h_n2_nccl_b5.err | tee gpurun_out/r02h_bench_n2_nccl_b5.json | python -c "$show" nccl_b5 >> gpurun_out/r02h_n2_summary.txt
true
