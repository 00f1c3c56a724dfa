# N = 2 validation run (gpurun --gpus 2): multi-process GPU test, collection variants under the
# NVLink load of N = 8 (--gather-fanout 7: every push is delivered seven times to the peer)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_distributed.py -x -q 2>&1 | tail -15 > gpurun_out/r02e_n2_tests.log
show='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); g=d.get("gather") or {}
        print(sys.argv[1], "%.4e"%d["value"], "ms/step %.2f"%d["ms_per_step"], "feat %.2f"%d["roofline"]["kernel_ms"], "alone %.2f"%g.get("alone_ms",0), "recv GB/s %.0f"%g.get("recv_gbs_per_rank",0), "base_chunks", g.get("chunks_as_base_rows"), g.get("rows_of_all_ranks_match_checksums"), (d.get("check") or {}).get("ok"))'
port=29800
run() {   # name, args...
  name=$1; shift
  port=$((port+1))
  $TR --master-port $port bench.py --gpus 2 --steps 6 --no-e2e --no-api --no-cpu "$@" 2>gpurun_out/r02e_n2_$name.err | tee gpurun_out/r02e_bench_n2_$name.json | python -c "$show" $name >> gpurun_out/r02e_n2_summary.txt
}
run ce_f7_b3 --gather ce --gather-fanout 7 --gather-base-chunks 3
run ce_f7_b4 --gather ce --gather-fanout 7 --gather-base-chunks 4
run ce_f7_b5 --gather ce --gather-fanout 7 --gather-base-chunks 5
run ce_f7_b8_c16 --gather ce --gather-fanout 7 --gather-base-chunks 8 --gather-chunks 16
run ce_f7_b6_c12 --gather ce --gather-fanout 7 --gather-base-chunks 6 --gather-chunks 12
run ce_f3_auto --gather ce --gather-fanout 3
port=$((port+1))
$TR --master-port $port bench.py --gpus 2 2>gpurun_out/r02e_n2_default.err > gpurun_out/r02e_bench_n2_default.json
python -c "$show" default < gpurun_out/r02e_bench_n2_default.json >> gpurun_out/r02e_n2_summary.txt
true
