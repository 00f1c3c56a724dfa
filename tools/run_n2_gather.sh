TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for c in 74 148 296; do $TR --master-port 2953$((c%10)) bench.py --gpus 2 --gather-ctas $c --no-e2e --no-api --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ctas $c', '%.3e'%d['value'], round(d['ms_per_step'],2), d['gather']['alone_ms'], d['gather']['rows_of_all_ranks_match_checksums'])
"; done
$TR --master-port 29540 bench.py --gpus 2 --gather nccl --no-e2e --no-api --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nccl', '%.3e'%d['value'], round(d['ms_per_step'],2), d['gather']['alone_ms'])
"
