for n in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err
tail -2 gpurun_out/r2_bench_n$n.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n$n.json')); print($n, d['value'], d['scaling'], d['e2e']['value'], d['independent_chains']['trajectories_per_s']); t=d['tau_slab']; print({k:t[k] for k in ('parity','cg_us_per_iter_1gpu','cg_us_per_iter','ranks_bit_identical')}); print(t['preconditioned'])"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29599 tools/shard_worker.py b80 2 1 2>&1 | tail -1 | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29598 tools/shard_worker.py cfg4 1 1 kpm 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('greens 8 gpus', d['greens'])"
