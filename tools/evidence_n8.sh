# Round-2 multi-GPU evidence on one 8 x B200 box (charged 8x: keep it short).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/ev8_pytest_multi.log 2>&1; tail -2 gpurun_out/ev8_pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 200 --warmup 10 > gpurun_out/ev8_bench_C3_n8.json 2> gpurun_out/ev8_bench_C3_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 4 --steps 200 --warmup 10 > gpurun_out/ev8_bench_C3_n4.json 2> gpurun_out/ev8_bench_C3_n4.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/ev8_bench_C3_n2.json 2> gpurun_out/ev8_bench_C3_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 8 --config C5 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/ev8_bench_C5_n8.json 2> gpurun_out/ev8_bench_C5_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29615 bench.py --gpus 8 --walkers-per-gpu 1024 --steps 100 --warmup 10 --no-cpu-baseline --no-check > gpurun_out/ev8_bench_C3_w1024_n8.json 2> gpurun_out/ev8_bench_C3_w1024_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29616 bench.py --gpus 8 --steps 200 --warmup 10 --no-e2e --no-flush --timeline gpurun_out/ev8_tl_ > gpurun_out/ev8_bench_C3_n8_tl.json 2> /dev/null
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus 8 --steps 200 --warmup 10 --no-e2e --no-check --transport nccl > gpurun_out/ev8_bench_C3_n8_nccl.json 2> /dev/null
for f in gpurun_out/ev8_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    e=d.get('e2e') or {}
    print(sys.argv[1].split('ev8_bench_')[1], 'ms/step', round(d.get('ms_per_step',0),5), 'value', round(d.get('value',0)), 'e2e', round(e.get('value',0)), 'bitwise', d.get('sharded_chain_bitwise'), 'mc', d.get('uses_multicast'), (d.get('clocks') or {}).get('sm_mhz'))
except Exception as ex: print(sys.argv[1], 'ERR', ex, open(sys.argv[1].replace('.json','.err')).read()[-800:] if sys.argv[1].endswith('.json') else '')
PY
done
python tools/timeline_report.py gpurun_out/ev8_tl_ 8 300 | head -30
