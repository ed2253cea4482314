mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29701 tests/multi/check_sharded.py > gpurun_out/r2x_check.log 2>&1; echo "check rc=$?"; tail -6 gpurun_out/r2x_check.log
timeout 300 $TR --master-port 29702 bench.py --gpus 2 > gpurun_out/r2x_C3_n2.json 2> gpurun_out/r2x_C3_n2.err; echo "bench rc=$?"
timeout 300 $TR --master-port 29703 bench.py --gpus 2 --config C4 --steps 40 > gpurun_out/r2x_C4_n2.json 2> gpurun_out/r2x_C4_n2.err; echo "bench C4 rc=$?"
timeout 300 $TR --master-port 29704 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/r2x_ref_n2.json 2>&1; echo "ref rc=$?"; tail -c 300 gpurun_out/r2x_ref_n2.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for f in gpurun_out/r2x_C*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    e=d.get('e2e') or {}
    print(sys.argv[1], 'ms/step', round(d.get('ms_per_step',0),5), 'value', round(d.get('value',0)), 'e2e', round(e.get('value',0)), 'bitwise', d.get('sharded_chain_bitwise'), 'mc', d.get('uses_multicast'))
except Exception as ex: print(sys.argv[1], 'ERR', ex, open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
