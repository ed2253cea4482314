mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
for v in "" "--no-blobs" "--no-multicast" "--no-blobs --no-flush" "--no-flush"; do
  tag=$(echo "x$v" | tr -d ' -')
  timeout 200 $TR bench.py --gpus 2 --steps 200 --warmup 10 --no-check --no-e2e $v > gpurun_out/r2i_n2_$tag.log 2>&1
done
python bench.py --steps 200 --warmup 10 --no-e2e --no-cpu-baseline --no-blobs > gpurun_out/r2i_n1_noblobs.log 2>&1
python bench.py --steps 200 --warmup 10 --no-e2e --no-cpu-baseline --no-flush > gpurun_out/r2i_n1_noflush.log 2>&1
for f in gpurun_out/r2i_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(round(d['ms_per_step'],5), d.get('step_ms_min_median_max'))
except Exception as e: print('ERR', open(sys.argv[1]).read()[-300:])
PY
done
