mkdir -p gpurun_out
./tools/probe/launch_gap > gpurun_out/r2l_launch_gap.txt 2>&1; cat gpurun_out/r2l_launch_gap.txt
B="python bench.py --steps 200 --warmup 12 --no-e2e --no-cpu-baseline"
$B --no-flush --timeline gpurun_out/r2l_base_ > gpurun_out/r2l_base.log 2>&1
$B --no-flush --carveout 100 --timeline gpurun_out/r2l_carve_ > gpurun_out/r2l_carve.log 2>&1
$B --no-flush --steps-per-graph 4 --timeline gpurun_out/r2l_spg4_ > gpurun_out/r2l_spg4.log 2>&1
$B --carveout 100 > gpurun_out/r2l_carve_flush.log 2>&1
$B --steps-per-graph 4 > gpurun_out/r2l_spg4_flush.log 2>&1
for f in gpurun_out/r2l_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(round(d['ms_per_step'],5), d.get('step_ms_min_median_max'))
except Exception as e: print('ERR', open(sys.argv[1]).read()[-600:])
PY
done
for t in base carve spg4; do echo "== $t"; python tools/timeline_report.py gpurun_out/r2l_${t}_ 1 300; done
python - <<'PY'
import numpy as np
t=np.load('gpurun_out/r2l_spg4_0.npy'); gen=int(t[0,5]); hs=np.arange(gen-320,gen-1)
a=t[hs].astype(float); b=t[hs+1].astype(float)
for pos in range(8):
    m=(hs%8)==pos
    print('pos',pos,'evaluate %.2f accept %.2f to_next %.2f half %.2f'%tuple(np.median(x[m])/1e3 for x in (a[:,2]-a[:,1], a[:,4]-a[:,2], b[:,0]-a[:,4], b[:,0]-a[:,0])))
PY
