mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tests/multi/check_sharded.py > gpurun_out/r2e_check_n$N.log 2>&1
echo "check rc=$?" >> gpurun_out/r2e_check_n$N.log
timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/r2e_bench_C3_n$N.log 2>&1
timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 10 --no-flush --no-check --no-e2e > gpurun_out/r2e_bench_C3_n${N}_noflush.log 2>&1
timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 10 --transport nccl --no-check > gpurun_out/r2e_bench_C3_n${N}_nccl.log 2>&1
timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 10 --config C5 > gpurun_out/r2e_bench_C5_n$N.log 2>&1
timeout 300 $TR bench.py --gpus $N --steps 100 --warmup 10 --walkers-per-gpu 1024 --no-check > gpurun_out/r2e_bench_C3_n${N}_w1024.log 2>&1
tail -3 gpurun_out/r2e_check_n$N.log
for f in gpurun_out/r2e_bench_*n$N*.log; do echo $f; tail -c 700 $f; echo; done
