# One multi-GPU gpurun call (NG ranks): the c3 bench line (with the real-collective parity check) and, with EPOCH=1,
# BASELINE.json configs[3] (full epoch at N = 100k).
cd $GRAFT_REPO_ROOT
NG=${NG:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $NG --steps 10 --warmup 3 --skip-cpu --skip-full --skip-c-entry > gpurun_out/bench_c3_n$NG.json 2> gpurun_out/bench_c3_n$NG.err; echo "bench rc=$?"
if [ "${EPOCH:-0}" = "1" ]; then
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 experiments/bench/epoch_dist.py 100000 > gpurun_out/epoch_100k_n$NG.json 2> gpurun_out/epoch_100k_n$NG.err; echo "epoch rc=$?"
fi
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c3_n$NG.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','stage_ms','gpu_launches','multi_gpu_check')}); print(d['e2e']['ms_per_step'], d['e2e']['sync_ms_per_step'], d['roofline']['frac'], d.get('structured_route',{}).get('ms_per_step'), d['clocks'])
PY
