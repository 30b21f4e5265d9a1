cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -x -q -k "two_rank or factor_first or graph" 2>&1 | tail -4
timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 --skip-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 10 --warmup 3 --skip-cpu --skip-full --skip-c-entry > gpurun_out/bench_c3_n2.json 2> gpurun_out/bench_c3_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
for wl in ('c2','c3_n2'):
    d=json.loads(open(f'gpurun_out/bench_{wl}.json').read().strip().splitlines()[-1])
    print(wl, {k:d.get(k) for k in ('value','ms_per_step','cuda_graph_ms_per_step','stage_ms','gpu_launches','multi_gpu_check')}); print('  e2e', d['e2e']['ms_per_step'], d['e2e'].get('sync_ms_per_step'), 'frac', d['roofline']['frac'], 'struct', d.get('structured_route',{}).get('ms_per_step'), d['clocks'])
PY
