cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/full_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/full_pytest.log
tail -4 gpurun_out/full_pytest.log
python experiments/bench/shard_stages.py 125000 2>&1 | grep "gram\|V^T Z (pl" | tee gpurun_out/shard_stages3.txt
for wl in c1 c2; do timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --skip-cpu > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"; done
timeout 900 python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_c3_n1.json 2> gpurun_out/bench_c3_n1.err; echo "bench c3 rc=$?"
python - <<'PY'
import json
for wl in ('c1','c2','c3_n1'):
    d=json.loads(open(f'gpurun_out/bench_{wl}.json').read().strip().splitlines()[-1])
    print(wl, {k:d.get(k) for k in ('value','ms_per_step','cuda_graph_ms_per_step','stage_ms','gpu_launches')}); print('  e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'struct', d.get('structured_route',{}).get('ms_per_step'), d['clocks'])
PY
