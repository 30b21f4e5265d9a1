cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/full_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/full_pytest.log
tail -5 gpurun_out/full_pytest.log
python experiments/bench/factor_only.py 4096 20 2>&1 | tee gpurun_out/full_factor.log
timeout 900 python bench.py --steps 5 --warmup 3 --skip-cpu > gpurun_out/bench_c3_n1.json 2> gpurun_out/bench_c3_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c3_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms','gpu_launches')}); print(d['e2e']['ms_per_step'], d['roofline']['frac'], d['structured_route']['ms_per_step'], d['structured_route']['stage_ms'])
PY
