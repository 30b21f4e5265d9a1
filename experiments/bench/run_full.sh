# One 1-GPU gpurun call: the GPU test suite, the driver-style bench line, the launch list of the same command.
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/full_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/full_pytest.log
tail -3 gpurun_out/full_pytest.log
timeout 900 python bench.py > gpurun_out/bench_c3_n1.json 2> gpurun_out/bench_c3_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c3_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms','gpu_launches')}); print(d['e2e']['ms_per_step'], d['roofline']['frac'], d['structured_route']['ms_per_step'], d['cpu_baseline']['value'], d['clocks'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-full --skip-c-entry --skip-structured > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
