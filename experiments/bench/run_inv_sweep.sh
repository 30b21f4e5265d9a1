cd $GRAFT_REPO_ROOT
echo "== baseline"; python experiments/bench/factor_only.py 4096 30
for f in 31 40 44 48 52 56; do for s in 0 32; do echo "== GPP_INV_OVERLAP=1 GPP_INV_FORK=$f GPP_INV_SIDE_SMS=$s"; GPP_INV_OVERLAP=1 GPP_INV_FORK=$f GPP_INV_SIDE_SMS=$s python experiments/bench/factor_only.py 4096 30; done; done
echo "== Q=1024"; python experiments/bench/factor_only.py 1024 50
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/factor4096_launches.csv python experiments/bench/factor_only.py 4096 1 > gpurun_out/factor4096_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/factor4096_launches.csv')) if len(r) > 5]
hdr = rows[0]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value'); iu = hdr.index('Metric Unit')
seq = [(r[ik], float(r[iv].replace(',', '')) / (1000.0 if r[iu] in ('ns', 'nsecond') else 1.0)) for r in rows[1:]]
# the last factor() + solve_w() calls: find last scal_init
names = [n for n, _ in seq]
last = max(i for i, n in enumerate(names) if 'scal_init' in n)
agg = collections.OrderedDict()
for n, t in seq[last:]:
    k = n.split('(')[0][:60]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += t
for k, (c, t) in agg.items(): print(f"{k:62s} {c:4d} {t:9.1f} us")
print("total", sum(t for _, t in seq[last:]))
PY
