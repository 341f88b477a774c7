cd $GRAFT_REPO_ROOT
for w in acoustics shallow sphere; do
B="python bench.py --workload $w --steps 2 --warmup 3 --no-cpu --no-e2e --no-other-build --no-quiescent-leg"
$B > gpurun_out/plain_l46_$w.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_${w}_final.csv $B > gpurun_out/ncu_l46_$w.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_r02_${w}_final.csv')) if len(r) > 5]
h = rows[0]; ik = h.index('Kernel Name'); iv = h.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ik][:48]].append(float(r[iv].replace(',', '')))
    except ValueError: pass
tot = sum(sum(v) for v in agg.values())
print('$w')
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:5]: print('   %-50s n=%3d mean %.1f us share %.1f %%' % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
PY
done
