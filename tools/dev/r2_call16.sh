# Round-2 GPU batch #16: new tests (submission writer, voxel lift, explicit-im2col backbone), backbone launch list
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/c16_*
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "submission or voxel_lift or explicit_im2col or pack_levels or bf16_feature" > gpurun_out/c16_tests.full 2>&1
tail -40 gpurun_out/c16_tests.full > gpurun_out/c16_tests.log; rm gpurun_out/c16_tests.full; tail -6 gpurun_out/c16_tests.log
timeout 300 python tools/dev/backbone_one.py 5 > gpurun_out/c16_backbone_time.log 2>&1; tail -2 gpurun_out/c16_backbone_time.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c16_backbone_launches.csv \
    python tools/dev/backbone_one.py 1 > gpurun_out/c16_backbone_ncu.log 2>&1
python - <<'P'
import csv, re, collections
rows = list(csv.reader(open('gpurun_out/c16_backbone_launches.csv')))
h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[h]; ki = H.index('Kernel Name'); vi = H.index('Metric Value')
data = [(re.sub(r'\(.*', '', r[ki]).replace('void occ::<unnamed>::', '').replace('occ::<unnamed>::', ''), float(r[vi].replace(',', ''))) for r in rows[h + 1:] if len(r) > vi]
half = len(data) // 2            # (warm-up forward + timed forward): keep the second
data = data[half:]
agg = collections.OrderedDict()
for k, v in data:
    a = agg.setdefault(k[:70], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in data)
print('one backbone forward: %d launches, %.1f us (serialised, cold cache)' % (len(data), tot / 1e3))
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-72s n=%3d %9.1f us %5.1f %%' % (k, n, v / 1e3, 100 * v / tot))
P
