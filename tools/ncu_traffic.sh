#!/bin/bash
# Evidence pass for bench.py's roofline: (1) the ncu launch list of one bench command (device time per launch),
# (2) DRAM bytes of every conv launch of one forward -> profiles-ready traffic json stamped with the library hash.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-mode"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG:-r02}_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:conv_tc|basic_block|bottleneck_tail|branch_kernel' -c 400 --csv --log-file gpurun_out/${TAG:-r02}_conv_dram.csv $CMD > gpurun_out/ncu_dram.log 2>&1
echo "dram rc=$?"
TAG=${TAG:-r02} python - <<'PY'
import csv, json, hashlib, collections, os
TAG = os.environ.get('TAG', 'r02')
def rows(path):
    lines = [l for l in open(path) if l.startswith('"')]
    return list(csv.DictReader(lines))
r = rows(f'gpurun_out/{TAG}_launches.csv')
agg = collections.defaultdict(lambda: [0, 0.0])
for x in r:
    if x.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    k = x['Kernel Name'].split('(')[0][-60:]
    v = float(x['Metric Value'].replace(',', ''))
    u = x['Metric Unit']
    v = v / 1e3 if u in ('ns', 'nsecond') else (v * 1e3 if u in ('ms', 'msecond') else v)
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
with open(f'gpurun_out/{TAG}_launches_summary.csv', 'w') as f:
    f.write('kernel,launches,total_us,share\n')
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f'"{k}",{n},{us:.1f},{us / tot:.4f}\n')
print(open(f'gpurun_out/{TAG}_launches_summary.csv').read()[:3000])
d = rows(f'gpurun_out/{TAG}_conv_dram.csv')
per = collections.defaultdict(lambda: collections.defaultdict(float))
for x in d:
    per[(x['ID'])][x['Metric Name']] = float(x['Metric Value'].replace(',', '')) * {'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'Gbyte': 1e9, 'us': 1, 'usecond': 1, 'ns': 1e-3, 'nsecond': 1e-3, 'ms': 1e3, 'msecond': 1e3}.get(x['Metric Unit'], 1)
    per[(x['ID'])]['name'] = x['Kernel Name']
lin = [v for v in per.values() if 'conv_tc_kernel<0' in v['name']]
gat = [v for v in per.values() if 'conv_tc_kernel<1' in v['name']]
def avg(vs, k): return sum(v[k] for v in vs) / max(1, len(vs))
sha = hashlib.sha256(open('poco_b200/libpoco_b200.so', 'rb').read()).hexdigest()[:16]
import sys
sys.path.insert(0, '.')
import bench
src_sha = bench.source_hash()
out = {'kernel': 'conv_tc_linear', 'launches_captured': len(lin),
       'dram_bytes_per_launch': round(avg(lin, 'dram__bytes_read.sum') + avg(lin, 'dram__bytes_write.sum')),
       'dram_read_bytes_per_launch': round(avg(lin, 'dram__bytes_read.sum')), 'dram_write_bytes_per_launch': round(avg(lin, 'dram__bytes_write.sum')),
       'avg_launch_us_under_ncu': round(avg(lin, 'gpu__time_duration.sum'), 2),
       'gather': {'launches_captured': len(gat), 'dram_bytes_per_launch': round(avg(gat, 'dram__bytes_read.sum') + avg(gat, 'dram__bytes_write.sum'))},
       'lib_sha256_16': sha, 'src_sha256_16': src_sha, 'command': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:conv_tc python bench.py --steps 2 --warmup 3 (cliff_w32, 256 crops, fp16; all conv launches of the eager + graph-capture forwards)'}
# the fused kernels of the same forward (one launch = a whole block / branch)
for key, pat in (('basic_block_tc', 'basic_block'), ('bottleneck_tail_tc', 'bottleneck_tail'), ('branch_tc', 'branch_kernel')):
    vs = [v for v in per.values() if pat in v['name']]
    out[key] = {'launches_captured': len(vs),
                'dram_bytes_per_launch': round(avg(vs, 'dram__bytes_read.sum') + avg(vs, 'dram__bytes_write.sum')),
                'avg_launch_us_under_ncu': round(avg(vs, 'gpu__time_duration.sum'), 2)}
json.dump(out, open(f'gpurun_out/{TAG}_traffic.json', 'w'), indent=1)
print(json.dumps(out))
PY
