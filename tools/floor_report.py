"""Where the step goes against its floors: reads a per-op table written by `bench.py --dump-ops` (eager, one launch at a
time, CUDA events) and puts next to every conv class the two floors that bound it on B200 --
  * HBM floor: algorithmic bytes (input once, output once, residual once, fp16) / measured copy bandwidth,
  * tensor floor: number of M=128, K=16 SS MMAs x max(N/2, 32 + N/4) cycles (the measured acceptance rate of the
    pipe with both operands in shared memory, profiles/r01b_mma_issue_rate.csv) / (148 SMs x SM clock),
then ranks the classes by the time they spend above max(floor).  CPU-only analysis of committed profiles:
    python tools/floor_report.py profiles/r01c_ops_b256.csv [batch] [sm_mhz] > profiles/<name>.md"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r01c_ops_b256.csv')
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
mhz = float(sys.argv[3]) if len(sys.argv) > 3 else 1912.0
peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) \
    else {'hbm_gbs': 6650.0}
hbm = peaks['hbm_gbs'] * 1e9
SMS = 148
rows = []
for r in csv.DictReader(open(path)):
    m = re.match(r'conv (\d+)->(\d+) k(\d) s(\d) (\d+)->(\d+)( \+res)?', r['label'])
    n, tot = int(r['count']), float(r['total_ms'])
    if not m:
        rows.append((r['label'], n, tot, None, None, None))
        continue
    cin, cout, k, s, hi, ho = (int(x) for x in m.groups()[:6])
    res = bool(m.group(7))
    by = 2.0 * B * (hi * hi * cin + ho * ho * cout * (2 if res else 1))
    t_hbm = by / hbm * 1e3                                              # ms per launch
    tiles = -(-B * (ho + 2) * (ho + 2) // 128)                          # 128 padded pixels per tile (halo included)
    ntile = min(cout, 256)
    mmas = tiles * (cout // ntile) * k * k * (max(cin, 16) // 16)
    t_mma = mmas * max(ntile / 2, 32 + ntile / 4) / (SMS * mhz * 1e6) * 1e3
    fl = 2.0 * B * ho * ho * cout * cin * k * k
    rows.append((r['label'], n, tot, t_hbm * n, t_mma * n, fl * n))
tot_ms = sum(r[2] for r in rows)
conv = [r for r in rows if r[3] is not None]
print(f'# floors of the eager per-op pass `{os.path.basename(path)}` (batch {B}, SM clock {mhz:.0f} MHz, HBM {hbm / 1e9:.0f} GB/s)\n')
print(f'total {tot_ms:.2f} ms; conv launches {sum(r[2] for r in conv):.2f} ms, their HBM floors sum to '
      f'{sum(r[3] for r in conv):.2f} ms, tensor floors to {sum(r[4] for r in conv):.2f} ms, '
      f'max(floor) per class to {sum(max(r[3], r[4]) for r in conv):.2f} ms\n')
print('| conv class | launches | measured ms | HBM floor ms | tensor floor ms | bound | x over floor | ms above floor | TFLOP/s |')
print('|---|---|---|---|---|---|---|---|---|')
for lab, n, t, th, tm, fl in sorted(conv, key=lambda r: -(r[2] - max(r[3], r[4])))[:24]:
    fl_ = max(th, tm)
    print(f'| {lab[5:]} | {n} | {t:.3f} | {th:.3f} | {tm:.3f} | {"hbm" if th > tm else "tensor"} | {t / fl_:.2f} | {t - fl_:.3f} | '
          f'{fl / t / 1e9:.0f} |')
other = [r for r in rows if r[3] is None]
print(f'\nnon-conv ops: {sum(r[2] for r in other):.2f} ms ({", ".join(f"{r[0]} {r[2]:.2f}" for r in sorted(other, key=lambda r: -r[2])[:6])} ...)')
