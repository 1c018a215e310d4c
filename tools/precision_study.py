"""CPU study (no GPU): how operand precision of the conv layers moves the outputs from the fp32 reference golden, and
which layers would have to run in a split-precision mode (x = x_hi + x_lo, W = W_hi + W_lo in fp16: three MMAs per
product, fp32 accumulation) for pred_pose to reach the north star's 1e-3.  It evaluates the pinned functional
restatement (oracle/poco_oracle.py) with its conv helper wrapped so that activations / weights are rounded per layer.
ANALYSIS AID ONLY (not a product path).    python tools/precision_study.py [preset] > profiles/<name>.md"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import poco_oracle as O  # noqa: E402
from synth import ckpt as S  # noqa: E402

preset = sys.argv[1] if len(sys.argv) > 1 else 'cliff_w32'
gd = os.path.join(ROOT, 'tests', 'golden')
meta = json.load(open(os.path.join(gd, f'spec_{preset}.json')))
gold = np.load(os.path.join(gd, f'golden_{preset}.npz'))
sd = S.synth_state_dict(S.template_from_spec(meta, 0), 0, np.load(os.path.join(gd, f'calib_{preset}.npz')))
bb, head = meta['kwargs']['backbone'].split('-')
uit = meta['kwargs']['uncert_inp_type']
batch = S.synthetic_batch(meta['test_b'], meta['test_seed'], 'cpu')
KEYS = ('pred_pose', 'pred_shape', 'pred_cam', 'var_pose')


def h16(t):
    return t.half().float()


def split(t):                       # hi + lo, both fp16: ~22 significant bits
    hi = h16(t)
    return hi + h16(t - hi)


def ident(t):
    return t


layers = []                         # (name, flops) in call order, filled by the first pass
state = {'i': 0, 'policy': None}
orig_conv = O.conv


def conv(x, s, name, stride=1, pad=None):
    i = state['i']
    state['i'] += 1
    w = s[name + '.weight']
    b = s[name + '.bias'] if s.has(name + '.bias') else None
    if pad is None:
        pad = w.shape[-1] // 2
    qx, qw = state['policy'](i) if state['policy'] else (ident, ident)
    y = F.conv2d(qx(x), qw(w), b, stride=stride, padding=pad)
    if len(layers) <= i:
        layers.append((s.p + name, 2.0 * y[0].numel() * w[0].numel()))
    return y


O.conv = conv


def run(policy):
    state['i'], state['policy'] = 0, policy
    with torch.no_grad():
        out = O.poco_forward(batch, sd, bb, head, uit)
    return {k: float(np.abs(out[k].numpy() - gold[k]).max() / np.abs(gold[k]).max()) for k in KEYS}


base = run(None)
n = len(layers)
tot = sum(f for _, f in layers)
cum = np.cumsum([f for _, f in layers]) / tot
print(f'# operand-precision study, {preset}, {meta["test_b"]} crops, {n} conv layers, {tot / 1e9:.2f} GFLOP / crop in convs\n')
print('error = max-abs-err / max-abs-ref against the reference golden (the e2e metric of tests/test_gpu_e2e.py)\n')
print('| conv operands | pred_pose | pred_shape | pred_cam | var_pose | MMA work |')
print('|---|---|---|---|---|---|')


def row(label, e, work):
    print(f'| {label} | ' + ' | '.join(f'{e[k]:.1e}' for k in KEYS) + f' | {work} |', flush=True)


row('fp32 (restatement as is)', base, '-')
row('fp16 activations x fp16 weights (what the kernels compute)', run(lambda i: (h16, h16)), '1x')
row('split activations x fp16 weights (2 MMAs)', run(lambda i: (split, h16)), '2x')
row('fp16 activations x split weights (2 MMAs)', run(lambda i: (h16, split)), '2x')
row('split x split without the lo x lo term (3 MMAs)', run(lambda i: (split, split)), '3x')
for frac in (0.25, 0.5, 0.75):
    k = int(np.searchsorted(cum, frac))
    row(f'split on the first {k} layers ({cum[k - 1] * 100 if k else 0:.0f} % of the FLOPs), fp16 after',
        run(lambda i, k=k: (split, split) if i < k else (h16, h16)), f'{1 + 2 * (cum[k - 1] if k else 0):.2f}x')
    row(f'fp16 on the first {k} layers, split on the last {n - k} ({(1 - cum[k - 1]) * 100 if k else 100:.0f} % of the FLOPs)',
        run(lambda i, k=k: (h16, h16) if i < k else (split, split)), f'{1 + 2 * (1 - (cum[k - 1] if k else 0)):.2f}x')

# finer sweep: where along the depth is the error injected?  (split on layers [a, b), fp16 elsewhere)
print('\n| split-precision window (layers) | share of FLOPs | pred_pose | pred_shape | pred_cam | var_pose |')
print('|---|---|---|---|---|---|')
edges = sorted({0, 2, n} | {max(2, int(round(n * q))) for q in (0.05, 0.13, 0.23, 0.39, 0.59, 0.80, 0.987)})
for a, b in zip(edges[:-1], edges[1:]):
    e = run(lambda i, a=a, b=b: (split, split) if a <= i < b else (h16, h16))
    sh = (cum[b - 1] - (cum[a - 1] if a else 0)) * 100
    print(f'| {a}..{b - 1} ({layers[a][0].replace("backbone.", "")} .. {layers[b - 1][0].replace("backbone.", "")}) | {sh:.1f} % | '
          + ' | '.join(f'{e[k]:.1e}' for k in KEYS) + ' |', flush=True)
# the complementary view: fp16 only inside the window, split everywhere else
print('\n| fp16 only in this window, split elsewhere | share of FLOPs | pred_pose | pred_shape | pred_cam | var_pose |')
print('|---|---|---|---|---|---|')
for a, b in zip(edges[:-1], edges[1:]):
    e = run(lambda i, a=a, b=b: (h16, h16) if a <= i < b else (split, split))
    sh = (cum[b - 1] - (cum[a - 1] if a else 0)) * 100
    print(f'| {a}..{b - 1} | {sh:.1f} % | ' + ' | '.join(f'{e[k]:.1e}' for k in KEYS) + ' |', flush=True)

print('\n| parity-mode candidates: split on layers [0, k), fp16 from k on | share of FLOPs split | pred_pose | pred_shape | pred_cam | var_pose | MMA work |')
print('|---|---|---|---|---|---|---|')
for k in sorted({max(2, int(round(n * q))) for q in (0.05, 0.13, 0.59, 0.80, 0.90, 0.987)}):
    e = run(lambda i, k=k: (split, split) if i < k else (h16, h16))
    print(f'| k = {k} | {cum[k - 1] * 100:.1f} % | ' + ' | '.join(f'{e[x]:.1e}' for x in KEYS) + f' | {1 + 2 * cum[k - 1]:.2f}x |', flush=True)
