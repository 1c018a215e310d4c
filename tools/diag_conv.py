"""GPU bring-up diagnostics for the tcgen05 conv kernel: runs increasingly complex cases through the
C ABI and prints where (channel block / row / pixel position) the first mismatches are.
usage: python tools/diag_conv.py [linear|gather|all]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from gpu_util import conv_reference, run_conv  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'all'
# (name, Cin, Cout, k, stride, H, N, residual, relu)
LINEAR = [
    ('1x1 id-ish 16->16 8x8', 16, 16, 1, 1, 8, 1, False, 0),
    ('1x1 32->32 12x12', 32, 32, 1, 1, 12, 2, False, 0),
    ('3x3 16->16 8x8', 16, 16, 3, 1, 8, 1, False, 0),
    ('3x3 32->32 56', 32, 32, 3, 1, 56, 2, True, 1),
    ('3x3 64->64 28', 64, 64, 3, 1, 28, 2, True, 1),
    ('3x3 128->128 14 (streamed)', 128, 128, 3, 1, 14, 3, True, 1),
    ('3x3 256->256 7 (streamed)', 256, 256, 3, 1, 7, 5, True, 1),
    ('3x3 480->256 28 (streamed)', 480, 256, 3, 1, 28, 1, False, 1),
    ('1x1 1024->2048 7 (8 N blocks)', 1024, 2048, 1, 1, 7, 2, False, 1),
    ('3x3 384->384 7 (2 N blocks)', 384, 384, 3, 1, 7, 2, True, 1),
]
GATHER = [
    ('3x3 s2 32->64 16', 32, 64, 3, 2, 16, 1, False, 0),
    ('3x3 s2 3->64 224', 3, 64, 3, 2, 224, 1, False, 1),
    ('3x3 s2 64->64 112', 64, 64, 3, 2, 112, 1, False, 1),
    ('7x7 s2 3->64 224', 3, 64, 7, 2, 224, 1, False, 1),
    ('1x1 s2 256->512 56', 256, 512, 1, 2, 56, 1, False, 0),
    ('3x3 s2 256->256 14 relu2+res', 256, 256, 3, 2, 14, 2, True, 2),
]


def analyse(out, ref):
    err = (out - ref).abs()
    print('    max err %.4g at %s ; ref max %.4g ; out max %.4g ; nan %d' % (
        float(err.max()), tuple(int(i) for i in torch.nonzero(err == err.max())[0]), float(ref.abs().max()),
        float(out.abs().max()), int(torch.isnan(out).sum())))
    N, C, H, W = out.shape
    bad = err > 1e-2 * ref.abs().max()
    print('    bad fraction %.4f' % float(bad.float().mean()))
    cb = bad.float().mean(dim=(0, 2, 3)).view(-1, 8).mean(1)
    print('    bad frac per 8-channel plane:', [round(float(v), 2) for v in cb][:32])
    rb = bad.float().mean(dim=(0, 1, 3))
    print('    bad frac per row:', [round(float(v), 2) for v in rb][:32])
    xb = bad.float().mean(dim=(0, 1, 2))
    print('    bad frac per col:', [round(float(v), 2) for v in xb][:32])
    print('    out[0,0,:3,:6]', out[0, 0, :3, :6].tolist())
    print('    ref[0,0,:3,:6]', ref[0, 0, :3, :6].tolist())


def run(cases):
    for name, cin, cout, k, s, H, N, res, relu in cases:
        g = torch.Generator().manual_seed(cin + cout + k)
        x = torch.randn(N, cin, H, H, generator=g)
        w = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
        b = 0.1 * torch.randn(cout, generator=g)
        Ho = (H + 2 * (k // 2) - k) // s + 1
        r = torch.randn(N, cout, Ho, Ho, generator=g) if res else None
        ref = conv_reference(x, w, b, s, None, relu, r)
        try:
            out = run_conv(x, w, b, s, None, relu, r, impl=0)
        except Exception as e:      # noqa: BLE001
            print('CASE', name, 'EXC', repr(e)[:300], flush=True)
            continue
        e = float((out - ref).abs().max() / ref.abs().max())
        print('CASE', name, 'rel_err %.3e' % e, 'OK' if e < 1.5e-3 else 'FAIL', flush=True)
        if not e < 1.5e-3:
            analyse(out, ref)
            ref1 = run_conv(x, w, b, s, None, relu, r, impl=1)
            print('    debug-kernel rel_err %.3e' % float((ref1 - ref).abs().max() / ref.abs().max()), flush=True)


if which in ('linear', 'all'):
    run(LINEAR)
if which in ('gather', 'all'):
    run(GATHER)
