"""helpers for -m gpu tests: every call goes through the C ABI (poco_b200._lib)"""
import os
import sys
import time

import torch
import torch.nn.functional as F

from poco_b200 import _lib as L
from poco_b200 import engine


def sync_or_die(seconds=60.0):
    """cudaStreamSynchronize with a deadline: a hung kernel must not take the GPU box down with it"""
    ev = torch.cuda.Event()
    ev.record()
    t0 = time.time()
    while not ev.query():
        if time.time() - t0 > seconds:
            sys.stderr.write(f'FATAL: GPU work did not finish within {seconds}s -- aborting the process\n')
            sys.stderr.flush()
            os._exit(3)
        time.sleep(0.002)


def stream():
    return torch.cuda.current_stream().cuda_stream


def run_conv(x, w, bias, stride=1, pad=None, relu=1, residual=None, impl=0, dxn=False, split=False, max_ctas=0):
    """x [N,Cin,H,W], w [Cout,Cin,k,k], bias [Cout] (CPU float) -> [N,Cout,Ho,Wo] float (CPU).
    Inputs are rounded to fp16 exactly as the engine does (split=True: to hi + lo fp16 pairs, the parity mode)."""
    dev = 'cuda'
    N, Cin, H, W = x.shape
    Cout, _, k, _ = w.shape
    pad = k // 2 if pad is None else pad
    cin_p = (Cin + 15) // 16 * 16
    a = engine.to_planar(x.to(dev), c_pad=cin_p, split=split)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    o = engine.alloc_act(Cout, N, Ho, Wo, dev, split)
    wfmt = 1 if dxn else (engine.split_wfmt(Cout, k, stride, pad) if (split and impl == 0) else 0)
    wp = engine.pack_conv_weight_dxn(w.to(dev).float()) if dxn else \
        engine.pack_conv_weight(w.to(dev).float(), cin_pad=cin_p, split=('ncat' if wfmt == 2 else split))
    b = bias.to(dev).float().contiguous()
    r = engine.to_planar(residual.to(dev), split=split) if residual is not None else None
    d = L.Conv(a.desc(), o.desc(), wp.data_ptr(), b.data_ptr(), r.ptr if r is not None else None,
               r.plane_stride if r is not None else 0, k, k, stride, pad, relu, impl, max_ctas, wfmt,
               r.ptr_lo if r is not None else None)
    L.run_op(d, stream())
    sync_or_die()
    out = engine.from_planar(o).cpu()
    for lo in ([False, True] if split else [False]):
        halo = engine.act_view(o, lo)
        assert float(halo[:, :, 0].abs().sum() + halo[:, :, -1].abs().sum() + halo[:, :, :, 0].abs().sum() +
                     halo[:, :, :, -1].abs().sum()) == 0.0, 'kernel wrote into the zero halo'
    return out


def conv_reference(x, w, bias, stride=1, pad=None, relu=1, residual=None):
    """fp32 oracle arithmetic on the fp16-rounded operands (what the kernel is specified to compute)"""
    k = w.shape[-1]
    pad = k // 2 if pad is None else pad
    y = F.conv2d(x.half().float(), w.half().float(), bias.float(), stride=stride, padding=pad)
    if relu == 2:
        y = F.relu(y)
    if residual is not None:
        y = y + residual.half().float()
    if relu == 1:
        y = F.relu(y)
    return y


def split16(t):
    """value a split-precision tensor holds for t: fp16(t) + fp16(t - fp16(t)), as float64"""
    hi = t.float().half()
    lo = (t.float() - hi.float()).half()
    return hi.double() + lo.double()


def conv_reference_split(x, w, bias, stride=1, pad=None, relu=1, residual=None):
    """float64 oracle arithmetic on the hi + lo operand values of the split-precision (parity) mode"""
    k = w.shape[-1]
    pad = k // 2 if pad is None else pad
    y = F.conv2d(split16(x), split16(w), bias.double(), stride=stride, padding=pad)
    if relu == 2:
        y = F.relu(y)
    if residual is not None:
        y = y + split16(residual)
    if relu == 1:
        y = F.relu(y)
    return y


def run_basic_block(x, w1, b1, w2, b2, max_ctas=0, s2d=False):
    """fused BasicBlock (poco_basic_block): x [N,C,H,W], w1 / w2 [C,C,3,3] (BN folded), b1 / b2 [C] -> [N,C,H,W] float (CPU);
    s2d=True: also the phase-split second output [N,4C,H/2,W/2] -> (out, out_s2d)"""
    dev = 'cuda'
    N, C_, H, W = x.shape
    a = engine.to_planar(x.to(dev))
    o = engine.alloc_act(C_, N, H, W, dev)
    p1, p2 = engine.pack_conv_weight(w1.to(dev).float()), engine.pack_conv_weight(w2.to(dev).float())
    c1, c2 = b1.to(dev).float().contiguous(), b2.to(dev).float().contiguous()
    d = L.BasicBlock(a.desc(), o.desc(), p1.data_ptr(), c1.data_ptr(), p2.data_ptr(), c2.data_ptr(), max_ctas, 0)
    o2 = None
    if s2d:
        o2 = engine.alloc_act(4 * C_, N, H // 2, W // 2, dev)
        d.out_s2d = o2.desc()
    L.run_op(d, stream())
    sync_or_die()
    for t in (o, o2) if s2d else (o,):
        halo = engine.act_view(t)
        assert float(halo[:, :, 0].abs().sum() + halo[:, :, -1].abs().sum() + halo[:, :, :, 0].abs().sum() +
                     halo[:, :, :, -1].abs().sum()) == 0.0, 'kernel wrote into the zero halo'
    if s2d:
        return engine.from_planar(o).cpu(), engine.from_planar(o2).cpu()
    return engine.from_planar(o).cpu()


def run_branch(x, ws, bs, max_ctas=0, in_place=True):
    """resident branch (poco_branch): x [N,128,H,W], ws / bs: 2 * n_blocks conv weights [C,C,3,3] (BN folded) / shifts [C]
    -> [N,C,H,W] float (CPU)"""
    dev = 'cuda'
    N, C_, H, W = x.shape
    a = engine.to_planar(x.to(dev))
    o = a if in_place else engine.alloc_act(C_, N, H, W, dev)
    keep = []
    d = L.Branch()
    d.in_, d.out = a.desc(), o.desc()
    for i, (w, b) in enumerate(zip(ws, bs)):
        keep += [engine.pack_conv_weight(w.to(dev).float()), b.to(dev).float().contiguous()]
        d.weight[i], d.bias[i] = keep[-2].data_ptr(), keep[-1].data_ptr()
    d.n_blocks, d.max_ctas = len(ws) // 2, max_ctas
    L.run_op(d, stream())
    sync_or_die()
    halo = engine.act_view(o)
    assert float(halo[:, :, 0].abs().sum() + halo[:, :, -1].abs().sum() + halo[:, :, :, 0].abs().sum() +
                 halo[:, :, :, -1].abs().sum()) == 0.0, 'kernel wrote into the zero halo'
    return engine.from_planar(o).cpu()


def run_bottleneck_tail(x, w2, b2, w3, b3, residual, max_ctas=0):
    """fused Bottleneck tail (poco_bottleneck_tail): x [N,64,H,W], w2 [64,64,3,3], w3 [256,64,1,1] (BN folded),
    residual [N,256,H,W] -> [N,256,H,W] float (CPU)"""
    dev = 'cuda'
    N, Cm, H, W = x.shape
    Co = w3.shape[0]
    a = engine.to_planar(x.to(dev))
    r = engine.to_planar(residual.to(dev))
    o = engine.alloc_act(Co, N, H, W, dev)
    p2, p3 = engine.pack_conv_weight(w2.to(dev).float()), engine.pack_conv_weight(w3.to(dev).float())
    c2, c3 = b2.to(dev).float().contiguous(), b3.to(dev).float().contiguous()
    d = L.BottleneckTail(a.desc(), o.desc(), r.ptr, r.plane_stride, p2.data_ptr(), c2.data_ptr(), p3.data_ptr(), c3.data_ptr(), max_ctas, 0)
    L.run_op(d, stream())
    sync_or_die()
    halo = engine.act_view(o)
    assert float(halo[:, :, 0].abs().sum() + halo[:, :, -1].abs().sum() + halo[:, :, :, 0].abs().sum() +
                 halo[:, :, :, -1].abs().sum()) == 0.0, 'kernel wrote into the zero halo'
    return engine.from_planar(o).cpu()
