"""Host-side plan builder: turns a POCO state dict + batch size into the static op schedule that
libpoco_b200.so replays (include/poco_b200.h `poco_plan`).

Tensor plumbing only (allowed to be torch): device buffers, BatchNorm folding, weight repacking
into the kernel layout, and the list of op descriptors.  No arithmetic of the hot path runs here.
"""
import ctypes as C
import os

import torch

from . import _lib as L

BN_EPS = 1e-5


# ------------------------------------------------------------------------------------------------
# planar-8 padded activation tensors
# ------------------------------------------------------------------------------------------------
class ActT:
    """[C/8][N][H+2][W+2][8] fp16 view into a torch buffer (see include/poco_b200.h)."""

    s2d = None      # the phase-split (space-to-depth) copy a producing conv wrote next to this tensor, if any

    def __init__(self, buf, ptr, C_, N, H, W, plane_stride, cap_planes, root=None, lo_off=0):
        self.buf, self.ptr = buf, ptr
        self.C, self.N, self.H, self.W = C_, N, H, W
        self.plane_stride = plane_stride
        self.cap_planes = cap_planes
        self.root = root or self
        self.lo_off = lo_off        # split-precision mode: byte offset from the hi tensor to its lo twin (0 = fp16 mode)

    @property
    def ptr_lo(self):
        return self.ptr + self.lo_off if self.lo_off else None

    def desc(self):
        return L.Act(self.ptr, self.plane_stride, self.C, self.N, self.H, self.W, self.ptr_lo)

    def channels(self, c0, c1):
        """channel-slice view (free torch.cat / torch.split: planes are contiguous)"""
        assert c0 % 8 == 0 and c1 % 8 == 0 and 0 <= c0 < c1 <= self.C
        return ActT(self.buf, self.ptr + (c0 // 8) * self.plane_stride * 16, c1 - c0, self.N, self.H, self.W,
                    self.plane_stride, 0, self.root, self.lo_off)

    def retype(self, C_):
        """same buffer, fewer channels (pool reuse)"""
        assert C_ % 8 == 0 and C_ // 8 <= self.cap_planes
        return ActT(self.buf, self.ptr, C_, self.N, self.H, self.W, self.plane_stride, self.cap_planes, None, self.lo_off)


def alloc_act(C_, N, H, W, device, split=False):
    """split=True (split-precision mode): one buffer holds the hi tensor and, behind its own guard, the lo tensor"""
    assert C_ % 8 == 0
    plane = N * (H + 2) * (W + 2)
    guard = L.ACT_GUARD_BYTES // 2
    one = guard + (C_ // 8) * plane * 8 + guard
    one = (one + 63) // 64 * 64                 # keep the lo tensor 128-byte aligned like the hi one
    buf = torch.zeros(one * (2 if split else 1), dtype=torch.float16, device=device)
    return ActT(buf, buf.data_ptr() + L.ACT_GUARD_BYTES, C_, N, H, W, plane, C_ // 8, None, one * 2 if split else 0)


def to_planar(x, c_pad=None, split=False):
    """torch reference of the layout (tests / debugging): NCHW float -> ActT on x.device"""
    N, C_, H, W = x.shape
    Cp = c_pad or ((C_ + 7) // 8 * 8)
    a = alloc_act(Cp, N, H, W, x.device, split)
    hi = x.to(torch.float16)
    parts = [(hi, False)] + ([((x.float() - hi.float()).to(torch.float16), True)] if split else [])
    for t, lo in parts:
        v = act_view(a, lo)
        xp = torch.zeros(N, Cp, H, W, dtype=torch.float16, device=x.device)
        xp[:, :C_] = t
        v[:, :, 1:H + 1, 1:W + 1, :] = xp.view(N, Cp // 8, 8, H, W).permute(1, 0, 3, 4, 2)
    return a


def act_view(a, lo=False):
    """[C/8, N, H+2, W+2, 8] torch view of an (unsliced) ActT (lo=True: its rounding-residual twin)"""
    guard = L.ACT_GUARD_BYTES // 2
    off = (a.ptr + (a.lo_off if lo else 0) - a.buf.data_ptr()) // 2
    assert off >= guard and (a.lo_off or not lo)
    n = (a.C // 8) * a.plane_stride * 8
    flat = a.buf[off:off + n]
    return flat.view(a.C // 8, a.plane_stride, 8)[:, :a.N * (a.H + 2) * (a.W + 2)].reshape(
        a.C // 8, a.N, a.H + 2, a.W + 2, 8)


def from_planar(a):
    """ActT -> NCHW float32 torch tensor (tests / debugging); split-precision tensors return hi + lo"""
    def one(lo):
        v = act_view(a, lo)[:, :, 1:a.H + 1, 1:a.W + 1, :]
        return v.permute(1, 0, 4, 2, 3).reshape(a.N, a.C, a.H, a.W).float()
    return one(False) + one(True) if a.lo_off else one(False)


# ------------------------------------------------------------------------------------------------
# weight preparation
# ------------------------------------------------------------------------------------------------
def fold_bn(w, conv_bias, bn):
    """w' = w * g/sqrt(v+eps),  b' = beta + (conv_bias - mean) * g/sqrt(v+eps)   (fp32, before the fp16 cast)"""
    w = w.float()
    cout = w.shape[0]
    b = conv_bias.float() if conv_bias is not None else torch.zeros(cout, device=w.device)
    if bn is not None:
        g, beta, mean, var = (t.float() for t in bn)
        s = g / torch.sqrt(var + BN_EPS)
        w = w * s.view(-1, 1, 1, 1)
        b = beta + (b - mean) * s
    return w, b


def pack_conv_weight(w, cin_pad=None, split=False):
    """[Cout, Cin, kh, kw] f32 -> fp16 [kh*kw][Cin/8][Cout][8] (UMMA no-swizzle K-major B slabs).
    split=True (split-precision mode): [2][kh*kw][Cin/8][Cout][8] = W_hi = fp16(W), then W_lo = fp16(W - W_hi)."""
    cout, cin, kh, kw = w.shape
    cp = cin_pad or cin
    if cp != cin:
        w = torch.cat([w, w.new_zeros(cout, cp - cin, kh, kw)], 1)
    assert cp % 16 == 0 and cout % 16 == 0, (cin, cout)
    p = w.permute(2, 3, 1, 0).reshape(kh * kw, cp // 8, 8, cout).permute(0, 1, 3, 2).contiguous()
    hi = p.to(torch.float16)
    if not split:
        return hi
    lo = (p.float() - hi.float()).to(torch.float16)
    if split == 'ncat':         # weight format 2: [kh*kw][Cin/8][2*Cout][8], slab rows W_hi then W_lo
        return torch.cat([hi, lo], 2).contiguous()
    return torch.stack([hi, lo]).contiguous()


def split_wfmt(cout, k, stride, pad):
    """split-precision convs with Cout <= 64 on the stride-1 path use the N-concatenated weight format (2)"""
    linear = stride == 1 and ((k == 3 and pad == 1) or (k == 1 and pad == 0))
    return 2 if (linear and cout <= 64 and os.environ.get('POCO_B200_NCAT', '1') != '0') else 0


def pack_conv_weight_dxn(w):
    """[Cout, Cin, 3, 3] f32 -> fp16 [3 (filter row r)][Cin/8][3*Cout (column s*Cout + co)][8]: the "dx in N"
    layout of poco_conv.wfmt = 1 (the three taps of a filter row share one MMA)"""
    cout, cin, kh, kw = w.shape
    assert (kh, kw) == (3, 3) and cin % 16 == 0 and cout % 32 == 0 and 3 * cout <= 256, tuple(w.shape)
    p = w.permute(2, 1, 3, 0).reshape(3, cin // 8, 8, 3 * cout).permute(0, 1, 3, 2)
    return p.contiguous().to(torch.float16)


def dxn_applies(cin, cout, k, stride, pad):
    """3x3 / stride 1 / pad 1 convs with 32 or 64 output channels CAN run in dx-in-N mode (POCO_B200_DXN=1).
    Off by default: measured at batch 256 (tools/conv_bench.py ... dxn) the mode cuts the MMA work of a tile
    from 18 to 6 instructions, but these layers are bound by the operand-load pipeline (26 us with MMAs and
    epilogue skipped) and the 3x wider TMEM read + shuffle epilogue costs more than the MMAs it saves
    (32->32 @56: 49 vs 36 us, 64->64 @28: 35 vs 31 us)."""
    return (os.environ.get('POCO_B200_DXN', '0') == '1' and k == 3 and stride == 1 and pad == 1 and
            cout in (32, 64) and cin % 16 == 0)


def pack_realnvp(sd, prefix='flow_head.flow.'):
    """flat fp32 parameter block of poco_realnvp (include/poco_b200.h)"""
    mask = sd[prefix + 'mask']
    parts = []
    for i in range(mask.shape[0]):
        parts.append(mask[i].reshape(-1))
        for net in ('s', 't'):
            for l in ('0', '2', '4'):
                parts.append(sd[f'{prefix}{net}.{i}.{l}.weight'].reshape(-1))
                parts.append(sd[f'{prefix}{net}.{i}.{l}.bias'].reshape(-1))
    return torch.cat([p.float() for p in parts]).contiguous()


def pack_realnvp_ctx(sd, num_rv, prefix='flow_head.flow.'):
    """the context columns of the first layer of every (coupling layer, s / t net), stacked for ONE GEMM:
    -> (W [L*2*HID, CTX], b [L*2*HID]) in the (layer, net = s then t, unit) order poco_realnvp.ctx_part uses"""
    n = sd[prefix + 'mask'].shape[0]
    ws, bs = [], []
    for i in range(n):
        for net in ('s', 't'):
            w0 = sd[f'{prefix}{net}.{i}.0.weight'].float()
            ws.append(w0[:, num_rv:])
            bs.append(sd[f'{prefix}{net}.{i}.0.bias'].float())
    return torch.cat(ws, 0).contiguous(), torch.cat(bs, 0).contiguous()


# ------------------------------------------------------------------------------------------------
# plan builder
# ------------------------------------------------------------------------------------------------
class PlanBuilder:
    """Backend of poco_b200.arch: every arch function call appends one op descriptor."""

    mode = 'plan'

    def __init__(self, sd, N, device, conv_impl=0, split=False, latency_mode=False):
        self.sd = sd
        # space-to-depth plumbing of the stride-2 convs (conv_bn s2d=...): fp16 mode only -- in split precision the four
        # phase blocks of [hi | lo] planes per K chunk leave too little shared memory (measured 28.2 vs 26.3 ms per step);
        # POCO_B200_S2D=0 / =2 force it off / on
        env_s2d = os.environ.get('POCO_B200_S2D', '1')
        self.use_s2d = conv_impl == 0 and (env_s2d == '2' or (env_s2d == '1' and not split))
        self.latency_mode = bool(latency_mode)      # small batches: branch convs as persistent chains (arch.chain_policy)
        self.split = bool(split)    # split-precision ("parity") mode: hi + lo fp16 activations and weights everywhere
        self.N = N
        self.device = torch.device(device)
        self.ops = []
        self.keep = []          # torch tensors the descriptors point into
        self.free_pool = {}     # (lane, H, W) -> [root ActT]; lane 'fork' = free before the current fork
        self.conv_impl = conv_impl
        self.conv_log = []
        self.lane = 0
        self.shares = None      # SM budget per lane inside a fork
        self.num_sms = 148
        if self.device.type == 'cuda':      # lane shares are fractions of the real SM count
            self.num_sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        self.use_lanes = os.environ.get('POCO_B200_LANES', '1') != '0'
        self.use_chains = os.environ.get('POCO_B200_CHAINS', '1') != '0' and not self.split
        self.chain = None       # pending conv descriptors of an open chain

    # -- buffers
    def act(self, C_, H, W):
        # a lane may reuse what it freed itself, or what was already free when the lanes forked
        for key in ((self.lane, H, W), ('fork', H, W)):
            pool = self.free_pool.get(key, [])
            best = None
            for a in pool:
                if a.cap_planes >= C_ // 8 and (best is None or a.cap_planes < best.cap_planes):
                    best = a
            if best is not None:
                pool.remove(best)
                return best.retype(C_) if best.C != C_ else best
        a = alloc_act(C_, self.N, H, W, self.device, self.split)
        self.keep.append(a.buf)
        return a

    def free(self, a):
        """return the buffer behind `a` to the pool (safe: ops run in program order on one stream)"""
        if a.s2d is not None:           # its phase-split copy goes with it
            s2d, a.s2d = a.s2d, None
            self.free(s2d)
        root = a.root
        if root.buf is None or (root is not a and a.ptr != root.ptr):      # (geometry-only tensor of an s2d='only' conv; a slice)
            return
        full = ActT(root.buf, root.ptr, root.cap_planes * 8, root.N, root.H, root.W, root.plane_stride, root.cap_planes,
                    None, root.lo_off)
        lst = self.free_pool.setdefault((self.lane, root.H, root.W), [])
        if all(x.buf is not full.buf for x in lst):
            lst.append(full)

    def f32(self, *shape):
        t = torch.zeros(*shape, dtype=torch.float32, device=self.device)
        self.keep.append(t)
        return t

    def dev(self, t, dtype=torch.float32):
        t = t.detach().to(device=self.device, dtype=dtype).contiguous()
        self.keep.append(t)
        return t

    def add(self, desc):
        self.ops.append(L.make_op(desc, lane=self.lane))

    def _share(self):
        """CTA budget of a conv launched now: 0 (all SMs) outside a fork, else POCO_B200_SHARE_SCALE (default 2) x the
        lane's cost-proportional share of the SMs.  The shares of the concurrent lanes deliberately add up to twice
        the SM count: a lane's SMs idle for ~10 us around each of its launches (launch gap + pipeline fill), and with
        oversubscription the hardware block scheduler hands those SMs to another lane's queued CTAs in the meantime
        (persistent CTAs that start late just take their strided share of the units later).  Measured (round 2, one
        box): cliff_w32 fp16 11.84 -> 11.66 ms, split 29.2 -> 27.0 ms, cliff_w48cls 19.0 -> 17.0 ms; scale 3 and
        "every launch asks for all SMs" (scale 0) are slower than 2."""
        if self.shares is None:
            return 0
        sc = float(os.environ.get('POCO_B200_SHARE_SCALE', '2'))
        if self.lane > 0:       # (the side lanes run on higher-priority streams: POCO_B200_SIDE_SCALE lets them ask for more)
            sc = float(os.environ.get('POCO_B200_SIDE_SCALE', sc))
        if self.chain is not None:      # chained launches spin on each other's tile flags: all their CTAs must be resident
            sc = min(sc, 1.0) if sc > 0 else 1.0
        return 0 if sc <= 0 else max(1, min(self.num_sms, int(round(self.shares[self.lane] * sc))))

    # -- lanes: independent op chains that the plan runs concurrently on internal streams
    def fork(self, costs):
        """start len(costs) concurrent lanes; `costs` (relative work) decide each lane's share of the SMs"""
        assert self.shares is None and self.lane == 0
        if not self.use_lanes:
            return
        n = len(costs)
        self.shares = self.shares_for(costs)
        self.ops.append(L.make_op(L.Sync(n), lane=0, kind=L.OP_FORK))
        for (lane, H, W), lst in list(self.free_pool.items()):     # everything free now is safe for any lane
            if lane == 0 and lst:
                self.free_pool.setdefault(('fork', H, W), []).extend(lst)
                lst.clear()

    def shares_for(self, costs):
        """SM share per lane for relative `costs` (every lane at least 4 SMs, together at most all of them)"""
        tot = float(sum(costs)) or 1.0
        shares = [max(4, int(round(self.num_sms * c / tot))) for c in costs]
        while sum(shares) > self.num_sms:
            shares[shares.index(max(shares))] -= 1
        return shares

    def set_shares(self, costs):
        """inside a fork: the ops emitted next take their CTA budgets from these costs (arch.hr_stage: a lane's fuse ops and
        its next-module branch ops share one fork / join region but not one cost profile)"""
        if self.use_lanes and self.shares is not None:
            assert len(costs) == len(self.shares)
            self.shares = self.shares_for(costs)

    def set_lane(self, k):
        if not self.use_lanes:
            return
        assert self.shares is not None and 0 <= k < len(self.shares)
        self.lane = k

    def join(self):
        if not self.use_lanes:
            return
        n = len(self.shares)
        self.lane = 0
        self.ops.append(L.make_op(L.Sync(n), lane=0, kind=L.OP_JOIN))
        for (lane, H, W), lst in list(self.free_pool.items()):
            if lane != 0 and lst:
                self.free_pool.setdefault((0, H, W), []).extend(lst)
                lst.clear()
        self.shares = None

    # -- ops
    def pack_image(self, img, H, W):
        out = self.act(16, H, W)
        self.add(L.PackImage(img.data_ptr(), out.desc(), 0, 0))
        return out

    def stem_conv(self, img, H, W, conv, bn, cout, s2d=None):
        """conv3x3(3 -> cout, stride 2, pad 1) + BN + ReLU straight from the f32 image (hrnet.py:299-301,
        :467-469): the packing kernel writes the 27-tap im2col (32 fp16 channels at half resolution) and the
        conv runs as a 1x1 conv with K = 32 -- half the activation bytes of a 3x3 conv over 16 zero-padded
        channels and no strided gather."""
        sd = self.sd
        col = self.act(32, H // 2, W // 2)
        self.add(L.PackImage(img.data_ptr(), col.desc(), 1, 0))
        # weight preparation (BN folding, repacking, fp16 cast) runs on the HOST: the only device work of a plan
        # build is the upload, so the first kernels a fresh process launches are the forward's own
        w = sd[conv + '.weight'].cpu()
        assert tuple(w.shape) == (cout, 3, 3, 3), (conv, tuple(w.shape))
        bnp = tuple(sd[bn + s].cpu() for s in ('.weight', '.bias', '.running_mean', '.running_var'))
        wf, bf = fold_bn(w, None, bnp)
        w1 = wf.permute(0, 2, 3, 1).reshape(cout, 27)               # k = (r*3+s)*3 + c
        w1 = torch.cat([w1, w1.new_zeros(cout, 5)], 1).reshape(cout, 32, 1, 1)
        wfmt = split_wfmt(cout, 1, 1, 0) if (self.split and self.conv_impl == 0) else 0
        wp = pack_conv_weight(w1, split=('ncat' if wfmt == 2 else self.split)).to(self.device)
        bf = bf.contiguous().to(self.device)
        self.keep += [wp, bf]
        use_s2d = s2d and self.use_s2d
        Ho, Wo = H // 2, W // 2
        out = self.act(cout, Ho, Wo) if not (use_s2d and s2d == 'only') else \
            ActT(None, 0, cout, self.N, Ho, Wo, self.N * (Ho + 2) * (Wo + 2), 0)
        d = L.Conv(col.desc(), out.desc(), wp.data_ptr(), bf.data_ptr(), None, 0, 1, 1, 1, 0, 1, self.conv_impl,
                   self._share(), wfmt)
        if use_s2d:
            out.s2d = self.act(4 * cout, Ho // 2, Wo // 2)
            d.out_s2d = out.s2d.desc()
            d.s2d_only = 1 if s2d == 'only' else 0
        self.add(d)
        self.conv_log.append((conv, 32, cout, 1, 1, H // 2, H // 2))
        self.free(col)
        return out

    def conv_bn(self, x, conv, bn, cin, cout, k, stride=1, relu=True, residual=None, out=None, pad=None, bias=False,
                s2d=None):
        """conv (+bias) + eval BatchNorm folded, + residual, ReLU.  `conv` / `bn` may be lists: several
        convs reading the same input run as one launch with their output channels concatenated.
        s2d='dual' / 'only': the conv also writes (only writes) the phase-split form of its output for a stride-2
        consumer (returned as out.s2d).  A 3x3 / stride 2 / pad 1 conv whose input carries such a copy runs on it
        (poco_conv.in_s2d: halo-run path instead of the 16-byte gather); a 1x1 / stride 2 conv reads phase block 0."""
        sd = self.sd
        convs = conv if isinstance(conv, (list, tuple)) else [conv]
        bns = bn if isinstance(bn, (list, tuple)) else [bn] * len(convs)
        ws, bs = [], []
        for cv, b_ in zip(convs, bns):
            w = sd[cv + '.weight'].cpu()            # (host-side weight preparation, see stem_conv)
            assert tuple(w.shape) == (cout // len(convs), cin, k, k), (cv, tuple(w.shape), (cout, cin, k, k))
            cb = sd.get(cv + '.bias')
            assert (cb is not None) == bool(bias), f'{cv}: conv bias presence mismatch'
            bnp = None
            if b_ is not None:
                bnp = tuple(sd[b_ + s].cpu() for s in ('.weight', '.bias', '.running_mean', '.running_var'))
            wf, bf = fold_bn(w, cb.cpu() if cb is not None else None, bnp)
            ws.append(wf)
            bs.append(bf)
        pad = k // 2 if pad is None else pad
        use_s2d = self.use_s2d and self.chain is None
        in_s2d = 0
        if use_s2d and stride == 2 and x.s2d is not None and x.C == cin and x.H % 2 == 0 and x.W % 2 == 0:
            if k == 3 and pad == 1 and cin % 16 == 0:
                xs2d, in_s2d = x.s2d, 1                     # the conv walks the phase-split copy
            elif k == 1 and pad == 0:
                x, stride = x.s2d.channels(0, cin), 1       # phase (0, 0) IS the stride-2 sampling: a plain 1x1 conv
        wfmt = 1 if (self.conv_impl == 0 and self.chain is None and x.C == cin and not self.split and
                     dxn_applies(cin, cout, k, stride, pad)) else 0
        if self.split and self.conv_impl == 0 and self.chain is None:
            wfmt = split_wfmt(cout, k, stride, pad)
        wp = pack_conv_weight_dxn(torch.cat(ws, 0)) if wfmt == 1 else \
            pack_conv_weight(torch.cat(ws, 0), cin_pad=x.C, split=('ncat' if wfmt == 2 else self.split))
        wp = wp.to(self.device)
        bf = torch.cat(bs, 0).contiguous().to(self.device)
        self.keep += [wp, bf]
        Ho = (x.H + 2 * pad - k) // stride + 1
        Wo = (x.W + 2 * pad - k) // stride + 1
        want_s2d = s2d if (use_s2d and s2d and Ho % 2 == 0 and Wo % 2 == 0 and
                           (stride == 1 and ((k == 3 and pad == 1) or (k == 1 and pad == 0)) or in_s2d)) else None
        if out is None:
            out = self.act(cout, Ho, Wo) if want_s2d != 'only' else ActT(None, 0, cout, self.N, Ho, Wo, self.N * (Ho + 2) * (Wo + 2), 0)
        assert (out.C, out.H, out.W) == (cout, Ho, Wo), (conv, (out.C, out.H, out.W), (cout, Ho, Wo))
        if residual is not None:
            assert (residual.C, residual.H, residual.W) == (cout, Ho, Wo), conv
        d = L.Conv((xs2d if in_s2d else x).desc(), out.desc(), wp.data_ptr(), bf.data_ptr(),
                   residual.ptr if residual is not None else None,
                   residual.plane_stride if residual is not None else 0,
                   k, k, stride, pad, int(relu), self.conv_impl,
                   self._share(), wfmt,
                   residual.ptr_lo if residual is not None else None)
        d.in_s2d = in_s2d
        if want_s2d:
            out.s2d = self.act(4 * cout, Ho // 2, Wo // 2)
            d.out_s2d = out.s2d.desc()
            d.s2d_only = 1 if want_s2d == 'only' else 0
        if self.chain is not None:
            self.chain.append(d)
        else:
            self.add(d)
        self.conv_log.append((convs[0], x.C, cout, k, stride, x.H, Ho))
        return out

    def basic_block_fusable(self, c, H, W):
        """would basic_block_fused take a block of c channels at H x W?  (arch.hr_module also asks for the lane costs)"""
        return (self.conv_impl == 0 and not self.split and self.chain is None and
                os.environ.get('POCO_B200_FUSE_BLOCK', '1') != '0' and hasattr(L.lib(), 'poco_basic_block_supported') and
                bool(L.lib().poco_basic_block_supported(c, H, W)) and
                not (c == 64 and os.environ.get('POCO_B200_FUSE_BLOCK64', '1') == '0'))

    def basic_block_fused(self, x, name, c, out=None, s2d=None):
        """BasicBlock `name` (conv1-bn1-ReLU-conv2-bn2, += x, ReLU; hrnet.py:42-58) as ONE poco_basic_block launch when the
        library takes the geometry (32 or 64 channels, W <= 61, fp16 mode); None otherwise (the caller emits two conv_bn ops).
        POCO_B200_FUSE_BLOCK=0 switches it off, POCO_B200_FUSE_BLOCK64=0 only the 64-channel flavour (conv2's weights streamed,
        csrc/bblock64_tc.cu: 53 us against 60 us for the two launches at batch 256)."""
        if x.C != c or not self.basic_block_fusable(c, x.H, x.W):
            return None
        want_s2d = bool(s2d) and self.use_s2d
        if want_s2d and (s2d != 'dual' or c != 32 or x.H % 2 or x.W % 2 or os.environ.get('POCO_B200_FUSE_BLOCK_S2D', '1') == '0'):
            return None     # (the phase-split second output exists for the 32-channel flavour only: two conv launches instead)
        sd = self.sd
        packed = []
        for cv, bn in ((name + '.conv1', name + '.bn1'), (name + '.conv2', name + '.bn2')):
            w = sd[cv + '.weight'].cpu()
            assert tuple(w.shape) == (c, c, 3, 3) and sd.get(cv + '.bias') is None, cv
            wf, bf = fold_bn(w, None, tuple(sd[bn + s_].cpu() for s_ in ('.weight', '.bias', '.running_mean', '.running_var')))
            packed += [pack_conv_weight(wf).to(self.device), bf.contiguous().to(self.device)]
        self.keep += packed
        if out is None:
            out = self.act(c, x.H, x.W)
        assert (out.C, out.H, out.W) == (c, x.H, x.W)
        d = L.BasicBlock(x.desc(), out.desc(), packed[0].data_ptr(), packed[1].data_ptr(), packed[2].data_ptr(),
                         packed[3].data_ptr(), self._share(), 0)
        if want_s2d:        # the block's output also feeds stride-2 fuse convs: epilogue 2 writes the phase-split copy as well
            out.s2d = self.act(4 * c, x.H // 2, x.W // 2)
            d.out_s2d = out.s2d.desc()
        self.add(d)
        self.conv_log.append((name + '.conv1', c, c, 3, 1, x.H, x.H))
        self.conv_log.append((name + '.conv2', c, c, 3, 1, x.H, x.H))
        return out

    def branch_fusable(self, c, H, W, n_blocks):
        """would branch_fused take n_blocks BasicBlocks of c channels at H x W?  (arch.hr_module also asks for the lane costs)"""
        return (self.conv_impl == 0 and not self.split and self.chain is None and n_blocks <= L.MAX_BRANCH_BLOCKS and
                os.environ.get('POCO_B200_FUSE_BRANCH', '1') != '0' and hasattr(L.lib(), 'poco_branch_supported') and
                bool(L.lib().poco_branch_supported(c, H, W, n_blocks)))

    def branch_fused(self, x, names, c):
        """the BasicBlocks `names` of one HRNet branch (hrnet.py:42-58 x4, hrnet.py:140-186) as ONE poco_branch launch with
        the crop resident in shared memory, in place (out aliases x), when the library takes the geometry (128 channels,
        padded crop <= 256 pixels, fp16 mode); None otherwise.  POCO_B200_FUSE_BRANCH=0 switches it off."""
        if x.C != c or not self.branch_fusable(c, x.H, x.W, len(names)):
            return None
        sd = self.sd
        d = L.Branch()
        d.in_ = x.desc()
        d.out = x.desc()
        for i, name in enumerate(names):
            for j, (cv, bn) in enumerate(((name + '.conv1', name + '.bn1'), (name + '.conv2', name + '.bn2'))):
                w = sd[cv + '.weight'].cpu()
                assert tuple(w.shape) == (c, c, 3, 3) and sd.get(cv + '.bias') is None, cv
                wf, bf = fold_bn(w, None, tuple(sd[bn + s_].cpu() for s_ in ('.weight', '.bias', '.running_mean', '.running_var')))
                wp, bp = pack_conv_weight(wf).to(self.device), bf.contiguous().to(self.device)
                self.keep += [wp, bp]
                d.weight[2 * i + j] = wp.data_ptr()
                d.bias[2 * i + j] = bp.data_ptr()
                self.conv_log.append((cv, c, c, 3, 1, x.H, x.H))
        d.n_blocks = len(names)
        d.max_ctas = self._share()
        self.add(d)
        return x

    def bottleneck_tail_supported(self, y1, planes, cout):
        """True when conv2 (3x3) -> conv3 (1x1) + residual of a Bottleneck runs as one poco_bottleneck_tail launch
        (64 -> 64 -> 256 channels, fp16 mode).  POCO_B200_FUSE_TAIL=0 switches it off."""
        return (self.conv_impl == 0 and not self.split and self.chain is None and y1.C == planes and
                os.environ.get('POCO_B200_FUSE_TAIL', '1') != '0' and hasattr(L.lib(), 'poco_bottleneck_tail_supported') and
                bool(L.lib().poco_bottleneck_tail_supported(planes, cout, y1.H, y1.W)))

    def bottleneck_tail_fused(self, y1, name, planes, cout, residual):
        """conv2-bn2-ReLU-conv3-bn3, += residual, ReLU of Bottleneck `name` (hrnet.py:88-99) on conv1's output y1"""
        sd = self.sd
        packed = []
        for cv, bn, ci, co, k in ((name + '.conv2', name + '.bn2', planes, planes, 3), (name + '.conv3', name + '.bn3', planes, cout, 1)):
            w = sd[cv + '.weight'].cpu()
            assert tuple(w.shape) == (co, ci, k, k) and sd.get(cv + '.bias') is None, cv
            wf, bf = fold_bn(w, None, tuple(sd[bn + s_].cpu() for s_ in ('.weight', '.bias', '.running_mean', '.running_var')))
            packed += [pack_conv_weight(wf).to(self.device), bf.contiguous().to(self.device)]
        self.keep += packed
        out = self.act(cout, y1.H, y1.W)
        assert (residual.C, residual.H, residual.W) == (cout, y1.H, y1.W) and residual.ptr != out.ptr
        d = L.BottleneckTail(y1.desc(), out.desc(), residual.ptr, residual.plane_stride, packed[0].data_ptr(), packed[1].data_ptr(),
                             packed[2].data_ptr(), packed[3].data_ptr(), self._share(), 0)
        self.add(d)
        self.conv_log.append((name + '.conv2', planes, planes, 3, 1, y1.H, y1.H))
        self.conv_log.append((name + '.conv3', planes, cout, 1, 1, y1.H, y1.H))
        return out

    # -- conv chains: consecutive same-geometry convs (the BasicBlocks of an HRNet branch) as ONE launch
    def begin_chain(self):
        assert self.chain is None
        if self.use_chains and self.conv_impl == 0:
            self.chain = []

    def end_chain(self, groups=1):
        """emit the pending convs as chain launches.  groups > 1: the batch is cut into that many contiguous
        crop ranges and the whole chain runs range by range, so that the activations one range passes from
        conv to conv stay in the 126 MB L2 instead of making an HBM round trip per conv (crops are
        independent; a range is a pointer offset and a smaller N in the planar layout)."""
        descs, self.chain = self.chain, None
        if not descs:
            return
        groups = max(1, min(int(groups), self.N)) if len(descs) > 1 else 1
        bounds = [self.N * g // groups for g in range(groups + 1)]
        for g in range(groups):
            n0, n1 = bounds[g], bounds[g + 1]
            sub = []
            for d in descs:
                e = L.Conv.from_buffer_copy(d)
                if groups > 1:
                    for a in (e.in_, e.out):
                        a.data += n0 * (a.H + 2) * (a.W + 2) * 16
                        a.N = n1 - n0
                    if e.residual:
                        e.residual += n0 * (e.out.H + 2) * (e.out.W + 2) * 16
                sub.append(e)
            for i in range(0, len(sub), L.MAX_CHAIN):
                grp = sub[i:i + L.MAX_CHAIN]
                if len(grp) == 1:
                    self.add(grp[0])
                    continue
                ch = L.ConvChain()
                for k, d in enumerate(grp):
                    ch.seg[k] = d
                ch.n_seg = len(grp)
                n_flags = int(L.lib().poco_conv_chain_flag_count(C.byref(ch)))
                flags = torch.zeros(max(1, n_flags), dtype=torch.int32, device=self.device)
                self.keep.append(flags)
                ch.flags = flags.data_ptr()
                self.add(ch)

    def fuse_sum(self, terms, relu, out=None):
        """terms: list of (ActT, shift)"""
        a0 = terms[0][0]
        H, W = a0.H << terms[0][1], a0.W << terms[0][1]
        if out is None:
            out = self.act(a0.C, H, W)
        d = L.FuseSum()
        d.out = out.desc()
        for i, (a, s) in enumerate(terms):
            d.in_[i] = a.desc()
            d.shift[i] = s
        d.n_in = len(terms)
        d.relu = 1 if relu else 0
        self.add(d)
        return out

    def upsample2x(self, x):
        out = self.act(x.C, 2 * x.H, 2 * x.W)
        self.add(L.Upsample2x(x.desc(), out.desc()))
        return out

    def maxpool(self, x):
        out = self.act(x.C, (x.H - 1) // 2 + 1, (x.W - 1) // 2 + 1)
        self.add(L.MaxPool(x.desc(), out.desc()))
        return out

    def avgpool(self, x, out, col=0):
        """-> columns [col, col+C) of the f32 matrix `out` [N, ld]"""
        self.add(L.AvgPool(x.desc(), out.data_ptr() + 4 * col, out.stride(0)))

    def unpack(self, x, c_valid=None):
        out = self.f32(self.N, c_valid or x.C, x.H, x.W)
        self.add(L.Unpack(x.desc(), out.data_ptr(), c_valid or x.C))
        return out

    def linear(self, x, xcol, I, wkey, y, ycol, act=0, res=None, rescol=0):
        """y[:, ycol:ycol+O] = act(x[:, xcol:xcol+I] W^T + b) (+ res[:, rescol:rescol+O])"""
        self.linear_w(x, xcol, I, self.sd[wkey + '.weight'], self.sd[wkey + '.bias'], y, ycol, act, res, rescol)

    def linear_w(self, x, xcol, I, w, b, y, ycol, act=0, res=None, rescol=0):
        w, b = self.dev(w), self.dev(b)
        O = w.shape[0]
        assert w.shape[1] == I, (tuple(w.shape), I)
        assert 0 <= ycol and ycol + O <= y.shape[1], f'linear writes columns [{ycol}, {ycol + O}) of a {tuple(y.shape)} matrix'
        # one split-K workspace for all linear ops of the plan (they run one after the other on the main lane)
        need = 8 * x.shape[0] * O
        if getattr(self, '_lin_scratch', None) is None or self._lin_scratch.numel() < need:
            self._lin_scratch = torch.empty(max(need, 8 * x.shape[0] * 1024), dtype=torch.float32, device=self.device)
            self.keep.append(self._lin_scratch)
        self.add(L.Linear(x.data_ptr() + 4 * xcol, x.stride(0), w.data_ptr(), b.data_ptr(),
                          (res.data_ptr() + 4 * rescol) if res is not None else None,
                          res.stride(0) if res is not None else 0,
                          y.data_ptr() + 4 * ycol, y.stride(0), x.shape[0], I, O, act,
                          self._lin_scratch.data_ptr(), self._lin_scratch.numel()))

    def copy2d(self, src, scol, dst, dcol, cols, bcast=False):
        self.add(L.Copy2d(src.data_ptr() + 4 * scol, src.stride(0), dst.data_ptr() + 4 * dcol, dst.stride(0),
                          dst.shape[0], cols, 1 if bcast else 0))

    def rot6d(self, x, xcol, per_row, out):
        self.add(L.Rot6d(x.data_ptr() + 4 * xcol, x.stride(0), per_row, x.shape[0] * per_row, out.data_ptr()))


class Plan:
    """Owns the C plan handle plus every device buffer its descriptors point into."""

    def __init__(self, builder):
        self.keep = builder.keep
        self.ops = builder.ops
        n = len(builder.ops)
        arr = (L.Op * n)(*builder.ops)
        h = C.c_void_p()
        L.check(L.lib().poco_plan_create(arr, n, C.byref(h)))
        self.handle = h
        self.num_ops = n
        self.flops = int(L.lib().poco_plan_flops(h))
        self.conv_log = builder.conv_log

    def run(self, stream=None):
        s = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        L.check(L.lib().poco_plan_run(self.handle, C.c_void_p(s)))

    def __del__(self):
        try:
            if self.handle:
                L.lib().poco_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
