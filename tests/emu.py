"""TEST INFRASTRUCTURE -- CPU interpreter of poco_op descriptors.

Replays an engine.Plan's op list with plain torch on CPU buffers, honouring the same layouts, fp16
storage roundings and in-place aliasing as the CUDA kernels.  It checks the *host logic* (graph
wiring, buffer-pool reuse, BN folding, weight packing, descriptor fields) without a GPU, and
predicts how far fp16 activation storage moves the outputs from the fp32 oracle.
It is not a product path: poco_b200 never imports it.
"""
import bisect
import ctypes as C

import torch
import torch.nn.functional as F

from poco_b200 import _lib as L


def space_to_depth(x):
    """[N, C, H, W] -> phase-split [N, 4C, H/2, W/2]: channel ((y & 1) * 2 + (x & 1)) * C + c of pixel (y / 2, x / 2)"""
    N, C_, H, W = x.shape
    return x.view(N, C_, H // 2, 2, W // 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(N, 4 * C_, H // 2, W // 2)


def depth_to_space(p):
    N, C4, H, W = p.shape
    C_ = C4 // 4
    return p.view(N, 2, 2, C_, H, W).permute(0, 3, 4, 1, 5, 2).reshape(N, C_, 2 * H, 2 * W)


class Emu:
    def __init__(self, keep):
        items = sorted(((t.data_ptr(), t) for t in keep if t.numel() > 0), key=lambda x: x[0])
        self.starts = [p for p, _ in items]
        self.tensors = [t for _, t in items]

    def flat(self, ptr, dtype):
        """1-D view (of `dtype`) starting at raw address ptr, to the end of its buffer"""
        i = bisect.bisect_right(self.starts, ptr) - 1
        t = self.tensors[i]
        base = t.data_ptr()
        nbytes = t.numel() * t.element_size()
        assert base <= ptr < base + nbytes, 'pointer outside every kept buffer'
        rem = t.view(-1).view(torch.uint8)[ptr - base:]
        es = torch.empty(0, dtype=dtype).element_size()
        return rem[:rem.numel() // es * es].view(dtype)

    def act(self, a, lo=False):
        f = self.flat(a.lo if lo else a.data, torch.float16)
        Hp, Wp = a.H + 2, a.W + 2
        return torch.as_strided(f, (a.C // 8, a.N, Hp, Wp, 8), (a.plane_stride * 8, Hp * Wp * 8, Wp * 8, 8, 1))

    def act_get(self, a):
        """value of a tensor: hi (+ lo in split-precision mode)"""
        def one(lo):
            v = self.act(a, lo)[:, :, 1:a.H + 1, 1:a.W + 1, :]
            return v.permute(1, 0, 4, 2, 3).reshape(a.N, a.C, a.H, a.W).float()
        return one(False) + one(True) if a.lo else one(False)

    def act_set(self, a, x):
        hi = x.to(torch.float16)
        parts = [(hi, False)] + ([((x.float() - hi.float()).to(torch.float16), True)] if a.lo else [])
        for t, lo in parts:
            v = self.act(a, lo)
            v[:, :, 1:a.H + 1, 1:a.W + 1, :] = t.view(a.N, a.C // 8, 8, a.H, a.W).permute(1, 0, 3, 4, 2)

    def mat(self, ptr, rows, cols, ld):
        return torch.as_strided(self.flat(ptr, torch.float32), (rows, cols), (ld, 1))

    # ------------------------------------------------------------------------------------------
    def run(self, op):
        k = op.kind
        getattr(self, '_op%d' % k)(getattr(op.u, L._FIELD_OF_KIND[k]))

    def _op1(self, d):      # pack image
        o = d.out
        if d.im2col:
            img = self.flat(d.img, torch.float32)[:o.N * 3 * 4 * o.H * o.W].view(o.N, 3, 2 * o.H, 2 * o.W)
            cols = F.unfold(img, 3, padding=1, stride=2).view(o.N, 3, 9, o.H, o.W)      # [N, c, r*3+s, H, W]
            x = torch.zeros(o.N, 32, o.H, o.W)
            x[:, :27] = cols.permute(0, 2, 1, 3, 4).reshape(o.N, 27, o.H, o.W)
            self.act_set(o, x)
            return
        img = self.flat(d.img, torch.float32)[:o.N * 3 * o.H * o.W].view(o.N, 3, o.H, o.W)
        x = torch.zeros(o.N, o.C, o.H, o.W)
        x[:, :3] = img
        self.act_set(o, x)

    def _op2(self, d):      # conv
        i, o = d.in_, d.out
        taps = d.kh * d.kw
        cin = i.C // 4 if d.in_s2d else i.C      # channels the weights see
        if d.wfmt == 1:     # dx-in-N layout [r][Cin/8][s*Cout + co][8]
            wp = self.flat(d.weight, torch.float16)[:taps * cin * o.C].view(3, cin // 8, 3, o.C, 8).float()
            w = wp.permute(3, 1, 4, 0, 2).reshape(o.C, cin, 3, 3)
        else:
            nw = taps * cin * o.C
            if d.wfmt == 2:     # split precision, N-concatenated: slab rows W_hi then W_lo
                w2 = self.flat(d.weight, torch.float16)[:2 * nw].view(taps, cin // 8, 2 * o.C, 8).float()
                wp = w2[:, :, :o.C] + w2[:, :, o.C:]
            else:
                wp = self.flat(d.weight, torch.float16)[:nw].view(taps, cin // 8, o.C, 8).float()
                if i.lo:        # split-precision mode: W_lo follows W_hi
                    wp = wp + self.flat(d.weight, torch.float16)[nw:2 * nw].view(taps, cin // 8, o.C, 8).float()
            w = wp.permute(2, 1, 3, 0).reshape(o.C, cin, d.kh, d.kw)
        b = self.flat(d.bias, torch.float32)[:o.C]
        xin = depth_to_space(self.act_get(i)) if d.in_s2d else self.act_get(i)      # (in_s2d: `in` is the phase-split input)
        y = F.conv2d(xin, w, b, stride=d.stride, padding=d.pad)
        if d.relu == 2:
            y = F.relu(y)
        if d.residual:
            r = L.Act(d.residual, d.res_plane_stride, o.C, o.N, o.H, o.W, d.residual_lo)
            y = y + self.act_get(r)
        if d.relu == 1:
            y = F.relu(y)
        if d.out_s2d.data:
            self.act_set(d.out_s2d, space_to_depth(y))
        if not d.s2d_only:
            self.act_set(o, y)

    def _op3(self, d):      # fuse sum
        y = None
        for k in range(d.n_in):
            t = self.act_get(d.in_[k])
            if d.shift[k]:
                t = F.interpolate(t, scale_factor=2 ** d.shift[k], mode='nearest')
            y = t if y is None else y + t
        self.act_set(d.out, F.relu(y) if d.relu else y)

    def _op4(self, d):      # bilinear x2
        self.act_set(d.out, F.interpolate(self.act_get(d.in_), scale_factor=2, mode='bilinear', align_corners=True))

    def _op5(self, d):      # maxpool
        self.act_set(d.out, F.max_pool2d(self.act_get(d.in_), 3, 2, 1))

    def _op6(self, d):      # avgpool
        x = self.act_get(d.in_)
        self.mat(d.out, x.shape[0], x.shape[1], d.ld)[:] = x.mean(dim=(2, 3))

    def _op7(self, d):      # unpack
        x = self.act_get(d.in_)[:, :d.c_valid]
        self.flat(d.out, torch.float32)[:x.numel()] = x.reshape(-1)

    def _op8(self, d):      # linear
        x = self.mat(d.x, d.M, d.I, d.ldx)
        w = self.flat(d.w, torch.float32)[:d.O * d.I].view(d.O, d.I)
        y = x @ w.t()
        if d.b:
            y = y + self.flat(d.b, torch.float32)[:d.O]
        if d.act == 1:
            y = torch.sigmoid(y)
        elif d.act == 2:
            y = F.softplus(y)
        if d.res:
            y = y + self.mat(d.res, d.M, d.O, d.ldres)
        self.mat(d.y, d.M, d.O, d.ldy)[:] = y

    def _op9(self, d):      # copy2d
        src = self.mat(d.src, 1 if d.bcast else d.rows, d.cols, d.lds)
        self.mat(d.dst, d.rows, d.cols, d.ldd)[:] = src

    def _op10(self, d):     # rot6d
        rows = d.n // d.per_row
        x = self.mat(d.x, rows, d.per_row * 6, d.ldx).reshape(-1, 3, 2)
        a1, a2 = x[:, :, 0], x[:, :, 1]
        b1 = F.normalize(a1)
        b2 = F.normalize(a2 - (b1 * a2).sum(1, keepdim=True) * b1)
        b3 = torch.linalg.cross(b1, b2, dim=1)
        self.flat(d.out, torch.float32)[:d.n * 9] = torch.stack((b1, b2, b3), -1).reshape(-1)

    def _op11(self, d):     # PARE head
        pf, sf = self.act_get(d.part_feats), self.act_get(d.smpl_feats)
        N, _, H, W = pf.shape
        f32 = lambda p, n: self.flat(p, torch.float32)[:n]
        wkp, bkp = f32(d.w_kp, 25 * 128).view(25, 128), f32(d.b_kp, 25)
        segm = torch.einsum('nchw,jc->njhw', pf, wkp) + bkp.view(1, 25, 1, 1)
        f32(d.segm, segm.numel())[:] = segm.reshape(-1)
        att = F.softmax(segm[:, 1:].reshape(N, 24, -1), -1)
        pl = torch.einsum('njp,ncp->ncj', att, sf.reshape(N, 128, -1))           # [N,128,24]
        f32(d.uncert_feat, N * 3072)[:] = pl.reshape(-1)
        wsf, bsf = f32(d.w_sf, 64 * 128).view(64, 128), f32(d.b_sf, 64)
        cs = torch.einsum('oc,ncj->noj', wsf, pl) + bsf.view(1, 64, 1)
        wp = f32(d.w_pose, 6 * 128 * 24).view(6, 128, 24)
        p6 = torch.einsum('ncj,ocj->njo', pl, wp)                               # [N,24,6]
        f32(d.pose6d, N * 144)[:] = p6.reshape(-1)
        flat = cs.reshape(N, -1)
        f32(d.shape, N * 10)[:] = (flat @ f32(d.w_shape, 15360).view(10, 1536).t() + f32(d.b_shape, 10)).reshape(-1)
        f32(d.cam, N * 3)[:] = (flat @ f32(d.w_cam, 4608).view(3, 1536).t() + f32(d.b_cam, 3)).reshape(-1)
        x = p6.reshape(-1, 3, 2)
        a1, a2 = x[:, :, 0], x[:, :, 1]
        b1 = F.normalize(a1)
        b2 = F.normalize(a2 - (b1 * a2).sum(1, keepdim=True) * b1)
        b3 = torch.linalg.cross(b1, b2, dim=1)
        f32(d.rotmat, N * 216)[:] = torch.stack((b1, b2, b3), -1).reshape(-1)


def _op15(self, d):         # conv chain = its segments in order
    for k in range(d.n_seg):
        self._op2(d.seg[k])


def _op19(self, d):         # fused BasicBlock: conv1 -> fp16 intermediate -> conv2 + input as residual
    i, o = d.in_, d.out
    c = o.C
    x = self.act_get(i)
    ws = [self.flat(w_, torch.float16)[:9 * c * c].view(9, c // 8, c, 8).float().permute(2, 1, 3, 0).reshape(c, c, 3, 3)
          for w_ in (d.weight1, d.weight2)]
    bs = [self.flat(b_, torch.float32)[:c] for b_ in (d.bias1, d.bias2)]
    mid = F.relu(F.conv2d(x, ws[0], bs[0], padding=1)).half().float()      # the kernel keeps it as fp16 in shared memory
    y = F.relu(F.conv2d(mid, ws[1], bs[1], padding=1) + x)
    self.act_set(o, y)
    if d.out_s2d.data:
        self.act_set(d.out_s2d, space_to_depth(y))


def _op20(self, d):         # fused Bottleneck tail: conv2 3x3 -> fp16 intermediate -> conv3 1x1 + residual
    i, o = d.in_, d.out
    cm, co = i.C, o.C
    w2 = self.flat(d.weight2, torch.float16)[:9 * cm * cm].view(9, cm // 8, cm, 8).float().permute(2, 1, 3, 0).reshape(cm, cm, 3, 3)
    w3 = self.flat(d.weight3, torch.float16)[:cm * co].view(1, cm // 8, co, 8).float().permute(2, 1, 3, 0).reshape(co, cm, 1, 1)
    b2, b3 = self.flat(d.bias2, torch.float32)[:cm], self.flat(d.bias3, torch.float32)[:co]
    mid = F.relu(F.conv2d(self.act_get(i), w2, b2, padding=1)).half().float()
    r = L.Act(d.residual, d.res_plane_stride, co, o.N, o.H, o.W, None)
    self.act_set(o, F.relu(F.conv2d(mid, w3, b3) + self.act_get(r)))


def _op21(self, d):         # resident branch: n_blocks BasicBlocks, every hand-off an fp16 tensor in shared memory
    i, o = d.in_, d.out
    c = o.C
    x = self.act_get(i)
    for blk in range(d.n_blocks):
        ws = [self.flat(d.weight[2 * blk + j], torch.float16)[:9 * c * c].view(9, c // 8, c, 8).float().permute(2, 1, 3, 0).reshape(c, c, 3, 3)
              for j in range(2)]
        bs = [self.flat(d.bias[2 * blk + j], torch.float32)[:c] for j in range(2)]
        mid = F.relu(F.conv2d(x, ws[0], bs[0], padding=1)).half().float()
        x = F.relu(F.conv2d(mid, ws[1], bs[1], padding=1) + x).half().float()
    self.act_set(o, x)


Emu._op15 = _op15
Emu._op21 = _op21
Emu._op19 = _op19
Emu._op20 = _op20
Emu._op13 = lambda self, d: None      # fork / join of plan lanes: the interpreter is sequential
Emu._op14 = lambda self, d: None


def run_plan_ops(ops, keep):
    emu = Emu(keep)
    with torch.no_grad():
        for op in ops:
            emu.run(op)
