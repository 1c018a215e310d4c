"""Architecture descriptions of the POCO hot path, written ONCE against a small backend protocol.

Every function below walks a network (backbone or head) and calls backend methods
(`conv_bn`, `fuse_sum`, `upsample2x`, `linear`, ...).  Two backends implement the protocol:

  * `SpecBackend` (this file)      -- records parameter / buffer names and shapes.  poco_b200.POCO
    uses it in __init__ to register tensors under the reference's state-dict names, so reference
    checkpoints load unchanged (poco.py:131-154, train_utils.py:69-90).
  * `engine.PlanBuilder`           -- emits op descriptors for libpoco_b200.so.

Reference structure being described (no code shared with it -- the reference builds nn.Modules,
we build a flat op schedule): hrnet.py:275-528, hrnet_cls.py:250-486, resnet.py:124-217,
pare_head.py:36-389/:669-928, cliff_head.py:9-127, poco_head.py:15-154, nf_head.py:33-76.
"""
from collections import OrderedDict


class SymAct:
    """shape-only activation used by SpecBackend"""

    def __init__(self, C, H, W):
        self.C, self.H, self.W = C, H, W

    def channels(self, c0, c1):
        return SymAct(c1 - c0, self.H, self.W)


class SpecBackend:
    mode = 'spec'

    def __init__(self):
        self.spec = OrderedDict()       # name -> (shape, kind)  kind in {'param', 'buffer', 'long'}

    def _p(self, name, shape, kind='param'):
        self.spec[name] = (tuple(shape), kind)

    def _bn(self, name, c):
        self._p(name + '.weight', (c,))
        self._p(name + '.bias', (c,))
        self._p(name + '.running_mean', (c,), 'buffer')
        self._p(name + '.running_var', (c,), 'buffer')
        self._p(name + '.num_batches_tracked', (), 'long')

    def pack_image(self, img, H, W):
        return SymAct(16, H, W)

    def stem_conv(self, img, H, W, conv, bn, cout, s2d=None):
        self._p(conv + '.weight', (cout, 3, 3, 3))
        self._bn(bn, cout)
        return SymAct(cout, H // 2, W // 2)

    def conv_bn(self, x, conv, bn, cin, cout, k, stride=1, relu=True, residual=None, out=None, pad=None, bias=False,
                s2d=None):
        convs = conv if isinstance(conv, (list, tuple)) else [conv]
        bns = bn if isinstance(bn, (list, tuple)) else [bn] * len(convs)
        each = cout // len(convs)
        for cv, b_ in zip(convs, bns):
            self._p(cv + '.weight', (each, cin, k, k))
            if bias:
                self._p(cv + '.bias', (each,))
            if b_ is not None:
                self._bn(b_, each)
        pad = k // 2 if pad is None else pad
        return SymAct(cout, (x.H + 2 * pad - k) // stride + 1, (x.W + 2 * pad - k) // stride + 1)

    def param_only_conv(self, name, cin, cout, k, bias=True):
        self._p(name + '.weight', (cout, cin, k, k))
        if bias:
            self._p(name + '.bias', (cout,))

    def param_only_linear(self, name, i, o):
        self._p(name + '.weight', (o, i))
        self._p(name + '.bias', (o,))

    def buffer(self, name, shape):
        self._p(name, shape, 'buffer')

    def param(self, name, shape):
        self._p(name, shape)

    def act(self, C, H, W):
        return SymAct(C, H, W)

    def fuse_sum(self, terms, relu, out=None):
        a, s = terms[0]
        return out or SymAct(a.C, a.H << s, a.W << s)

    def upsample2x(self, x):
        return SymAct(x.C, 2 * x.H, 2 * x.W)

    def maxpool(self, x):
        return SymAct(x.C, (x.H - 1) // 2 + 1, (x.W - 1) // 2 + 1)

    def free(self, a):
        pass

    N = 1

    def fork(self, costs):
        pass

    def set_lane(self, k):
        pass

    def join(self):
        pass

    def begin_chain(self):
        pass

    def end_chain(self, groups=1):
        pass


def conv_cost(N, cin, cout, k, Hout, Wout):
    """relative SM-cycles of one conv launch, used only to split the SMs between concurrent lanes.
    tiles x MMAs x effective cycles per MMA + a fixed launch cost; the per-MMA constants (N <= 32, N = 64,
    N = 128, N = 256 with weights re-streamed from L2 per tile) were calibrated on per-tile times at batch 256.
    Re-weighting them from the measured lane end times of the HR modules (tools/region_times.py: in eager replay
    the 256-channel lane of a stage-4 module ends at 1.00 ms, the 32-channel one at 0.73 ms) did NOT improve the
    captured-graph step (A/B: 20.18 k vs 20.13-20.19 k crops/s, more aggressive weights 19.3-19.7 k), so they
    stay.  Round 2, after the fused block kernels and the prioritised lanes: the 256-channel 7x7 lane ends the four-branch
    modules (0.80 ms against 0.60-0.65 for the others, tools/region_times.py) because its eight dependent launches queue
    behind the other lanes' long-lived persistent CTAs; c256 260 -> 400 gives it a larger share: 10.74 -> 10.65 ms per
    step (A/B twice on one box; 550: the same).  POCO_B200_COST="c32,c64,c128,c256" overrides them."""
    import os
    env = os.environ.get('POCO_B200_COST')
    c32, c64, c128, c256 = [float(v) for v in env.split(',')] if env else (105.0, 105.0, 170.0, 400.0)
    tiles = -(-N * (Hout + 2) * (Wout + 2) // 128)
    n_tile = min(cout, 256)
    streamed = k * k * cin * n_tile * 2 > 112 * 1024
    if n_tile <= 32:
        cyc = c32
    elif n_tile <= 64:
        cyc = c64
    elif n_tile <= 128:
        cyc = c128 if streamed else 0.6 * c128
    else:
        cyc = c256 if streamed else 0.5 * c256
    return tiles * (k * k * -(-cin // 16)) * -(-cout // n_tile) * cyc + 1500000


# ------------------------------------------------------------------------------------------------
# residual blocks
# ------------------------------------------------------------------------------------------------
def basic_block(b, x, name, cin, cout, free_input=True, s2d_out=None):
    """conv3x3-BN-ReLU, conv3x3-BN, += x, ReLU  (hrnet.py:42-58); HRNet branches never downsample.
    s2d_out: the block's output also feeds stride-2 convs -> its last conv writes the phase-split copy too.
    A backend may run the whole block as one launch (PlanBuilder.basic_block_fused: the intermediate stays on the SM)."""
    fused = getattr(b, 'basic_block_fused', None)
    if fused is not None and cin == cout:
        o = fused(x, name, cin, s2d=s2d_out)
        if o is not None:
            if free_input:
                b.free(x)
            return o
    y = b.conv_bn(x, name + '.conv1', name + '.bn1', cin, cout, 3)
    o = b.conv_bn(y, name + '.conv2', name + '.bn2', cout, cout, 3, residual=x, s2d=s2d_out)
    b.free(y)
    if free_input:
        b.free(x)
    return o


def bottleneck(b, x, name, cin, planes, stride=1, downsample=False, free_input=True, s2d_out=None):
    """1x1 -> 3x3(stride) -> 1x1 x4, residual (v1.5)  (hrnet.py:79-99, resnet.py:100-121).
    s2d_out: see basic_block (a strided block reads phase-split copies: of conv1's output, and of the block input
    for the strided 1x1 downsample when the previous block wrote one)"""
    cout = planes * 4
    y1 = b.conv_bn(x, name + '.conv1', name + '.bn1', cin, planes, 1)      # (phase-split hand-off to a strided conv2 measured slower on ResNet-50)
    can_fuse = getattr(b, 'bottleneck_tail_supported', None)
    if can_fuse is not None and stride == 1 and s2d_out is None and can_fuse(y1, planes, cout):
        # conv2 -> conv3 + residual as one launch (PlanBuilder.bottleneck_tail_fused: conv2's tile never leaves the SM)
        r = x
        if downsample:
            r = b.conv_bn(x, name + '.downsample.0', name + '.downsample.1', cin, cout, 1, 1, relu=False)
        o = b.bottleneck_tail_fused(y1, name, planes, cout, r)
        b.free(y1)
        if r is not x:
            b.free(r)
        if free_input:
            b.free(x)
        return o
    y2 = b.conv_bn(y1, name + '.conv2', name + '.bn2', planes, planes, 3, stride)
    b.free(y1)
    r = x
    if downsample:
        r = b.conv_bn(x, name + '.downsample.0', name + '.downsample.1', cin, cout, 1, stride, relu=False)
    o = b.conv_bn(y2, name + '.conv3', name + '.bn3', planes, cout, 1, residual=r, s2d=s2d_out)
    b.free(y2)
    if r is not x:
        b.free(r)
    if free_input:
        b.free(x)
    return o


# ------------------------------------------------------------------------------------------------
# HRNet trunk (shared by the pose variant and the classification variant)
# ------------------------------------------------------------------------------------------------
def chain_groups(branch, n_branches):
    """crop ranges per branch chain (PlanBuilder.end_chain): POCO_B200_GROUPS="g0,g1,g2,g3" overrides"""
    import os
    env = os.environ.get('POCO_B200_GROUPS')
    table = [int(v) for v in env.split(',')] if env else [1, 1, 1, 1]
    return table[branch] if branch < len(table) else 1


def chain_policy(channels, N=256, latency_mode=False):
    """Run the eight BasicBlock convs of an HRNet branch as ONE persistent chained launch?
    Small batches (N <= 16 crops, the video-stream case: a handful of detections per frame) are launch-latency bound --
    a conv is ~10 us of launch gap + pipeline fill around ~1 us of work -- and the chain removes that per conv
    (1080p stream, 8 detections: 2.15 -> 1.92 ms per frame).  At batch 256 the flag protocol's gpu-scope fences cost
    what the saved launches give back (profiles/r01b_chain_vs_separate.csv), so large batches launch conv by conv.
    Opt-in (POCO(latency_mode=True), what the stream driver uses): a chain accumulates its K chunks in another order
    than the separate launches, so switching it by batch size would break "a crop's result is bitwise independent of
    the batch it travels in" across the threshold.
    POCO_B200_CHAIN_MIN_C=<c> overrides: chain the branches with at least c channels at every batch size."""
    import os
    env = os.environ.get('POCO_B200_CHAIN_MIN_C')
    if env is not None:
        return channels >= int(env)
    return bool(latency_mode) and N <= 16


def _branch_phase(b, xs, dims, name, chans):
    """the four BasicBlocks of every branch of one HighResolutionModule (hrnet.py:140-186):
    -> (per-lane costs, emit(i)); emit(i) emits branch i on xs[i] in the current lane and stores its output in xs[i]"""
    import os
    nb = len(chans)
    N = b.N
    # The branch's eight convs share one geometry and CAN run as one persistent chained launch
    # (POCO_B200_CHAIN_MIN_C, off by default: see chain_policy).
    chain_on = [chain_policy(chans[i], N, getattr(b, 'latency_mode', False)) for i in range(nb)]
    # A backend may run the four blocks as ONE launch with the crop resident in shared memory (PlanBuilder.branch_fused:
    # the 128-channel 14x14 branch); not when the last block must also write a phase-split copy.
    fusable = getattr(b, 'branch_fusable', None)
    resident = [fusable is not None and not chain_on[i] and not (nb - 1 - i >= 2) and fusable(chans[i], dims[i][0], dims[i][1], 4)
                for i in range(nb)]
    # (a resident branch needs about half the SM time of its eight separate launches: tools/branch_bench.py)
    rcost = float(os.environ.get('POCO_B200_BRANCH_COST', '0.5'))
    # (and a fused 64-channel block ~0.85 of its two launches: tools/bblock_bench.py; a block writing a phase-split copy is not fused)
    fusable64 = getattr(b, 'basic_block_fusable', None)
    c64 = float(os.environ.get('POCO_B200_BLOCK64_COST', '0.85'))
    scale = [rcost if resident[i] else (c64 if (chans[i] == 64 and fusable64 is not None and not chain_on[i] and
                                                fusable64(64, dims[i][0], dims[i][1])) else 1.0) for i in range(nb)]
    costs = [8 * conv_cost(N, chans[i], chans[i], 3, dims[i][0], dims[i][1]) * scale[i] for i in range(nb)]

    def emit(i):
        x = xs[i]
        chained = chain_on[i]
        if resident[i]:
            o = b.branch_fused(x, [f'{name}.branches.{i}.{k}' for k in range(4)], chans[i])
            if o is not None:
                if o is not x:
                    b.free(x)
                xs[i] = o
                return o
        if chained:
            b.begin_chain()
        for k in range(4):
            # A branch's output feeds the stride-2 fuse convs of every lower-resolution output.  Its last conv writes the
            # phase-split copy as well when at least two of them read it: the second write of the tensor costs about what
            # ONE consumer gains by leaving the gather path (measured: +16 us on 32->32 @56, -11...-15 us per consumer).
            x = basic_block(b, x, f'{name}.branches.{i}.{k}', chans[i], chans[i],
                            s2d_out='dual' if (k == 3 and nb - 1 - i >= 2 and not chained) else None)
        if chained:
            b.end_chain(groups=chain_groups(i, nb))
        xs[i] = x
        return x
    return costs, emit


def _fuse_phase(b, xs, dims, name, chans, out0=None):
    """the multi-resolution fuse of one HighResolutionModule (hrnet.py:213-266): -> (per-lane costs, emit(i)); emit(i) emits
    output i from the branch outputs xs in the current lane and returns it (xs stay allocated: the caller frees them)"""
    nb = len(chans)
    N = b.N

    def fuse_cost(i):
        c = 0
        for j in range(i + 1, nb):
            c += conv_cost(N, chans[j], chans[i], 1, *dims[j])
        for j in range(i):
            for k in range(i - j):
                co = chans[i] if k == i - j - 1 else chans[j]
                c += conv_cost(N, chans[j], co, 3, dims[j][0] >> (k + 1), dims[j][1] >> (k + 1)) * 2   # gather mode
        return c + N * dims[i][0] * dims[i][1] * chans[i] // 4          # + the element-wise sum

    def emit(i):
        ups = []
        for j in range(i + 1, nb):      # 1x1 conv + BN at low resolution; nearest upsample folded into the sum
            z = b.conv_bn(xs[j], f'{name}.fuse_layers.{i}.{j}.0', f'{name}.fuse_layers.{i}.{j}.1',
                          chans[j], chans[i], 1, relu=False)
            ups.append((z, j - i))
        if i == 0:
            o = b.fuse_sum([(xs[0], 0)] + ups, relu=True, out=out0)        # (out0: the caller's buffer slice)
        else:
            acc = b.fuse_sum([(xs[i], 0)] + ups, relu=False) if ups else xs[i]
            for j in range(i):          # stride-2 conv chains; the running sum rides on the residual input
                t = xs[j]
                for k in range(i - j):
                    f = f'{name}.fuse_layers.{i}.{j}.{k}'
                    if k != i - j - 1:      # (read by the next stride-2 conv only: phase-split form only)
                        t2 = b.conv_bn(t, f + '.0', f + '.1', chans[j], chans[j], 3, 2, relu=True, s2d='only')
                    else:
                        t2 = b.conv_bn(t, f + '.0', f + '.1', chans[j], chans[i], 3, 2, relu=(j == i - 1), residual=acc)
                        if acc is not xs[i]:
                            b.free(acc)
                        acc = t2
                    if t is not xs[j]:
                        b.free(t)
                    t = t2
            o = acc
        for z, _ in ups:
            b.free(z)
        return o
    return [fuse_cost(i) for i in range(nb)], emit


def _lane_order(nb):
    """emission order of the lanes = creation order of their graph nodes: POCO_B200_LANE_ORDER=rev enqueues the low-resolution
    branches (dependent chains of short launches, the lanes that end a module) before the 32-channel lane's long kernels
    (measured slightly slower: not the default)"""
    import os
    return list(range(nb))[::-1] if os.environ.get('POCO_B200_LANE_ORDER', 'fwd') == 'rev' else list(range(nb))


def hr_module(b, xs, name, chans, out0=None):
    """HighResolutionModule: 4 BasicBlocks per branch, then the multi-resolution fuse
    (hrnet.py:188-266).  Inputs are consumed (freed).  The branches, and afterwards the per-output
    fuse chains, are independent: they are emitted as concurrent plan lanes, each with a share of the
    SMs proportional to its work (the 14x14 / 7x7 branches cannot fill 148 SMs on their own)."""
    import os
    nb = len(xs)
    dims = [(x.H, x.W) for x in xs]
    costs, emit = _branch_phase(b, xs, dims, name, chans)
    b.fork(costs)
    for i in _lane_order(nb):
        b.set_lane(i)
        emit(i)
    b.join()
    if nb == 1:
        return xs
    costs, emit = _fuse_phase(b, xs, dims, name, chans, out0)
    b.fork(costs)
    outs = [None] * nb
    for i in (_lane_order(nb) if os.environ.get('POCO_B200_LANE_ORDER_FUSE', '0') == '1' else range(nb)):
        b.set_lane(i)
        outs[i] = emit(i)
    b.join()
    for x in xs:
        b.free(x)
    return outs


def hr_stage(b, xs, names, chans, out0=None):
    """The HighResolutionModules `names` of one HRNet stage (hrnet.py:386-412) with the fuse of module m and the branches of
    module m + 1 in ONE fork / join region: branch i of the next module reads fuse output i only, so lane i runs
    fuse_i(m) -> branch_i(m + 1) without waiting for the other lanes; only the fuse needs every branch, i.e. a real join.
    One barrier per module instead of two (POCO_B200_MERGE_PHASES, see hrnet_trunk).  out0: see hr_module (last module)."""
    nb = len(xs)
    dims = [(x.H, x.W) for x in xs]
    set_shares = getattr(b, 'set_shares', None)        # (PlanBuilder: SM shares of the ops emitted next, per sub-phase)
    bcosts, bemit = _branch_phase(b, xs, dims, names[0], chans)
    b.fork(bcosts)
    for i in range(nb):
        b.set_lane(i)
        bemit(i)
    b.join()
    for m, name in enumerate(names):
        last = m == len(names) - 1
        fcosts, femit = _fuse_phase(b, xs, dims, name, chans, out0 if last else None)
        outs = [None] * nb
        if last:
            b.fork(fcosts)
            for i in range(nb):
                b.set_lane(i)
                outs[i] = femit(i)
        else:
            bcosts, bemit = _branch_phase(b, outs, dims, names[m + 1], chans)
            b.fork([f + c for f, c in zip(fcosts, bcosts)])
            for i in range(nb):
                b.set_lane(i)
                if set_shares is not None:
                    set_shares(fcosts)
                outs[i] = femit(i)
                if set_shares is not None:
                    set_shares(bcosts)
                bemit(i)                    # (runs on outs[i] and replaces it)
        b.join()
        for x in xs:
            b.free(x)
        xs = outs
    return xs


def hrnet_trunk(b, img, widths, H=224, W=224, prefix='backbone.', final_out0=None):
    """final_out0(H, W) -> activation slice the LAST module writes its branch-0 output into (hrnet_pose: the
    first 32 channels of the 480-channel feature buffer, which saves a copy kernel)"""
    p = prefix
    y = b.stem_conv(img, H, W, p + 'conv1', p + 'bn1', 64, s2d='only')      # read by the stride-2 conv2 only
    x = b.conv_bn(y, p + 'conv2', p + 'bn2', 64, 64, 3, 2)
    b.free(y)
    for k in range(4):
        x = bottleneck(b, x, f'{p}layer1.{k}', 64 if k == 0 else 256, 64, downsample=(k == 0))
    # transition1: 256 -> [w0 @56 (3x3 s1), w1 @28 (3x3 s2)]
    x0 = b.conv_bn(x, p + 'transition1.0.0', p + 'transition1.0.1', 256, widths[0], 3)
    x1 = b.conv_bn(x, p + 'transition1.1.0.0', p + 'transition1.1.0.1', 256, widths[1], 3, 2)
    b.free(x)
    ys = hr_module(b, [x0, x1], p + 'stage2.0', widths[:2])
    n_modules = {3: 4, 4: 3}
    for st in (3, 4):
        nbr = st
        new = b.conv_bn(ys[-1], f'{p}transition{st - 1}.{nbr - 1}.0.0', f'{p}transition{st - 1}.{nbr - 1}.0.1',
                        widths[nbr - 2], widths[nbr - 1], 3, 2)
        ys = ys + [new]
        import os
        if os.environ.get('POCO_B200_MERGE_PHASES', '0') == '1' and nbr > 1:
            ys = hr_stage(b, ys, [f'{p}stage{st}.{m}' for m in range(n_modules[st])], widths[:nbr],
                          out0=final_out0(ys[0].H, ys[0].W) if (st == 4 and final_out0 is not None) else None)
            continue
        for m in range(n_modules[st]):
            last = st == 4 and m == n_modules[st] - 1 and final_out0 is not None
            ys = hr_module(b, ys, f'{p}stage{st}.{m}', widths[:nbr],
                           out0=final_out0(ys[0].H, ys[0].W) if last else None)
    return ys


def hrnet_pose(b, img, width=32, prefix='backbone.'):
    """PoseHighResolutionNet with use_conv=True / downsample=False -> [N, 15*width, 56, 56]
    (hrnet.py:466-528).  Each branch lands in its channel slice of one buffer, so torch.cat is free."""
    widths = [width, 2 * width, 4 * width, 8 * width]
    holder = {}

    def out0(H, W):
        holder['feats'] = b.act(sum(widths), H, W)
        return holder['feats'].channels(0, widths[0])
    ys = hrnet_trunk(b, img, widths, prefix=prefix, final_out0=out0)
    feats = holder['feats']             # branch 0 of the last module already landed in its first channels
    c0 = widths[0]
    # The three up-sampling chains are independent.  POCO_B200_OUT_LANES=1 runs them as concurrent lanes, the longest
    # (256 channels, three steps) on the lane with the highest stream priority, so that the HBM-bound upsample kernels
    # and the short convs of the other two fill in around its tensor-bound convs.
    import os
    lanes = os.environ.get('POCO_B200_OUT_LANES', '0') == '1' and hasattr(b, 'fork') and getattr(b, 'use_lanes', False)
    if lanes:
        def chain_cost(br):
            h, c = ys[br].H, 0
            for k in range(br):
                h *= 2
                c += conv_cost(b.N, widths[br], widths[br], 3, h, h) + b.N * h * h * widths[br] // 2
            return c
        b.fork([chain_cost(br) for br in (1, 2, 3)])
    for br in range(1, 4):
        if lanes:
            b.set_lane(br - 1)
        t = ys[br]
        for k in range(br):
            u = b.upsample2x(t)
            b.free(t)
            name = f'{prefix}upsample_stage_{br + 1}'
            last = k == br - 1
            t = b.conv_bn(u, f'{name}.{4 * k + 1}', f'{name}.{4 * k + 2}', widths[br], widths[br], 3,
                          out=feats.channels(c0, c0 + widths[br]) if last else None)
            b.free(u)
        c0 += widths[br]
    if lanes:
        b.join()
    if b.mode == 'spec':
        b.param_only_conv(prefix + 'final_layer', widths[0], 24, 1)      # present in checkpoints, never executed
    return feats


def hrnet_cls(b, img, width=48, prefix='backbone.'):
    """HighResolutionNet + classification head up to the 2048-channel 7x7 map (hrnet_cls.py:438-477);
    the caller pools it."""
    widths = [width, 2 * width, 4 * width, 8 * width]
    ys = hrnet_trunk(b, img, widths, prefix=prefix)
    head = [32, 64, 128, 256]
    # (every y below is read by the next stride-2 downsamp conv only: phase-split form only)
    y = bottleneck(b, ys[0], f'{prefix}incre_modules.0.0', widths[0], head[0], downsample=True, s2d_out='only')
    for i in range(3):
        inc = bottleneck(b, ys[i + 1], f'{prefix}incre_modules.{i + 1}.0', widths[i + 1], head[i + 1], downsample=True)
        d = f'{prefix}downsamp_modules.{i}'
        # y = incre(x_{i+1}) + ReLU(BN(conv3x3 s2 (y)))  -> ReLU *before* the residual add (relu=2)
        y2 = b.conv_bn(y, d + '.0', d + '.1', head[i] * 4, head[i + 1] * 4, 3, 2, relu=2, residual=inc, bias=True,
                       s2d='only' if i < 2 else None)
        b.free(y)
        b.free(inc)
        y = y2
    f = prefix + 'final_layer'
    o = b.conv_bn(y, f + '.0', f + '.1', 1024, 2048, 1, bias=True)
    b.free(y)
    if b.mode == 'spec':
        b.param_only_linear(prefix + 'classifier', 2048, 1000)           # unused (hrnet_cls.py:484)
    return o


def resnet50(b, img, prefix='backbone.'):
    """torchvision-style ResNet-50 v1.5 trunk -> [N, 2048, 7, 7] (resnet.py:201-217)"""
    p = prefix
    x = b.pack_image(img, 224, 224)
    y = b.conv_bn(x, p + 'conv1', p + 'bn1', 3, 64, 7, 2, pad=3)
    b.free(x)
    x = b.maxpool(y)
    b.free(y)
    cin = 64
    for li, (planes, blocks) in enumerate(((64, 3), (128, 4), (256, 6), (512, 3)), start=1):
        for k in range(blocks):
            stride = 2 if (k == 0 and li > 1) else 1
            x = bottleneck(b, x, f'{p}layer{li}.{k}', cin, planes, stride, downsample=(k == 0))
            cin = planes * 4
    return x


# name -> (builder, output channels)
BACKBONES = {
    'hrnet_w32': (lambda b, img: hrnet_pose(b, img, 32), 480),
    'hrnet_w48_cls': (lambda b, img: hrnet_cls(b, img, 48), 2048),
    'resnet50': (resnet50, 2048),
}


# ------------------------------------------------------------------------------------------------
# head parameter specs (the op emission for heads lives in poco.py, next to the output plumbing)
# ------------------------------------------------------------------------------------------------
def pare_head_convs(b, feats, cin, prefix='head.'):
    """the two conv branches of pare_head (_make_conv_layer, pare_head.py:468-491).  Their first
    convs read the same input, so they run as ONE conv with the output channels concatenated."""
    kd, sd = prefix + 'keypoint_deconv_layers', prefix + 'smpl_deconv_layers'
    t = b.conv_bn(feats, [kd + '.0', sd + '.0'], [kd + '.1', sd + '.1'], cin, 256, 3)
    part = b.conv_bn(t.channels(0, 128), kd + '.3', kd + '.4', 128, 128, 3)
    smpl = b.conv_bn(t.channels(128, 256), sd + '.3', sd + '.4', 128, 128, 3)
    b.free(t)
    return part, smpl


def pare_head_spec(b, prefix='head.'):
    b.param_only_conv(prefix + 'keypoint_final_layer', 128, 25, 1)
    b.param_only_conv(prefix + 'smpl_final_layer', 128, 64, 1)
    b.buffer(prefix + 'temperature', ())
    b.buffer(prefix + 'init_pose', (1, 144))
    b.buffer(prefix + 'init_shape', (1, 10))
    b.buffer(prefix + 'init_cam', (1, 3))
    b.param_only_linear(prefix + 'shape_mlp', 64 * 24, 10)
    b.param_only_linear(prefix + 'cam_mlp', 64 * 24, 3)
    b.param(prefix + 'pose_mlp.weight', (1, 6, 128, 24, 1, 1))


def cliff_head_spec(b, nfeat, prefix='head.'):
    b.param_only_linear(prefix + 'fc1', nfeat + 3 + 144 + 13, 1024)
    b.param_only_linear(prefix + 'fc2', 1024, 1024)
    b.param_only_linear(prefix + 'decpose', 1024, 144)
    b.param_only_linear(prefix + 'decshape', 1024, 10)
    b.param_only_linear(prefix + 'deccam', 1024, 3)
    b.buffer(prefix + 'init_pose', (1, 144))
    b.buffer(prefix + 'init_shape', (1, 10))
    b.buffer(prefix + 'init_cam', (1, 3))


def poco_head_layers(nfeat, num_neurons, uncert_inp_type, n_out=24):
    """Layer list of poco_head (poco_head.py:15-82): returns (pose_net, [(name, in, out), ...])."""
    inp = nfeat + (216 if uncert_inp_type == 'feat-pose' else 0)
    nn_ = [inp] + list(num_neurons) + [n_out]
    pre = []
    if 'pose-net' in uncert_inp_type:
        pre = [('uncert_fc_poseNet', 216, nn_[1]), ('uncert_fc_featNet', nn_[0], nn_[1])]
        nn_ = nn_[1:]
        nn_[0] *= 2
    layers = [(f'uncert_fc{i + 1}', nn_[i], nn_[i + 1]) for i in range(len(nn_) - 1)]
    return pre, layers


def poco_head_spec(b, nfeat, num_neurons, uncert_inp_type, prefix='uncert_head.', n_out=24):
    pre, layers = poco_head_layers(nfeat, num_neurons, uncert_inp_type, n_out)
    for name, i, o in pre + layers:
        b.param_only_linear(prefix + name, i, o)


def flow_head_spec(b, nfeat, context_dim, cond, num_flow_layers, num_rv=9, hid=64, prefix='flow_head.'):
    ctx = context_dim if cond else 0
    if cond:
        b.param_only_linear(prefix + 'cond_layer', nfeat, context_dim)
    b.buffer(prefix + 'flow.mask', (2 * num_flow_layers, num_rv))
    for net in ('t', 's'):
        for i in range(2 * num_flow_layers):
            b.param_only_linear(f'{prefix}flow.{net}.{i}.0', num_rv + ctx, hid)
            b.param_only_linear(f'{prefix}flow.{net}.{i}.2', hid, hid)
            b.param_only_linear(f'{prefix}flow.{net}.{i}.4', hid, num_rv)
