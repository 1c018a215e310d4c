"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch fp32, functional) of POCO's per-crop
inference hot path.  Nothing under poco_b200/ may import this file; only tests/, bench.py's
cpu_baseline / --impl reference leg and __graft_entry__.smoke() use it, and only as the checker.

Parity status: PINNED.  tests/test_oracle_golden.py checks every function here against golden
vectors produced by the unmodified reference forward (oracle/make_golden.py, run in the build
container where /root/reference is importable through oracle/ref_loader.py).  The reference has no
tests / golden vectors of its own, so those generated vectors are the pin.

The restatement is *data driven*: it walks the reference's state-dict key names (the checkpoint
contract of poco.py:131-154 / train_utils.py:69-90) and infers widths, block counts and the branch
structure from tensor shapes, so one function covers HRNet-W32 (hrnet.py), HRNet-W48-cls
(hrnet_cls.py) and any other width.  All arithmetic is fp32, eval-mode BatchNorm (eps=1e-5).
"""
import math

import torch
import torch.nn.functional as F

EPS = 1e-5


class _SD:
    """state-dict view with a key prefix"""

    def __init__(self, sd, prefix=''):
        self.sd, self.p = sd, prefix

    def sub(self, name):
        return _SD(self.sd, f'{self.p}{name}.')

    def has(self, name):
        return (self.p + name) in self.sd

    def __getitem__(self, name):
        return self.sd[self.p + name]

    def count(self, name=''):
        """number of consecutive integer children  <prefix><name>.0, .1, ..."""
        base = self.p + (name + '.' if name else '')
        n = 0
        while any(k.startswith(f'{base}{n}.') for k in self.sd):
            n += 1
        return n


def conv(x, s, name, stride=1, pad=None):
    w = s[name + '.weight']
    b = s[name + '.bias'] if s.has(name + '.bias') else None
    if pad is None:
        pad = w.shape[-1] // 2
    return F.conv2d(x, w, b, stride=stride, padding=pad)


def bn(x, s, name):
    return F.batch_norm(x, s[name + '.running_mean'], s[name + '.running_var'],
                        s[name + '.weight'], s[name + '.bias'], False, 0.0, EPS)


# ------------------------------------------------------------------------------------------------
# residual blocks  (hrnet.py:29-99, hrnet_cls.py:29-99, resnet.py:75-121)
# ------------------------------------------------------------------------------------------------
def basic_block(x, s, stride=1):
    out = F.relu(bn(conv(x, s, 'conv1', stride), s, 'bn1'))
    out = bn(conv(out, s, 'conv2'), s, 'bn2')
    res = x
    if s.has('downsample.0.weight'):
        res = bn(conv(x, s, 'downsample.0', stride), s, 'downsample.1')
    return F.relu(out + res)


def bottleneck(x, s, stride=1):
    out = F.relu(bn(conv(x, s, 'conv1'), s, 'bn1'))
    out = F.relu(bn(conv(out, s, 'conv2', stride), s, 'bn2'))      # v1.5: stride on the 3x3
    out = bn(conv(out, s, 'conv3'), s, 'bn3')
    res = x
    if s.has('downsample.0.weight'):
        res = bn(conv(x, s, 'downsample.0', stride), s, 'downsample.1')
    return F.relu(out + res)


def block(x, s, stride=1):
    return bottleneck(x, s, stride) if s.has('conv3.weight') else basic_block(x, s, stride)


def block_seq(x, s, stride=1):
    for i in range(s.count()):
        x = block(x, s.sub(str(i)), stride if i == 0 else 1)
    return x


# ------------------------------------------------------------------------------------------------
# HRNet  (hrnet.py:102-266 module, :466-528 forward; hrnet_cls.py:438-486)
# ------------------------------------------------------------------------------------------------
def hr_module(xs, s):
    nb = s.count('branches')
    xs = [block_seq(xs[i], s.sub(f'branches.{i}')) for i in range(nb)]
    if nb == 1:
        return xs
    outs = []
    for i in range(s.count('fuse_layers')):
        y = None
        for j in range(nb):
            f = s.sub(f'fuse_layers.{i}.{j}')
            if j == i:
                t = xs[j]
            elif j > i:      # 1x1 conv + BN + nearest upsample x2^(j-i)   (hrnet.py:198-209)
                t = bn(conv(xs[j], f, '0'), f, '1')
                t = F.interpolate(t, scale_factor=2 ** (j - i), mode='nearest')
            else:            # chain of 3x3 stride-2 convs, ReLU on all but the last (hrnet.py:213-240)
                t = xs[j]
                for k in range(i - j):
                    t = bn(conv(t, f, f'{k}.0', 2), f, f'{k}.1')
                    if k != i - j - 1:
                        t = F.relu(t)
            y = t if y is None else y + t
        outs.append(F.relu(y))
    return outs


def hr_transition(s, name, prev, nb_cur):
    """hrnet.py:345-384 + the call sites :476-502 (new branches are fed from prev[-1])"""
    xs = []
    for i in range(nb_cur):
        t = s.sub(f'{name}.{i}')
        if i < len(prev):
            if t.has('0.weight'):
                xs.append(F.relu(bn(conv(prev[i] if len(prev) > 1 else prev[0], t, '0'), t, '1')))
            else:
                xs.append(prev[i])
        else:
            y = prev[-1]
            for j in range(t.count()):
                y = F.relu(bn(conv(y, t, f'{j}.0', 2), t, f'{j}.1'))
            xs.append(y)
    return xs


def hr_stage(xs, s, name):
    for m in range(s.count(name)):
        xs = hr_module(xs, s.sub(f'{name}.{m}'))
    return xs


def hrnet_trunk(x, s):
    x = F.relu(bn(conv(x, s, 'conv1', 2), s, 'bn1'))
    x = F.relu(bn(conv(x, s, 'conv2', 2), s, 'bn2'))
    x = block_seq(x, s.sub('layer1'))
    ys = [x]
    for st, tr in ((2, 'transition1'), (3, 'transition2'), (4, 'transition3')):
        nb = s.count(f'stage{st}.0.branches')
        xs = hr_transition(s, tr, ys, nb)
        ys = hr_stage(xs, s, f'stage{st}')
    return ys


def hrnet_pose(x, sd, prefix='backbone.'):
    """PoseHighResolutionNet.forward with use_conv=True, downsample=False (hrnet.py:466-528,
    :437-450, :515-519): bilinear(align_corners) x2 + conv3x3 + BN + ReLU per level, then concat."""
    s = _SD(sd, prefix)
    ys = hrnet_trunk(x, s)
    outs = [ys[0]]
    for b in range(1, len(ys)):
        u = s.sub(f'upsample_stage_{b + 1}')
        t = ys[b]
        for k in range(b):
            t = F.interpolate(t, scale_factor=2, mode='bilinear', align_corners=True)
            t = F.relu(bn(conv(t, u, str(4 * k + 1)), u, str(4 * k + 2)))
        outs.append(t)
    return torch.cat(outs, 1)


def hrnet_cls(x, sd, prefix='backbone.'):
    """HighResolutionNet.forward (hrnet_cls.py:438-486) incl. classification head (:306-353)."""
    s = _SD(sd, prefix)
    ys = hrnet_trunk(x, s)
    y = block_seq(ys[0], s.sub('incre_modules.0'))
    for i in range(s.count('downsamp_modules')):
        d = s.sub(f'downsamp_modules.{i}')
        y = block_seq(ys[i + 1], s.sub(f'incre_modules.{i + 1}')) + F.relu(bn(conv(y, d, '0', 2), d, '1'))
    f = s.sub('final_layer')
    y = F.relu(bn(conv(y, f, '0'), f, '1'))
    return y.mean(dim=(2, 3))


def resnet(x, sd, prefix='backbone.'):
    """ResNet._forward_impl (resnet.py:201-217), avgpool/fc removed."""
    s = _SD(sd, prefix)
    x = F.relu(bn(conv(x, s, 'conv1', 2, 3), s, 'bn1'))
    x = F.max_pool2d(x, 3, 2, 1)
    for l in range(1, 5):
        x = block_seq(x, s.sub(f'layer{l}'), 1 if l == 1 else 2)
    return x


def backbone(x, sd, name):
    if name.startswith('hrnet') and name.endswith('_cls'):
        return hrnet_cls(x, sd)
    if name.startswith('hrnet'):
        return hrnet_pose(x, sd)
    if name.startswith('resnet'):
        return resnet(x, sd)
    raise ValueError(name)


# ------------------------------------------------------------------------------------------------
# small math
# ------------------------------------------------------------------------------------------------
def rot6d_to_rotmat(x):
    """geometry.py:247-261 -- NB the 6 numbers are a 3x2 row-major matrix (a1 = elems 0,2,4)."""
    x = x.reshape(-1, 3, 2)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = a1 / a1.norm(dim=1, keepdim=True).clamp_min(1e-12)
    u = a2 - (b1 * a2).sum(1, keepdim=True) * b1
    b2 = u / u.norm(dim=1, keepdim=True).clamp_min(1e-12)
    b3 = torch.linalg.cross(b1, b2, dim=1)
    return torch.stack((b1, b2, b3), dim=-1)


def linear(x, s, name):
    return F.linear(x, s[name + '.weight'], s[name + '.bias'])


# ------------------------------------------------------------------------------------------------
# heads
# ------------------------------------------------------------------------------------------------
def pare_head(feats, sd, prefix='head.'):
    """pare_head.forward with POCO's 3-arg construction (poco.py:67): part_segm attention,
    non-iterative regression (pare_head.py:669-752, :754-826, :855-928)."""
    s = _SD(sd, prefix)
    B = feats.shape[0]

    def branch(x, name):         # _make_conv_layer: [conv3x3, BN, ReLU] x2  (pare_head.py:468-491)
        b = s.sub(name)
        x = F.relu(bn(conv(x, b, '0'), b, '1'))
        return F.relu(bn(conv(x, b, '3'), b, '4'))

    part_feats = branch(feats, 'keypoint_deconv_layers')
    segm = conv(part_feats, s, 'keypoint_final_layer')             # [B,25,H,W]
    heat = segm[:, 1:]                                             # drop background (:796)
    smpl_feats = branch(feats, 'smpl_deconv_layers')
    cam_shape = conv(smpl_feats, s, 'smpl_final_layer')            # [B,64,H,W]

    def attend(f):               # KeypointAttention.forward (keypoint_attention.py:34-56)
        J = heat.shape[1]
        a = F.softmax(heat.reshape(B, J, -1), dim=-1)
        return torch.matmul(a, f.reshape(B, f.shape[1], -1).transpose(2, 1)).transpose(2, 1)

    point_local = attend(smpl_feats)                               # [B,128,24]
    cam_shape = attend(cam_shape)                                  # [B,64,24]
    w = s['pose_mlp.weight'][0, :, :, :, 0, 0]                     # [6,128,24]  (locallyconnected2d.py:27-37)
    pose6 = torch.einsum('bcj,ocj->boj', point_local, w)           # [B,6,24]
    shape_feats = cam_shape.flatten(1)                             # [B,64*24] (c-major, then joint)
    pred_cam = linear(shape_feats, s, 'cam_mlp')
    pred_shape = linear(shape_feats, s, 'shape_mlp')
    pred_pose6d = pose6.transpose(2, 1)                            # [B,24,6]
    return {
        'pred_segm_mask': segm,
        'pred_pose': rot6d_to_rotmat(pred_pose6d).reshape(B, 24, 3, 3),
        'pred_pose6d': pred_pose6d,
        'pred_cam': pred_cam,
        'pred_shape': pred_shape,
        'uncert_feat': point_local.reshape(B, -1),
    }


def cliff_head(feats, bbox_info, sd, prefix='head.', n_iter=3):
    """cliff_head.forward (cliff_head.py:74-127): no activation between fc1 and fc2."""
    s = _SD(sd, prefix)
    B = feats.shape[0]
    if feats.dim() > 2:
        feats = feats.mean(dim=(2, 3))
    pose = s['init_pose'].expand(B, -1)
    shape = s['init_shape'].expand(B, -1)
    cam = s['init_cam'].expand(B, -1)
    for _ in range(n_iter):
        xc = torch.cat([feats, bbox_info, pose, shape, cam], 1)
        xc = linear(linear(xc, s, 'fc1'), s, 'fc2')
        pose = linear(xc, s, 'decpose') + pose
        shape = linear(xc, s, 'decshape') + shape
        cam = linear(xc, s, 'deccam') + cam
    return {
        'pred_pose': rot6d_to_rotmat(pose).view(B, 24, 3, 3),
        'pred_cam': cam,
        'pred_shape': shape,
        'pred_pose_6d': pose,
        'uncert_feat': feats,
        'body_feat2': xc,
    }


def poco_head(head_out, sd, uncert_inp_type, prefix='uncert_head.', act='sigmoid'):
    """poco_head.forward, inference branch (poco_head.py:96-154)."""
    s = _SD(sd, prefix)
    a = torch.sigmoid if act == 'sigmoid' else F.softplus
    x = head_out['uncert_feat']
    B = x.shape[0]
    if 'pose' in uncert_inp_type:
        pose = head_out['pred_pose'].reshape(B, -1)
        if 'pose-net' in uncert_inp_type:
            pf = torch.sigmoid(linear(pose, s, 'uncert_fc_poseNet'))
            ff = torch.sigmoid(linear(x, s, 'uncert_fc_featNet'))
            x = torch.cat([ff, pf], 1)
        else:
            x = torch.cat([x, pose], 1)
    i = 1
    while s.has(f'uncert_fc{i}.weight'):
        x = a(linear(x, s, f'uncert_fc{i}'))
        i += 1
    return {'var_pose': x, 'gt_pose_cond_idx': []}


# ------------------------------------------------------------------------------------------------
# conditional RealNVP  (real_nvp.py:25-65, nets nf_head.py:13-17)
# ------------------------------------------------------------------------------------------------
def _st_net(x, s, tanh):
    h = F.leaky_relu(linear(x, s, '0'), 0.01)
    h = F.leaky_relu(linear(h, s, '2'), 0.01)
    h = linear(h, s, '4')
    return torch.tanh(h) if tanh else h


def realnvp_backward(x, ctx, sd, prefix='flow_head.flow.'):
    s = _SD(sd, prefix)
    mask = s['mask']
    z = x
    logdet = x.new_zeros(x.shape[0])
    for i in reversed(range(mask.shape[0])):
        m = mask[i]
        z_ = m * z
        inp = torch.cat((z_, ctx), 1) if ctx is not None else z_
        sc = _st_net(inp, s.sub(f's.{i}'), True) * (1 - m)
        tr = _st_net(inp, s.sub(f't.{i}'), False) * (1 - m)
        z = (1 - m) * (z - tr) * torch.exp(-sc) + z_
        logdet = logdet - sc.sum(1)
    return z, logdet


def realnvp_log_prob(x, ctx, sd, prefix='flow_head.flow.'):
    z, logdet = realnvp_backward(x, ctx, sd, prefix)
    d = z.shape[1]
    return -0.5 * (z * z).sum(1) - 0.5 * d * math.log(2 * math.pi) + logdet


def realnvp_forward(z, ctx, sd, prefix='flow_head.flow.'):
    s = _SD(sd, prefix)
    mask = s['mask']
    x = z
    for i in range(mask.shape[0]):
        m = mask[i]
        x_ = x * m
        inp = torch.cat((x_, ctx), 1) if ctx is not None else x_
        sc = _st_net(inp, s.sub(f's.{i}'), True) * (1 - m)
        tr = _st_net(inp, s.sub(f't.{i}'), False) * (1 - m)
        x = x_ + (1 - m) * (x * torch.exp(sc) + tr)
    return x


def flow_context(uncert_feat, sd, prefix='flow_head.'):
    """cond_layer (nf_head.py:82); computed and discarded at inference by the reference."""
    return linear(uncert_feat, _SD(sd, prefix), 'cond_layer')


# ------------------------------------------------------------------------------------------------
# whole path  (poco.py:99-129, SMPL mesh stage excluded -- host side, parity unpinned)
# ------------------------------------------------------------------------------------------------
def poco_forward(batch, sd, backbone_name, head_name, uncert_inp_type):
    feats = backbone(batch['img'], sd, backbone_name)
    if 'cliff' in head_name:
        out = cliff_head(feats, batch['bbox_info'], sd)
    else:
        out = pare_head(feats, sd)
    out.update(poco_head(out, sd, uncert_inp_type))
    out['log_phi'] = None
    return out
