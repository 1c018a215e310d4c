#!/usr/bin/env python
"""bench.py -- crops/sec of the POCO per-crop inference hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--preset cliff_w32] [--batch 256]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      # the reference's CPU implementation of the same path

A step = one POCO forward (backbone + SMPL-parameter head + uncertainty head; SMPL mesh stage
excluded) over one batch of synthetic 224x224 crops per GPU.  `value` is timed on the
device with inputs resident in HBM (CUDA-graph replay of the op schedule); `e2e` goes through
POCO.forward with pinned HOST buffers (H2D of the crops and D2H of the packed results inside the
timed region).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic work per crop, 2*MAC of every conv/linear/matmul (probed on the reference)
GFLOP_PER_CROP = {'cliff_w32': 22.04, 'pare_w32': 30.9, 'cliff_w48cls': 34.58, 'pare_r50': 8.67}
WORKLOAD = {
    'cliff_w32': 'POCO-CLIFF HRNet-W32 forward, 224x224 synthetic crops, fp16 (BASELINE configs[1])',
    'pare_w32': 'POCO-PARE HRNet-W32 + part-attention head forward (BASELINE configs[2])',
    'cliff_w48cls': 'POCO-CLIFF HRNet-W48-cls forward (shipped demo_poco_cliff.yaml)',
    'pare_r50': 'POCO-PARE ResNet-50 forward (BASELINE configs[0])',
}


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'tflops': d.get('bf16_tflops_sustained', d.get('bf16_tflops')), 'tflops_burst': d.get('bf16_tflops'),
                'hbm_gbs': d.get('hbm_gbs'), 'src': 'measured (MEASURED_PEAKS.json, sustained cuBLAS bf16)'}
    return {'tflops': 1400.0, 'tflops_burst': 1590.0, 'hbm_gbs': 6650.0, 'src': 'fallback (B200_PROFILING.md)'}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        if not sm:
            return None
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == 'Active' for r in self.rows)]
        smax = max(int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit())
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': smax, 'reasons': reasons, 'samples': len(sm)}


def build_inputs(preset, B, device):
    from synth import ckpt as S
    return S.synthetic_batch(B, 1, device)


def load_model_and_sd(preset):
    import numpy as np
    from synth import ckpt as S
    from poco_b200 import POCO
    gd = os.path.join(ROOT, 'tests', 'golden')
    meta = json.load(open(os.path.join(gd, f'spec_{preset}.json')))
    sd = S.synth_state_dict(S.template_from_spec(meta, 0), 0, np.load(os.path.join(gd, f'calib_{preset}.npz')))
    model = POCO(**meta['kwargs'], smpl_mean_params=S.smpl_mean_params(0))
    model.load_state_dict(sd)
    return model, sd, meta


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (oracle port when the reference tree is absent)
# ------------------------------------------------------------------------------------------------
def cpu_forward_fn(preset):
    """returns (fn(batch) -> out, kind).  kind 'reference' = unmodified pocolib via oracle/ref_loader,
    'port' = oracle/poco_oracle.py (the pinned CPU restatement)."""
    import numpy as np
    import torch
    from oracle import poco_oracle as O
    from oracle import ref_loader as R
    from synth import ckpt as S
    gd = os.path.join(ROOT, 'tests', 'golden')
    meta = json.load(open(os.path.join(gd, f'spec_{preset}.json')))
    sd = S.synth_state_dict(S.template_from_spec(meta, 0), 0, np.load(os.path.join(gd, f'calib_{preset}.npz')))
    if R.find_reference_root() is not None:
        try:
            model, _ = R.build_reference(preset)
            model.load_state_dict(sd)
            model.eval()
            return (lambda b: model(b)), 'reference'
        except Exception as e:      # noqa: BLE001
            sys.stderr.write(f'[bench] reference import failed ({e}); using the oracle port\n')
    bb, head = meta['kwargs']['backbone'].split('-')
    uit = meta['kwargs']['uncert_inp_type']
    return (lambda b: O.poco_forward(b, sd, bb, head, uit)), 'port'


def time_cpu(preset, sample_b, steps, warmup):
    import torch
    torch.set_num_threads(os.cpu_count())
    fn, kind = cpu_forward_fn(preset)
    batch = build_inputs(preset, sample_b, 'cpu')
    with torch.no_grad():
        for _ in range(warmup):
            fn(batch)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn(batch)
        dt = time.perf_counter() - t0
    return sample_b * steps / dt, dt / steps, kind, torch.get_num_threads()


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sample_b = args.cpu_sample
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    cps, spstep, kind, cores = time_cpu(args.preset, sample_b, steps, warmup)
    line = {
        'impl': 'reference', 'metric': 'crops/sec', 'value': round(cps, 3), 'unit': 'crops/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup, 'ms_per_step': round(spstep * 1e3, 3),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD[args.preset], 'preset': args.preset, 'sample_batch': sample_b},
        'cpu_baseline': {'value': round(cps, 3), 'unit': 'crops/s', 'cores': cores, 'kind': kind,
                         'sample': f'{steps} forwards of {sample_b} crops (same synthetic workload, fp32, torch CPU)'},
        'e2e': {'value': round(cps, 3), 'unit': 'crops/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def op_label(op):
    from poco_b200 import _lib as L
    if op.kind == L.OP_CONV:
        c = op.u.conv
        return f'conv {c.in_.C}->{c.out.C} k{c.kh} s{c.stride} {c.in_.H}->{c.out.H}' + (' +res' if c.residual else '')
    if op.kind == L.OP_CONV_CHAIN:
        c = op.u.conv_chain.seg[0]
        return f'chain x{op.u.conv_chain.n_seg} conv {c.in_.C}->{c.out.C} k{c.kh} {c.in_.H}'
    if op.kind == L.OP_LINEAR:
        return f'linear {op.u.linear.M}x{op.u.linear.I}->{op.u.linear.O}'
    if op.kind == L.OP_FUSE_SUM:
        return f'fuse_sum x{op.u.fuse_sum.n_in} c{op.u.fuse_sum.out.C} h{op.u.fuse_sum.out.H}'
    if op.kind == L.OP_UPSAMPLE2X:
        return f'upsample2x c{op.u.upsample2x.in_.C} h{op.u.upsample2x.in_.H}'
    return L._FIELD_OF_KIND[op.kind]


def per_kernel_pass(eng, torch, reps=2, dump=None):
    """eager op-by-op replay with CUDA events: time share and algorithmic FLOPs per kernel class"""
    from poco_b200 import _lib as L
    s = torch.cuda.current_stream().cuda_stream
    ops = eng.plan.ops
    classes = {}
    for _ in range(reps):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(ops) + 1)]
        evs[0].record()
        for i, op in enumerate(ops):
            if op.kind not in (L.OP_FORK, L.OP_JOIN):       # (op-by-op pass is sequential)
                L.run_op(op, s)
            evs[i + 1].record()
        torch.cuda.synchronize()
        classes = {}
        rows = []
        for i, op in enumerate(ops):
            if op.kind in (L.OP_FORK, L.OP_JOIN):
                continue
            ms = evs[i].elapsed_time(evs[i + 1])
            rows.append((i, op_label(op), ms))
            by = 0.0
            if op.kind in (L.OP_CONV, L.OP_CONV_CHAIN):
                segs = [op.u.conv] if op.kind == L.OP_CONV else [op.u.conv_chain.seg[k] for k in range(op.u.conv_chain.n_seg)]
                c = segs[0]
                lin = c.stride == 1 and c.in_.H == c.out.H and ((c.kh == 3 and c.pad == 1) or (c.kh == 1 and c.pad == 0))
                name = 'conv_tc_linear' if lin else 'conv_tc_gather'
                fl = 0.0
                for c in segs:      # a chain launch does the work of all its segments
                    by += 2.0 * c.out.N * (c.in_.H * c.in_.W * c.in_.C + c.out.H * c.out.W * c.out.C * (2 if c.residual else 1))
                    fl += 2.0 * c.out.N * c.out.H * c.out.W * c.out.C * c.in_.C * c.kh * c.kw
            else:
                name, fl = L._FIELD_OF_KIND[op.kind], 0.0
            e = classes.setdefault(name, {'ms': 0.0, 'flops': 0.0, 'launches': 0, 'bytes': 0.0})
            e['ms'] += ms
            e['flops'] += fl
            e['bytes'] += by
            e['launches'] += 1
    if dump:
        agg = {}
        for i, lab, ms in rows:
            a = agg.setdefault(lab, [0, 0.0])
            a[0] += 1
            a[1] += ms
        with open(dump, 'w') as f:
            f.write('label,count,total_ms,avg_us\n')
            for lab, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f'{lab},{n},{ms:.4f},{ms / n * 1e3:.2f}\n')
    return classes


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from poco_b200 import dist as pdist
    from poco_b200 import kernel_launches
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B, K, W = args.batch, args.steps, max(3, args.warmup)
    preset = args.preset
    peaks = load_peaks()

    model, sd, meta = load_model_and_sd(preset)
    if args.with_smpl:      # mesh stage on the device inside the e2e leg (seeded stand-in for the licensed model)
        from synth import smpl_model
        from poco_b200.smpl import DeviceSmplStage
        model.smpl = DeviceSmplStage(model.head_name, smpl_model.synthetic_model(0))
    model = model.to(dev).eval()
    batch = build_inputs(preset, B, dev)
    clocks = ClockSampler(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`)
    with torch.no_grad():
        out = model.hot_path(batch)              # eager first call (also counts launches per forward)
        n0 = kernel_launches()
        model.use_cuda_graph = False
        out = model.hot_path(batch)
        launches_per_fwd = kernel_launches() - n0
        model.use_cuda_graph = True
        eng = model._engine(B, dev)
        for _ in range(W):
            eng.run(True)
            if world > 1:
                pdist.all_gather_outputs({k: eng.out[k] for k in ('pred_pose', 'pred_shape', 'pred_cam', 'var_pose')})
        barrier()
        if rank == 0:
            clocks.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            eng.run(True)
            if world > 1:       # the one collective of the path: all-gather of the packed per-crop records
                pdist.all_gather_outputs({k: eng.out[k] for k in ('pred_pose', 'pred_shape', 'pred_cam', 'var_pose')})
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * K / (ms * 1e-3)

    # ---------------- end to end through POCO.forward with host buffers
    e2e_steps = max(2, min(K, 10))
    host = {k: v.cpu().pin_memory() for k, v in batch.items()}
    h2d = sum(host[k].numel() * host[k].element_size() for k in host)
    rec_host = torch.empty(B, pdist.RECORD_WIDTH, dtype=torch.float32).pin_memory()
    d2h = rec_host.numel() * 4
    verts_host = torch.empty(B, 6890, 3, dtype=torch.float32).pin_memory() if args.with_smpl else None
    d2h += verts_host.numel() * 4 if args.with_smpl else 0

    # The caller-side pipeline a serving loop uses (the reference's DataLoader does the same with
    # pin_memory + non_blocking, tester.py:394-405): a copy stream uploads step i+1's crops from pinned host
    # memory into the other of two device buffers while POCO.forward runs step i; every step's H2D and the
    # D2H of its packed results are inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    dev_bufs = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])
            for k, v in host.items():
                dev_bufs[i % 2][k].copy_(v, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def run_steps(n):
        for e in consumed:
            e.record()
        upload(0)
        for i in range(n):
            if i + 1 < n:
                upload(i + 1)
            torch.cuda.current_stream().wait_event(ready[i % 2])
            o = model(dev_bufs[i % 2])
            consumed[i % 2].record()
            rec_host.copy_(pdist.pack_record(o), non_blocking=True)
            if verts_host is not None:
                verts_host.copy_(o['smpl_vertices'], non_blocking=True)
        torch.cuda.synchronize()

    with torch.no_grad():
        run_steps(2)
        barrier()
        t0 = time.perf_counter()
        run_steps(e2e_steps)
        wall = time.perf_counter() - t0
    tw = torch.tensor([wall], device=dev)
    if world > 1:
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(tw.item())
    clk = clocks.stop() if rank == 0 else None

    # ---------------- roofline of the dominant kernel (rank 0, eager op-by-op timing with CUDA events)
    roofline = cpu_base = None
    shares = {}
    if rank == 0:
        with torch.no_grad():
            # per-kernel timing wants every launch alone on the full GPU: rebuild the schedule without
            # concurrent lanes (whose convs are capped to a share of the SMs) for this pass only
            os.environ['POCO_B200_LANES'] = '0'
            eng_seq = model._build_engine(B, dev)
            os.environ.pop('POCO_B200_LANES')
            eng_seq.img.copy_(batch['img'])
            if eng_seq.bbox is not None:
                eng_seq.bbox.copy_(batch['bbox_info'])
            eng_seq.plan.run()
            classes = per_kernel_pass(eng_seq, torch, dump=args.dump_ops)
            del eng_seq
        tot = sum(c['ms'] for c in classes.values())
        dom = max(classes, key=lambda k: classes[k]['ms'])
        d = classes[dom]
        ach = d['flops'] / (d['ms'] * 1e-3) / 1e12 if d['flops'] > 0 else 0.0
        traffic = None
        tf = os.path.join(ROOT, 'profiles', 'r01_traffic.json')
        if os.path.exists(tf) and B == 256 and preset == 'cliff_w32':
            tj = json.load(open(tf))
            traffic = tj['dram_bytes_per_launch'] if tj.get('kernel') == dom else None
        roofline = {'bound': 'tensor', 'kernel': dom, 'achieved': round(ach, 2), 'peak': peaks['tflops'],
                    'unit': 'TFLOP/s', 'frac': round(ach / peaks['tflops'], 4), 'traffic': traffic,
                    'traffic_unit': 'DRAM bytes per launch (ncu, profiles/r01_traffic.json)',
                    'algorithmic_bytes_per_launch': round(d['bytes'] / d['launches']),
                    'algorithmic_flops_per_launch': round(d['flops'] / d['launches']),
                    'launches_per_step': d['launches'], 'avg_launch_ms': round(d['ms'] / d['launches'], 4),
                    'peak_source': peaks['src'], 'frac_of_burst_peak': round(ach / peaks['tflops_burst'], 4)}
        shares = {k: round(c['ms'] / tot, 4) for k, c in sorted(classes.items(), key=lambda kv: -kv[1]['ms'])}
        if world == 1 and not args.no_cpu_baseline:
            cps, spstep, kind, cores = time_cpu(preset, args.cpu_sample, 3, 1)
            cpu_base = {'value': round(cps, 3), 'unit': 'crops/s', 'cores': cores, 'kind': kind,
                        'sample': f'3 forwards of {args.cpu_sample} crops of the same workload (fp32 torch CPU)'}
    if rank == 0:
        gf = GFLOP_PER_CROP[preset]
        line = {
            'metric': 'crops/sec', 'value': round(value, 2), 'unit': 'crops/s', 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': round(ms / K, 4), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'fp16', 'data': 'synthetic',
            'config': {'workload': WORKLOAD[preset], 'preset': preset, 'crops_per_gpu': B, 'global_batch': B * world,
                       'parallelism': f'dp{world}',
                       'smpl_mesh_stage': ('in the e2e leg: poco_smpl_run on a seeded synthetic SMPL model, vertices read back '
                                           '(SURVEY 8 f4); not in `value`') if args.with_smpl else 'excluded (SURVEY 8d)',
                       'l2': 'inputs (154 MB f32 at B=256) and activations exceed the 126 MB L2 every step',
                       'cuda_graph': True, 'weights': 'calibrated synthetic checkpoint seed 0'},
            'tensor_peak_frac_end_to_end': round(value / world * gf / 1e3 / peaks['tflops_burst'], 4),
            'achieved_tflops_per_gpu': round(value / world * gf / 1e3, 2),
            'e2e': {'value': round(e2e_value, 2), 'unit': 'crops/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'steps': e2e_steps, 'api': 'poco_b200.POCO.forward(batch); crops uploaded from pinned host buffers on a copy stream (double buffered), packed results read back every step'},
            'gpu_launches': launches_per_fwd * K,
            'launches_per_forward': launches_per_fwd,
            'roofline': roofline, 'kernel_time_share': shares, 'cpu_baseline': cpu_base, 'clocks': clk,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='poco_b200', choices=['poco_b200', 'reference'])
    ap.add_argument('--preset', default='cliff_w32', choices=sorted(GFLOP_PER_CROP))
    ap.add_argument('--batch', type=int, default=256, help='crops per GPU (weak scaling)')
    ap.add_argument('--cpu-sample', type=int, default=16, help='crops per CPU-arm forward (bounded sample)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--with-smpl', action='store_true', help='run the device SMPL mesh stage (f4) inside the e2e leg')
    ap.add_argument('--dump-ops', default=None, help='write the per-op timing table (eager, CUDA events) to this CSV')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == '__main__':
    main()
