#!/usr/bin/env python
"""Benchmark of the elastic-MBConv SR hot path on B200 (contract: see the task brief / DESIGN.md §6).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA library)
    python bench.py --impl reference [...]                         # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[1]): OFAMobileNetS4 (2 shuffle stages), max sub-network
set_active_subnet(ks=7, e=6, d=4, pixel_d=2), 4x SR of one synthetic LR frame 960x540 -> 3840x2160
per step per GPU (frames are independent units: no data-path collective; "weak" scaling).
Metric: SR output Mpix/s, whole job (all ranks).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, 'ofa-for-super-resolution_b200'))
ORACLE_DIR = os.path.join(ROOT, 'oracle')     # imported ONLY by the reference / cpu_baseline legs (cpu_forward_sample)

import numpy as np
import torch

METRIC = 'SR output Mpix/s (4x, max subnet ks=7 e=6 d=4)'
UNIT = 'Mpix/s'
NET_CFG = dict(ks_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], depth_list=[2, 3, 4], pixelshuffle_depth_list=[1, 2])
SUBNET = dict(ks=7, e=6, d=4, pixel_d=2)
WEIGHT_SEED = 1234


def workload_text(W, H):
    return ('OFAMobileNetS4 max subnet (ks=7,e=6,d=4,pixel_d=2 -> 14 MBConv blocks), 4x SR forward, LR %dx%d -> %dx%d, '
            '1 frame per step per GPU, frames sharded across ranks (no collective)' % (W, H, 4 * W, 4 * H))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--lr-h', type=int, default=540)
    ap.add_argument('--lr-w', type=int, default=960)
    ap.add_argument('--cpu-tile', type=int, default=256, help='LR tile edge of the bounded CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-train', action='store_true', help='skip the training-step measurements (the `train` object)')
    ap.add_argument('--no-extras', action='store_true', help='skip the secondary configurations (the `extra` object)')
    ap.add_argument('--eager', action='store_true', help='launch the forward eagerly instead of replaying its CUDA graph')
    ap.add_argument('--train-steps', type=int, default=20)
    ap.add_argument('--dtype', default='f16', choices=['f16', 'bf16', 'fp32'],
                    help='activation storage of our arm: f16 (default; meets the 0.01 dB PSNR criterion), bf16 (same speed) or fp32 (exact CUDA-core path)')
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {'hbm': d['hbm_gbs'], 'tensor_burst': d['bf16_tflops'], 'tensor_sustained': d['bf16_tflops_sustained'],
                'source': 'MEASURED_PEAKS.json'}
    return {'hbm': 6650.0, 'tensor_burst': 1590.0, 'tensor_sustained': 1400.0, 'source': 'fallback (B200_PROFILING.md)'}


# ncu --set full captures of the hot kernels (profiles/*.csv, written by tools/ncu_summary.py): DRAM bytes per launch
NCU_PROFILES = {'dw7x7 C384 planar': 'r2_ncu_dw7_planar_dual_issuer.csv', 'mbconv expand 64->384 planar': 'r1_ncu_expand_planar_v1.csv',
                'mbconv project 384->64 planar': 'r1_ncu_project_planar_v1.csv', 'conv5x5 64->3 rows-tc': 'r1_ncu_conv_out_rows_v1.csv'}


def ncu_traffic(tag):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `tag` from the committed ncu summary, or None."""
    import csv
    name = NCU_PROFILES.get(tag)
    path = os.path.join(ROOT, 'profiles', name) if name else None
    if not path or not os.path.exists(path):
        return None
    rows = list(csv.reader(open(path)))
    hdr, units, last = rows[0], rows[1], rows[-1]
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    tot = 0.0
    for key in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        i = hdr.index(key)
        tot += float(last[i].replace(',', '')) * scale.get(units[i], 1.0)
    return tot


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region: through NVML (nvidia_ml_py, ~10 ms period, so a
    100 ms timed region yields ~10 samples) when it is importable, else by spawning nvidia-smi (~0.2 s period)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = torch.cuda.get_device_properties(index).uuid           # CUDA order may differ from NVML order
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + str(uuid)).encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        bits = [n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown,
                n.nvmlClocksEventReasonSwThermalSlowdown, n.nvmlClocksEventReasonSwPowerCap]
        self.rows.append([str(mhz), str(self.max_mhz)] + ['Active' if mask & b else 'Not Active' for b in bits])

    def run(self):
        while not self._halt.is_set():
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits'], capture_output=True, text=True,
                                         timeout=5).stdout
                    self.rows.append([c.strip() for c in out.strip().split(',')])
            except Exception:
                pass
            self._halt.wait(0.01 if self.nvml is not None else 0.2)

    def finish(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace('.', '').isdigit()]
        reasons = sorted({self.NAMES[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == 'active'})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm), 'source': 'nvml' if self.nvml is not None else 'nvidia-smi'}


# =====================================================================================================
# reference arm / cpu baseline: the oracle port on the host cores
# =====================================================================================================
def cpu_forward_sample(tile, iters, warm):
    if ORACLE_DIR not in sys.path:
        sys.path.insert(0, ORACLE_DIR)
    import ofa_sr_oracle as O
    torch.set_num_threads(os.cpu_count())
    spec = O.SuperNetSpec('s4', NET_CFG['ks_list'], NET_CFG['expand_ratio_list'], NET_CFG['depth_list'],
                          NET_CFG['pixelshuffle_depth_list'])
    sd = O.synth_state_dict(spec.param_shapes(), WEIGHT_SEED)
    spec.set_active_subnet(**SUBNET)
    x = torch.from_numpy(np.random.RandomState(0).rand(1, 3, tile, tile).astype(np.float32))
    times = []
    with torch.no_grad():
        for i in range(warm + iters):
            t0 = time.perf_counter()
            y = O.supernet_forward(x, sd, spec)
            t1 = time.perf_counter()
            if i >= warm:
                times.append(t1 - t0)
    out_pix = y.shape[0] * y.shape[2] * y.shape[3]
    return out_pix, times


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    out_pix, times = cpu_forward_sample(args.cpu_tile, args.steps, args.warmup)
    ms = 1e3 * float(np.mean(times))
    val = out_pix / 1e6 / (ms / 1e3)
    sample = ('oracle port (torch CPU fp32, same ops the reference calls) of the S4 max-subnet forward on one '
              '%dx%d LR tile -> %dx%d per step; Mpix/s is size-normalised' % (args.cpu_tile, args.cpu_tile,
                                                                               4 * args.cpu_tile, 4 * args.cpu_tile))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_text(args.lr_w, args.lr_h),
                   'reference_sample': 'each step = one %dx%d LR tile of that frame (Mpix/s is size-normalised)' % (args.cpu_tile, args.cpu_tile)},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


# =====================================================================================================
# our arm
# =====================================================================================================
def synth_weights(net, seed):
    """Synthetic parameters of the SURVEY §8d recipe written straight into the product net (no checkpoint exists
    offline): he_fout conv weights, BN gamma ~ U(.5,1.5), beta ~ N(0,.1), running mean ~ N(0,.1), running var ~
    U(.5,1.5), kernel-transform matrices eye + 0.05 N(0,1) — so BN folding and the 7->5->3 transform are not vacuous."""
    rs = np.random.RandomState(seed)
    sd = {}
    for key, t in net.state_dict().items():
        shape = tuple(t.shape)
        if key.endswith('num_batches_tracked'):
            sd[key] = torch.zeros((), dtype=torch.int64)
        elif key.endswith('_matrix'):
            sd[key] = torch.from_numpy((np.eye(shape[0]) + 0.05 * rs.randn(*shape)).astype(np.float32))
        elif key.endswith('conv.weight'):
            fan = shape[0] * shape[2] * shape[3]
            sd[key] = torch.from_numpy((rs.randn(*shape) * np.sqrt(2.0 / fan)).astype(np.float32))
        elif key.endswith('running_var') or key.endswith('bn.weight'):
            sd[key] = torch.from_numpy(rs.uniform(0.5, 1.5, size=shape).astype(np.float32))
        else:                                            # running_mean, bn.bias
            sd[key] = torch.from_numpy((0.1 * rs.randn(*shape)).astype(np.float32))
    net.load_state_dict(sd)


def build_net(dev):
    from ofa_b200.elastic_nn.modules.dynamic_op import DynamicSeparableConv2d
    DynamicSeparableConv2d.KERNEL_TRANSFORM_MODE = 1
    from ofa_b200.elastic_nn.networks import OFAMobileNetS4
    net = OFAMobileNetS4(**{k: list(v) for k, v in NET_CFG.items()})
    synth_weights(net, WEIGHT_SEED)
    net.set_active_subnet(**SUBNET)
    return net.to(dev).eval()


def pin_rank_to_cores(local_rank, local_world):
    """Several ranks share the box's host cores: keep each rank on the cores NVML reports as local to its GPU (same NUMA
    node: the pinned frame buffers are first-touched there, and the device -> host copies do not cross the socket
    link).  No-op when it cannot be done."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
        groups = None
        try:
            import pynvml
            pynvml.nvmlInit()
            words = (max(allowed) // 64) + 1
            masks = []
            for r in range(local_world):
                h = pynvml.nvmlDeviceGetHandleByIndex(r)
                m = pynvml.nvmlDeviceGetCpuAffinity(h, words)
                cores = [64 * w + b for w in range(words) for b in range(64) if (int(m[w]) >> b) & 1]
                masks.append(tuple(c for c in cores if c in allowed))
            if all(masks):
                groups = masks
        except Exception:
            groups = None
        if groups is None:
            groups = [tuple(allowed)] * local_world
        mine = groups[local_rank]
        # the whole NUMA-local set, NOT an exclusive slice of it: a rank runs its Python loop, autograd's device thread,
        # the NCCL proxy and the driver's helper threads; 16 cores / 8 ranks = 2-core slices starved the launch-heavy C4
        # step (54.7 ms at 4 ranks against 40.0 ms unpinned); the node's scheduler balances the bursty threads better
        if os.environ.get('OFA_BENCH_PIN_SLICES', '0') == '1':
            peers = [r for r in range(local_world) if groups[r] == mine]
            per = max(1, len(mine) // len(peers))
            k = peers.index(local_rank)
            mine = tuple(mine[k * per:(k + 1) * per]) or mine
        os.sched_setaffinity(0, list(mine))
    except Exception:
        pass


def run_train(dev, dist, rank, world, which, steps, warmup):
    """BASELINE.json configs[2] (C3) / configs[3] (C4) as one progressive-shrinking optimizer step per timed step, on
    every rank count the driver runs: batch 64 of 96x96 HR patches PER GPU (weak scaling), bf16 compute with fp32 master
    weights, `dynamic_batch_size` 2 (two sampled sub-networks accumulate gradients, progressive_shrinking.py:158-203;
    seed rule :164), fused Adam (sr_run_manager.py:115-133) and -- with several ranks -- ONE flat NCCL all-reduce of the
    gradients per step, its tail segment overlapped with backward (distributed_run_manager.py:72-75).
      c3: OFAMobileNetS4, inputs = the 2x / 4x bicubic-downscaled patches the sampled pixel_d selects
      c4: OFAMobileNetX4 (HR in, HR out), one 2x and one 4x student per step, frozen max-sub-network teacher at both
          scales under no_grad, loss = MSE(out, HR) + 0.5 MSE(out, teacher_out)"""
    import copy
    import random
    import ofa_b200
    from ofa_b200 import backend as B, optim, parallel as P
    from ofa_b200.elastic_nn.modules.dynamic_op import DynamicSeparableConv2d
    from ofa_b200.elastic_nn.networks import OFAMobileNetS4, OFAMobileNetX4
    from ofa_b200.elastic_nn.training import train_step, subnet_seed
    ofa_b200.set_train_dtype(torch.bfloat16)
    DynamicSeparableConv2d.KERNEL_TRANSFORM_MODE = 1
    Net = OFAMobileNetS4 if which == 'c3' else OFAMobileNetX4
    cfg = {k: list(v) for k, v in NET_CFG.items()}
    if which == 'c3':
        # S4 runs runtime_depth[0] PixelShuffle stages whatever pixel_d says (always 4x), so the
        # reference's S4 training is the 4x task: pixel_d = 2 -> the 4x-downscaled input (progressive_shrinking.py:178-180)
        cfg['pixelshuffle_depth_list'] = [2]
    net = Net(**cfg)
    synth_weights(net, WEIGHT_SEED + 1)
    net = net.to(dev).train()
    per_gpu = 64
    g = torch.Generator(device='cpu').manual_seed(100 + rank)
    hr = torch.rand(per_gpu, 3, 96, 96, generator=g).to(dev)
    batch = {'image': hr,
             '2x_down_image': torch.nn.functional.interpolate(hr, scale_factor=0.5, mode='bilinear', align_corners=False),
             '4x_down_image': torch.nn.functional.interpolate(hr, scale_factor=0.25, mode='bilinear', align_corners=False)}
    teacher = copy.deepcopy(net).eval() if which == 'c4' else None
    decay, no_decay = optim.split_no_decay(net.named_parameters())
    opt = optim.FusedAdam(decay, no_decay, lr=1e-4, weight_decay=3e-5)
    reducer, boundary = None, None
    if world > 1:
        P.broadcast_parameters(net, src=0)
        if which == 'c3':
            tail, boundary = P.s4_tail_parameters(net)
            reducer = P.FlatGradAllReduce(net.parameters(), n_buckets=2, tail_params=tail)
        else:
            reducer = P.FlatGradAllReduce(net.parameters(), n_buckets=2)
    n_batch = 1000
    ar_events = []

    def step(i, time_allreduce=False):
        lr = optim.cosine_lr(1e-4, 120, 0, i, n_batch)
        if which == 'c3':
            if time_allreduce and reducer is not None:
                inner = reducer.reduce

                def timed_reduce():
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    inner()
                    e1.record()
                    ar_events.append((e0, e1))
                reducer.reduce = timed_reduce
            try:
                train_step(net, opt, batch, 0, i, n_batch, dynamic_batch_size=2, reducer=reducer, lr=lr,
                           boundary_module=boundary)
            finally:
                if time_allreduce and reducer is not None:
                    reducer.reduce = inner
            return
        # c4: teacher at both scales, then one 2x and one 4x student (tools/bench_train_x4.py)
        opt.set_lr(lr)
        opt.zero_grad(set_to_none=True)
        soft = {}
        with torch.no_grad():
            for pd in (1, 2):
                teacher.set_active_subnet(ks=7, e=6, d=4, pixel_d=pd)
                soft[pd] = teacher(hr)
        for k, pd in enumerate((1, 2)):
            random.seed(subnet_seed(0, n_batch, i, k))
            net.sample_active_subnet()
            net.set_active_subnet(pixel_d=pd)
            out = net(hr)
            loss = torch.nn.functional.mse_loss(out, hr) + 0.5 * torch.nn.functional.mse_loss(out, soft[pd])
            loss.backward()
        if reducer is not None:
            if time_allreduce:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                reducer.reduce()
                e1.record()
                ar_events.append((e0, e1))
            else:
                reducer.reduce()
        opt.step()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        step(i)
    barrier()
    B.launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(warmup + i)
    e1.record()
    barrier()
    launches = B.launch_count() // steps
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / steps
    ar_ms = None
    if reducer is not None:
        for i in range(4):
            step(warmup + steps + i, time_allreduce=True)
        torch.cuda.synchronize()
        ar_ms = float(np.mean([a.elapsed_time(b) for a, b in ar_events]))
    barrier()
    del net, teacher, opt, reducer
    torch.cuda.empty_cache()
    return {'config': ('C3: OFAMobileNetS4 progressive-shrinking step' if which == 'c3' else
                       'C4: OFAMobileNetX4 joint 2x/4x step with teacher distillation (kd 0.5)') +
                      ', batch 64 x 96x96 HR patches per GPU, dynamic_batch_size 2, bf16 + fp32 masters, fused Adam' +
                      ((', flat NCCL all-reduce' + (' (tail overlapped with backward)' if which == 'c3' else '')) if world > 1 else ''),
            'ms_per_step': ms_per_step, 'patches_per_s': world * per_gpu / (ms_per_step / 1e3),
            'hr_mpix_per_s': world * per_gpu * 96 * 96 / 1e6 / (ms_per_step / 1e3),
            'library_launches_per_step': int(launches), 'allreduce_ms': ar_ms, 'graphed': False,
            'steps': steps, 'warmup': warmup, 'scaling': 'weak', 'timing': 'CUDA events around the timed steps, max over ranks'}


def run_extras(args, dev, dist, rank, world, net, x_host, flush):
    """Secondary measurements, each a handful of forwards (max over ranks where ranks take part):
      serial_latency : one frame, pinned host -> device -> net -> fp32 image -> pinned host, nothing overlapped
      e2e_uint8      : the pipelined e2e loop with the uint8 image output (set_output_dtype), 4x less D2H
      lr1080_to_8k   : LR 1920x1080 -> 7680x4320 (second frame size)
      c1             : S4 smallest sub-network (ks 3, e 3, d 2, one PixelShuffle stage), 1x3x256x256 -> 512x512, eager and
                       as a CUDA graph (BASELINE.json configs[0])
      tile_sharded   : (N > 1) ONE 960x540 frame cut into N tiles with a 64-px LR halo, one tile per rank, no collective
                       (parallel.tiled_forward): single-frame latency, strong scaling"""
    import ofa_b200
    from ofa_b200 import parallel as P
    from ofa_b200.elastic_nn.networks import OFAMobileNetS4
    out = {}
    H, W = args.lr_h, args.lr_w

    def maxr(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def ev_ms(fn, reps=5, warm=2):
        with torch.no_grad():
            for _ in range(warm):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    # serial per-frame latency (no overlap between copies and compute)
    y_host = torch.empty(1, 3, 4 * H, 4 * W, dtype=torch.float32).pin_memory()
    x_d = torch.empty(1, 3, H, W, device=dev)

    def serial():
        x_d.copy_(x_host, non_blocking=True)
        y_host.copy_(net(x_d), non_blocking=True)
    out['serial_latency'] = {'ms_per_frame': maxr(ev_ms(serial)), 'what': 'H2D (pinned) + forward + D2H of the fp32 image, one stream'}

    # uint8 image output: same forward, the last epilogue writes tensor2img_np's uint8
    net.set_output_dtype(torch.uint8)
    y8_host = [torch.empty(1, 3, 4 * H, 4 * W, dtype=torch.uint8).pin_memory() for _ in range(2)]
    cs = torch.cuda.Stream(device=dev)
    st = {'i': 0}

    def e2e_u8():
        i = st['i'] & 1
        st['i'] += 1
        x_d.copy_(x_host, non_blocking=True)
        y = net(x_d)
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(cs):
            cs.wait_event(ready)
            y8_host[i].copy_(y, non_blocking=True)
            y.record_stream(cs)
    with torch.no_grad():
        for _ in range(3):
            e2e_u8()
        torch.cuda.current_stream().wait_stream(cs)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            e2e_u8()
        torch.cuda.current_stream().wait_stream(cs)
        e1.record()
        torch.cuda.synchronize()
    ms = maxr(e0.elapsed_time(e1)) / args.steps
    out['e2e_uint8'] = {'value': world * 16 * H * W / 1e6 / (ms / 1e3), 'unit': UNIT, 'ms_per_step': ms,
                        'h2d_bytes_per_step': 12 * H * W, 'd2h_bytes_per_step': 3 * 16 * H * W,
                        'what': 'as e2e, uint8 image output (tensor2img_np in the last epilogue); L2 not flushed between steps'}
    net.set_output_dtype(torch.float32)

    # LR 1920x1080 -> 8K
    if (H, W) == (540, 960):
        xb = torch.rand(1, 3, 1080, 1920, device=dev)
        ms = maxr(ev_ms(lambda: net(xb), reps=3, warm=1))
        out['lr1080_to_8k'] = {'ms_per_step': ms, 'value': world * 16 * 1080 * 1920 / 1e6 / (ms / 1e3), 'unit': UNIT}
        del xb

    # C1: smallest sub-network, 2x, 256x256
    if rank == 0:
        c1 = OFAMobileNetS4(ks_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], depth_list=[2, 3, 4], pixelshuffle_depth_list=[1])
        synth_weights(c1, WEIGHT_SEED + 2)
        c1.set_active_subnet(ks=3, e=3, d=2, pixel_d=1)
        c1 = c1.to(dev).eval()
        x1 = torch.rand(1, 3, 256, 256, device=dev)
        eager = ev_ms(lambda: c1(x1), reps=20, warm=5)
        with torch.no_grad():
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                c1(x1)
            torch.cuda.current_stream().wait_stream(s)
            with torch.cuda.graph(g):
                y1 = c1(x1)
        graph = ev_ms(g.replay, reps=20, warm=5)
        out['c1'] = {'workload': 'S4 smallest sub-network (ks 3, e 3, d 2, 1 PixelShuffle stage), 1x3x256x256 -> %dx%d' % (y1.shape[2], y1.shape[3]),
                     'eager_ms': eager, 'graph_ms': graph, 'value': y1.shape[2] * y1.shape[3] / 1e6 / (graph / 1e3), 'unit': UNIT}
        del c1, g

    # one frame, tile-sharded over the ranks (strong scaling of the latency)
    if world > 1:
        ty = 2 if world >= 4 else 1
        tx = world // ty
        xf = x_host.to(dev)

        def tiled():
            P.tiled_forward(net, xf, ty, tx, halo=P.S4_HALO_LR, scale=4, rank=rank, world=world)
        ms = maxr(ev_ms(tiled))
        out['tile_sharded'] = {'tiles': [ty, tx], 'halo_lr_px': P.S4_HALO_LR, 'latency_ms': ms,
                               'value': 16 * H * W / 1e6 / (ms / 1e3), 'unit': UNIT, 'scaling': 'strong'}
    return out


def run_ours(args):
    import ofa_b200
    from ofa_b200 import backend as B, functional as OF
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    assert torch.cuda.is_available(), 'bench.py (our arm) needs a CUDA device: there is no CPU path'
    if world > 1:
        pin_rank_to_cores(local_rank, int(os.environ.get('LOCAL_WORLD_SIZE', world)))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    B.lib()  # fail loudly if the extension is missing
    ofa_b200.set_compute_dtype({'f16': torch.float16, 'bf16': torch.bfloat16, 'fp32': torch.float32}[args.dtype])
    net = build_net(dev)
    H, W = args.lr_h, args.lr_w
    out_pix = 16 * H * W
    g = torch.Generator(device='cpu').manual_seed(rank)
    x_host = torch.rand(1, 3, H, W, generator=g).pin_memory()
    x_dev = x_host.to(dev)
    y_host = torch.empty(1, 3, 4 * H, 4 * W, dtype=torch.float32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, drain=None):
        with torch.no_grad():
            for _ in range(warmup):
                fn()
            if drain:
                drain()
            barrier()
            evs = []
            for k in range(steps):
                flush.zero_()                       # L2 flush between timed iterations (not timed)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                if drain and k == steps - 1:
                    drain()                         # the last frame's copy-out is inside the timed region
                e1.record()
                evs.append((e0, e1))
            barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # The forward is launched as a CUDA-graph replay (ofa_b200.GraphedModule: same kernels, captured once; the
    # multi-job weight pack is part of the graph); the eager launch path is reported in `extra.eager`.
    fast = ofa_b200.GraphedModule(net) if not args.eager else net
    # e2e: two captures with their own output buffers -- frame k's device -> host copy reads one while frame k+1 fills the other
    fast2 = ofa_b200.GraphedModule(net, copies=2) if not args.eager else net

    def step_resident():
        return fast(x_dev)

    def step_eager():
        return net(x_dev)

    # e2e: the call a user makes (net(x)) with the frame coming from pinned host memory and the fp32 SR image
    # going back to pinned host memory, every step.  The device -> host copy of frame i runs on a copy stream
    # while frame i+1 computes (two result buffers on each side); the timed region ends when the LAST copy has
    # landed (the compute stream waits for the copy stream before the closing event).
    copy_stream = torch.cuda.Stream(device=dev)
    y_hosts = [y_host, torch.empty_like(y_host).pin_memory()]
    e2e_state = {'i': 0, 'done': [None, None]}

    # host -> device: frame k+1 is uploaded on its own stream while frame k computes (two device input buffers); every
    # step still uploads exactly one frame inside the timed region
    h2d_stream = torch.cuda.Stream(device=dev)
    x_devs = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    h2d_ready = [None, None]

    def upload(slot):
        with torch.cuda.stream(h2d_stream):
            x_devs[slot].copy_(x_host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        h2d_ready[slot] = ev

    def step_e2e():
        i = e2e_state['i'] & 1
        e2e_state['i'] += 1
        if h2d_ready[i] is None:
            upload(i)                                   # very first frame
        torch.cuda.current_stream().wait_event(h2d_ready[i])
        h2d_stream.wait_stream(torch.cuda.current_stream())   # the other buffer's last reader has been queued
        upload(i ^ 1)                                   # next frame's copy overlaps this frame's compute
        xd = x_devs[i]
        if e2e_state['done'][i] is not None:            # this graph copy's output buffer: frame k-2's copy-out has read it
            torch.cuda.current_stream().wait_event(e2e_state['done'][i])
        y = fast2(xd)
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            y_hosts[i].copy_(y, non_blocking=True)
            if args.eager:
                y.record_stream(copy_stream)
            done = torch.cuda.Event()
            done.record()
        e2e_state['done'][i] = done

    def e2e_drain():
        torch.cuda.current_stream().wait_stream(copy_stream)
        torch.cuda.current_stream().wait_stream(h2d_stream)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    # kernels of ours inside the timed region: the launches of one forward (counted by the library on an eager pass; a
    # graph replay executes the same kernel nodes) x the timed steps
    with torch.no_grad():
        net(x_dev)
        B.launch_count_reset()
        net(x_dev)
    torch.cuda.synchronize()
    launches = B.launch_count() * args.steps
    total_ms = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.finish() if sampler else None
    e2e_ms = timed(step_e2e, args.steps, args.warmup, drain=e2e_drain)
    eager_ms = timed(step_eager, args.steps, args.warmup) if not args.eager else total_ms

    ms_per_step = total_ms / args.steps
    value = world * out_pix / 1e6 / (ms_per_step / 1e3)
    e2e_val = world * out_pix / 1e6 / (e2e_ms / args.steps / 1e3)

    # ---- live per-kernel timing over the same workload: pick the dominant kernel, roofline it ---------
    roof = None
    if rank == 0:
        rec = []
        OF.set_profiler(rec)
        with torch.no_grad():
            for _ in range(3):
                flush.zero_()
                net(x_dev)
        torch.cuda.synchronize()
        OF.set_profiler(None)
        # a tag can cover launches of different sizes (the PixelShuffle conv runs at 1x and at 2x resolution): flops and
        # bytes are SUMMED over the launches, so the rates are true averages (round 1 kept the first launch's figures)
        agg = {}
        for tag, flops, nbytes, e0, e1 in rec:
            a = agg.setdefault(tag, [0.0, 0, 0.0, 0.0])
            a[0] += e0.elapsed_time(e1)
            a[1] += 1
            a[2] += flops
            a[3] += nbytes
        total = sum(a[0] for a in agg.values())
        pk = peaks()
        table = []
        for tag, (ms, cnt, flops, nbytes) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            table.append({'kernel': tag, 'share': ms / total, 'launches_per_step': cnt // 3, 'avg_ms': ms / cnt,
                          'GBps': nbytes / (ms / 1e3) / 1e9, 'TFLOPs': flops / (ms / 1e3) / 1e12})
        top = table[0]
        tag = top['kernel']
        flops, nbytes = agg[tag][2] / agg[tag][1], agg[tag][3] / agg[tag][1]
        tensor_bound = tag.startswith('conv') and tag.endswith('tc') and (flops / nbytes) > 247
        if tensor_bound:
            roof = {'bound': 'tensor', 'achieved': top['TFLOPs'], 'peak': pk['tensor_sustained'], 'unit': 'TFLOP/s',
                    'frac': top['TFLOPs'] / pk['tensor_sustained'], 'traffic': None}
        else:
            roof = {'bound': 'hbm', 'achieved': top['GBps'], 'peak': pk['hbm'], 'unit': 'GB/s',
                    'frac': top['GBps'] / pk['hbm'], 'traffic': None}
        roof['traffic'] = ncu_traffic(tag)
        roof.update({'kernel': tag, 'share_of_step': top['share'], 'peak_source': pk['source'],
                     'algorithmic_bytes': nbytes, 'kernels': table[:8]})

    # ---- secondary configurations: reported, not the headline ---------------------------------------
    extra = run_extras(args, dev, dist, rank, world, net, x_host, flush) if not args.no_extras else None
    if extra is not None:
        extra['eager'] = {'ms_per_step': eager_ms / args.steps, 'value': world * out_pix / 1e6 / (eager_ms / args.steps / 1e3),
                          'unit': UNIT, 'what': 'the same forward launched eagerly (net(x)) instead of as a CUDA-graph replay'}

    # ---- the workload WITH a collective: progressive-shrinking training steps (C3 on S4, C4 on X4) -------------
    train = None
    if not args.no_train:
        del net
        torch.cuda.empty_cache()
        train = {'c3': run_train(dev, dist, rank, world, 'c3', args.train_steps, 8),
                 'c4': run_train(dev, dist, rank, world, 'c4', max(4, args.train_steps // 2), 8)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        pix, times = cpu_forward_sample(args.cpu_tile, 6, 1)
        v = pix / 1e6 / float(np.mean(times))
        cpu = {'value': v, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port',
               'sample': 'oracle port, S4 max subnet, one %dx%d LR tile -> 4x, 1 warm-up + 6 timed forwards (~10 s), '
                         'torch CPU fp32 on all host threads' % (args.cpu_tile, args.cpu_tile)}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': args.dtype, 'data': 'synthetic',
            'config': {'workload': workload_text(W, H),
                       'l2': 'flushed between timed iterations (256 MiB write); per-step activations are also >> L2',
                       'timing': 'CUDA events per step on the launch stream, summed, max over ranks',
                       'launch': ('eager net(x)' if args.eager else
                                  'CUDA-graph replay of net.forward (ofa_b200.GraphedModule), one graph launch per frame')},
            'e2e': {'value': e2e_val, 'unit': UNIT, 'h2d_bytes_per_step': x_host.numel() * 4,
                    'd2h_bytes_per_step': y_host.numel() * 4},
            'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roof, 'cpu_baseline': cpu,
            'train': train, 'extra': extra,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    # exactly ONE line may reach stdout (the JSON): libraries that print there (NCCL's version banner) are
    # sent to stderr, the JSON line is written to the saved descriptor
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    _print = print

    def print(*args, **kw):  # noqa: A001
        sys.stdout.flush()
        os.write(_real_stdout, (' '.join(str(x) for x in args) + '\n').encode())

    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
