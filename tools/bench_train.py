"""C3 (BASELINE.json configs[2]) on one GPU: progressive-shrinking training step of OFAMobileNetS4 on a batch of
64 synthetic 96x96 HR patches (24x24 LR in), forward + backward + fused Adam step through the library's kernels.  Not the
headline bench; prints ms/step and patches/s.
    python tools/bench_train.py [--batch 64] [--steps 5]
"""
import argparse, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ofa-for-super-resolution_b200'))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch
import ofa_b200
import ofa_sr_oracle as O
from ofa_b200 import backend as B
from ofa_b200.elastic_nn.modules.dynamic_op import DynamicSeparableConv2d
from ofa_b200.elastic_nn.networks import OFAMobileNetS4

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--steps', type=int, default=30)
ap.add_argument('--warmup', type=int, default=8, help='untimed steps: sampled sub-networks change the tensor sizes every step, so the allocator needs a few steps to settle')
ap.add_argument('--max-subnet', dest='max', action='store_true', help='max subnet instead of sampled ones')
ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
ap.add_argument('--verbose', action='store_true')
ap.add_argument('--weak', action='store_true', help='weak scaling: --batch patches PER GPU (default: the batch is split over the ranks)')
ap.add_argument('--graph', action='store_true', help='capture the whole step (fwd + bwd + Adam) of the fixed max subnet in a CUDA graph')
a = ap.parse_args()
rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
local_rank = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local_rank)
dev = torch.device('cuda', local_rank)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=dev)
ofa_b200.set_train_dtype(torch.bfloat16 if a.dtype == 'bf16' else torch.float32)
DynamicSeparableConv2d.KERNEL_TRANSFORM_MODE = 1
cfg = dict(ks_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], depth_list=[2, 3, 4], pixelshuffle_depth_list=[1, 2])
net = OFAMobileNetS4(**{k: list(v) for k, v in cfg.items()})
spec = O.SuperNetSpec('s4', cfg['ks_list'], cfg['expand_ratio_list'], cfg['depth_list'], [1, 2])
net.load_state_dict(O.synth_state_dict(spec.param_shapes(), 7))
net = net.to(dev).train()
per_rank = a.batch if a.weak else a.batch // world     # batch-sharded data parallelism (default: strong scaling of one step)
lr_img = torch.rand(per_rank, 3, 24, 24, device=dev)
hr_img = torch.rand(per_rank, 3, 96, 96, device=dev)
from ofa_b200 import optim
decay, no_decay = optim.split_no_decay(net.named_parameters())       # 'bn#bias' keys: no weight decay
opt = optim.FusedAdam(decay, no_decay, lr=1e-4, weight_decay=3e-5)       # Adam(lr 1e-4, wd 3e-5), SURVEY §8d C3
reducer = None
if world > 1:
    from ofa_b200 import parallel as P
    P.broadcast_parameters(net, src=0)
    tail, boundary_module = P.s4_tail_parameters(net)
    reducer = P.FlatGradAllReduce(net.parameters(), n_buckets=2, tail_params=tail)
    # the tail segment's all-reduce starts as soon as backward reaches the input of the tail layers
    boundary_module.register_forward_pre_hook(lambda m, inp: reducer.watch(inp[0]) if inp[0].requires_grad else None)


def step(i):
    net.zero_grad(set_to_none=True)
    random.seed(i)
    if a.max:
        net.set_active_subnet(ks=7, e=6, d=4, pixel_d=2)
    else:
        net.sample_active_subnet()
        net.set_active_subnet(pixel_d=2)
    loss = torch.nn.functional.mse_loss(net(lr_img), hr_img)
    loss.backward()
    if reducer is not None:
        reducer.reduce()
    opt.set_lr(optim.cosine_lr(1e-4, 120, 0, i, 1000))
    opt.step()
    return loss


for i in range(a.warmup):
    step(i)
torch.cuda.synchronize()
if a.graph:
    # whole-step capture (forward, backward, gradient all-reduce, fused Adam) of a FIXED sub-network -- the regime of
    # train_teacher_net_sr_simple.py (max network only); sampled sub-networks change the launch sequence every step
    assert a.max and world == 1, '--graph captures a fixed sub-network on one GPU (capturing the NCCL exchange hung on 2 GPUs)'

    def step_body():
        torch.nn.functional.mse_loss(net(lr_img), hr_img).backward()
        if reducer is not None:
            reducer.reduce()
        opt.step()
    net.zero_grad(set_to_none=True)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step_body()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    net.zero_grad(set_to_none=True)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, capture_error_mode='thread_local'):
        step_body()
    step = lambda i: graph.replay()
B.launch_count_reset()
t0 = time.perf_counter()
per_step = []
for i in range(a.steps):
    ts = time.perf_counter()
    step(100 + i)
    if a.verbose:
        torch.cuda.synchronize()
        per_step.append((time.perf_counter() - ts) * 1e3)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / a.steps
if a.verbose and rank == 0:
    print('per-step ms:', ' '.join('%.1f' % t for t in per_step))
if world > 1:
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
if rank == 0:
    print('%d GPU(s)' % world, a.dtype, 'train step: %.2f ms  %.1f patches/s  (%d library launches/step/rank)' % (dt * 1e3, per_rank * world / dt, B.launch_count() // a.steps))
if world > 1:
    dist.destroy_process_group()
