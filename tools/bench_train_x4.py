"""C4 (BASELINE.json configs[3]): the task-aware downscale -> upscale net OFAMobileNetX4 trained at 2x and 4x in the same
step with teacher distillation (progressive_shrinking.py:158-203, kd_type != 'ce'): per step, the frozen teacher (max
sub-network, eval mode, no_grad) runs at both scales, then a 2x and a 4x student are sampled, loss = MSE(out, HR) +
kd * MSE(out, teacher_out), gradients accumulate over the two students (dynamic_batch_size = 2), one fused Adam step.
Batch of 96x96 HR patches (HR in, HR out), bf16 training path.  One process per GPU; with several ranks the batch is
per GPU (weak scaling) and the gradients are all-reduced once per step.
    python tools/bench_train_x4.py [--batch 64] [--steps 5]"""
import argparse, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ofa-for-super-resolution_b200'))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import copy
import torch
import ofa_b200
import ofa_sr_oracle as O
from ofa_b200 import backend as B, optim
from ofa_b200.elastic_nn.modules.dynamic_op import DynamicSeparableConv2d
from ofa_b200.elastic_nn.networks import OFAMobileNetX4

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--steps', type=int, default=5)
ap.add_argument('--kd', type=float, default=0.5)
a = ap.parse_args()
rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
local_rank = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local_rank)
dev = torch.device('cuda', local_rank)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=dev)
ofa_b200.set_train_dtype(torch.bfloat16)
ofa_b200.set_compute_dtype(torch.float16)          # the teacher's inference path
DynamicSeparableConv2d.KERNEL_TRANSFORM_MODE = 1
cfg = dict(ks_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], depth_list=[2, 3, 4], pixelshuffle_depth_list=[1, 2])
net = OFAMobileNetX4(**{k: list(v) for k, v in cfg.items()})
spec = O.SuperNetSpec('x4', cfg['ks_list'], cfg['expand_ratio_list'], cfg['depth_list'], [1, 2])
net.load_state_dict(O.synth_state_dict(spec.param_shapes(), 7))
net = net.to(dev).train()
teacher = copy.deepcopy(net).eval()                # frozen copy, as the reference loads its teacher checkpoint
hr = torch.rand(a.batch, 3, 96, 96, device=dev)
decay, no_decay = optim.split_no_decay(net.named_parameters())
opt = optim.FusedAdam(decay, no_decay, lr=1e-4, weight_decay=3e-5)
reducer = None
if world > 1:
    from ofa_b200 import parallel as P
    P.broadcast_parameters(net, src=0)
    reducer = P.FlatGradAllReduce(net.parameters(), n_buckets=2)


def step(i):
    net.zero_grad(set_to_none=True)
    random.seed(i)
    with torch.no_grad():
        soft = {}
        for pd in (1, 2):
            teacher.set_active_subnet(ks=7, e=6, d=4, pixel_d=pd)
            soft[pd] = teacher(hr)
    for pd in (1, 2):                                # one 2x and one 4x student per step
        net.sample_active_subnet()
        net.set_active_subnet(pixel_d=pd)
        out = net(hr)
        loss = torch.nn.functional.mse_loss(out, hr) + a.kd * torch.nn.functional.mse_loss(out, soft[pd])
        loss.backward()
    if reducer is not None:
        reducer.reduce()
    opt.step()


for i in range(2):
    step(i)
torch.cuda.synchronize()
B.launch_count_reset()
t0 = time.perf_counter()
for i in range(a.steps):
    step(10 + i)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / a.steps
if world > 1:
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
if rank == 0:
    print('%d GPU(s) X4 joint 2x/4x distillation step (teacher x2 + 2 students, batch %d per GPU): %.2f ms  %.1f patches/s  '
          '(%d library launches/step/rank)' % (world, a.batch, dt * 1e3, a.batch * world / dt, B.launch_count() // a.steps))
if world > 1:
    dist.destroy_process_group()
