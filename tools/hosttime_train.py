"""Host enqueue time vs device time of the eager C3 training step (is the step bounded by the Python launch loop?).
    python tools/hosttime_train.py [--max-subnet]
"""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ofa-for-super-resolution_b200'))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch
import ofa_b200
import ofa_sr_oracle as O
from ofa_b200 import backend as B, optim
from ofa_b200.elastic_nn.modules.dynamic_op import DynamicSeparableConv2d
from ofa_b200.elastic_nn.networks import OFAMobileNetS4
mx = '--max-subnet' in sys.argv
dev = torch.device('cuda', 0)
ofa_b200.set_train_dtype(torch.bfloat16)
DynamicSeparableConv2d.KERNEL_TRANSFORM_MODE = 1
cfg = dict(ks_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], depth_list=[2, 3, 4], pixelshuffle_depth_list=[1, 2])
net = OFAMobileNetS4(**{k: list(v) for k, v in cfg.items()})
spec = O.SuperNetSpec('s4', cfg['ks_list'], cfg['expand_ratio_list'], cfg['depth_list'], [1, 2])
net.load_state_dict(O.synth_state_dict(spec.param_shapes(), 7))
net = net.to(dev).train()
lr_img = torch.rand(64, 3, 24, 24, device=dev)
hr_img = torch.rand(64, 3, 96, 96, device=dev)
decay, no_decay = optim.split_no_decay(net.named_parameters())
opt = optim.FusedAdam(decay, no_decay, lr=1e-4, weight_decay=3e-5)
T = {'zero+sample': 0.0, 'fwd': 0.0, 'bwd': 0.0, 'opt': 0.0}


def step(i, rec):
    t0 = time.perf_counter()
    net.zero_grad(set_to_none=True)
    random.seed(i)
    if mx:
        net.set_active_subnet(ks=7, e=6, d=4, pixel_d=2)
    else:
        net.sample_active_subnet()
        net.set_active_subnet(pixel_d=2)
    t1 = time.perf_counter()
    loss = torch.nn.functional.mse_loss(net(lr_img), hr_img)
    t2 = time.perf_counter()
    loss.backward()
    t3 = time.perf_counter()
    opt.set_lr(1e-4)
    opt.step()
    t4 = time.perf_counter()
    if rec:
        T['zero+sample'] += t1 - t0; T['fwd'] += t2 - t1; T['bwd'] += t3 - t2; T['opt'] += t4 - t3


for i in range(40):
    step(i, False)
torch.cuda.synchronize()
N = 20
# (a) host enqueue time with the device kept idle-free: sync before every step so no back-pressure from the launch queue
for i in range(N):
    torch.cuda.synchronize()
    step(100 + i, True)
torch.cuda.synchronize()
print('host enqueue per step (device drained before each step): ' + '  '.join('%s %.2f ms' % (k, v / N * 1e3) for k, v in T.items()),
      ' total %.2f ms' % (sum(T.values()) / N * 1e3))
# (b) free-running
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
B.launch_count_reset()
t0 = time.perf_counter(); e0.record()
for i in range(N):
    step(100 + i, False)
e1.record(); th = time.perf_counter() - t0
torch.cuda.synchronize()
print('free-running: host loop %.2f ms/step, device %.2f ms/step, %d library calls/step' % (th / N * 1e3, e0.elapsed_time(e1) / N, B.launch_count() // N))
