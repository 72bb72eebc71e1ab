"""per-parameter gradient cosine (bf16 mixed-precision product vs fp32 oracle autograd), X4 or S4."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ofa-for-super-resolution_b200')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np, torch
import ofa_b200, ofa_sr_oracle as O
from ofa_b200.elastic_nn.modules.dynamic_op import DynamicSeparableConv2d
from ofa_b200.elastic_nn.networks import OFAMobileNetS4, OFAMobileNetX4
kind = sys.argv[1] if len(sys.argv) > 1 else 'x4'
tdt = {'bf16': torch.bfloat16, 'fp32': torch.float32}[sys.argv[2] if len(sys.argv) > 2 else 'bf16']
dev = torch.device('cuda:0')
DynamicSeparableConv2d.KERNEL_TRANSFORM_MODE = 1
FULL = dict(ks_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], depth_list=[2, 3, 4])
cls = OFAMobileNetS4 if kind == 's4' else OFAMobileNetX4
net = cls(pixelshuffle_depth_list=[1, 2], **{k: list(v) for k, v in FULL.items()})
spec = O.SuperNetSpec(kind, FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], [1, 2])
sd = O.synth_state_dict(spec.param_shapes(), 81)
net.load_state_dict(sd); net = net.to(dev).train()
sd_ref = {k: (v.clone().requires_grad_(True) if v.dtype == torch.float32 and 'running' not in k else v.clone()) for k, v in sd.items()}
rs = np.random.RandomState(6)
if kind == 's4':
    x = torch.from_numpy(rs.rand(4, 3, 12, 16).astype(np.float32)); tgt = torch.from_numpy(rs.rand(4, 3, 48, 64).astype(np.float32))
else:
    x = torch.from_numpy(rs.rand(4, 3, 32, 48).astype(np.float32)); tgt = x.clone()
sub = dict(ks=7, e=6, d=4, pixel_d=2) if kind == 's4' else dict(ks=5, e=4, d=3, pixel_d=2)
net.set_active_subnet(**sub); spec.set_active_subnet(**sub)
ofa_b200.set_train_dtype(tdt)
out = net(x.to(dev)); loss = torch.nn.functional.mse_loss(out.float(), tgt.to(dev)); loss.backward()
out_ref = O.supernet_forward(x, sd_ref, spec, training=True); loss_ref = torch.nn.functional.mse_loss(out_ref, tgt); loss_ref.backward()
print('loss', float(loss.detach()), float(loss_ref.detach()))
for pname, p in net.named_parameters():
    g = sd_ref[pname].grad
    if g is None or float(g.norm()) == 0 or p.grad is None: continue
    a, b = p.grad.cpu().flatten(), g.flatten()
    print('%-70s numel %6d  norm ratio %.3f  cos %.4f' % (pname, b.numel(), float(a.norm() / b.norm()), float(a @ b / (a.norm() * b.norm()))))
