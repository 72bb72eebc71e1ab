"""Host cost of one library call from Python (cProfile over many calls): inference conv and the training conv + BN pair."""
import cProfile, pstats, io, os, sys, time, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'ofa-for-super-resolution_b200'))
import ofa_b200
from ofa_b200 import functional as OF, backend as B
dev = torch.device('cuda:0')
w = torch.randn(64, 64, 1, 1, device=dev) * 0.05
x = torch.randn(4, 64, 24, 24, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
cache = OF.PackedWeightCache()
f = lambda: OF.conv_bn_act_infer(x, w, 64, 64, 1, None, B.ACT_NONE, cache=cache, out_dtype=torch.bfloat16)
for _ in range(10):
    f()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2000):
    f()
torch.cuda.synchronize()
print('conv_bn_act_infer: %.1f us per call (wall, tiny kernel)' % ((time.perf_counter() - t0) / 2000 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(2000):
    f()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(14); print(s.getvalue()[:3500])

# training pair: conv2d (autograd) + bn_act, forward + backward
ofa_b200.set_train_dtype(torch.bfloat16)
wp = torch.nn.Parameter(w.clone())
bn = torch.nn.BatchNorm2d(64).to(dev).train()
def g():
    y = OF.bn_act(OF.conv2d(x, wp, 64, 64, 1), bn, 64, B.ACT_RELU6)
    y.backward(y.detach())
    wp.grad = None; bn.weight.grad = None; bn.bias.grad = None
for _ in range(10):
    g()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(500):
    g()
torch.cuda.synchronize()
print('conv2d + bn_act fwd+bwd: %.1f us per pair (wall)' % ((time.perf_counter() - t0) / 500 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(500):
    g()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(22); print(s.getvalue()[:5000])
