"""Read-only / write-only / copy HBM bandwidth of this GPU with plain torch ops (context for the roofline fractions of
kernels whose traffic is not 50/50 read/write).  1 GiB fp16 buffers, CUDA events, best of 10."""
import torch
dev = torch.device('cuda:0')
n = 512 * 1024 * 1024
a = torch.empty(n, dtype=torch.float16, device=dev).normal_()
b = torch.empty_like(a)


def best(fn, nbytes, reps=10):
    fn(); torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return nbytes / min(t) / 1e6


print('copy  (read+write) %.0f GB/s' % best(lambda: b.copy_(a), 2 * n * 2))
print('write only (fill)  %.0f GB/s' % best(lambda: b.fill_(1.0), n * 2))
print('read only (sum)    %.0f GB/s' % best(lambda: a.sum(), n * 2))
print('read only (max)    %.0f GB/s' % best(lambda: a.max(), n * 2))
