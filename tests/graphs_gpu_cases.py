"""Bodies of tests/test_zz_graphs_gpu.py - each runs in its own interpreter (a capture that is made to fail must not be able to
disturb the CUDA context the rest of the GPU suite shares)."""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from panic3d_b200 import graphs

DEV = 'cuda:0'


class Net(torch.nn.Module):
    """A miniature of what gets graphed: convs, p3d ops, per-call noise, dict-valued input, tuple-valued output."""

    def __init__(self):
        super().__init__()
        self.c1 = torch.nn.Conv2d(3, 8, 3, padding=1)
        self.c2 = torch.nn.Conv2d(8, 3, 3, padding=1)
        from panic3d_b200.torch_utils.ops import upfirdn2d
        self.register_buffer('f', upfirdn2d.setup_filter([1, 3, 3, 1]))          # like the networks: filters are buffers, built once
                                                                                 # (setup_filter inside forward = a host->device copy per call,
                                                                                 #  which no capture allows - the wrapper then stays eager)

    def forward(self, x, cond, noise_mode='random'):
        from panic3d_b200.torch_utils.ops import bias_act, upfirdn2d
        h = self.c1(x + cond['img'])
        if noise_mode == 'random':
            h = h + torch.randn([x.shape[0], 1, x.shape[2], x.shape[3]], device=x.device) * 0.1
        h = bias_act.bias_act(h, self.c1.bias, act='lrelu')
        h = upfirdn2d.upsample2d(h, self.f)
        return self.c2(h), {'feat': h}


def capture_replay_equals_eager_and_noise_is_fresh():
    torch.manual_seed(0)
    net = Net().to(DEV).eval().requires_grad_(False)
    g = graphs.GraphedCallable(net.forward, name='net')
    x1, x2 = torch.randn(2, 3, 16, 16, device=DEV), torch.randn(2, 3, 16, 16, device=DEV)
    cond = {'img': torch.randn(2, 3, 16, 16, device=DEV)}
    with torch.no_grad():
        for x in (x1, x2, x1):
            y, d = g(x, cond, noise_mode='const')
            ye, de = net(x, cond, noise_mode='const')
            assert torch.equal(y, ye) and torch.equal(d['feat'], de['feat'])          # same kernels, same data: bit-identical
        assert g.captures == 1 and g.hits == 3
        a = g(x1, cond)[0].clone()                                                 # noise_mode='random': a second signature
        b = g(x1, cond)[0]
        assert g.captures == 2 and not torch.equal(a, b)                            # every replay draws new noise
        y3, _ = g(torch.randn(1, 3, 8, 8, device=DEV), {'img': torch.zeros(1, 3, 8, 8, device=DEV)}, noise_mode='const')
        assert g.captures == 3 and tuple(y3.shape) == (1, 3, 16, 16)
        out1 = g(x1, cond, noise_mode='const')[0]
        out2 = g(x2, cond, noise_mode='const')[0]
        assert not torch.equal(out1, out2) and torch.equal(out1, net(x1, cond, noise_mode='const')[0])   # returned tensors are clones


def autograd_calls_and_failing_captures_stay_eager():
    net = Net().to(DEV)
    g = graphs.GraphedCallable(net.forward, name='net')
    x = torch.randn(1, 3, 8, 8, device=DEV, requires_grad=True)
    cond = {'img': torch.zeros(1, 3, 8, 8, device=DEV)}
    y, _ = g(x, cond, noise_mode='const')
    y.sum().backward()
    assert x.grad is not None and g.captures == 0 and g.bypassed == 1

    def syncing(t):
        return t * float(t.sum().item())                                            # a host sync: illegal during capture

    bad = graphs.GraphedCallable(syncing, name='syncing', warmup=1)
    t = torch.ones(4, device=DEV)
    with warnings.catch_warnings(record=True) as w, torch.no_grad():
        warnings.simplefilter('always')
        r1 = bad(t)
        r2 = bad(t)
    assert torch.equal(r1, t * 4) and torch.equal(r2, t * 4)
    assert bad.captures == 0 and bad.bypassed == 2 and len(bad.failed) == 1
    assert sum('stays on the eager path' in str(m.message) for m in w) == 1        # said once
    assert torch.equal(net.c1.weight * 2, net.c1.weight + net.c1.weight)            # the context is still healthy


if __name__ == '__main__':
    globals()[sys.argv[1]]()
    print('CASE_OK', sys.argv[1])
