"""CPU: bookkeeping of panic3d_b200.graphs (argument keying, bypass rules); capture / replay is tested on the GPU."""
import torch

from panic3d_b200 import graphs


def test_flatten_rebuild_round_trip_and_key_stability():
    a, b, c = torch.zeros(2, 3), torch.ones(4), torch.full((1,), 7.0)
    args = ([a, {'z': b, 'k': 3, 'y': (c, 'const')}], {'noise_mode': 'const', 'cond': {'img': a}})
    t1, t2 = [], []
    k1 = graphs._flatten(args, t1)
    k2 = graphs._flatten(([a.clone(), {'y': (c, 'const'), 'k': 3, 'z': b}], {'cond': {'img': a}, 'noise_mode': 'const'}), t2)
    assert k1 == k2 and hash(k1) == hash(k2)                      # dict order does not matter, values of tensors do not enter the key
    assert len(t1) == 4
    assert graphs._flatten(([a, {'z': b, 'k': 4, 'y': (c, 'const')}], {'noise_mode': 'const', 'cond': {'img': a}}), []) != k1
    assert graphs._flatten(([a.double(), {'z': b, 'k': 3, 'y': (c, 'const')}], {'noise_mode': 'const', 'cond': {'img': a}}), []) != k1
    new = [t + 1 for t in t1]
    back = graphs._rebuild(args, new, [0])
    t3 = []
    graphs._flatten(back, t3)
    assert all(x is y for x, y in zip(t3, new)) and back[0][1]['k'] == 3 and back[1]['noise_mode'] == 'const'


def test_bypass_rules_without_cuda():
    calls = []

    def fn(x, scale=1.0, return_more=False, obj=None):
        calls.append(1)
        return x * scale

    g = graphs.GraphedCallable(fn, name='fn')
    x = torch.arange(4.0)
    assert torch.equal(g(x, scale=2.0), x * 2)                    # CPU tensor: eager
    assert torch.equal(g(x, obj=object()), x)                     # unkeyable argument: eager
    assert g.bypassed == 2 and g.captures == 0 and g.hits == 0 and len(calls) == 2


def test_enable_is_idempotent_and_keeps_module_state():
    class Syn(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(3, 3)

        def forward(self, x):
            return self.lin(x)

    class G(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.backbone = torch.nn.Module()
            self.backbone.synthesis = Syn()
            self.superresolution = Syn()

    m = G()
    keys = list(m.state_dict())
    w1 = graphs.enable_cuda_graphs(m)
    w2 = graphs.enable_cuda_graphs(m)
    assert w1['backbone'] is w2['backbone'] and w1['superresolution'] is w2['superresolution']
    assert list(m.state_dict()) == keys
    x = torch.randn(2, 3)
    assert torch.equal(m.backbone.synthesis(x), w1['backbone'].fn(x))     # CPU call goes straight through


def test_lean_return_more_replaces_locals_with_an_empty_dict():
    def fn(x, return_more=False):
        return (x + 1, locals()) if return_more else x + 1

    x = torch.zeros(3)
    keep = graphs.GraphedCallable(fn)
    y, loc = keep(x, return_more=True)
    assert torch.equal(y, x + 1) and 'x' in loc                      # default: the reference's behaviour, eager
    lean = graphs.GraphedCallable(fn, lean_return_more=True)
    y, loc = lean(x, return_more=True)
    assert torch.equal(y, x + 1) and loc == {}
    assert torch.equal(lean(x), x + 1)


def test_composes_with_the_plane_memo_in_either_order():
    import panic3d_b200.dropin as dropin

    class Syn(torch.nn.Module):
        def forward(self, ws, cond=None, **kw):
            return ws * 2

    def make():
        g = torch.nn.Module()
        g.backbone = torch.nn.Module()
        g.backbone.synthesis = Syn()
        return g

    ws = torch.ones(1, 4)
    a = make()
    memo = dropin.enable_plane_reuse(a)
    w = graphs.enable_cuda_graphs(a, superresolution=False)['backbone']
    assert a.backbone.synthesis.forward is memo and memo.synthesis is w        # graph underneath the memo
    b = make()
    w2 = graphs.enable_cuda_graphs(b, superresolution=False)['backbone']
    memo2 = dropin.enable_plane_reuse(b)
    assert b.backbone.synthesis.forward is memo2 and memo2.synthesis is w2
    for g_, m_ in ((a, memo), (b, memo2)):
        with torch.no_grad():
            assert torch.equal(g_.backbone.synthesis(ws, None, noise_mode='const'), ws * 2)
            assert torch.equal(g_.backbone.synthesis(ws, None, noise_mode='const'), ws * 2)
        assert m_.hits == 1 and m_.misses == 1
