"""CPU, world_size 2 over gloo: the view-sharding host logic (panic3d_b200.views).  The renderer itself needs a
GPU, so the stand-in 'renderer' here is the CPU oracle with the same call shape; what is under test is the
sharding arithmetic, the depth-bound all-reduce protocol and the gather order."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_is_a_balanced_partition():
    sys.path.insert(0, ROOT)
    from panic3d_b200.views import shard_range
    for n in (0, 1, 3, 8, 16, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


class _OracleRenderer:
    """Same call surface as ImportanceRenderer.forward, two-phase depth clamp like the CUDA path."""

    def __init__(self, case, u_c, u_f):
        self.case, self.u_c, self.u_f = case, u_c, u_f
        self.depth_bounds_reduce = None
        self.offset = 0

    def __call__(self, planes, dec, ro, rd, opts, **flags):
        from oracle import renderer_oracle as orc
        n, M = ro.shape[0], ro.shape[1]
        a = self.offset
        u_c, u_f = self.u_c[a:a + n], self.u_f[a * M:(a + n) * M]
        kw = dict(use_triplane=self.case.get('use_triplane', True))
        res, (lo, hi) = orc.render(planes, dec, ro, rd, opts, u_c, u_f, return_bounds=True, **kw)
        if self.depth_bounds_reduce is not None:
            b2 = torch.stack([lo, hi])
            self.depth_bounds_reduce(b2)
            res = orc.render(planes, dec, ro, rd, opts, u_c, u_f, depth_bounds=(b2[0], b2[1]), **kw)
        return res


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from panic3d_b200 import views
        from tests.golden.cases import RENDER_CASES, build_case_inputs
        from tests.helpers import case_rays, load_golden
        # gather of uneven shards keeps view order
        n = 5
        a, b = views.shard_range(n, world, rank)
        local = torch.arange(a, b, dtype=torch.float32)[:, None].expand(-1, 3).contiguous()
        full = views.gather_views(local, n, dst=None)
        assert torch.equal(full[:, 0], torch.arange(n, dtype=torch.float32))
        only0 = views.gather_views(local, n, dst=0)
        assert (only0 is not None) == (rank == 0)
        # PeerGather falls back to the collective gather off-GPU (on the box it is a CUDA-IPC peer copy): same result contract
        pg = views.PeerGather((2, 3), torch.float32, 'cpu', dst=0)
        pg.push(torch.full((2, 3), float(rank)), step=0)
        pg.fence()
        res = pg.result(0)
        assert (res is not None) == (rank == 0)
        if rank == 0:
            assert res.shape == (2 * world, 3) and torch.equal(res[:, 0], torch.arange(world, dtype=torch.float32).repeat_interleave(2))
        # depth-bound protocol
        b2 = torch.tensor([1.0 + rank, 2.0 + rank])
        views.all_reduce_depth_bounds(b2)
        assert b2.tolist() == [1.0, 2.0 + world - 1]
        # sharded render of the 3-view batch == the reference's single-process batch render (fixture)
        case = RENDER_CASES['small_batch3']
        planes, dec, c2w, K, u_c, u_f, opts = build_case_inputs(case)
        ro, rd = case_rays(case, c2w, K)
        r = _OracleRenderer(case, u_c, u_f)
        r.offset = views.shard_range(3, world, rank)[0]
        out = views.render_sharded(r, planes, dec, ro, rd, opts, dst=0, exact_depth=True)
        if rank == 0:
            g = load_golden('render', 'small_batch3')
            errs = [float((o - g[k]).abs().max()) for o, k in zip(out, ('rgb', 'depth', 'wsum', 'xyz'))]
            assert max(errs) < 2e-5, errs
        else:
            assert all(o is None for o in out)
        q.put((rank, 'ok'))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_sharded_render_over_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == 'ok', f'rank {rank}: {msg}'
