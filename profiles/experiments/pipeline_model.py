"""Performance model of the fused renderers built on the protocol simulation (tests/test_fused_protocol_sim_cpu.py):
same control flow, but every work item gets the duration measured on hardware (P3D_WS_TIMING cycle accounts in
profiles/r1_fused_summary.md, in k-cycles) instead of a random one.  The simulated time per ray group is the pipeline's
steady-state period if nothing but the dependency structure and the role occupancy limited it.

    python profiles/experiments/pipeline_model.py

Round-1 numbers: shipped kernel 42.5 k cycles/group modelled vs 49.3 k measured (the model has no issue-slot contention
between roles); depth-3 kernel (render_fused_ws3.cu) 31 k modelled - the gather role's 2 tiles x 15 k per group."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import tests.test_fused_protocol_sim_cpu as P          # noqa: E402


def period(simcls, n_groups, d):
    sim = simcls(n_groups, 0)
    table = {(5, 20): d['gather_tile'], (0.2, 2): d['mma'], (0.2, 1): d['tmem_read'], (1, 4): d['tile_convert'],
             (2, 10): d['per_ray_phase'], (2, 12): d['per_ray_phase'], (1, 6): d['colours'], (0.1, 0.5): 0.2}
    sim.dur = lambda lo, hi: table[(lo, hi)]
    sim.run()
    return sim.now / n_groups


if __name__ == '__main__':
    shipped = dict(gather_tile=15.0, mma=0.3, tmem_read=0.45, tile_convert=1.7, per_ray_phase=5.9, colours=4.1)
    depth3 = dict(gather_tile=15.0, mma=0.3, tmem_read=0.45, tile_convert=1.7, per_ray_phase=7.5, colours=7.0)
    for n in (16, 32):
        print(f'{n} groups/CTA: shipped {period(P.SimV3, n, shipped):.1f} k cycles/group (measured 49.3), '
              f'depth-3 {period(P.Sim, n, depth3):.1f}, depth-3 with a 20 % faster gather '
              f'{period(P.Sim, n, dict(depth3, gather_tile=12.0)):.1f}')
