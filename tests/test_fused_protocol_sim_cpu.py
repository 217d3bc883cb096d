"""CPU: discrete-event simulation of the mbarrier protocols of the fused renderers - the experimental pipeline-depth-3
kernel (csrc/render_fused_ws3.cu, P3D_FUSED_IMPL=v5) and the shipped one (csrc/render_fused_ws.cu).

The kernel has never run on hardware (see its header), so this test checks what can be checked without a GPU: the pass
schedule and the hand-off protocol between its 25 warps.  Every role's control flow is transcribed from the CUDA source
(same loops, same barrier, parity and count for every wait / arrive / tcgen05.commit); work items get random durations
and the runnable warps are stepped in random order, over many seeds and group counts.  Asserted:
  * no deadlock and no lost mbarrier phase (a waiter that a barrier overtakes by two phases blocks forever);
  * TMEM hazards: layer 2 never overwrites a colour area whose previous occupant has not been reduced yet, the colour
    reduction of a group only starts after all six of its tiles were written, D1 and the sigma accumulator are never
    overwritten before they were read;
  * shared-memory hazards: a per-group state slot is not rewritten by the gather before its previous group is finished,
    the omega slot not before its previous group's colours are done, fine depths are read only after importance.
It mirrors the protocol, not the arithmetic; parity on hardware is what tests/test_render_gpu.py (P3D_TEST_V5=1) is for."""
import heapq
import random

import pytest

K_NA, K_EW, K_RW, K_TEAMS = 4, 8, 4, 3


def pass_at(j, N):                                   # render_fused_ws3.cu: pass_at
    if j == 0:
        return (0, 0)
    if N == 1:
        return (0, 1)
    if j == 1:
        return (1, 0)
    k, body = j - 2, 2 * (N - 2)
    if k < body:
        return ((k >> 1) + 2, 0) if (k & 1) else (k >> 1, 1)
    return (N - 2 + (k - body), 1)


def tile_at(q, N):
    j = q // 3
    n, p = pass_at(j, N)
    return (n, p, q - 3 * j)


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def passed(self, parity):                        # mbarrier.try_wait.parity
        return (self.phase & 1) != parity


class Sim:
    def __init__(self, N, seed, gcol=False, altsig=False):
        # gcol / altsig: two protocol experiments that were built, parity-tested and measured slower than the shipped kernel
        # (profiles/experiments/render_fused_ws3_switches.cu.txt, profiles/r2_kernel_log.md); both False = csrc/render_fused_ws3.cu.
        # gcol: colour reduction shared by the gather warps, finalised by the ray warps
        self.gcol = gcol
        # altsig: the two column halves of the epilogue role read sigma back in turns (tile parity), d2_full is
        # one barrier per tile parity so that each half observes every phase of its own, sigc/sigf_ready take one arrival per warp-tile
        self.altsig = altsig
        self.N, self.T, self.rng = N, 6 * N, random.Random(seed)
        B = Bar
        self.a1_full = [B(4) for _ in range(K_NA)]; self.a1_empty = [B(1) for _ in range(K_NA)]
        self.a2_full = [B(K_EW) for _ in range(2)]; self.a2_empty = [B(1) for _ in range(2)]
        self.d1_full, self.d1_empty, self.d2_full, self.dsig_empty = B(1), B(K_EW), ([B(1), B(1)] if altsig else B(1)), B(4)
        self.fine_ready = [B(K_RW) for _ in range(4)]; self.state_free = [B(K_RW if gcol else 1) for _ in range(4)]   # gcol: released by the ray warps' finalize
        self.fa_free = B(12)                          # the twelve gather warps' colour shares of a group have left TMEM
        self.share_count = {}
        self.sigc_ready = [B(12 if altsig else 4) for _ in range(4)]; self.sigf_ready = [B(12 if altsig else 4) for _ in range(4)]
        self.omega_ready = [B(K_RW) for _ in range(2)]
        self.sig_tile = None
        self.now, self.events, self.seq = 0.0, [], 0
        # hazard bookkeeping
        self.area_owner = {}                          # TMEM colour area -> (group, tiles written)
        self.colours_done, self.importance_done, self.composite_done = set(), set(), set()
        self.tiles_written = {}                       # group -> set of (pass, k) whose layer 2 completed
        self.state_owner = [None] * 4
        self.slot_owner = [None] * 2
        self.d1_unread, self.sig_unread = False, False
        self.part_owner, self.finalized = None, (set() if gcol else self.colours_done)
        self.ebar_count, self.ebar_gen = 0, 0

    def dur(self, lo, hi):
        return self.rng.uniform(lo, hi)

    def later(self, dt, fn):
        self.seq += 1
        heapq.heappush(self.events, (self.now + dt, self.seq, fn))

    # ---------------------------------------------------------------- roles (generators yield ('wait', bar, parity) | ('work', dt) | ('ebar',))
    def gather_warp(self, gw):
        team, it = gw >> 2, 0

        def colour_share(n):                          # this warp's share of colours(n): tiles 2*team, 2*team+1 of its TMEM lane quarter
            yield ('wait', self.omega_ready[n & 1], (n >> 1) & 1)
            assert n in self.composite_done and self.slot_owner[n & 1] == n
            assert self.tiles_written.get(n, set()) == {(p, k) for p in (0, 1) for k in range(3)}, f'colour share of {n} before all six tiles'
            assert self.part_owner is None or self.part_owner == n or self.part_owner in self.finalized, 'partial sums overwritten before finalize'
            self.part_owner = n
            yield ('work', self.dur(0.5, 3))
            self.share_count[n] = self.share_count.get(n, 0) + 1
            if self.share_count[n] == 12:
                self.colours_done.add(n)
            self.fa_free.arrive()
        for q in range(self.T):
            n, p, k = tile_at(q, self.N)
            my_it = it; it += 1
            if my_it % K_TEAMS != team:
                continue
            if self.gcol and p == 1 and n >= 1:
                yield from colour_share(n - 1)
            if p == 0:
                yield ('wait', self.state_free[n & 3], ((n >> 2) & 1) ^ 1)
                prev = self.state_owner[n & 3]
                assert prev is None or prev == n or prev in self.finalized, f'state slot of group {prev} rewritten for {n}'
                self.state_owner[n & 3] = n
            else:
                yield ('wait', self.fine_ready[n & 3], (n >> 2) & 1)
                assert n in self.importance_done, f'fine gather of {n} before its importance sampling'
            stage = my_it % K_NA
            yield ('wait', self.a1_empty[stage], ((my_it // K_NA) & 1) ^ 1)
            yield ('work', self.dur(5, 20))
            self.a1_full[stage].arrive()
        if self.gcol:
            yield from colour_share(self.N - 1)

    def mma_thread(self):
        it, prev = 0, None

        def layer2(it2, tp):
            n, p, k = tp
            buf = it2 & 1
            yield ('wait', self.a2_full[buf], (it2 >> 1) & 1)
            yield ('wait', self.dsig_empty, (it2 & 1) ^ 1)
            if self.gcol and p == 1 and k == 0 and n >= 1:
                yield ('wait', self.fa_free, (n - 1) & 1)
            area = ('CA', n % 3, k) if p == 0 else ('FA', k)
            owner = self.area_owner.get(area)
            assert owner is None or owner in self.colours_done, f'layer 2 of {tp} overwrites {area} of group {owner} before its colours'
            assert not self.sig_unread, 'sigma accumulator overwritten before it was read'
            self.area_owner[area] = n

            def done():
                self.tiles_written.setdefault(n, set()).add((p, k))
                self.sig_unread = True
                self.sig_tile = tp
                (self.d2_full[it2 & 1] if self.altsig else self.d2_full).arrive(); self.a2_empty[buf].arrive()
            self.later(self.dur(0.2, 2), done)
        for q in range(self.T):
            td = tile_at(q, self.N)
            stage = it % K_NA
            yield ('wait', self.a1_full[stage], (it // K_NA) & 1)
            yield ('wait', self.d1_empty, (it & 1) ^ 1)
            assert not self.d1_unread, 'D1 overwritten before the epilogue read it'

            def done1(stage=stage):
                self.d1_unread = True
                self.d1_full.arrive(); self.a1_empty[stage].arrive()
            self.later(self.dur(0.2, 2), done1)
            if prev is not None:
                yield from layer2(it - 1, prev)
            prev = td; it += 1
            if td[2] == 2:
                yield from layer2(it - 1, prev)
                prev = None
        if prev is not None:
            yield from layer2(it - 1, prev)

    def epilogue_warp(self, e):
        chunk, it, prev = e >> 2, 0, None
        reads = {'d1': 0}

        def sigma_read(it_prev, tp):
            n, p, k = tp
            if self.altsig:
                if chunk != (it_prev & 1):
                    return
                yield ('wait', self.d2_full[it_prev & 1], (it_prev >> 1) & 1)
            else:
                if chunk != 0:
                    return
                yield ('wait', self.d2_full, it_prev & 1)
            assert self.sig_tile == tp, f'sigma read-back of {tp} found the accumulator of {self.sig_tile}'
            yield ('work', self.dur(0.2, 1))
            self.sig_unread_readers = getattr(self, 'sig_unread_readers', 0) + 1
            if self.sig_unread_readers == 4:
                self.sig_unread_readers, self.sig_unread = 0, False
            self.dsig_empty.arrive()
            if self.altsig or k == 2:
                (self.sigc_ready if p == 0 else self.sigf_ready)[n & 3].arrive()

        def colours(n):
            yield ('wait', self.omega_ready[n & 1], (n >> 1) & 1)
            assert n in self.composite_done and self.slot_owner[n & 1] == n
            assert self.tiles_written.get(n, set()) == {(p, k) for p in (0, 1) for k in range(3)}, f'colours({n}) before all six tiles'
            yield ('work', self.dur(1, 6))
            yield ('ebar',)
            yield ('work', self.dur(0.1, 0.5))
            yield ('ebar',)
            if e == 0:
                self.colours_done.add(n)
                self.state_free[n & 3].arrive()
        for q in range(self.T):
            td = tile_at(q, self.N)
            if not self.gcol and td[1] == 1 and td[2] == 0 and td[0] >= 1:
                yield from colours(td[0] - 1)
            yield ('wait', self.d1_full, it & 1)
            yield ('work', self.dur(0.2, 1))
            reads['d1'] += 1
            self.d1_readers = getattr(self, 'd1_readers', 0) + 1
            if self.d1_readers == K_EW:
                self.d1_readers, self.d1_unread = 0, False
            self.d1_empty.arrive()
            buf = it & 1
            yield ('wait', self.a2_empty[buf], ((it >> 1) & 1) ^ 1)
            yield ('work', self.dur(1, 4))
            self.a2_full[buf].arrive()
            if prev is not None:
                yield from sigma_read(it - 1, prev)
            prev = td; it += 1
            if td[2] == 2:
                yield from sigma_read(it - 1, prev)
                prev = None
        if prev is not None:
            yield from sigma_read(it - 1, prev)
        if not self.gcol:
            yield from colours(self.N - 1)

    def ray_warp(self, rw):
        def finalize(n):                              # fixed-order sum of the three per-team partials, out_rgb, release the group's state
            yield ('wait', self.fa_free, n & 1)
            assert n in self.colours_done and self.part_owner == n
            yield ('work', self.dur(0.05, 0.3))
            self.fin_count = getattr(self, 'fin_count', {})
            self.fin_count[n] = self.fin_count.get(n, 0) + 1
            if self.fin_count[n] == K_RW:
                self.finalized.add(n)
            self.state_free[n & 3].arrive()
        for j in range(2 * self.N):
            n, p = pass_at(j, self.N)
            si, par = n & 3, (n >> 2) & 1
            if p == 0:
                yield ('wait', self.sigc_ready[si], par)
                yield ('work', self.dur(2, 12))
                if rw == 0:
                    self.importance_done.add(n)
                self.fine_ready[si].arrive()
            else:
                yield ('wait', self.sigf_ready[si], par)
                if self.gcol and n >= 1:
                    yield from finalize(n - 1)
                if n >= 2:
                    if not self.gcol:
                        yield ('wait', self.state_free[(n - 2) & 3], ((n - 2) >> 2) & 1)
                    assert (n - 2) in self.colours_done
                prev = self.slot_owner[n & 1]
                assert prev is None or prev == n or prev in self.colours_done, f'omega slot of group {prev} rewritten for {n}'
                self.slot_owner[n & 1] = n
                yield ('work', self.dur(2, 12))
                if rw == 0:
                    self.composite_done.add(n)
                self.omega_ready[n & 1].arrive()
        if self.gcol:
            yield from finalize(self.N - 1)

    def procs(self):
        return ([('G%d' % g, self.gather_warp(g)) for g in range(12)] + [('M', self.mma_thread())] +
                [('E%d' % e, self.epilogue_warp(e)) for e in range(K_EW)] + [('R%d' % r, self.ray_warp(r)) for r in range(K_RW)])

    # ---------------------------------------------------------------- scheduler
    def run(self):
        procs = self.procs()
        state = {name: ('ready', None) for name, _ in procs}          # ready | wait(bar, parity) | sleep | ebar(gen) | done
        gens = dict(procs)

        def step(name):
            try:
                op = next(gens[name])
            except StopIteration:
                state[name] = ('done', None)
                return
            if op[0] == 'wait':
                state[name] = ('wait', (op[1], op[2]))
            elif op[0] == 'work':
                state[name] = ('sleep', None)
                self.later(op[1], lambda name=name: state.__setitem__(name, ('ready', None)))
            else:
                self.ebar_count += 1
                state[name] = ('ebar', self.ebar_gen)
                if self.ebar_count == K_EW:
                    self.ebar_count = 0
                    self.ebar_gen += 1
        guard = 0
        while True:
            guard += 1
            assert guard < 5_000_000
            runnable = []
            for name, (st, arg) in state.items():
                if st == 'ready' or (st == 'wait' and arg[0].passed(arg[1])) or (st == 'ebar' and arg != self.ebar_gen):
                    runnable.append(name)
            if runnable:
                step(self.rng.choice(runnable))
                continue
            if self.events:
                t, _, fn = heapq.heappop(self.events)
                self.now = t
                fn()
                continue
            break
        stuck = {n: s for n, s in state.items() if s[0] != 'done'}
        assert not stuck, f'deadlock with N={self.N}: {sorted(stuck)[:6]}'
        assert self.colours_done == set(range(self.N))


# ------------------------------------------------------------------------------------------------------------------
# the shipped kernel (csrc/render_fused_ws.cu): two groups in flight, per-ray phases on the epilogue warps
# ------------------------------------------------------------------------------------------------------------------
def tile_at_v3(q):                                   # render_fused_ws.cu: tile_at -  C(A) C(B) F(A) F(B) | ...
    u, j = divmod(q, 12)
    sub = j // 3
    return (2 * u + (sub & 1), sub >> 1, j - sub * 3)


class SimV3(Sim):
    """Same machinery, protocol of the default kernel: schedule C(A) C(B) F(A) F(B), colour logits in two 6-tile slots,
    importance / merge / colours inline on the eight epilogue warps (named barrier), sigma read-back deferred by one tile
    except for the odd tail group, fine_ready / state_free with one arrival."""

    def __init__(self, N, seed):
        super().__init__(N, seed)
        self.T = 12 * ((N + 1) // 2)
        self.fine_ready = [Bar(1) for _ in range(4)]
        self.state_free = [Bar(1) for _ in range(4)]        # released by epilogue warp 0 after colours(n)
        self.altsig, self.d2_full = False, Bar(1)
        self.sigc_ready = [Bar(4) for _ in range(4)]; self.sigf_ready = [Bar(4) for _ in range(4)]
        self.finalized = self.colours_done                   # that kernel has no separate finalize step

    def tiles(self):
        return [t for t in (tile_at_v3(q) for q in range(self.T)) if t[0] < self.N]

    def is_tail(self, td):
        return (self.N & 1) and td[0] == self.N - 1 and td[1] == 0 and td[2] == 2

    def gather_warp(self, gw):
        team = gw >> 2
        for my_it, (n, p, k) in enumerate(self.tiles()):
            if my_it % K_TEAMS != team:
                continue
            if p == 0:
                yield ('wait', self.state_free[n & 3], ((n >> 2) & 1) ^ 1)
                prev = self.state_owner[n & 3]
                assert prev is None or prev == n or prev in self.finalized, f'state slot of group {prev} rewritten for {n}'
                self.state_owner[n & 3] = n
            else:
                yield ('wait', self.fine_ready[n & 3], (n >> 2) & 1)
                assert n in self.importance_done, f'fine gather of {n} before its importance sampling'
            stage = my_it % K_NA
            yield ('wait', self.a1_empty[stage], ((my_it // K_NA) & 1) ^ 1)
            yield ('work', self.dur(5, 20))
            self.a1_full[stage].arrive()

    def mma_thread(self):
        it, prev = 0, None

        def layer2(it2, tp):
            n, p, k = tp
            buf = it2 & 1
            yield ('wait', self.a2_full[buf], (it2 >> 1) & 1)
            yield ('wait', self.dsig_empty, (it2 & 1) ^ 1)
            area = (n & 1, p * 3 + k)
            owner = self.area_owner.get(area)
            assert owner is None or owner in self.colours_done, f'layer 2 of {tp} overwrites slot tile {area} of group {owner} before its colours'
            assert not self.sig_unread, 'sigma accumulator overwritten before it was read'
            self.area_owner[area] = n

            def done():
                self.tiles_written.setdefault(n, set()).add((p, k))
                self.sig_unread = True
                self.d2_full.arrive(); self.a2_empty[buf].arrive()
            self.later(self.dur(0.2, 2), done)
        for td in self.tiles():
            stage = it % K_NA
            yield ('wait', self.a1_full[stage], (it // K_NA) & 1)
            yield ('wait', self.d1_empty, (it & 1) ^ 1)
            assert not self.d1_unread, 'D1 overwritten before the epilogue read it'

            def done1(stage=stage):
                self.d1_unread = True
                self.d1_full.arrive(); self.a1_empty[stage].arrive()
            self.later(self.dur(0.2, 2), done1)
            if prev is not None:
                yield from layer2(it - 1, prev)
            prev = td; it += 1
            if self.is_tail(td):
                yield from layer2(it - 1, prev)
                prev = None
        if prev is not None:
            yield from layer2(it - 1, prev)

    def epilogue_warp(self, e):
        chunk, it, prev = e >> 2, 0, None

        def sigma_read(it_prev, tp):
            if chunk != 0:
                return
            yield ('wait', self.d2_full, it_prev & 1)
            yield ('work', self.dur(0.2, 1))
            self.sig_unread_readers = getattr(self, 'sig_unread_readers', 0) + 1
            if self.sig_unread_readers == 4:
                self.sig_unread_readers, self.sig_unread = 0, False
            self.dsig_empty.arrive()

        def after_sigma(tp):
            n, p, k = tp
            if k != 2:
                return
            if p == 0:                                                   # importance(n)
                yield ('ebar',)
                yield ('work', self.dur(2, 10))
                yield ('ebar',)
                if e == 0:
                    self.importance_done.add(n)
                    self.fine_ready[n & 3].arrive()
            else:                                                        # composite(n); colours(n)
                yield ('ebar',)
                yield ('work', self.dur(2, 10))
                yield ('ebar',)
                assert self.tiles_written.get(n, set()) == {(pp, kk) for pp in (0, 1) for kk in range(3)}, f'colours({n}) before all six tiles'
                yield ('work', self.dur(1, 6))
                yield ('ebar',)
                yield ('work', self.dur(0.1, 0.5))
                yield ('ebar',)
                if e == 0:
                    self.colours_done.add(n)
                    self.state_free[n & 3].arrive()
        for td in self.tiles():
            yield ('wait', self.d1_full, it & 1)
            yield ('work', self.dur(0.2, 1))
            self.d1_readers = getattr(self, 'd1_readers', 0) + 1
            if self.d1_readers == K_EW:
                self.d1_readers, self.d1_unread = 0, False
            self.d1_empty.arrive()
            buf = it & 1
            yield ('wait', self.a2_empty[buf], ((it >> 1) & 1) ^ 1)
            yield ('work', self.dur(1, 4))
            self.a2_full[buf].arrive()
            if prev is not None:
                yield from sigma_read(it - 1, prev)
                yield from after_sigma(prev)
            prev = td; it += 1
            if self.is_tail(td):
                yield from sigma_read(it - 1, prev)
                yield from after_sigma(prev)
                prev = None
        if prev is not None:
            yield from sigma_read(it - 1, prev)
            yield from after_sigma(prev)

    def procs(self):
        return ([('G%d' % g, self.gather_warp(g)) for g in range(12)] + [('M', self.mma_thread())] +
                [('E%d' % e, self.epilogue_warp(e)) for e in range(K_EW)])


def test_pass_schedule_visits_every_group_once_in_a_valid_order():
    for N in range(1, 12):
        seq = [pass_at(j, N) for j in range(2 * N)]
        assert sorted(seq) == sorted([(n, p) for n in range(N) for p in (0, 1)])
        pos = {np_: i for i, np_ in enumerate(seq)}
        for n in range(N):
            assert pos[(n, 0)] < pos[(n, 1)]                                       # F(n) after C(n)
            if n + 1 < N:
                assert pos[(n, 1)] < pos[(n + 1, 1)] and pos[(n, 0)] < pos[(n + 1, 0)]
            if n + 3 < N:
                assert pos[(n + 1, 1)] < pos[(n + 3, 0)]                           # CA[n % 3] is free again before C(n+3)
            if N > 2 and 1 <= n and n + 1 < N:
                assert pos[(n + 1, 0)] < pos[(n, 1)]                               # >= one pass of slack for importance(n)


@pytest.mark.parametrize('N', [1, 2, 3, 4, 5, 7, 8, 13])
def test_protocol_has_no_deadlock_or_hazard(N):
    for seed in range(12):
        Sim(N, seed * 101 + N).run()


@pytest.mark.parametrize('N', [1, 2, 3, 5, 8])
def test_alternating_sigma_read_back_variant_has_no_deadlock_or_hazard(N):
    """P3D_W3_ALTSIG=1 of the experiments file: the sigma read-back alternates between the column halves of the epilogue role;
    d2_full is one barrier per tile parity (a single barrier would alias phases for a half that skips every other one - the
    sig_tile assertion in sigma_read is what catches that)."""
    for seed in range(6):
        Sim(N, seed * 17 + N, altsig=True).run()


@pytest.mark.parametrize('N', [1, 2, 3, 4, 5, 8, 11])
def test_gather_side_colour_share_variant_has_no_deadlock_or_hazard(N):
    """P3D_W3_GCOL=1 of the experiments file (colour reduction on the gather warps, finalize on the ray warps, fa_free hand-off to
    the MMA thread)."""
    for seed in range(8):
        Sim(N, seed * 131 + N, gcol=True).run()


@pytest.mark.parametrize('N', [1, 2, 3, 4, 5, 6, 7, 9, 12])
def test_shipped_kernel_protocol_has_no_deadlock_or_hazard(N):
    """The default kernel's hand-offs (validated on hardware for the sizes the GPU tests use) for every small group count,
    odd and even - the odd tail group takes a different path through the sigma read-back."""
    for seed in range(10):
        SimV3(N, seed * 77 + N).run()


def test_shipped_kernel_needs_its_odd_tail_flush():
    """Sanity of the simulator itself: without the immediate sigma flush for an odd tail group the protocol deadlocks
    (the group's fine tiles wait for importance depths that wait for a sigma nobody reads)."""
    class NoFlush(SimV3):
        def is_tail(self, td):
            return False
    with pytest.raises(AssertionError, match='deadlock'):
        NoFlush(3, 1).run()
