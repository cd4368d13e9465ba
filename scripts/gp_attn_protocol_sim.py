"""Discrete-event model of the mbarrier / TMEM / shared-memory-ring protocol of gp_attn_tc_kernel (csrc/gp_attn_tc.cu, the fused
DeAOT long-term attention).  Same machinery as scripts/lt_ahead_protocol_sim.py; agents: K producer, V producer, MMA issuer
(in-order tensor pipe, tcgen05.commit semantics), the 16 softmax warps as one lockstep agent.  Checks for T = 0..14 key tiles:
no deadlock; every score buffer read / overwritten in the right state; K and V stages re-loaded only after their consumer
completed and read only while they hold the expected tile; O' rescaled / read only when every earlier PV has completed.
Run: python scripts/gp_attn_protocol_sim.py"""
import heapq
import random
import sys


class MBar:
    def __init__(self, count=1):
        self.count, self.pending, self.phase, self.waiters = count, count, 0, []

    def arrive(self, sim):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count
            w, self.waiters = self.waiters, []
            for agent in w:
                sim.wake(agent)

    def test(self, parity):
        return (self.phase & 1) != parity


class Sim:
    def __init__(self, T, seed):
        self.T = T
        self.rng = random.Random(seed)
        self.now, self.events, self.seq = 0.0, [], 0
        self.q_full = MBar()
        self.k_full, self.k_free = [MBar(), MBar()], [MBar(), MBar()]
        self.v_full, self.v_free = [MBar(), MBar()], [MBar(), MBar()]
        self.s_full, self.p_full = [MBar() for _ in range(3)], [MBar() for _ in range(3)]
        self.o_done = [MBar(), MBar()]
        self.o_final = MBar()
        self.pipe_free_at = 0.0
        self.sbuf = [None] * 3
        self.kst, self.vst = [None, None], [None, None]
        self.s_done, self.pv_done, self.done, self.agents = set(), set(), set(), {}

    def at(self, t, fn):
        self.seq += 1
        heapq.heappush(self.events, (t, self.seq, fn))

    def wake(self, agent):
        self.at(self.now, lambda: self.step(agent))

    def step(self, agent):
        gen = self.agents[agent]
        try:
            req = next(gen)
        except StopIteration:
            self.done.add(agent)
            return
        if req[0] == "delay":
            self.at(self.now + req[1], lambda: self.step(agent))
        else:
            bar, parity = req[1], req[2]
            if bar.test(parity):
                self.at(self.now, lambda: self.step(agent))
            else:
                bar.waiters.append(agent)
                self.agents[agent] = self._rewait(gen, bar, parity)

    def _rewait(self, gen, bar, parity):
        while not bar.test(parity):
            yield ("wait", bar, parity)
        yield from gen

    def mma(self, dur, on_complete):
        start = max(self.now, self.pipe_free_at)
        self.pipe_free_at = start + dur
        self.at(self.pipe_free_at, on_complete)

    def commit(self, bar):
        self.at(max(self.now, self.pipe_free_at) + 1e-6, lambda: bar.arrive(self))

    # ---- agents
    def producer(self, which):
        full, free, st = (self.k_full, self.k_free, self.kst) if which == "K" else (self.v_full, self.v_free, self.vst)
        if which == "K":
            yield ("delay", self.rng.uniform(50, 400))
            self.q_full.arrive(self)
        for j in range(self.T):
            s = j & 1
            if j >= 2:
                yield ("wait", free[s], ((j >> 1) - 1) & 1)
                consumed = self.s_done if which == "K" else self.pv_done
                assert (j - 2) in consumed, f"{which} stage {s} reloaded before its consumer of tile {j - 2} completed"

            def landed(j=j, s=s):
                st[s] = j
                full[s].arrive(self)
            self.at(self.now + self.rng.uniform(300, 3000), landed)
            yield ("delay", 10)

    def issuer(self):
        T = self.T
        yield ("wait", self.q_full, 0)

        def issue_S(n):
            s, b = n & 1, n % 3
            yield ("wait", self.k_full[s], (n >> 1) & 1)
            assert self.kst[s] == n, f"S({n}) reads K stage {s} holding tile {self.kst[s]}"

            def done(n=n, b=b, s=s):
                assert self.kst[s] == n, f"K stage {s} overwritten while S({n}) was executing"
                prev = self.sbuf[b]
                assert prev is None or prev == ("Pused", n - 3), f"S({n}) overwrote {prev}"
                self.sbuf[b] = ("S", n)
                self.s_done.add(n)
            self.mma(self.rng.uniform(600, 900), done)
            self.commit(self.k_free[s])
            self.commit(self.s_full[b])
            yield ("delay", 5)

        def issue_PV(n):
            s, b = n & 1, n % 3
            yield ("wait", self.v_full[s], (n >> 1) & 1)
            assert self.vst[s] == n, f"PV({n}) reads V stage {s} holding tile {self.vst[s]}"
            assert self.sbuf[b] == ("P", n), f"PV({n}) found {self.sbuf[b]}"

            def done(n=n, b=b, s=s):
                assert self.vst[s] == n, f"V stage {s} overwritten while PV({n}) was executing"
                self.sbuf[b] = ("Pused", n)
                self.pv_done.add(n)
            # occasionally a very slow PV: exposes waits that only hold when the tensor pipe keeps up
            self.mma(self.rng.uniform(800, 1200) * (8 if self.rng.random() < 0.15 else 1), done)
            self.commit(self.v_free[s])
            self.commit(self.o_done[n & 1])
            if n + 1 == T:
                self.commit(self.o_final)
            yield ("delay", 5)

        for n in range(min(3, T)):
            yield from issue_S(n)
        for n in range(T):
            yield ("wait", self.p_full[n % 3], (n // 3) & 1)
            yield from issue_PV(n)
            if n + 3 < T:
                yield from issue_S(n + 3)

    def softmax(self):
        T = self.T
        if T > 0:
            yield ("wait", self.s_full[0], 0)
            assert self.sbuf[0] == ("S", 0)
            yield ("delay", self.rng.uniform(100, 600))
        regs, b, par = 0, 0, 1
        for n in range(T):
            bn = 0 if b == 2 else b + 1
            has_next = n + 1 < T
            next_par = (par >> bn) & 1
            assert regs == n
            yield ("delay", self.rng.uniform(100, 400))
            if n > 0 and self.rng.random() < 0.3:
                yield ("wait", self.o_done[(n - 1) & 1], ((n - 1) >> 1) & 1)
                assert all(k in self.pv_done for k in range(n)), f"O' rescaled at tile {n} before PV({n - 1}) completed"
                assert n not in self.pv_done
                yield ("delay", 120)
            if has_next:
                yield ("wait", self.s_full[bn], next_par)
                assert self.sbuf[bn] == ("S", n + 1), f"prefetch of tile {n + 1} found {self.sbuf[bn]}"
            yield ("delay", self.rng.uniform(300, 900))
            assert self.sbuf[b] == ("S", n)
            self.sbuf[b] = ("P", n)
            self.p_full[b].arrive(self)
            if has_next:
                assert self.sbuf[bn] == ("S", n + 1)
                regs = n + 1
            par ^= 1 << bn
            b = bn
        if T > 0:
            yield ("wait", self.o_final, 0)
            assert self.pv_done == set(range(T)), "epilogue read O' early"

    def run(self):
        self.agents = {"K": self.producer("K"), "V": self.producer("V"), "mma": self.issuer(), "softmax": self.softmax()}
        for a in list(self.agents):
            self.wake(a)
        steps = 0
        while self.events:
            t, _, fn = heapq.heappop(self.events)
            self.now = max(self.now, t)
            fn()
            steps += 1
            assert steps < 2_000_000
        missing = set(self.agents) - self.done
        assert not missing, f"deadlock: {sorted(missing)} never finished (T = {self.T})"


def main():
    n = 0
    for T in range(0, 15):
        for seed in range(200):
            Sim(T, seed * 104729 + T).run()
            n += 1
    print(f"gp_attn_tc protocol model: {n} randomised schedules (T = 0..14), no deadlock, no buffer / stage hazard")


if __name__ == "__main__":
    sys.exit(main())
