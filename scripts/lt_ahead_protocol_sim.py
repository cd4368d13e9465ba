"""Discrete-event model of the mbarrier / TMEM-buffer protocol of lt_attn_tc3_kernel (the "ahead" layout, csrc/lt_attn_tc.cu).

No GPU needed.  Three agents follow the kernel source statement by statement -- the TMA producer, the MMA issuer (whose
tcgen05 ops complete IN ORDER after a latency, tcgen05.commit arriving on an mbarrier once everything issued before it has
completed) and the 16 softmax warps (modelled as one agent: they move in lockstep through `bar.sync 1, 512`) -- with
phase-accurate mbarriers (`try_wait.parity P` succeeds iff the barrier's current phase parity != P).  Random latencies are
drawn per run.  Checked for T = 1..12 key tiles per CTA:
  * no deadlock (every agent terminates);
  * every score buffer is read by the softmax only when it holds the tile the softmax expects, and is overwritten by S(n+3)
    only after PV(n) consumed its P;
  * a K/V stage is re-loaded only after both PVs of its key tile completed;
  * the O_i rescale and the epilogue read O_i only when every PV into it has completed.
Run: python scripts/lt_ahead_protocol_sim.py            (exit code 0 = all schedules fine)
"""
import heapq
import random
import sys

STAGES = 4


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0
        self.waiters = []

    def arrive(self, sim, n=1):
        self.pending -= n
        assert self.pending >= 0, "too many arrivals in one phase"
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count
            w, self.waiters = self.waiters, []
            for agent in w:
                sim.wake(agent)

    def test(self, parity):            # try_wait.parity: true iff the phase with this parity has completed
        return (self.phase & 1) != parity


class Sim:
    def __init__(self, T, seed):
        self.T, self.nT = T, 2 * T
        self.rng = random.Random(seed)
        self.now = 0.0
        self.events = []               # (time, seq, callable)
        self.seq = 0
        self.q_full = MBar(1)
        self.kv_full = [MBar(1) for _ in range(STAGES)]
        self.kv_free = [MBar(1) for _ in range(STAGES)]
        self.s_full = [MBar(1) for _ in range(3)]
        self.p_full = [MBar(1) for _ in range(3)]      # 512 thread arrivals modelled as one (lockstep agent)
        self.o_done = [MBar(1) for _ in range(2)]
        self.o_final = [MBar(1) for _ in range(2)]
        self.pipe_free_at = 0.0        # in-order tensor pipe
        self.sbuf = [None, None, None]  # what each score buffer holds: ("S", n) | ("P", n) | None
        self.kv_stage = [None] * STAGES  # key tile resident in each smem stage
        self.pv_done = set()
        self.s_done = set()
        self.done = set()
        self.agents = {}

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
        kind = req[0]
        if kind == "delay":
            self.at(self.now + req[1], lambda: self.step(agent))
        elif kind == "wait":
            bar, parity = req[1], req[2]
            if bar.test(parity):
                self.at(self.now, lambda: self.step(agent))
            else:
                bar.waiters.append(agent)
                self.agents[agent] = self._rewait(gen, bar, parity)

    def _rewait(self, gen, bar, parity):
        # re-test after wake-up (another phase flip may be needed)
        while not bar.test(parity):
            yield ("wait", bar, parity)
        yield from gen

    # ---- tensor pipe: ops complete in order; commit = arrive when everything issued before has completed
    def mma(self, dur, on_complete):
        start = max(self.now, self.pipe_free_at)
        self.pipe_free_at = start + dur
        self.at(self.pipe_free_at, on_complete)

    def commit(self, bar):
        self.at(max(self.now, self.pipe_free_at) + 1e-6, lambda: bar.arrive(self))

    # ---- agents -----------------------------------------------------------------------------------------------
    def tma(self):
        yield ("delay", self.rng.uniform(50, 400))
        self.q_full.arrive(self)
        for j in range(self.T):
            s = j % STAGES
            if j >= STAGES:
                yield ("wait", self.kv_free[s], ((j // STAGES) - 1) & 1)
                old = self.kv_stage[s]
                assert (2 * old in self.pv_done) and (2 * old + 1 in self.pv_done), f"K/V stage {s} reloaded before PV({old}) done"
            lat = self.rng.uniform(300, 3000)
            jj = j

            def landed(jj=jj, s=s):
                self.kv_stage[s] = jj
                self.kv_full[s].arrive(self)
            self.at(self.now + lat, landed)
            yield ("delay", 10)

    def issuer(self):
        T, nT = self.T, self.nT
        kv_ready = -1
        yield ("wait", self.q_full, 0)

        def issue_S(n):
            nonlocal kv_ready
            j, s, b = n >> 1, (n >> 1) % STAGES, n % 3
            if j > kv_ready:
                yield ("wait", self.kv_full[s], (j // STAGES) & 1)
                kv_ready = j
            assert self.kv_stage[s] == j, f"S({n}) reads stage {s} holding tile {self.kv_stage[s]}"

            def done(n=n, b=b):
                prev = self.sbuf[b]
                assert prev is None or (prev[0] == "Pused" and prev[1] == n - 3), f"S({n}) overwrote {prev} in buffer {b}"
                self.sbuf[b] = ("S", n)
                self.s_done.add(n)
            self.mma(self.rng.uniform(300, 450), done)
            self.commit(self.s_full[b])
            yield ("delay", 5)

        def issue_PV(n):
            i, j, s, b = n & 1, n >> 1, (n >> 1) % STAGES, n % 3
            assert self.sbuf[b] == ("P", n), f"PV({n}) found {self.sbuf[b]} in buffer {b}"
            assert self.kv_stage[s] == j

            def done(n=n, b=b):
                self.sbuf[b] = ("Pused", n)
                self.pv_done.add(n)
            # occasionally a very slow PV: exposes waits that only hold when the tensor pipe keeps up
            self.mma(self.rng.uniform(450, 600) * (8 if self.rng.random() < 0.15 else 1), done)
            self.commit(self.o_done[i])
            if n + 2 >= nT:
                self.commit(self.o_final[i])
            if i == 1:
                self.commit(self.kv_free[s])
            yield ("delay", 5)

        for n in range(min(3, nT)):
            yield from issue_S(n)
        for n in range(nT):
            yield ("wait", self.p_full[n % 3], (n // 3) & 1)
            yield from issue_PV(n)
            if n + 3 < nT:
                yield from issue_S(n + 3)

    def softmax(self):
        T, nT = self.T, self.nT
        if T > 0:
            yield ("wait", self.s_full[0], 0)
            assert self.sbuf[0] == ("S", 0)
            yield ("delay", self.rng.uniform(200, 1100))       # tcgen05.ld of tile 0
        regs = 0                                               # tile whose scores are in the current register set
        b, par = 0, 1
        for j in range(T):
            for i in (0, 1):
                n = 2 * j + i
                bn = 0 if b == 2 else b + 1
                has_next = (i == 0) or (j + 1 < T)
                next_par = (par >> bn) & 1
                assert regs == n, f"register set holds tile {regs}, expected {n}"
                yield ("delay", self.rng.uniform(100, 400))    # max + bar.sync + exchange
                if j > 0 and self.rng.random() < 0.3:          # the rescale path
                    yield ("wait", self.o_done[i], (j - 1) & 1)
                    assert (n - 2) in self.pv_done, f"O_{i} rescaled before PV({n - 2}) completed"
                    assert n not in self.pv_done
                    yield ("delay", 60)
                if has_next:
                    yield ("wait", self.s_full[bn], next_par)
                    assert self.sbuf[bn] == ("S", n + 1), f"prefetch of tile {n + 1} found {self.sbuf[bn]} in buffer {bn}"
                    regs_next = n + 1
                yield ("delay", self.rng.uniform(600, 1200))   # ex2 pass + tcgen05.st of P (hi and lo) over S(n)
                assert self.sbuf[b] == ("S", n), f"P({n}) written over {self.sbuf[b]}"
                self.sbuf[b] = ("P", n)
                self.p_full[b].arrive(self)
                if has_next:
                    assert self.sbuf[bn] == ("S", n + 1)        # still intact when the loads complete
                    regs = regs_next
                par ^= 1 << bn
                b = bn
        for i in (0, 1):
            if T > 0:
                yield ("wait", self.o_final[i], 0)
                assert all((2 * j + i) in self.pv_done for j in range(T)), f"epilogue read O_{i} early"

    def run(self):
        self.agents = {"tma": self.tma(), "mma": self.issuer(), "softmax": self.softmax()}
        for a in list(self.agents):
            self.wake(a)
        steps = 0
        while self.events:
            t, _, fn = heapq.heappop(self.events)
            self.now = max(self.now, t)
            fn()
            steps += 1
            assert steps < 2_000_000, "runaway"
        missing = set(self.agents) - self.done
        assert not missing, f"deadlock: {sorted(missing)} never finished (T = {self.T})"
        assert self.pv_done == set(range(self.nT))


def main():
    n = 0
    for T in range(0, 13):
        for seed in range(200):
            Sim(T, seed * 7919 + T).run()
            n += 1
    print(f"lt_attn_tc3 protocol model: {n} randomised schedules (T = 0..12), no deadlock, no buffer hazard")


if __name__ == "__main__":
    sys.exit(main())
