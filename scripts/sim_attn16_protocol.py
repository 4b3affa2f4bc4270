"""Discrete-event model of the mbarrier protocol of pb_attn16_sm100.cu (producer, two score-MMA warps, two accumulate-MMA
warps, eight compute warps) -- finds dead-locks and phase aliasing over (KC, nj, ring depths, role) without a GPU.

    python scripts/sim_attn16_protocol.py            # sweep
Each agent is a generator that yields ("wait", bar, parity) | ("arrive", bar) | ("commit", [bars]) | ("tma", bar) | ("work", clocks).
An mbarrier completes a phase after `count` arrivals (+ its pending TMA transactions); tcgen05.commit arrives after the
MMAs issued so far by that agent are done (modelled as a delay)."""
import heapq, itertools, sys


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase, self.tx = name, count, count, 0, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"over-arrival on {self.name}"
        self._check()

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def done(self, parity):            # try_wait.parity: true once the phase with this parity has completed
        return (self.phase & 1) != parity


class Ring:
    def __init__(self, idx=0):
        self.idx, self.ph = idx, 0

    def next(self, n):
        self.idx += 1
        if self.idx == n:
            self.idx, self.ph = 0, self.ph ^ 1

    def next2(self, n):
        self.idx += 2
        if self.idx >= n:
            self.idx, self.ph = self.idx - n, self.ph ^ 1


def simulate(KC, nj, NS, NSH, NPC, NT, has_c2, nbpc, tma_lat=1500, mma_lat=300, verbose=False, skew=None):
    skew = skew or {}                                             # agent name -> multiplier of its 'work' durations
    LAG = 2 if has_c2 else 0
    nu = nj * KC
    pc_cnt = (1 if nbpc else 0) + (1 if has_c2 else 0)
    has_pc = pc_cnt > 0
    B = dict(a_full=Bar("a_full", 1), acc_full=Bar("acc_full", 2), acc_zeroed=Bar("acc_zeroed", 8))
    for i in range(NSH):
        B[f"sh_full{i}"] = Bar(f"sh_full{i}", 1); B[f"sh_empty{i}"] = Bar(f"sh_empty{i}", 12)
    for i in range(max(NPC, 1)):
        B[f"pc_full{i}"] = Bar(f"pc_full{i}", 1); B[f"pc_empty{i}"] = Bar(f"pc_empty{i}", max(pc_cnt, 1))
    for i in range(NS):
        B[f"s_full{i}"] = Bar(f"s_full{i}", 1); B[f"s_free{i}"] = Bar(f"s_free{i}", 4)
    for i in range(NT):
        B[f"t_full{i}"] = Bar(f"t_full{i}", 4); B[f"t_empty{i}"] = Bar(f"t_empty{i}", 1)
    log = []

    def producer():
        yield ("tma", "a_full")
        rsh, rpc = Ring(), Ring()
        msh = mpc = 0
        while msh < nj or (has_pc and mpc < nu):
            any_ = False
            if has_pc and mpc < nu and B[f"pc_empty{rpc.idx}"].done(rpc.ph ^ 1):
                yield ("tma", f"pc_full{rpc.idx}")
                rpc.next(NPC); mpc += 1; any_ = True
            if msh < nj and B[f"sh_empty{rsh.idx}"].done(rsh.ph ^ 1):
                yield ("tma", f"sh_full{rsh.idx}")
                rsh.next(NSH); msh += 1; any_ = True
            if not any_:
                yield ("work", 50)

    def score(w):
        rsh, rpc, rs = Ring(), Ring(w), Ring(w)
        j, c = w // KC, w % KC
        jr, jdone, jw = 0, -1, -1
        yield ("wait", "a_full", 0)
        u = w
        while u < nu:
            while jr < j:
                if jdone != jr:
                    yield ("wait", f"sh_full{rsh.idx}", rsh.ph)
                    yield ("arrive", f"sh_empty{rsh.idx}")
                rsh.next(NSH); jr += 1
            if jw != j:
                yield ("wait", f"sh_full{rsh.idx}", rsh.ph); jw = j
            if nbpc:
                yield ("wait", f"pc_full{rpc.idx}", rpc.ph)
            yield ("wait", f"s_free{rs.idx}", rs.ph ^ 1)
            last = c + 2 >= KC
            bars = ([f"pc_empty{rpc.idx}"] if nbpc else []) + [f"s_full{rs.idx}"] + ([f"sh_empty{rsh.idx}"] if last else [])
            yield ("commit", bars)
            yield ("work", 200)
            if last:
                jdone = j
            if nbpc:
                rpc.next2(NPC)
            rs.next2(NS)
            c += 2
            while c >= KC:
                c -= KC; j += 1
            u += 2
        while jr < nj:
            if jdone != jr:
                yield ("wait", f"sh_full{rsh.idx}", rsh.ph)
                yield ("arrive", f"sh_empty{rsh.idx}")
            rsh.next(NSH); jr += 1

    def accum(w):
        rshp, rsha, rpc, rt = Ring(), Ring(), Ring(w), Ring(w)
        jp, cp, jrp, jwp = w // KC, w % KC, 0, -1
        ja, ca, jra, jwa, jdone = w // KC, w % KC, 0, -1, -1
        yield ("wait", "acc_zeroed", 0)
        u = w
        while u < nu + LAG:
            if u >= LAG:
                while jra < ja:
                    if jdone != jra:
                        yield ("wait", f"sh_full{rsha.idx}", rsha.ph)
                        yield ("arrive", f"sh_empty{rsha.idx}")
                    rsha.next(NSH); jra += 1
            if has_c2 and u < nu:
                while jrp < jp:
                    rshp.next(NSH); jrp += 1
                if jwp != jp:
                    yield ("wait", f"sh_full{rshp.idx}", rshp.ph); jwp = jp
                yield ("wait", f"pc_full{rpc.idx}", rpc.ph)
                yield ("commit", [f"pc_empty{rpc.idx}"])
                yield ("work", 100)
                rpc.next2(NPC)
                cp += 2
                while cp >= KC:
                    cp -= KC; jp += 1
            if u >= LAG:
                if jwa != ja:
                    yield ("wait", f"sh_full{rsha.idx}", rsha.ph); jwa = ja
                yield ("wait", f"t_full{rt.idx}", rt.ph)
                last = ca + 2 >= KC
                yield ("commit", [f"t_empty{rt.idx}"] + ([f"sh_empty{rsha.idx}"] if last else []))
                yield ("work", 100)
                if last:
                    jdone = ja
                rt.next2(NT)
                ca += 2
                while ca >= KC:
                    ca -= KC; ja += 1
            u += 2
        while jra < nj:
            if jdone != jra:
                yield ("wait", f"sh_full{rsha.idx}", rsha.ph)
                yield ("arrive", f"sh_empty{rsha.idx}")
            rsha.next(NSH); jra += 1
        yield ("commit", ["acc_full"])

    def compute(grp, q):
        yield ("arrive", "acc_zeroed")
        rsh, rs, rt = Ring(), Ring(grp), Ring(grp)
        jn = 0
        c, j = grp % KC, grp // KC
        u = grp
        while u < nu:
            while jn <= j:
                yield ("wait", f"sh_full{rsh.idx}", rsh.ph)
                yield ("arrive", f"sh_empty{rsh.idx}")
                rsh.next(NSH); jn += 1
            yield ("wait", f"s_full{rs.idx}", rs.ph)
            yield ("work", 150)
            yield ("arrive", f"s_free{rs.idx}")
            yield ("work", 250)
            yield ("wait", f"t_empty{rt.idx}", rt.ph ^ 1)
            yield ("work", 100)
            yield ("arrive", f"t_full{rt.idx}")
            rs.next2(NS); rt.next2(NT)
            c += 2
            while c >= KC:
                c -= KC; j += 1
            u += 2
        while jn < nj:
            yield ("wait", f"sh_full{rsh.idx}", rsh.ph)
            yield ("arrive", f"sh_empty{rsh.idx}")
            rsh.next(NSH); jn += 1
        yield ("wait", "acc_full", 0)

    agents = {"prod": producer(), "s0": score(0), "s1": score(1), "a0": accum(0), "a1": accum(1)}
    for g in range(2):
        for q in range(4):
            agents[f"c{g}{q}"] = compute(g, q)
    now = 0
    ready = [(0, i, name) for i, name in enumerate(agents)]      # (time, tiebreak, agent)
    heapq.heapify(ready)
    seq = itertools.count(len(agents))
    blocked = {}                                                  # agent -> (bar, parity)
    events = []                                                   # (time, seq, fn)
    finished = set()

    def wake():
        for name, (bar, par) in list(blocked.items()):
            if B[bar].done(par):
                del blocked[name]
                heapq.heappush(ready, (now, next(seq), name))

    while ready or events:
        if now > 30_000_000:                                      # the polling producer never blocks: cap simulated time
            break
        if ready and (not events or ready[0][0] <= events[0][0]):
            now, _, name = heapq.heappop(ready)
            try:
                op = next(agents[name])
            except StopIteration:
                finished.add(name)
                continue
            kind = op[0]
            if kind == "wait":
                if B[op[1]].done(op[2]):
                    heapq.heappush(ready, (now + 20, next(seq), name))
                else:
                    blocked[name] = (op[1], op[2])
            elif kind == "arrive":
                B[op[1]].arrive(); wake()
                heapq.heappush(ready, (now + 5, next(seq), name))
            elif kind == "commit":
                bars = op[1]
                def fire(bars=bars):
                    for b in bars:
                        B[b].arrive()
                heapq.heappush(events, (now + mma_lat, next(seq), fire))
                heapq.heappush(ready, (now + 10, next(seq), name))
            elif kind == "tma":
                bar = B[op[1]]
                bar.tx += 1
                bar.pending -= 1                                  # arrive.expect_tx
                def land(bar=bar):
                    bar.tx -= 1
                    bar._check()
                heapq.heappush(events, (now + tma_lat, next(seq), land))
                heapq.heappush(ready, (now + 10, next(seq), name))
            elif kind == "work":
                heapq.heappush(ready, (now + int(op[1] * skew.get(name, skew.get(name[0], 1))), next(seq), name))
        else:
            now, _, fn = heapq.heappop(events)
            fn(); wake()
    ok = len(finished) == len(agents)
    if not ok and verbose:
        print("  blocked:", {k: v for k, v in blocked.items()})
    return ok, now


SKEWS = [None, {"c1": 12}, {"c0": 12}, {"s1": 8}, {"a0": 8, "c1": 5}, {"p": 20}, {"c00": 30}]


if __name__ == "__main__":
    bad = 0
    for KC in range(1, 9):
        for has_c2, nbpc in ((True, 1), (False, 0), (True, 0), (False, 1)):
            kmin_ok = lambda nsh: (not has_c2) or KC * (nsh - 1) >= 3
            for NSH in (2, 3, 4):
                if not kmin_ok(NSH):
                    continue
                for NS in (2, 3, 4):
                    for NPC in ((2, 3, 4, 5, 6) if (nbpc or has_c2) else (0,)):
                        for NT in (2, 3, 4):
                            for nj, sk in itertools.product((1, 2, 3, 7, 16), SKEWS):
                                if (NS | NT | NPC) & 1:
                                    continue                      # rings shared by the two parity classes must be even (phase aliasing)
                                ok, t = simulate(KC, nj, NS, NSH, NPC, NT, has_c2, nbpc, skew=sk)
                                if not ok:
                                    bad += 1
                                    if bad < 15:
                                        print("DEADLOCK", dict(KC=KC, nj=nj, NS=NS, NSH=NSH, NPC=NPC, NT=NT, has_c2=has_c2, nbpc=nbpc))
                                        simulate(KC, nj, NS, NSH, NPC, NT, has_c2, nbpc, verbose=True)
    print("dead-locking configurations:", bad)
    sys.exit(1 if bad else 0)
