"""Discrete-event model of the barrier protocol of csrc/icnv_smooth_banded.cu (developer aid, CPU only).

Warps are coroutines that yield barrier operations; a random scheduler interleaves them.  Named barriers follow the
PTX rules (a generation completes when `count` participants — arrivals and syncs alike — have shown up; syncs block
until then).  The model asserts the memory-ordering claims of the kernel's header comment:
  * a store into range X of pair i happens after every owner of band X has read range X of pair i-1;
  * the owners of band X read range X of pair i only after every unit of band X of pair i has been stored;
  * the staged rows of pair i+1 are issued only after every unit of pair i has been gathered;
and that no run dead-locks.      python tools/banded_protocol_model.py [n_seeds]
"""
import random
import sys

NW = 16  # warps per CTA; 0-7 own band A, 8-15 band B


class Barrier:
    def __init__(self, count):
        self.count, self.arrived, self.waiting, self.generation = count, 0, [], 0

    def arrive(self):
        self.arrived += 1
        self._maybe_complete()

    def sync(self, w):
        self.arrived += 1
        self.waiting.append(w)
        self._maybe_complete()

    def _maybe_complete(self):
        assert self.arrived <= self.count, "more participants than the barrier expects"
        if self.arrived == self.count:
            self.arrived, released, self.waiting = 0, self.waiting, []
            self.generation += 1
            return released
        return []


def run(seed, n_iter=5, units=(8, 8)):
    rng = random.Random(seed)
    bars = {0: Barrier(NW), 1: Barrier(NW), 2: Barrier(NW // 2), 3: Barrier(NW), 4: Barrier(NW // 2), 5: Barrier(NW)}
    counters = [[0, 0], [0, 0]]
    stored = [[0, 0] for _ in range(n_iter)]          # units stored per (iteration, band)
    read_done = [[0, 0] for _ in range(n_iter)]       # owners that finished reading per (iteration, band)
    gathered_all = [0] * n_iter                       # warps past barrier 0 of iteration i
    tma_issued = [False] * (n_iter + 1)
    tma_issued[0] = True
    blocked = {}

    def warp(w):
        band = w >> 3
        for it in range(n_iter):
            while not tma_issued[it]:
                yield ("spin",)
            for b, bar_id in ((0, 1), (1, 3)):
                handed = it == 0 or band == b
                while True:
                    u = counters[it & 1][b]
                    counters[it & 1][b] += 1
                    if u >= units[b]:
                        break
                    yield ("work",)  # gathers into registers
                    if not handed:
                        yield ("sync", bar_id)
                        handed = True
                    if it > 0:
                        assert read_done[it - 1][b] == NW // 2, f"store into range {b} before its owners read pair {it-1}"
                    stored[it][b] += 1
                    yield ("work",)
                if not handed:
                    yield ("sync", bar_id)
                if b == 0:
                    if band == 0:
                        yield ("sync", 5)
                        assert stored[it][0] == units[0], "band A read before all its units are stored"
                        yield ("work",)
                        read_done[it][0] += 1
                        yield ("arrive", 1)
                        yield ("sync", 2)
                        yield ("work",)  # global stores
                    else:
                        yield ("arrive", 5)
            yield ("sync", 0)
            assert stored[it][0] == units[0] and stored[it][1] == units[1]
            gathered_all[it] += 1
            if w == NW - 1:
                counters[it & 1] = [0, 0]
                assert gathered_all[it] >= 1
                tma_issued[it + 1] = True
            if band == 1:
                yield ("work",)
                read_done[it][1] += 1
                yield ("arrive", 3)
                yield ("sync", 4)
                yield ("work",)

    gens = {w: warp(w) for w in range(NW)}
    runnable = set(gens)
    steps = 0
    while gens:
        if not runnable:
            raise RuntimeError(f"deadlock (seed {seed}): blocked = {blocked}")
        w = rng.choice(sorted(runnable))
        try:
            op = next(gens[w])
        except StopIteration:
            del gens[w]
            runnable.discard(w)
            continue
        steps += 1
        if op[0] == "sync":
            b = bars[op[1]]
            b.arrived += 1
            b.waiting.append(w)
            runnable.discard(w)
            blocked[w] = op[1]
            if b.arrived == b.count:
                for x in b.waiting:
                    runnable.add(x)
                    blocked.pop(x, None)
                b.arrived, b.waiting = 0, []
            assert b.arrived <= b.count
        elif op[0] == "arrive":
            b = bars[op[1]]
            b.arrived += 1
            assert b.arrived <= b.count, f"barrier {op[1]} over-subscribed"
            if b.arrived == b.count:
                for x in b.waiting:
                    runnable.add(x)
                    blocked.pop(x, None)
                b.arrived, b.waiting = 0, []
        elif op[0] == "spin" and steps > 10_000_000:
            raise RuntimeError("livelock")
    return steps


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    for seed in range(n):
        run(seed, n_iter=rng_iter if (rng_iter := 2 + seed % 5) else 2, units=(1 + seed % 9, 1 + (seed // 3) % 9))
    print(f"{n} random schedules: no deadlock, no ordering violation")
