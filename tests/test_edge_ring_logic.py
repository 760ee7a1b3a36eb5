"""CPU: the bookkeeping of the per-thread cp.async rings in the edge kernels (csrc/edge.cu, EDGE_*_RING), replayed in Python.

A thread issues one commit group per potential load (empty groups included) and waits with `cp.async.wait_group NB-1` before it
reads a slot.  The model below replays the kernels' loop for every chunk length and checks the two things the scheme rests on:
the group that filled a slot has landed when the slot is read (only "all but the newest N groups" are known to have landed), and
a slot is never refilled before it has been read.  (The kernels themselves are checked on the GPU: tests/test_gpu_parity.py.)"""
import pytest


class _Thread:
    def __init__(self):
        self.groups = []          # commit groups in order: list of (slot, first_row) or None for an empty group
        self.landed = 0           # groups [0, landed) are known to have completed
        self.slot = {}            # slot -> (first_row, group index, consumed?)

    def ring_load(self, b, k, on):
        if on:
            prev = self.slot.get(b)
            assert prev is None or prev[2], f"slot {b} refilled before rows {prev[0]}.. were consumed"
            self.slot[b] = [k, len(self.groups), False]
            self.groups.append((b, k))
        else:
            self.groups.append(None)

    def wait(self, n):
        self.landed = max(self.landed, len(self.groups) - n)

    def consume(self, b, k):
        first, g, done = self.slot[b]
        assert first == k, f"slot {b} holds rows {first}.., wanted {k}.."
        assert g < self.landed, f"rows {k}.. read before their group ({g}) is known to have landed ({self.landed})"
        assert not done
        self.slot[b][2] = True


@pytest.mark.parametrize("nb", [1, 2, 3, 4, 6])
@pytest.mark.parametrize("u", [1, 2, 4])
def test_ring_schedule(nb, u):
    for cmax in range(0, 33):                      # rows of the chunk (warp-uniform trip count)
        t = _Thread()
        for b in range(nb):                        # prologue
            t.ring_load(b, b * u, b * u < cmax)
        consumed = []
        k = 0
        while k < cmax:
            for b in range(nb):
                if k + b * u < cmax:
                    t.wait(nb - 1)
                    t.consume(b, k + b * u)
                    consumed.append(k + b * u)
                    t.ring_load(b, k + (b + nb) * u, k + (b + nb) * u < cmax)
            k += nb * u
        t.wait(0)                                  # nothing is left in flight when the next chunk reuses the slots
        assert consumed == list(range(0, cmax, u)), (nb, u, cmax)
        assert t.landed == len(t.groups)
