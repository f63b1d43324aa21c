"""Host logic of the S3Gen batcher and the T3 slot semaphore (no GPU): dependency chaining inside a batch, finished
dependencies passed as tensors, dropped jobs, submission order, priority order."""
import os
import sys
import threading
import time

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "chatterbox-tts_b200"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)

from cbx_b200.engine import PrioritySlots, S3GenBatcher, _Cancelled  # noqa: E402
from fake_backend import FakeModel  # noqa: E402


class BatchNative:
    """FakeNative's S3Gen with the batch entry point; records how every batch was composed."""
    is_fake = True
    device = 0

    def __init__(self, gate: threading.Event = None):
        self.batches, self.gate = [], gate

    def s3gen_infer(self, voice, tokens, cache_source=None, seed=0, **kw):
        return FakeModel.s3gen(tokens, cache_source)

    def s3gen_infer_batch(self, calls):
        if self.gate is not None:
            self.gate.wait(5)
        self.batches.append([(len(t), ("chain", c) if isinstance(c, int) else ("tensor" if c is not None else None)) for _, t, c, *_ in calls])
        outs = []
        for voice, toks, cache, seed, *_ in calls:
            if isinstance(cache, int):
                cache = outs[cache][1]            # the vocoder of a batch runs in call order: the earlier source exists
            outs.append(FakeModel.s3gen(toks, cache))
        return outs


def _toks(n, k=0):
    return [(7 * i + k) % 6561 for i in range(n)]


def test_chained_slices_share_a_batch_and_match_sequential():
    gate = threading.Event()
    nat = BatchNative(gate)
    b = S3GenBatcher(nat, max_batch=8, workers=1)
    try:
        first = b.submit(0, _toks(35), None, 1)           # occupies the worker (blocked on the gate) ...
        time.sleep(0.1)
        j1 = b.submit(0, _toks(70), first, 2)             # ... while three dependent slices and one unrelated call pile up
        j2 = b.submit(0, _toks(105), j1, 3)
        other = b.submit(0, _toks(20, 5), None, 4)
        j3 = b.submit(0, _toks(140), j2, 5)
        gate.set()
        outs = [j.wait() for j in (first, j1, j2, other, j3)]
    finally:
        b.stop()
    # the reference order: every slice synthesised with the previous slice's source as cache_source
    src, ref = None, []
    for n in (35, 70, 105):
        w, src = FakeModel.s3gen(_toks(n), src)
        ref.append((w, src))
    w4, s4 = FakeModel.s3gen(_toks(140), src)
    for got, want in zip((outs[0], outs[1], outs[2]), ref):
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
    assert torch.equal(outs[4][0], w4)
    assert torch.equal(outs[3][0], FakeModel.s3gen(_toks(20, 5), None)[0])
    # first ran alone; the second batch holds j1 (finished dependency -> tensor), j2 and j3 chained to earlier members
    assert nat.batches[0] == [(35, None)]
    assert nat.batches[1] == [(70, "tensor"), (105, ("chain", 0)), (20, None), (140, ("chain", 1))]


def test_dropped_job_and_its_dependents_are_cancelled():
    gate = threading.Event()
    nat = BatchNative(gate)
    b = S3GenBatcher(nat, max_batch=8, workers=1)
    try:
        first = b.submit(0, _toks(35), None, 1)
        time.sleep(0.1)
        j1 = b.submit(0, _toks(70), first, 2)
        j2 = b.submit(0, _toks(105), j1, 3)
        keep = b.submit(0, _toks(12), None, 4)
        j1.dropped = True                                  # what a cancelled request does to its pending jobs
        gate.set()
        first.wait()
        with pytest.raises(_Cancelled):
            j1.wait()
        with pytest.raises(_Cancelled):
            j2.wait()
        assert torch.equal(keep.wait()[0], FakeModel.s3gen(_toks(12), None)[0])
    finally:
        b.stop()


def test_batch_size_is_bounded_and_order_is_kept():
    gate = threading.Event()
    nat = BatchNative(gate)
    b = S3GenBatcher(nat, max_batch=4, workers=1)
    try:
        head = b.submit(0, _toks(3), None, 0)
        time.sleep(0.1)
        jobs = [b.submit(0, _toks(4 + i), None, i) for i in range(9)]
        gate.set()
        head.wait()
        for j in jobs:
            j.wait()
    finally:
        b.stop()
    sizes = [len(x) for x in nat.batches]
    assert sizes[0] == 1 and max(sizes) <= 4 and sum(sizes) == 10
    assert [n for batch in nat.batches[1:] for n, _ in batch] == [4 + i for i in range(9)]


def test_priority_slots_serve_lowest_priority_first():
    slots = PrioritySlots(1)
    assert slots.acquire(0)
    order, ths = [], []

    def waiter(prio):
        if slots.acquire(prio):
            order.append(prio)
            slots.release()

    for prio in (5, 1, 3):                                 # arrive in this order while the slot is taken
        t = threading.Thread(target=waiter, args=(prio,))
        t.start()
        ths.append(t)
        time.sleep(0.05)
    slots.release()
    for t in ths:
        t.join(5)
    assert order == [1, 3, 5]
    assert slots.acquire(9, cancelled=lambda: False)       # free again
    assert not slots.acquire(0, cancelled=lambda: True)    # a cancelled waiter gives up instead of blocking


def test_urgent_job_gets_a_consumer_window_before_the_next_batch():
    """The first audio of a request: the batcher does not start the next batch until the consumer has signalled that it
    took the output (or the window expired), so the emitter's PCM kernel never queues behind the next call's launch burst."""
    nat = BatchNative()
    b = S3GenBatcher(nat, max_batch=8, workers=1)
    b.urgent_window_s = 0.5
    try:
        first = b.submit(0, _toks(35), None, 1, urgent=True)
        first.wait()
        nxt = b.submit(0, _toks(70), first, 2)
        time.sleep(0.15)
        assert not nxt.done.is_set() and len(nat.batches) == 1      # still inside the window
        first.consumed.set()
        nxt.wait()
        assert len(nat.batches) == 2
        # a consumer that never answers (cancelled request) only costs the window
        b.urgent_window_s = 0.05
        u = b.submit(0, _toks(35, 3), None, 3, urgent=True)
        u.wait()
        t0 = time.time()
        b.submit(0, _toks(35, 4), None, 4).wait()
        assert time.time() - t0 < 2.0
        assert not b.submit(0, _toks(3), None, 5).urgent
    finally:
        b.stop()


def test_scheduler_can_align_streams_that_open_together(monkeypatch):
    """CBX_T3_ALIGN_OPENS_MS (opt-in): streams whose prefills overlap in time start decoding in the same round."""
    from cbx_b200.engine import T3Scheduler, SamplingDefaults
    from fake_backend import FakeNative

    class SlowOpen(FakeNative):
        def __init__(self):
            super().__init__()
            self.first_batches, self.mu = [], threading.Lock()

        def t3_open(self, *a, **kw):
            with self.mu:                      # prefills are serialised on the device
                time.sleep(0.03)
                return super().t3_open(*a, **kw)

        def t3_step(self, slots, n_steps=1, noise=None):
            self.first_batches.append(sorted(slots))
            return super().t3_step(slots, n_steps, noise)

    def run(align_ms):
        monkeypatch.setenv("CBX_T3_ALIGN_OPENS_MS", str(align_ms))
        nat = SlowOpen()
        sch = T3Scheduler(nat, max_batch=8)
        try:
            out = [None] * 4
            def opener(i):
                out[i] = sch.open(0, [255] + _toks(20, i) + [0], 0.5, 0.8, SamplingDefaults(), i, 40)
            th = [threading.Thread(target=opener, args=(i,)) for i in range(4)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            for s in out:
                with s.cv:
                    while not s.finished:
                        s.cv.wait(0.05)
            return nat.first_batches[0], [list(s.tokens) for s in out]
        finally:
            sch.running = False
            with sch.lock:
                sch.lock.notify_all()

    first_aligned, toks_aligned = run(500)
    first_plain, toks_plain = run(0)
    assert len(first_aligned) == 4            # all four streams in the first decode round
    assert len(first_plain) < 4               # without alignment the first stream runs ahead of the others' prefills
    assert sorted(map(tuple, toks_aligned)) == sorted(map(tuple, toks_plain))   # same tokens either way


def test_gather_window_waits_for_imminent_first_slices():
    """A first slice that opens a gather window is held (bounded) while other requests are about to submit theirs: requests that
    decode one T3 round apart still share ONE batch; a lone request (nothing imminent) is not delayed beyond the short window;
    the extension is capped."""
    nat = BatchNative()
    b = S3GenBatcher(nat, max_batch=8, workers=1)
    b.gather_s, b.urgent_gather_s = 0.003, 0.2
    pending = {"n": 1}
    b.imminent = lambda: pending["n"]
    try:
        t0 = time.time()
        j0 = b.submit(0, _toks(35), None, 1, urgent=True)
        time.sleep(0.03)                       # ten short windows later the second request's first slice arrives
        j1 = b.submit(0, _toks(35, 3), None, 2, urgent=True)
        pending["n"] = 0
        j0.consumed.set(); j1.consumed.set()
        j0.wait(); j1.wait()
        assert nat.batches[0] == [(35, None), (35, None)], nat.batches
        # nothing imminent: only the short window
        t1 = time.time()
        j2 = b.submit(0, _toks(35, 4), None, 3, urgent=True)
        j2.consumed.set()
        j2.wait()
        assert time.time() - t1 < 0.1 and len(nat.batches) == 2
        # imminent forever: the cap ends the wait
        pending["n"] = 5
        t2 = time.time()
        j3 = b.submit(0, _toks(35, 5), None, 4, urgent=True)
        j3.consumed.set()
        j3.wait()
        assert 0.15 < time.time() - t2 < 1.0
        # a window without a first slice in it does not wait for imminent ones
        t3 = time.time()
        j4 = b.submit(0, _toks(35, 6), None, 5)
        j4.wait()
        assert time.time() - t3 < 0.1
    finally:
        b.stop()
