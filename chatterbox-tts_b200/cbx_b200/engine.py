"""Streaming engine: the host-side mirror of the reference's `TextToSpeechEngine`
(src/tts_streaming.py:161-968) over the native B200 path.

Same surface the worker touches (src/worker.py:30-44, :76-77, :100, :113, :129-131):
`TextToSpeechEngine(device)`, `await ainit()`, `async stream(...)`, `prepare_conditionals(path)`,
`clear_voice_cache(id)`, `shutdown()`, `.voice_manager`, `.voice_conditioning_executor`.
Same audio semantics: 35-token slices with a look-ahead of max(3, 0.2*slice) tokens (:499-501, :535),
"full"/"zero" overlap (:655-659, :694-699), EOS(0) append + <6561 filter + pad to 3 (:661-677),
lead/trail trim (:702-707), equal-power crossfade state machine (:710-746, :756-760), clamp +
int16 truncation (:149-155).

What is different (B200-first): T3 decode steps of all concurrent requests are batched
into one GEMV pass per step by a scheduler thread; the text chunks of one request decode concurrently
(they are independent: the reference re-primes T3 and resets the S3Gen cache per chunk, :453-491, :648-653)
and their audio is emitted in order; S3Gen slices of all chunks / requests that are ready at the same time go
through ONE batched token->mel pass (cbx_s3gen_infer_batch); crossfade + PCM conversion are one device kernel.
"""
import asyncio
import collections
import concurrent.futures
import contextlib
import heapq
import os
import queue
import threading
import time
import zlib
from dataclasses import dataclass, field
from pathlib import Path
from typing import AsyncGenerator, List, Optional

import numpy as np
import torch

from .config import ModelConfig, S3GEN_SR
from .native import NativeEngine
from .text_processing import split_text_into_chunks, SyntheticTokenizer, JsonTokenizer

SPEECH_VOCAB = 6561


class _Cancelled(Exception):
    pass


class CancellationToken:
    """asyncio.Event-backed token with the reference's interface (src/tts_streaming.py:88-104)."""

    def __init__(self, loop: asyncio.AbstractEventLoop = None):
        self._event = asyncio.Event()
        self._flag = threading.Event()

    def cancel(self):
        self._flag.set()
        self._event.set()

    def is_cancelled(self) -> bool:
        return self._flag.is_set()

    async def wait(self):
        await self._event.wait()


class VoiceManager:
    """voice file lookup: user directory overrides preloaded (reference src/voice_manager.py:39-52)."""

    def __init__(self, voices_dir: str = None, preloaded_voices_dir: str = None):
        self.voices_dir = voices_dir or os.environ.get("VOICES_DIR", "voices/")
        self.preloaded_voices_dir = preloaded_voices_dir or os.environ.get("PRELOADED_VOICES_DIR", "preloaded-voices/")

    def get_voice_path(self, voice_id: str):
        for d in (self.voices_dir, self.preloaded_voices_dir):
            p = os.path.join(d, voice_id)
            if os.path.exists(p):
                return p
        return None

    def list_voices(self) -> List[str]:
        out = set()
        for d in (self.voices_dir, self.preloaded_voices_dir):
            if os.path.isdir(d):
                out.update(f for f in os.listdir(d) if os.path.isfile(os.path.join(d, f)))
        return sorted(out)


@dataclass
class SamplingDefaults:
    """The reference passes only temperature and cfg_weight (:483-491); the fork's other defaults are
    unknown, so they are engine settings."""
    repetition_penalty: float = 1.2
    min_p: float = 0.05
    top_p: float = 0.95
    max_new_tokens: int = 1000
    tokens_per_word: Optional[int] = None   # benchmark length rule: max_new = tokens_per_word * words


class _T3Stream:
    def __init__(self, slot, max_new):
        self.slot, self.max_new = slot, max_new
        self.tokens: List[int] = []
        self.finished = False
        self.cancelled = False
        self.error: Optional[BaseException] = None
        self.cv = threading.Condition()
        self.on_closed = None      # called once, after the native slot has been closed (releases the request's slot permit)


class T3Scheduler(threading.Thread):
    """Continuous batching: every round advances all open streams by `steps_per_round` tokens in one
    batched decode pass (rows = 2 x streams), then hands the new tokens to their requests."""

    def __init__(self, native: NativeEngine, steps_per_round: int = 7, max_batch: int = 8):
        super().__init__(daemon=True, name="cbx-t3-scheduler")
        self.native, self.k, self.max_batch = native, steps_per_round, max_batch
        self.on_gpu = not getattr(native, "is_fake", False)
        self.active: List[_T3Stream] = []
        self.lock = threading.Condition()
        self.running = True
        self.rounds = 0
        self.busy_s = 0.0                              # wall time inside decode rounds (stats)
        self.row_hist = collections.Counter()          # rows per decode round -> count
        # CBX_T3_ALIGN_OPENS_MS (default 30, 0 = off): while every open stream is still at token 0 and other streams are being
        # opened (their prefills run one after the other, ~3 ms each), the first decode round waits up to this long, so that
        # requests that arrive together decode in lockstep and their first slices share ONE S3Gen batch instead of "first
        # alone, rest in the next batch".  Measured on a B200 (tools/conc_bench.py, 8 streams x 100 words, second wave): first
        # chunk p50 468 ms without, 255 ms at 5 ms, 207 ms at 15 ms.  A lone request never waits (nothing else is opening).
        self.align_s = float(os.environ.get("CBX_T3_ALIGN_OPENS_MS", "30")) * 1e-3
        self.opening = 0
        self.open_q = collections.deque()
        self.open_gather_s = float(os.environ.get("CBX_T3_OPEN_GATHER_MS", "0.5")) * 1e-3
        self.unhealthy: Optional[BaseException] = None   # a native slot could not be closed: its pages are lost, refuse new work
        # T3 stream priority per decode round (native.t3_set_priority): high_priority() is asked before every round.  The engine
        # wires it to "no first slice of a request is waiting for S3Gen": T3 runs ahead of the wide S3Gen grids (throughput) except
        # while a request's first audio is being synthesised (first-chunk latency).
        self.high_priority = None
        self._prio = None
        self.opener = threading.Thread(target=self._opener, daemon=True, name="cbx-t3-opener")
        self.opener.start()
        self.start()

    def open(self, voice, text_ids, cfg_w, temp, sd: SamplingDefaults, seed, max_new, on_closed=None) -> _T3Stream:
        """Prefill + stream slot.  `on_closed` runs exactly once after the NATIVE slot has been closed (natural end, cancel or
        error): the caller's slot permit must not be handed to another request before that, or the next t3_open finds no
        free native slot."""
        if self.unhealthy is not None:
            raise RuntimeError(f"T3 engine is unhealthy after an unrecoverable error: {self.unhealthy}")
        if not self.running:
            raise RuntimeError("engine is shutting down")
        # Opens are handed to the opener thread: the ones that are pending at the same moment (8 requests arriving together, the
        # text chunks of one request) are prefilled in ONE pass (cbx_t3_open_batch) instead of queueing behind each other.
        req = {"args": (voice, text_ids, cfg_w, temp, sd.repetition_penalty, sd.min_p, sd.top_p, seed, max_new), "max_new": max_new,
               "on_closed": on_closed, "done": threading.Event(), "stream": None, "err": None}
        with self.lock:
            self.opening += 1
            self.open_q.append(req)
            self.lock.notify_all()
        req["done"].wait()
        if req["err"] is not None:
            raise req["err"]
        return req["stream"]

    def _opener(self):
        if self.on_gpu:
            torch.cuda.set_device(self.native.device)
        batchable = hasattr(self.native, "t3_open_batch")
        while True:
            with self.lock:
                while self.running and not self.open_q:
                    self.lock.wait(0.5)
                if not self.running:
                    pend, self.open_q = list(self.open_q), collections.deque()
                    self.opening -= len(pend)
                    for r in pend:
                        r["err"] = RuntimeError("engine is shutting down")
                        r["done"].set()
                    return
                if batchable and self.open_gather_s > 0 and len(self.open_q) < 8:
                    t_end = time.time() + self.open_gather_s       # companions of a burst arrive within a fraction of a millisecond
                    while self.running and len(self.open_q) < 8:
                        left = t_end - time.time()
                        if left <= 0:
                            break
                        self.lock.wait(left)
                batch = [self.open_q.popleft() for _ in range(min(8 if batchable else 1, len(self.open_q)))]
            slots = None
            if len(batch) > 1:
                try:
                    with self._ctx():
                        slots = self.native.t3_open_batch([r["args"] for r in batch])
                except BaseException:
                    slots = None        # e.g. not enough slots / pages for all of them: open one by one, each with its own verdict
            for i, r in enumerate(batch):
                try:
                    if slots is not None:
                        slot = slots[i]
                    else:
                        with self._ctx():
                            slot = self.native.t3_open(*r["args"])
                    st = _T3Stream(slot, r["max_new"])
                    st.on_closed = r["on_closed"]
                    r["stream"] = st
                except BaseException as ex:
                    r["err"] = ex
            with self.lock:
                self.opening -= len(batch)
                for r in batch:
                    if r["stream"] is not None:
                        self.active.append(r["stream"])
                self.lock.notify_all()
            for r in batch:
                r["done"].set()

    def cancel(self, s: _T3Stream):
        s.cancelled = True

    _tls = threading.local()

    def _ctx(self):
        """Per-thread CUDA stream context (a no-op for the host-logic tests' fake backend)."""
        if not self.on_gpu:
            return contextlib.nullcontext()
        if not hasattr(self._tls, "st"):
            self._tls.st = torch.cuda.Stream()
        return torch.cuda.stream(self._tls.st)

    def run(self):
        if self.on_gpu:
            torch.cuda.set_device(self.native.device)
        while self.running:
            with self.lock:
                while self.running and not self.active:
                    self.lock.wait(0.5)
                if not self.running:
                    return
                if self.align_s > 0 and self.opening > 0 and all(len(s.tokens) == 0 for s in self.active):
                    t_end = time.time() + self.align_s
                    while self.running and self.opening > 0 and time.time() < t_end:
                        self.lock.wait(0.001)
                dead = [s for s in self.active if s.cancelled]      # wherever they sit in the rotation
                batch = [s for s in self.active if not s.cancelled][: self.max_batch]
            for s in dead:
                self._retire(s)
            live = batch
            if not live:
                continue
            try:
                t_busy = time.time()
                with self._ctx():
                    if self.high_priority is not None:
                        want = bool(self.high_priority())
                        if want != self._prio:
                            self.native.t3_set_priority(want)
                            self._prio = want
                    self.native.t3_step([s.slot for s in live], self.k)
                    for s in live:
                        n, done = self.native.t3_poll(s.slot)
                        new = self.native.t3_tokens(s.slot, len(s.tokens), n - len(s.tokens)).tolist() if n > len(s.tokens) else []
                        with s.cv:
                            s.tokens.extend(new)
                            if done:
                                s.finished = True
                            s.cv.notify_all()
                        if done:
                            self._retire(s)
                self.rounds += 1
                self.busy_s += time.time() - t_busy
                self.row_hist[2 * len(live)] += 1
            except BaseException as ex:  # surface engine failures to every waiting request
                for s in live:
                    with s.cv:
                        s.error = ex
                    self._retire(s)      # always give the native slot and its KV pages back (or mark the engine unhealthy)
            # round-robin fairness when more streams than the batch size are open
            with self.lock:
                if len(self.active) > self.max_batch:
                    self.active = self.active[self.max_batch:] + self.active[: self.max_batch]

    def _retire(self, s: _T3Stream):
        with self.lock:
            if s not in self.active:
                return
            self.active.remove(s)
        try:
            self.native.t3_close(s.slot)
        except BaseException as ex:      # the slot and its pages are lost: stop handing out work instead of failing later opens
            self.unhealthy = ex
        cb, s.on_closed = s.on_closed, None
        if cb is not None:
            cb()
        with s.cv:
            s.finished = True
            s.cv.notify_all()

    def stop(self):
        """Stops the scheduler and fails every stream that is still open, so no request thread waits for tokens forever."""
        self.running = False
        with self.lock:
            self.lock.notify_all()
        if threading.current_thread() is not self:
            self.join(timeout=10.0)
        with self.lock:
            rest = list(self.active)
        for s in rest:
            with s.cv:
                if s.error is None and not s.finished:
                    s.error = RuntimeError("engine is shutting down")
            self._retire(s)


class _S3Job:
    """One S3Gen call in flight: `dep` is the job whose source output is this call's cache_source (the previous slice of
    the same text chunk under the "full" overlap strategy), or None."""
    __slots__ = ("voice", "toks", "dep", "seed", "done", "out", "err", "dropped", "urgent", "consumed", "emit_from")

    def __init__(self, voice, toks, dep, seed, urgent=False, emit_from=0):
        self.voice, self.toks, self.dep, self.seed = voice, toks, dep, seed
        self.emit_from = emit_from      # samples: the consumer reads wav[emit_from:] only ("full" overlap), the vocoder may decode a window
        self.done, self.out, self.err, self.dropped = threading.Event(), None, None, False
        # urgent: the first audio of a request.  The batcher gives its consumer a short exclusive window before it starts the
        # next batch: the PCM kernel + D2H of the emitter otherwise queue inside the driver behind the ~5 000-launch burst of the
        # next S3Gen call (measured: the first chunk intermittently arrived 60-150 ms late, 4 ms late at best)
        self.urgent, self.consumed = urgent, (threading.Event() if urgent else None)

    def wait(self):
        self.done.wait()
        if self.err is not None:
            raise self.err
        return self.out


class S3GenBatcher:
    """Collects the S3Gen calls that are pending at the same moment (slices of concurrent requests, of the text chunks
    of one request, and consecutive slices of one chunk) and runs them as one batch: while a batch is on the GPU the
    next one accumulates.  Jobs are taken in submission
    order; a job runs only when its dependency is finished or an earlier member of the same batch (chained on the
    device).  `workers` threads can run batches concurrently (the native side has two batch workspaces); measured on the
    pipelined workloads one worker is better (two fragment the batches: 41 vs 54 audio-s/s at 8 concurrent streams)."""

    def __init__(self, native, max_batch: int = None, workers: int = None):
        self.native = native
        self.max_batch = max_batch or int(os.environ.get("CBX_S3GEN_MAX_BATCH", "8"))
        self.max_batch_deep = max(self.max_batch, int(os.environ.get("CBX_S3GEN_MAX_BATCH_DEEP", "16"))) if max_batch is None else max_batch
        self.on_gpu = not getattr(native, "is_fake", False)
        self.can_batch = hasattr(native, "s3gen_infer_batch")
        self.jobs = collections.deque()
        self.cv = threading.Condition()
        self.running = True
        self.batches = collections.Counter()   # batch size -> count (bench / tests)
        self.busy_s = 0.0                      # wall time inside batched calls, including the stream sync (stats)
        self.pad_stats = [0, 0]                # new tokens requested, new tokens computed after padding to the batch maximum
        n = workers or int(os.environ.get("CBX_S3GEN_WORKERS", "1"))
        # after the first pending job shows up, wait this long for companions (slices of concurrent requests become ready
        # within a decode round of each other): one batch of 8 beats a single call followed by a batch of 7
        self.gather_s = float(os.environ.get("CBX_S3GEN_GATHER_MS", "3")) * 1e-3
        # "full" overlap emits wav[previous_length:] only: the vocoder decodes a window that ends the call (exact from emit_from on)
        self.window = os.environ.get("CBX_HIFT_WINDOW", "1") != "0"
        self.gather_always = os.environ.get("CBX_S3GEN_GATHER_ALWAYS", "0") == "1"      # experiment: also gather when jobs were already pending
        self.urgent_window_s = float(os.environ.get("CBX_S3GEN_URGENT_MS", "15")) * 1e-3   # upper bound; the emitter ends it
        self._urgent = []                      # urgent jobs (first audio of a request) submitted and not finished yet
        # imminent(): how many OTHER requests are about to submit their first slice (their T3 streams are within two decode rounds
        # of it).  A gather window that holds a first slice stays open for them (bounded): requests that arrived together but
        # whose prefills fell into different passes decode one round apart, and "one alone, seven in the next batch" costs the
        # seven a whole extra call (first chunk 107 / 181 ms instead of ~125 for all)
        self.imminent = None
        self.lone = None                       # lone(): True while a single request is in flight (no gather window then)
        self.urgent_gather_s = float(os.environ.get("CBX_S3GEN_URGENT_GATHER_MS", "15")) * 1e-3
        self.threads = [threading.Thread(target=self.run, daemon=True, name=f"cbx-s3gen-batcher-{i}") for i in range(n if self.can_batch else 1)]
        for t in self.threads:
            t.start()

    def submit(self, voice, toks, dep: Optional[_S3Job], seed, urgent=False, emit_from=0) -> _S3Job:
        job = _S3Job(voice, toks, dep, seed, urgent, emit_from if self.window else 0)
        with self.cv:
            self.jobs.append(job)
            if urgent:
                self._urgent.append(job)
            self.cv.notify_all()
        return job

    def urgent_inflight(self) -> bool:
        """True while some request's first slice is queued or on the GPU (the T3 scheduler then yields the SMs to it)."""
        with self.cv:
            if self._urgent:
                self._urgent = [j for j in self._urgent if not j.done.is_set()]
            return bool(self._urgent)

    def infer(self, voice, toks, cache_source, seed):
        """Blocking single call (cache_source given as a finished tensor)."""
        dep = None
        if cache_source is not None:
            dep = _S3Job(voice, None, None, 0)
            dep.out = (None, cache_source)
            dep.done.set()
        return self.submit(voice, toks, dep, seed).wait()

    def _take(self):
        """Under self.cv: the longest prefix-ordered set of runnable jobs, at most max_batch."""
        batch, members, keep = [], set(), collections.deque()
        # deep queue (many concurrent requests): the native capacity of 16 calls per batch pays; otherwise smaller batches keep
        # the slice-to-slice latency chain of a single request short
        limit = (self.max_batch_deep if len(self.jobs) >= self.max_batch_deep else self.max_batch) if self.can_batch else 1
        while self.jobs:
            j = self.jobs.popleft()
            ready = j.dep is None or j.dep.done.is_set() or id(j.dep) in members
            if ready and len(batch) < limit:
                batch.append(j)
                members.add(id(j))
            else:
                keep.append(j)      # its dependency is in another worker's batch (or the batch is full): stays queued, in order
        self.jobs = keep
        return batch

    def run(self):
        st = None
        if self.on_gpu:
            torch.cuda.set_device(self.native.device)
            st = torch.cuda.Stream()
        while True:
            with self.cv:
                batch = []
                idle = (not self.jobs) or self.gather_always      # nothing accumulated while the previous batch ran: the next job opens a gather window
                while self.running:
                    # the first slice of a request that is alone on the engine has nobody to wait for (its other chunks start after it)
                    alone = self.lone is not None and len(self.jobs) == 1 and self.jobs[0].urgent and self.lone()
                    if idle and self.jobs and self.can_batch and self.gather_s > 0 and len(self.jobs) < self.max_batch and not alone:
                        # every submit wakes this wait: keep gathering until the window (counted from the first job) closes
                        t_end = time.time() + self.gather_s
                        t_cap = time.time() + self.urgent_gather_s
                        while self.running and len(self.jobs) < self.max_batch:
                            now = time.time()
                            left = t_end - now
                            if left <= 0:
                                if self.imminent is None or now >= t_cap or not any(j.urgent for j in self.jobs) or self.imminent() <= 0:
                                    break
                                left = min(0.001, t_cap - now)
                            self.cv.wait(left)
                    batch = self._take() if self.jobs else []
                    if batch:
                        break
                    self.cv.wait(0.01 if self.jobs else 0.5)
                if not self.running:
                    for j in list(self.jobs) + batch:
                        j.err = RuntimeError("engine is shutting down")
                        j.done.set()
                    return
            live = []
            for j in batch:
                # a dropped job (cancelled request) or one whose dependency failed is not run
                if j.dropped or (j.dep is not None and (j.dep.dropped or j.dep.err is not None)):
                    j.dropped, j.err = True, (j.dep.err if j.dep is not None and j.dep.err is not None else _Cancelled())
                    j.done.set()
                else:
                    live.append(j)
            if live:
                t_busy = time.time()
                try:
                    with (torch.cuda.stream(st) if st is not None else contextlib.nullcontext()):
                        def cache_of(j, pos):
                            if j.dep is None:
                                return None
                            if id(j.dep) in pos:
                                return pos[id(j.dep)]       # earlier member of this batch: chained on the device
                            return j.dep.out[1]
                        if not self.can_batch:
                            j = live[0]
                            outs = [self.native.s3gen_infer(j.voice, j.toks, cache_source=cache_of(j, {}), seed=j.seed)]
                        else:
                            # one call also goes through the batch entry point (plain launches): it is as fast as the per-lane
                            # CUDA graph of cbx_s3gen_infer (the GPU, not the launch rate, bounds a call) and never stalls on
                            # capturing a graph for a token count it has not seen yet
                            pos, calls = {}, []
                            for i, j in enumerate(live):
                                calls.append((j.voice, j.toks, cache_of(j, pos), j.seed, j.emit_from))
                                pos[id(j)] = i
                            outs = self.native.s3gen_infer_batch(calls)
                        if st is not None:
                            st.synchronize()
                    with self.cv:
                        self.busy_s += time.time() - t_busy
                        self.batches[len(live)] += 1
                        lens = [len(j.toks) for j in live]
                        self.pad_stats[0] += sum(lens)
                        self.pad_stats[1] += max(lens) * len(lens)
                    for j, o in zip(live, outs):
                        j.out = o
                except BaseException as ex:
                    for j in live:
                        j.err = ex
                for j in live:
                    j.done.set()
                for j in live:
                    if j.urgent and j.err is None and not j.dropped:
                        j.consumed.wait(self.urgent_window_s)
            with self.cv:
                self.cv.notify_all()     # jobs that were waiting for this batch's outputs are runnable now

    def stop(self):
        self.running = False
        with self.cv:
            self.cv.notify_all()


class PrioritySlots:
    """Counting semaphore whose waiters are served lowest (priority, arrival) first: T3 stream slots go to the earliest
    text chunks, so a request's later chunks never starve another request's first chunk."""

    def __init__(self, n: int):
        self.free, self.cv, self.waiters, self.seq = n, threading.Condition(), [], 0

    def acquire(self, prio: int, cancelled=None) -> bool:
        with self.cv:
            self.seq += 1
            me = (prio, self.seq)
            heapq.heappush(self.waiters, me)
            while not (self.free > 0 and self.waiters[0] == me):
                if cancelled is not None and cancelled():
                    self.waiters.remove(me)
                    heapq.heapify(self.waiters)
                    self.cv.notify_all()
                    return False
                self.cv.wait(0.05)
            heapq.heappop(self.waiters)
            self.free -= 1
            self.cv.notify_all()
            return True

    def release(self):
        with self.cv:
            self.free += 1
            self.cv.notify_all()


def drop_invalid_tokens(x: List[int], sos=6561, eos=6562) -> List[int]:
    """chatterbox.models.s3tokenizer.drop_invalid_tokens (reference call site :667): keep what lies
    between the first SOS (exclusive) and the first EOS (exclusive)."""
    s = x.index(sos) + 1 if sos in x else 0
    e = x.index(eos) if eos in x else None
    return x[s:e]


class TextToSpeechEngine:
    ENC_COND_LEN = 6 * 16000
    DEC_COND_LEN = 10 * S3GEN_SR

    def __init__(self, device: str, cfg: ModelConfig = None, state_dict=None, concurrent_requests: int = None,
                 sampling: SamplingDefaults = None, native_kwargs: dict = None, seed: int = 0, device_sink: bool = False,
                 backend=None, encoder_state_dict=None, alignment_eos: bool = None):
        """`backend` injects an object with NativeEngine's interface (host-logic tests only); the product path
        always builds a NativeEngine on `device` and refuses anything that is not cuda:N."""
        self.device = device
        self.gpu_id = int(device.split(":")[-1]) if "cuda" in device else -1
        self._backend = backend
        if self.gpu_id < 0 and backend is None:
            raise RuntimeError("TextToSpeechEngine (B200 path) needs a cuda:N device; there is no CPU fallback")
        self.cfg = cfg or ModelConfig()
        self._state_dict = state_dict
        self.sampling = sampling or SamplingDefaults()
        self.sr = S3GEN_SR
        self.voice_manager = VoiceManager()
        self.voice_cache = {}            # voice id -> native slot (mirrors the reference dict of Conditionals)
        self.default_conds = None
        n = concurrent_requests or int(os.environ.get("CONCURRENT_REQUESTS_PER_WORKER", "1"))
        self.concurrent_requests = n
        self.tts_semaphore = None
        self.request_executor = concurrent.futures.ThreadPoolExecutor(max_workers=n, thread_name_prefix="cbx-req")
        self.voice_conditioning_executor = concurrent.futures.ThreadPoolExecutor(max_workers=n, thread_name_prefix="cbx-voice")
        # two S3Gen lanes: measured on B200, more lanes do not raise aggregate throughput (the per-call kernels are
        # latency-bound and already span the SMs) and cost workspace + graph captures per lane
        # T3 stream slots: more than one decode batch (8 streams = 16 rows) may be open; the scheduler rotates batches, so the
        # ninth text chunk of a paragraph does not have to wait for a whole chunk to finish decoding
        self.native_kwargs = dict(max_streams=max(int(os.environ.get("CBX_MAX_STREAMS", "16")), n), n_lanes=2)
        # text chunks of one request that may be in flight at once (T3 decoding + S3Gen), ahead of the chunk being emitted
        self.chunk_parallelism = int(os.environ.get("CBX_CHUNK_PARALLELISM", "8"))
        # a request whose chunks ALL fit in this many streams starts them all (after chunk 0's first slice): there is no steady
        # state to protect, T3 finishes in one go and S3Gen batches fill.  Measured: 200-word paragraph (10 chunks) 71.4 -> 74.0
        # audio-s/s end to end; a 2000-word document is better off with 8 in flight (64.3 vs 60.4 at 16)
        self.chunk_parallelism_short = int(os.environ.get("CBX_CHUNK_PARALLELISM_SHORT", "16"))
        # later chunks of a request start when chunk 0's first slice is on its way: "audio" = its PCM has been sent (lowest
        # first-chunk latency), "tokens" = its tokens are decoded (their prefills overlap the first S3Gen call: +throughput)
        # "auto": "audio" for a request that is alone on the GPU, "tokens" when others are in flight.  Measured at 8 streams:
        # with 8 T3 stream slots "tokens" wins (first chunk p50 190 vs 251 ms), with the default 16 slots it loses badly
        # (669 vs 227 ms: the released chunks rotate with the other requests' first chunks in the decode batches), so the
        # default stays "audio"
        self.hold_until = os.environ.get("CBX_HOLD_UNTIL", "audio")
        self._inflight = 0
        self._inflight_lock = threading.Lock()
        # one worker thread per text chunk in flight: every request can have chunk_parallelism of them (most just wait for a T3
        # slot or for tokens), so the pool is sized for that product and a new request never queues behind waiting chunks
        self.chunk_executor = concurrent.futures.ThreadPoolExecutor(max_workers=max(32, n * (max(self.chunk_parallelism, self.chunk_parallelism_short) + 1)), thread_name_prefix="cbx-chunk")
        self.s3gen: Optional[S3GenBatcher] = None
        self._pinned_pool: queue.Queue = queue.Queue()
        self.t3_slots: Optional[PrioritySlots] = None
        self.native_kwargs.update(native_kwargs or {})
        self.native: Optional[NativeEngine] = None
        self.scheduler: Optional[T3Scheduler] = None
        self.tokenizer = None
        self.seed = seed
        self.weights_source = None
        self._encoder_sd = encoder_state_dict     # tokenizer.* / speaker_encoder.* / ve.* (conditioning encoders), else from the checkpoint
        self._cond, self._cond_lock = None, threading.Lock()
        self.device_sink = device_sink   # bench `value` leg: PCM stays in HBM, emit() receives sample counts
        # alignment-based EOS control of the T3 generators (the model package's AlignmentStreamAnalyzer): whether the
        # reference's fork runs it inside inference_stream is unknown, and with random-init weights its verdicts mean nothing, so
        # it is opt-in (CBX_ALIGNMENT_EOS=1 / alignment_eos=True; probe layer CBX_ALIGNMENT_LAYER, upstream 9)
        self.alignment_eos = (os.environ.get("CBX_ALIGNMENT_EOS", "0") == "1") if alignment_eos is None else bool(alignment_eos)
        self._first_pending, self._fp_lock = {}, threading.Lock()   # chunk-0 T3 streams whose first slice is not submitted yet
        self._seq = 0
        self._seq_lock = threading.Lock()
        self._ready = False
        self.stats = {"first_chunk_ms": []}

    # ------------------------------------------------------------------ lifecycle (reference ainit :209-339)
    async def ainit(self):
        loop = asyncio.get_running_loop()
        await loop.run_in_executor(self.request_executor, self._init_blocking)
        self.tts_semaphore = asyncio.Semaphore(self.concurrent_requests)
        return self

    def _init_blocking(self):
        from .weights import random_state_dict, synthetic_conditionals
        if self._backend is not None:
            self.native = self._backend
            self.tokenizer = SyntheticTokenizer(self.cfg.t3.text_vocab)
            self.voice_cache["default"] = 0
            self.scheduler = T3Scheduler(self.native, max_batch=8)
            self.s3gen = S3GenBatcher(self.native)
            self.t3_slots = PrioritySlots(self.native_kwargs["max_streams"])
            self._ready = True
            return
        torch.cuda.set_device(self.gpu_id)
        self.native = NativeEngine(self.cfg, device=self.gpu_id, **self.native_kwargs)
        sd = self._state_dict
        model_path = os.environ.get("MODEL_PATH", "models")
        self.weights_source = "state_dict"
        if sd is None:
            # merged file, or the upstream t3_cfg / s3gen files converted on the fly; a missing checkpoint is an ERROR as in
            # the reference (from_local, :252-258) unless CBX_ALLOW_RANDOM_WEIGHTS=1 (checkpoint.py)
            from .checkpoint import load_checkpoint
            sd, enc_sd, self.weights_source = load_checkpoint(model_path, self.cfg, self.seed)
            if self._encoder_sd is None:
                self._encoder_sd = enc_sd
        self.native.load_state_dict(sd)
        self._state_dict = None
        if self.alignment_eos:
            self.native.t3_set_alignment_eos(True, int(os.environ.get("CBX_ALIGNMENT_LAYER", "9")))
        tj = os.path.join(model_path, "tokenizer.json")
        self.tokenizer = JsonTokenizer(tj) if os.path.exists(tj) else SyntheticTokenizer(self.cfg.t3.text_vocab)
        cp = os.path.join(model_path, "conds.pt")
        if os.path.exists(cp):
            from .checkpoint import load_conds
            self.default_conds = load_conds(cp)                    # tts.conds (:399-404)
        elif self.weights_source in ("state_dict", "random"):
            self.default_conds = synthetic_conditionals(self.cfg)  # benchmark / tests: seeded tensors of trump.wav's shapes
        else:
            raise RuntimeError(f"{cp} is missing: a real checkpoint needs its default voice conditioning (reference :399-404)")
        self.voice_cache["default"] = self.native.voice_put("default", self.default_conds["t3"], self.default_conds["gen"])
        # up to 16 streams (32 rows) decode in ONE pass over the weights; more are rotated in groups
        self.scheduler = T3Scheduler(self.native, max_batch=min(int(os.environ.get("CBX_T3_MAX_BATCH", "16")), self.native_kwargs["max_streams"]))
        self.s3gen = S3GenBatcher(self.native)
        self.s3gen.imminent = self._imminent_first_slices
        self.s3gen.lone = lambda: self._inflight <= 1
        # CBX_T3_PRIORITY_MODE: auto (default) = T3 on its high-priority stream except while a first slice is in S3Gen, high / low = pinned
        mode = os.environ.get("CBX_T3_PRIORITY_MODE", "auto")
        if mode == "auto":
            self.scheduler.high_priority = lambda: not self.s3gen.urgent_inflight()
        elif mode == "high":
            self.scheduler.high_priority = lambda: True
        self.t3_slots = PrioritySlots(self.native_kwargs["max_streams"] - 1)   # one slot stays free for warm-up / direct opens
        # warm-up, as the reference does (:274-326): 4 T3 tokens with cfg 0, one tiny S3Gen call
        text = [self.cfg.t3.start_text_token] + self.tokenizer.text_to_tokens("compiling")[0].tolist() + [self.cfg.t3.stop_text_token]
        s = self.scheduler.open(self.voice_cache["default"], text, 0.0, 0.8, self.sampling, 0, 4)
        self._wait_tokens(s, 4, None)
        self.native.s3gen_infer(self.voice_cache["default"], [0, 0, 0])
        torch.cuda.synchronize()
        self._ready = True

    def _imminent_first_slices(self) -> int:
        with self._fp_lock:
            return sum(1 for st, need in self._first_pending.values() if st.error is None and not st.cancelled and len(st.tokens) >= need - 14)

    def healthy(self) -> bool:
        """False once the GPU context is unusable (sticky device fault); backends without the probe count as healthy."""
        probe = getattr(self.native, "healthy", None)
        return True if probe is None else bool(probe())

    def shutdown(self):
        if self.scheduler:
            self.scheduler.stop()
        if self.s3gen:
            self.s3gen.stop()
        self.chunk_executor.shutdown(wait=True)
        self.request_executor.shutdown(wait=True)
        self.voice_conditioning_executor.shutdown(wait=True)
        if self.native:
            self.native.close()

    # ------------------------------------------------------------------ voices (reference :349-406)
    def clear_voice_cache(self, voice_id: str):
        if voice_id in self.voice_cache and voice_id != "default":
            del self.voice_cache[voice_id]
            self.native.voice_drop(voice_id)

    def put_conditionals(self, voice_id: str, t3_cond: dict, gen: dict):
        self.voice_cache[voice_id] = self.native.voice_put(voice_id, t3_cond, gen)

    def conditioning_encoders(self):
        """The GPU conditioning encoders (S3Tokenizer-v2, CAMPPlus, VoiceEncoder, mel front ends), built on first use from the
        checkpoint's `tokenizer.*` / `speaker_encoder.*` / `ve.*` tensors.  Raises when the checkpoint has none: a voice is
        never invented."""
        with self._cond_lock:
            if self._cond is None:
                if self._backend is not None:
                    raise RuntimeError("prepare_conditionals needs the CUDA engine")
                if not self._encoder_sd:
                    raise RuntimeError("prepare_conditionals: the checkpoint holds no conditioning-encoder weights (tokenizer.*, speaker_encoder.*, "
                                       "ve.*); pass them (encoder_state_dict=) or use put_conditionals() with tensors computed elsewhere")
                from .conditioning import ConditioningEncoders
                self._cond = ConditioningEncoders(self._encoder_sd, self.cfg.cond, device=self.gpu_id)
                self._encoder_sd = None
                # one pass over a 10 s dummy clip (the longest the reference conditions on): the encoders' stream gets its pool of
                # temporaries from the caching allocator now, not inside the first request that brings a new voice
                self._cond.prepare_conditionals(np.zeros(10 * S3GEN_SR, dtype=np.float32) + 1e-3 * np.sin(np.arange(10 * S3GEN_SR) * 0.05).astype(np.float32), S3GEN_SR)
            return self._cond

    def prepare_conditionals(self, wav_fpath: str):
        """Reference :357-384: read the clip, run the conditioning encoders on it (24 kHz prompt mel, S3Tokenizer prompt tokens,
        CAMPPlus x-vector, VoiceEncoder speaker embedding -- all on the GPU, cbx_b200/conditioning.py) and cache the result
        under the file's base name."""
        from .conditioning import load_wav
        wav, sr = load_wav(wav_fpath)
        if self._backend is not None and hasattr(self._backend, "prepare_conditionals"):
            # host-logic tests only (injected fake backend): the fake supplies its own deterministic conditioning
            c = self._backend.prepare_conditionals(wav, sr)
            self.put_conditionals(Path(wav_fpath).name, c["t3"], c["gen"])
            return
        enc = self.conditioning_encoders()
        with torch.cuda.device(self.gpu_id):
            c = enc.prepare_conditionals(wav, sr, speech_cond_prompt_len=self.cfg.t3.speech_cond_prompt_len,
                                         exaggeration=float(os.environ.get("TTS_VOICE_EXAGGERATION_FACTOR", "0.5")))
        self.put_conditionals(Path(wav_fpath).name, c["t3"], c["gen"])

    # ------------------------------------------------------------------ helpers
    def _pinned_get(self):
        """A pinned int16 staging buffer (one S3Gen call of PCM at most) from the engine's pool, or None without a GPU."""
        if self._backend is not None or self.device_sink:
            return None, None
        try:
            return self._pinned_pool.get_nowait()
        except queue.Empty:
            # (pinned host buffer, device buffer): the pair lives in the pool for the life of the engine, so the emit path of a
            # request never allocates; every use is followed by a stream sync, so reuse across requests / streams is safe
            return (torch.empty(960 * 1100, dtype=torch.int16, pin_memory=True),
                    torch.empty(960 * 1100, dtype=torch.int16, device=f"cuda:{self.gpu_id}"))

    def _pinned_put(self, pair):
        if pair is not None and pair[0] is not None:
            self._pinned_pool.put(pair)

    def _wait_tokens(self, s: _T3Stream, n: int, token: Optional[CancellationToken], stop: Optional[threading.Event] = None):
        """Blocks until the stream holds >= n tokens or is finished; gives up when the request is cancelled / stopped or the
        scheduler has shut down (its streams would never finish)."""
        with s.cv:
            while len(s.tokens) < n and not s.finished:
                if (token is not None and token.is_cancelled()) or (stop is not None and stop.is_set()):
                    self.scheduler.cancel(s)
                    return False
                if not self.scheduler.running:
                    raise RuntimeError("engine is shutting down")
                s.cv.wait(0.05)
            if s.error:
                raise s.error
        return True

    # ------------------------------------------------------------------ the request pipeline (one thread per request)
    def _run_request(self, text, voice_id, cfg_w, temp, chunk_size, slice_len, trim_tail_ms, trim_lead_ms, overlap, fade_ms,
                     request_id, token: Optional[CancellationToken], emit, t_start):
        nat = self.native
        if self._backend is None:
            torch.cuda.set_device(self.gpu_id)
        with (torch.cuda.stream(torch.cuda.Stream()) if self._backend is None else contextlib.nullcontext()):
            vid = Path(voice_id).name if voice_id else "default"
            if vid not in self.voice_cache or (hasattr(nat, "voice_slot") and nat.voice_slot(vid) is None):   # never seen, or evicted (LRU)
                path = self.voice_manager.get_voice_path(voice_id)
                if path is None:
                    raise ValueError(f"voice '{voice_id}' not found")
                self.prepare_conditionals(path)
            # pinned for the life of the request: the native voice cache neither evicts nor rewrites a slot that an open
            # request still reads (its T3 prefills and S3Gen jobs); released when every job of the request has retired
            pinned_voice = hasattr(nat, "voice_acquire")
            voice = nat.voice_acquire(vid) if pinned_voice else self.voice_cache[vid]
            jobs = []                                         # every S3Gen job of this request (dropped on cancel)
            try:
                self._run_request_pinned(text, voice, jobs, cfg_w, temp, chunk_size, slice_len, trim_tail_ms, trim_lead_ms, overlap, fade_ms,
                                         request_id, token, emit, t_start)
            finally:
                if pinned_voice:
                    for j in jobs:       # dropped jobs are retired by the batcher's next pass; running ones finish their batch
                        j.done.wait(5.0)
                    nat.voice_release(voice)

    def _run_request_pinned(self, text, voice, jobs, cfg_w, temp, chunk_size, slice_len, trim_tail_ms, trim_lead_ms, overlap, fade_ms,
                            request_id, token: Optional[CancellationToken], emit, t_start):
        nat, sched = self.native, self.scheduler
        if True:
            chunks = split_text_into_chunks(text, chunk_size)
            if not chunks:
                emit(b"")
                return
            fade_len = int(self.sr * (fade_ms / 1000.0))
            look_ahead = max(3, int(0.2 * slice_len))
            lead = (trim_lead_ms * self.sr) // 1000
            trail = (trim_tail_ms * self.sr) // 1000
            prev_tail = None
            first_sent = False
            with self._seq_lock:      # concurrent arrivals must not swap or share sequence numbers (they feed the seed)
                self._seq += 1
                seq = self._seq
            base_seed = (self.seed << 20) ^ (zlib.crc32(str(request_id).encode()) & 0xFFFFF) ^ (seq << 8)
            fade_in = fade_out = None
            if fade_len > 0 and self._backend is None:
                from .native import fade_curves
                fade_in, fade_out = fade_curves(fade_len, torch.device("cuda", self.gpu_id))   # per request, as the reference (:867-871)

            def send(cur, n_out, tail):
                """crossfade (optional) + PCM on device, then D2H and hand the bytes to the event loop."""
                if n_out <= 0:
                    return
                kw = dict(fade_in=fade_in, fade_out=fade_out) if fade_in is not None else {}
                if pinned is not None and n_out <= pinned.shape[0]:
                    kw["out"] = pcm_dev
                pcm = nat.crossfade_pcm(cur, n_out, tail, fade_len if tail is not None else 0, **kw)
                if not first_sent:
                    trace("pcm_enqueued")
                if self.device_sink:
                    emit(int(n_out))
                    return
                # stream-ordered D2H into a pinned staging buffer: one async copy + one stream sync (a pageable .cpu() is a
                # multi-call staged copy that queues behind the launch bursts of the S3Gen thread inside the driver)
                if pinned is not None and n_out <= pinned.shape[0]:
                    pinned[:n_out].copy_(pcm, non_blocking=True)
                    if not first_sent:
                        trace("d2h_enqueued")
                    torch.cuda.current_stream().synchronize()
                    data = pinned[:n_out].numpy().tobytes()
                else:
                    data = pcm.cpu().numpy().tobytes()
                if not first_sent:
                    trace("first_pcm_host")
                emit(data)

            pinned, pcm_dev = self._pinned_get()
            cancelled = (lambda: token is not None and token.is_cancelled())
            tr = self.stats.setdefault("trace", collections.deque(maxlen=64))
            trace = (lambda label: tr.append((label, round((time.time() - t_start) * 1e3, 1))))
            trace("start")
            streams = [None] * len(chunks)
            outq = [queue.Queue() for _ in chunks]           # per chunk: (cur, last) items, then None (or an exception)
            first_slice_ready = threading.Event()            # chunk 0's first slice has been synthesised (or chunk 0 is done)
            par = len(chunks) if self.chunk_parallelism < len(chunks) <= self.chunk_parallelism_short else self.chunk_parallelism
            window = threading.Semaphore(max(1, par))   # chunks in flight ahead of the emitter
            stop = threading.Event()

            def chunk_worker(ci):
                """T3 stream + S3Gen slices of one text chunk (reference _t3_producer_task / _s3gen_producer_task per chunk)."""
                have_slot = False
                try:
                    if self._backend is None:
                        torch.cuda.set_device(self.gpu_id)
                    if not self.t3_slots.acquire(ci, lambda: cancelled() or stop.is_set()):
                        return
                    have_slot = True
                    ids = self.tokenizer.text_to_tokens(chunks[ci])[0].tolist()
                    ids = [self.cfg.t3.start_text_token] + ids + [self.cfg.t3.stop_text_token]
                    max_new = self.sampling.max_new_tokens
                    if self.sampling.tokens_per_word:
                        max_new = max(1, self.sampling.tokens_per_word * len(chunks[ci].split()))
                    # the permit travels with the stream: the scheduler releases it after the native slot is closed (natural
                    # end, cancel or error), never earlier
                    s = streams[ci] = sched.open(voice, ids, cfg_w, temp, self.sampling, base_seed + ci, max_new, on_closed=self.t3_slots.release)
                    have_slot = False
                    if ci == 0:
                        trace("t3_open")
                        with self._fp_lock:
                            self._first_pending[id(s)] = (s, slice_len + look_ahead)
                    is_first_chunk, is_last_chunk = ci == 0, ci == len(chunks) - 1
                    consumed, slice_idx, acc, prev_job = 0, 0, [], None
                    while not stop.is_set():
                        if not self._wait_tokens(s, consumed + slice_len + look_ahead, token, stop):
                            break
                        avail = len(s.tokens) - consumed
                        if avail >= slice_len + look_ahead:
                            new, last = s.tokens[consumed: consumed + slice_len], False
                        elif s.finished:
                            if avail <= 0:
                                break
                            new, last = s.tokens[consumed:], True
                        else:
                            continue
                        consumed += len(new)
                        slice_idx += 1
                        first_slice = slice_idx == 1
                        acc = acc + new if overlap == "full" else new
                        toks = list(acc)
                        if last:
                            toks = toks + [self.cfg.t3.stop_text_token]          # reference appends hp.stop_text_token (0)
                        toks = [t for t in drop_invalid_tokens(toks) if t < SPEECH_VOCAB]
                        if len(toks) == 0:
                            if last:
                                break
                            continue
                        if len(toks) < 3:
                            toks = toks + [0] * (3 - len(toks))
                        # submitted without waiting for the previous slice: its source cache is chained (job dependency), so
                        # consecutive slices of this chunk can ride in the same batch when T3 runs ahead of S3Gen
                        if ci == 0 and first_slice:
                            trace("slice1_tokens")
                            if self.hold_until == "tokens" or (self.hold_until == "auto" and self._inflight > 1):
                                first_slice_ready.set()
                        job = self.s3gen.submit(voice, toks, prev_job if overlap == "full" else None, base_seed + 7919 * ci + slice_idx,
                                                urgent=(ci == 0 and first_slice),
                                                emit_from=960 * len(prev_job.toks) if (overlap == "full" and prev_job is not None) else 0)
                        if ci == 0 and first_slice:
                            with self._fp_lock:
                                self._first_pending.pop(id(s), None)
                        prev_job = job
                        jobs.append(job)
                        outq[ci].put((job, first_slice, last))
                        if last:
                            break
                except BaseException as ex:
                    outq[ci].put(ex)
                finally:
                    if ci == 0:
                        first_slice_ready.set()
                        if streams[ci] is not None:
                            with self._fp_lock:
                                self._first_pending.pop(id(streams[ci]), None)
                    if have_slot:
                        self.t3_slots.release()
                    if streams[ci] is not None and not streams[ci].finished:
                        sched.cancel(streams[ci])
                    outq[ci].put(None)

            def opener():
                """Starts the chunk workers in order, at most `chunk_parallelism` ahead of the emitter; chunks after the
                first wait until chunk 0's first slice has been synthesised, so the first audio does not compete with their
                prefills and decode steps for the GPU."""
                for ci in range(len(chunks)):
                    while not window.acquire(timeout=0.05):
                        if stop.is_set() or cancelled():
                            break
                    else:
                        if ci == 1:
                            while not first_slice_ready.wait(0.05):
                                if stop.is_set() or cancelled():
                                    break
                        if not (stop.is_set() or cancelled()):
                            self.chunk_executor.submit(chunk_worker, ci)
                            continue
                    outq[ci].put(None)      # never started (cancelled / stopped)

            threading.Thread(target=opener, daemon=True, name="cbx-opener").start()
            try:
                for ci in range(len(chunks)):
                    prev_len = 0
                    while True:
                        item = outq[ci].get()
                        if item is None:
                            break
                        if isinstance(item, BaseException):
                            raise item
                        if cancelled():
                            continue
                        job, first_slice, last = item
                        try:
                            wav, src = job.wait()
                        except BaseException:
                            first_slice_ready.set()
                            raise
                        if not first_slice_ready.is_set():
                            trace("slice1_audio")
                        cur = wav[0]
                        if overlap == "full":                 # keep only what the previous slice has not already produced
                            full_len = cur.shape[0]
                            if not first_slice:
                                cur = cur[prev_len:]
                            prev_len = full_len
                        if ci == 0 and first_slice and lead > 0 and cur.shape[0] > lead:
                            cur = cur[lead:]
                        if ci == len(chunks) - 1 and last and trail > 0 and cur.shape[0] > trail:
                            cur = cur[:-trail]
                        n = cur.shape[0]
                        # crossfade state machine (reference :710-746)
                        if not first_sent:
                            if fade_len > 0 and n > fade_len:
                                send(cur, n - fade_len, None)
                                prev_tail = cur[n - fade_len:]
                            else:
                                send(cur, n, None)
                                prev_tail = None
                            first_sent = True
                        else:
                            can_fade = fade_len > 0 and prev_tail is not None and prev_tail.shape[0] == fade_len and n > fade_len
                            if can_fade:
                                body = n - 2 * fade_len if n > 2 * fade_len else 0
                                send(cur, fade_len + body, prev_tail)
                                prev_tail = cur[n - fade_len:]
                            else:
                                if prev_tail is not None:
                                    send(prev_tail, prev_tail.shape[0], None)
                                prev_tail = cur[n - fade_len:] if (fade_len > 0 and n > fade_len) else cur
                        if job.urgent:
                            job.consumed.set()
                        first_slice_ready.set()          # later chunks may start: the first audio no longer competes with them
                    first_slice_ready.set()
                    window.release()
                    if cancelled():
                        break
                if prev_tail is not None and prev_tail.shape[0] > 0 and not cancelled():
                    send(prev_tail, prev_tail.shape[0], None)     # reference `finally` flush (:756-760)
            finally:
                stop.set()
                self._pinned_put((pinned, pcm_dev))
                for j in jobs:
                    if not j.done.is_set():
                        j.dropped = True
                for s in streams:
                    if s is not None and not s.finished:
                        sched.cancel(s)

    # ------------------------------------------------------------------ public API (reference stream :815-968)
    async def stream(self, text: str, output_format: str, voice_id: Optional[str], cfg_guidance_weight: float,
                     synthesis_temperature: float, text_processing_chunk_size: int, audio_tokens_per_slice: int,
                     remove_trailing_milliseconds: int, remove_leading_milliseconds: int, chunk_overlap_strategy: str,
                     crossfade_duration_milliseconds: int, request_id: str,
                     cancellation_token: Optional[CancellationToken] = None) -> AsyncGenerator[bytes, None]:
        if not self._ready:
            raise RuntimeError(f"TTS Engine on GPU {self.gpu_id} is not ready")
        if not self.healthy():
            # a device fault is sticky for the process: fail the request up front (the worker's per-job handler logs it and keeps
            # answering, src/worker.py:54-56) instead of queueing work that can only fail slice by slice
            raise RuntimeError(f"TTS Engine on GPU {self.gpu_id} lost its CUDA context (device fault): the worker process must be restarted")
        if output_format not in ("raw_pcm", "wav"):
            raise ValueError(f"Unsupported format on the B200 path: {output_format} (containers are muxed by the caller)")
        if cancellation_token is None:
            # an abandoned generator (consumer stops iterating) must stop the worker too: internal token, set in `finally`
            cancellation_token = CancellationToken()
        async with self.tts_semaphore:
            loop = asyncio.get_running_loop()
            q: asyncio.Queue = asyncio.Queue(maxsize=int(os.environ.get("TTS_PCM_CHUNK_QUEUE_MAX_SIZE", "3")))
            t_start = time.time()
            DONE = object()

            def emit(b):   # called from the request thread; blocks it when the consumer is slow (bounded queue)
                f = asyncio.run_coroutine_threadsafe(q.put(b), loop)
                while True:
                    try:
                        return f.result(timeout=0.1)
                    except concurrent.futures.TimeoutError:
                        if cancellation_token is not None and cancellation_token.is_cancelled():
                            f.cancel()
                            raise _Cancelled()

            def work():
                with self._inflight_lock:
                    self._inflight += 1
                try:
                    self._run_request(text, voice_id, cfg_guidance_weight, synthesis_temperature, text_processing_chunk_size,
                                      audio_tokens_per_slice, remove_trailing_milliseconds, remove_leading_milliseconds,
                                      chunk_overlap_strategy, crossfade_duration_milliseconds, request_id, cancellation_token,
                                      emit, t_start)
                    emit(DONE)
                except _Cancelled:
                    pass
                except BaseException as ex:
                    try:
                        emit(ex)
                    except _Cancelled:
                        pass
                finally:
                    with self._inflight_lock:
                        self._inflight -= 1

            fut = loop.run_in_executor(self.request_executor, work)
            first = True
            try:
                if output_format == "wav":
                    yield _wav_header(self.sr)
                while True:
                    item = await q.get()
                    if item is DONE:
                        break
                    if isinstance(item, BaseException):
                        raise item
                    if first and len(item):
                        self.stats["first_chunk_ms"].append((time.time() - t_start) * 1e3)
                        first = False
                    yield item
            finally:
                if cancellation_token is not None and not fut.done():
                    cancellation_token.cancel()
                while not fut.done():        # drain so the worker thread is never blocked on a full queue
                    try:
                        q.get_nowait()
                    except asyncio.QueueEmpty:
                        await asyncio.sleep(0.005)


def _wav_header(sr: int, channels: int = 1, bits: int = 16) -> bytes:
    """Streaming WAV header with unknown size 0xFFFFFFFF (reference src/audio_encoding.py:97-115)."""
    import struct
    byte_rate = sr * channels * bits // 8
    h = struct.pack("<4sL4s", b"RIFF", 0xFFFFFFFF, b"WAVE")
    h += struct.pack("<4sLHHLLHH", b"fmt ", 16, 1, channels, sr, byte_rate, channels * bits // 8, bits)
    h += struct.pack("<4sL", b"data", 0xFFFFFFFF)
    return h
