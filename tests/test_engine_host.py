"""Host logic of the streaming engine (slice schedule, overlap, EOS/filter/pad rules, trims, crossfade state
machine, PCM conversion) pinned bit-exactly against PCM produced by the UNMODIFIED reference engine
(src/tts_streaming.py) driving the same deterministic FakeModel -- see tests/golden/make_engine_golden.py."""
import asyncio
import os
import zlib

import numpy as np
import pytest

from fake_backend import FakeNative, SCENARIOS, scenario_text

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "engine_pcm.npz"))


async def _run(sc, cancel_after=None):
    from cbx_b200.engine import TextToSpeechEngine, CancellationToken
    eng = TextToSpeechEngine("cpu", backend=FakeNative(), concurrent_requests=2)
    await eng.ainit()
    tok = CancellationToken()
    out, n = b"", 0
    agen = eng.stream(text=scenario_text(sc["words"]), output_format="raw_pcm", voice_id=None, cfg_guidance_weight=0.5,
                      synthesis_temperature=0.8, text_processing_chunk_size=sc["chunk"], audio_tokens_per_slice=sc["slice"],
                      remove_trailing_milliseconds=sc["trail"], remove_leading_milliseconds=sc["lead"],
                      chunk_overlap_strategy=sc["overlap"], crossfade_duration_milliseconds=sc["fade"],
                      request_id=sc["name"], cancellation_token=tok)
    try:
        async for chunk in agen:
            out += chunk
            n += 1
            if cancel_after is not None and n >= cancel_after:
                tok.cancel()     # what the worker does on a cancel_request broadcast (reference src/worker.py:116-122)
                break
    finally:
        await agen.aclose()
        eng.shutdown()
    return np.frombuffer(out, dtype=np.int16)


@pytest.mark.parametrize("sc", SCENARIOS, ids=[s["name"] for s in SCENARIOS])
def test_engine_matches_reference_pcm(sc):
    pcm = asyncio.run(_run(sc))
    k = sc["name"]
    assert pcm.shape[0] == int(GOLD[k + "_len"][0])
    assert np.array_equal(pcm[:4000], GOLD[k + "_head"])
    assert np.array_equal(pcm[::53], GOLD[k + "_stride"])
    assert zlib.crc32(pcm.tobytes()) == int(GOLD[k + "_crc"][0])


def test_cancel_stops_stream():
    sc = SCENARIOS[-1]
    pcm = asyncio.run(_run(sc, cancel_after=2))
    assert 0 < pcm.shape[0] < int(GOLD[sc["name"] + "_len"][0])


def test_concurrent_requests_are_independent():
    async def both():
        from cbx_b200.engine import TextToSpeechEngine
        eng = TextToSpeechEngine("cpu", backend=FakeNative(), concurrent_requests=2)
        await eng.ainit()

        async def one(sc):
            out = b""
            async for c in eng.stream(text=scenario_text(sc["words"]), output_format="raw_pcm", voice_id=None, cfg_guidance_weight=0.5,
                                      synthesis_temperature=0.8, text_processing_chunk_size=sc["chunk"], audio_tokens_per_slice=sc["slice"],
                                      remove_trailing_milliseconds=sc["trail"], remove_leading_milliseconds=sc["lead"],
                                      chunk_overlap_strategy=sc["overlap"], crossfade_duration_milliseconds=sc["fade"], request_id=sc["name"]):
                out += c
            return np.frombuffer(out, dtype=np.int16)
        r = await asyncio.gather(one(SCENARIOS[0]), one(SCENARIOS[3]))
        eng.shutdown()
        return r
    a, b = asyncio.run(both())
    assert zlib.crc32(a.tobytes()) == int(GOLD[SCENARIOS[0]["name"] + "_crc"][0])
    assert zlib.crc32(b.tobytes()) == int(GOLD[SCENARIOS[3]["name"] + "_crc"][0])


def test_wav_header_and_empty_text():
    async def go():
        from cbx_b200.engine import TextToSpeechEngine
        eng = TextToSpeechEngine("cpu", backend=FakeNative())
        await eng.ainit()
        chunks = [c async for c in eng.stream(text="hello there friend.", output_format="wav", voice_id=None, cfg_guidance_weight=0.5,
                                              synthesis_temperature=0.8, text_processing_chunk_size=150, audio_tokens_per_slice=35,
                                              remove_trailing_milliseconds=0, remove_leading_milliseconds=0, chunk_overlap_strategy="full",
                                              crossfade_duration_milliseconds=30, request_id="w")]
        empty = [c async for c in eng.stream(text="   ", output_format="raw_pcm", voice_id=None, cfg_guidance_weight=0.5,
                                             synthesis_temperature=0.8, text_processing_chunk_size=150, audio_tokens_per_slice=35,
                                             remove_trailing_milliseconds=0, remove_leading_milliseconds=0, chunk_overlap_strategy="full",
                                             crossfade_duration_milliseconds=30, request_id="e")]
        eng.shutdown()
        return chunks, empty
    chunks, empty = asyncio.run(go())
    h = chunks[0]
    assert len(h) == 44 and h[:4] == b"RIFF" and h[8:12] == b"WAVE" and h[4:8] == b"\xff\xff\xff\xff"   # reference audio_encoding.py:97-115
    assert int.from_bytes(h[24:28], "little") == 24000 and int.from_bytes(h[34:36], "little") == 16
    assert empty == [b""]


def test_opens_pending_together_are_prefilled_in_one_pass():
    """The scheduler's opener thread hands opens that are pending at the same moment to ONE t3_open_batch call (8 requests arriving
    together, chunks 1.. of a request): the audio is what one-by-one opens produce, and a batch that cannot be opened as a whole
    falls back to individual opens, each with its own verdict."""
    class BatchNative(FakeNative):
        def __init__(self, fail_batches=False):
            super().__init__()
            self.batch_sizes, self.fail_batches = [], fail_batches

        def t3_open_batch(self, reqs):
            self.batch_sizes.append(len(reqs))
            if self.fail_batches:
                raise RuntimeError("not enough pages for the whole batch")
            return [self.t3_open(*r) for r in reqs]

    async def many(native):
        from cbx_b200.engine import TextToSpeechEngine
        eng = TextToSpeechEngine("cpu", backend=native, concurrent_requests=4)
        await eng.ainit()
        eng.scheduler.open_gather_s = 0.05          # a wide window: the four requests below must meet in it

        async def one(sc):
            out = b""
            async for c in eng.stream(text=scenario_text(sc["words"]), output_format="raw_pcm", voice_id=None, cfg_guidance_weight=0.5,
                                      synthesis_temperature=0.8, text_processing_chunk_size=sc["chunk"], audio_tokens_per_slice=sc["slice"],
                                      remove_trailing_milliseconds=sc["trail"], remove_leading_milliseconds=sc["lead"],
                                      chunk_overlap_strategy=sc["overlap"], crossfade_duration_milliseconds=sc["fade"], request_id=sc["name"]):
                out += c
            return np.frombuffer(out, dtype=np.int16)
        scs = [SCENARIOS[0], SCENARIOS[3], SCENARIOS[1], SCENARIOS[0]]
        r = await asyncio.gather(*[one(sc) for sc in scs])
        eng.shutdown()
        return scs, r

    for fail in (False, True):
        native = BatchNative(fail_batches=fail)
        scs, res = asyncio.run(many(native))
        for sc, pcm in zip(scs, res):
            assert zlib.crc32(pcm.tobytes()) == int(GOLD[sc["name"] + "_crc"][0]), (sc["name"], fail)
        assert max(native.batch_sizes) >= 2, native.batch_sizes


def test_a_lost_device_context_refuses_new_requests():
    """cbx_engine_health: after a sticky device fault the engine refuses requests up front with a clear error (the worker's
    per-job handler logs it and keeps answering, reference src/worker.py:54-56) instead of failing them slice by slice."""
    async def go():
        from cbx_b200.engine import TextToSpeechEngine
        nat = FakeNative()
        state = {"ok": True}
        nat.healthy = lambda: state["ok"]
        eng = TextToSpeechEngine("cpu", backend=nat, concurrent_requests=1)
        await eng.ainit()
        sc = SCENARIOS[0]
        kw = dict(text=scenario_text(sc["words"]), output_format="raw_pcm", voice_id=None, cfg_guidance_weight=0.5, synthesis_temperature=0.8,
                  text_processing_chunk_size=sc["chunk"], audio_tokens_per_slice=sc["slice"], remove_trailing_milliseconds=sc["trail"],
                  remove_leading_milliseconds=sc["lead"], chunk_overlap_strategy=sc["overlap"], crossfade_duration_milliseconds=sc["fade"], request_id="h")
        try:
            n = 0
            async for c in eng.stream(**kw):
                n += len(c)
            assert n > 0 and eng.healthy()
            state["ok"] = False
            with pytest.raises(RuntimeError, match="lost its CUDA context"):
                async for c in eng.stream(**kw):
                    pass
        finally:
            eng.shutdown()
    asyncio.run(go())


def test_t3_priority_follows_the_first_slice():
    """The scheduler asks `high_priority()` before every decode round and tells the native engine only about CHANGES; wired as in
    the product path (high unless a request's first slice is queued or running in S3Gen) a request starts its decode on the
    high-priority stream, drops to low while its first slice is synthesised and returns to high afterwards."""
    async def go():
        from cbx_b200.engine import TextToSpeechEngine
        nat = FakeNative()
        calls = []
        nat.t3_set_priority = lambda high: calls.append(bool(high))
        eng = TextToSpeechEngine("cpu", backend=nat, concurrent_requests=1)
        await eng.ainit()
        seen = []

        def high():
            v = not eng.s3gen.urgent_inflight()
            seen.append(v)
            return v
        eng.scheduler.high_priority = high
        sc = SCENARIOS[-1]
        try:
            async for _ in eng.stream(text=scenario_text(sc["words"]), output_format="raw_pcm", voice_id=None, cfg_guidance_weight=0.5,
                                      synthesis_temperature=0.8, text_processing_chunk_size=sc["chunk"], audio_tokens_per_slice=sc["slice"],
                                      remove_trailing_milliseconds=sc["trail"], remove_leading_milliseconds=sc["lead"],
                                      chunk_overlap_strategy=sc["overlap"], crossfade_duration_milliseconds=sc["fade"], request_id="prio"):
                pass
        finally:
            eng.shutdown()
        assert calls and calls[0] is True, "the first decode rounds run on the high-priority stream"
        assert all(a != b for a, b in zip(calls, calls[1:])), "only changes are sent to the native engine"
        assert len(seen) > len(calls), "asked every round, told only on change"
        assert not eng.s3gen.urgent_inflight()
    asyncio.run(go())
