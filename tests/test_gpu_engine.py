"""End-to-end GPU tests through the public engine API (host text in, host PCM bytes out)."""
import asyncio

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(tiny_cfg):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cbx_b200.engine import TextToSpeechEngine, SamplingDefaults
    from cbx_b200.weights import random_state_dict
    # The CFM block tail has two kernel paths chosen by the ROWS of a batched call (fused tail from 4096 rows, csrc/flow.cu):
    # which requests share a batch depends on timing, so bit-exact reproducibility is a property of a PINNED path.  These
    # tests pin the fused tail for every call (that also puts it on the end-to-end path of a small model).
    import os
    old = os.environ.get("CBX_CFM_TAIL_MIN_ROWS")
    os.environ["CBX_CFM_TAIL_MIN_ROWS"] = "0"
    eng = TextToSpeechEngine("cuda:0", cfg=tiny_cfg, state_dict=random_state_dict(tiny_cfg, 0), concurrent_requests=4,
                             sampling=SamplingDefaults(tokens_per_word=10), seed=0,
                             native_kwargs=dict(max_s3_tokens=400, n_lanes=4))
    asyncio.run(eng.ainit())
    yield eng
    eng.shutdown()
    if old is None:
        os.environ.pop("CBX_CFM_TAIL_MIN_ROWS", None)
    else:
        os.environ["CBX_CFM_TAIL_MIN_ROWS"] = old


REQ = dict(output_format="raw_pcm", voice_id=None, cfg_guidance_weight=0.5, synthesis_temperature=0.8, text_processing_chunk_size=150,
           audio_tokens_per_slice=35, remove_trailing_milliseconds=0, remove_leading_milliseconds=0, chunk_overlap_strategy="full",
           crossfade_duration_milliseconds=30)


async def _collect(eng, text, rid, **over):
    kw = dict(REQ)
    kw.update(over)
    out = b""
    async for c in eng.stream(text=text, request_id=rid, **kw):
        out += c
    return np.frombuffer(out, dtype=np.int16)


def test_stream_is_deterministic_and_sized(engine):
    text = "alpha bravo charlie delta echo foxtrot golf hotel india juliet kilo lima. mike november oscar papa."
    a = asyncio.run(_collect(engine, text, "req-1"))
    engine._seq = 0
    b = asyncio.run(_collect(engine, text, "req-1"))
    engine._seq = 0
    assert a.shape[0] > 24000 and a.shape[0] % 2 == 0
    assert np.array_equal(a, b), "fixed seed + fixed request id must reproduce the PCM bit for bit"
    assert np.abs(a).max() <= 32767
    assert engine.healthy() and engine.native.healthy()      # cbx_engine_health: no sticky device fault


def test_concurrent_streams_match_sequential(engine):
    """T3 rows of concurrent requests are batched and S3Gen calls run on parallel lanes: results must not change."""
    texts = [f"stream number {i} says hello to the world and keeps talking for a while." for i in range(4)]
    seq = []
    for i, t in enumerate(texts):
        engine._seq = 100 + i
        seq.append(asyncio.run(_collect(engine, t, f"r{i}")))

    async def both():
        outs = []
        tasks = []
        for i, t in enumerate(texts):
            engine._seq = 100 + i      # the sequence number is folded into the seed when the request starts
            tasks.append(asyncio.create_task(_collect(engine, t, f"r{i}")))
            await asyncio.sleep(0.05)
        for tk in tasks:
            outs.append(await tk)
        return outs
    conc = asyncio.run(both())
    for a, b in zip(seq, conc):
        assert a.shape == b.shape
        # T3 is batch invariant (same tokens) and a batched S3Gen pass reproduces the single call to float rounding, so the
        # audio of a request must not depend on what else is running: int16 PCM within a few LSB
        d = np.abs(a.astype(np.int32) - b.astype(np.int32))
        assert d.max() <= 64 and d.mean() < 1.0, f"concurrent vs sequential PCM: max {d.max()} mean {d.mean():.3f}"


def test_zero_overlap_and_wav(engine):
    text = "one two three four five six seven eight nine ten eleven twelve."
    z = asyncio.run(_collect(engine, text, "z", chunk_overlap_strategy="zero"))
    assert z.shape[0] > 0
    w = asyncio.run(_collect(engine, text, "w", output_format="wav"))
    assert w.tobytes()[:4] == b"RIFF"


def test_alignment_eos_control_through_the_engine(engine):
    """cbx_t3_set_alignment_eos on the engine's own native handle (what `alignment_eos=True` / CBX_ALIGNMENT_EOS=1 do at load):
    with random-init weights the attention over the text is flat, so the analyzer only ever SUPPRESSES the stop token -- which
    these weights never sample -- and the stream's PCM must be bit-identical to the run without the control, while the analyzer
    has followed the generated frames of the stream's T3 slots."""
    text = "alpha bravo charlie delta echo foxtrot golf hotel india juliet kilo lima."
    engine._seq = 7
    a = asyncio.run(_collect(engine, text, "req-al"))
    engine.native.t3_set_alignment_eos(True, 9)
    try:
        engine._seq = 7
        b = asyncio.run(_collect(engine, text, "req-al"))
        states = [engine.native.t3_alignment_peek(s, rows=False)[0] for s in range(engine.native_kwargs["max_streams"])]
        engine._seq = 7
    finally:
        engine.native.t3_set_alignment_eos(False, 9)
    c = asyncio.run(_collect(engine, text, "req-al"))
    assert np.array_equal(a, b) and np.array_equal(a, c)
    followed = [st for st in states if st["on"] == 1 and st["rows"] > 20]
    assert followed, f"no T3 slot was followed by the analyzer: {states}"
    assert all(st["frame_pos"] == st["rows"] - 1 and st["i0"] == 34 for st in followed)
