"""Level-2 drop-in: the UNMODIFIED reference engine (/root/reference/src/tts_streaming.py -- ainit() with its
torch.compile wrappers and warm-up calls, prepare_conditionals(), stream() with its three pipeline tasks) runs on the
`chatterbox`-named shim of this repository (chatterbox-tts_b200/chatterbox/), i.e. on NativeEngine's interface.  Here that
interface is the deterministic FakeNative (no GPU in this container), so what is pinned is the shim's protocol: generator
semantics of T3.inference_stream, token / tensor shapes, cache_source threading, S3Gen call conventions, conditionals.
The PCM must equal the golden PCM the same reference engine produced over the reference-side FakeTTS.
Needs /root/reference: skipped where it does not exist.  Only librosa and pysbd (absent third-party modules) are stubbed."""
import asyncio
import os
import sys
import tempfile
import types
import zlib

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference checkout not present on this machine")
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "engine_pcm.npz"))


@pytest.fixture(scope="module")
def ref_engine_module():
    from fake_backend import FakeNative
    from cbx_b200.text_processing import _sentences
    os.environ.setdefault("API_KEY", "test-key")
    if REF not in sys.path:
        sys.path.insert(0, REF)

    class Segmenter:
        def __init__(self, language="en", clean=False):
            pass

        def segment(self, text):
            return _sentences(text)

    def load(path, sr=None):
        import scipy.io.wavfile as wavfile
        r, d = wavfile.read(path)
        d = d.astype(np.float32) / 32768.0
        if sr and sr != r:
            d = np.interp(np.arange(0, len(d), r / sr), np.arange(len(d)), d).astype(np.float32)
        return d, sr or r

    def resample(y, orig_sr, target_sr):
        return np.interp(np.arange(0, len(y), orig_sr / target_sr), np.arange(len(y)), y).astype(np.float32)

    saved = {k: sys.modules.get(k) for k in ("librosa", "pysbd", "src.tts_streaming")}
    sys.modules["librosa"] = types.SimpleNamespace(load=load, resample=resample)
    sys.modules["pysbd"] = types.SimpleNamespace(Segmenter=Segmenter)
    sys.modules.pop("src.tts_streaming", None)        # another test may have installed its own stand-in
    for k in [k for k in sys.modules if k == "chatterbox" or k.startswith("chatterbox.")]:
        del sys.modules[k]
    import chatterbox                                   # THIS repository's shim (chatterbox-tts_b200/ is on sys.path)
    assert "chatterbox-tts_b200" in chatterbox.__file__
    chatterbox.set_backend_factory(lambda ckpt, device: FakeNative())
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    import src.tts_streaming as M                       # the reference, unmodified
    yield M
    os.chdir(cwd)
    chatterbox.set_backend_factory(None)
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


async def _collect(M, eng, sc, text, voice_id=None):
    out = b""
    tok = M.CancellationToken(asyncio.get_running_loop())
    async for chunk in eng.stream(text=text, output_format="raw_pcm", voice_id=voice_id, cfg_guidance_weight=0.5,
                                  synthesis_temperature=0.8, text_processing_chunk_size=sc["chunk"], audio_tokens_per_slice=sc["slice"],
                                  remove_trailing_milliseconds=sc["trail"], remove_leading_milliseconds=sc["lead"],
                                  chunk_overlap_strategy=sc["overlap"], crossfade_duration_milliseconds=sc["fade"],
                                  request_id=sc["name"], cancellation_token=tok):
        out += chunk
    return np.frombuffer(out, dtype=np.int16)


def test_unmodified_reference_engine_runs_on_the_shim(ref_engine_module, tmp_path):
    from fake_backend import SCENARIOS, scenario_text
    M = ref_engine_module

    async def run():
        eng = M.TextToSpeechEngine("cpu")
        await eng.ainit()                       # from_local -> shim, torch.compile wrappers, T3 + S3Gen warm-up: all through the shim
        assert eng.tts.sr == 24000 and eng.tts.t3.hp.start_text_token == 255 and eng.tts.t3.hp.stop_text_token == 0
        res = {}
        for name in ("full_fade30", "zero_fade30", "trims_slice20", "long_eos"):
            sc = next(s for s in SCENARIOS if s["name"] == name)
            res[name] = await _collect(M, eng, sc, scenario_text(sc["words"]))
        # voice conditioning (reference prepare_conditionals :357-384) runs the GPU encoders behind embed_ref / tokenizer.forward /
        # embeds_from_wavs: on this CPU-only fake backend it must fail loudly, never invent a voice (GPU: tests/test_gpu_shim.py)
        import scipy.io.wavfile as wavfile
        wav = tmp_path / "bob.wav"
        wavfile.write(str(wav), 24000, (np.sin(np.arange(24000 * 3) * 0.03) * 9000).astype(np.int16))
        with pytest.raises(RuntimeError, match="conditioning-encoder weights"):
            eng.prepare_conditionals(str(wav))
        assert "bob.wav" not in eng.voice_cache
        return res

    res = asyncio.run(run())
    for k, pcm in res.items():
        assert pcm.shape[0] == int(GOLD[k + "_len"][0]), k
        assert np.array_equal(pcm[:4000], GOLD[k + "_head"]), k
        assert zlib.crc32(pcm.tobytes()) == int(GOLD[k + "_crc"][0]), k


def test_generator_close_releases_the_stream(ref_engine_module):
    """An abandoned generator (cancel path, reference :505-519) must close its T3 stream: KV pages go back to the pool."""
    import torch
    from chatterbox.tts import ChatterboxTTS
    tts = ChatterboxTTS.from_local("", "cpu")
    nat = tts.backend.native
    text = torch.tensor([[255, 5, 685, 9, 685, 11, 0]] * 2)
    gen = tts.t3.inference_stream(t3_cond=tts.conds.t3, text_tokens=text, max_new_tokens=30, temperature=0.8, cfg_weight=0.5)
    first = next(gen)
    assert first.shape == (1, 1) and len(nat.streams) == 1
    gen.close()
    assert len(nat.streams) == 0
    toks = torch.cat(list(tts.t3.inference_stream(t3_cond=tts.conds.t3, text_tokens=text, max_new_tokens=30)), dim=1)
    assert toks.shape == (1, 30) and len(nat.streams) == 0
    wav, src = tts.s3gen.inference(speech_tokens=toks[toks < 6561][:12], ref_dict=tts.conds.gen, cache_source=torch.zeros(1, 1, 0))
    assert wav.shape == (1, 960 * 12) and src.shape == (1, 1, 960 * 12)
