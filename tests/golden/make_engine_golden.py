#!/usr/bin/env python
"""Generates tests/golden/engine_pcm.npz by running the UNMODIFIED reference engine
(/root/reference/src/tts_streaming.py: stream(), the three pipeline tasks, _AudioProcessor.to_pcm) on the CPU
with the deterministic FakeModel behind the `chatterbox` import surface.  Needs /root/reference (this
container only); the committed .npz is what travels.  Missing third-party modules (librosa, pysbd, chatterbox)
are stubbed in sys.modules; no reference file is copied or modified."""
import asyncio
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "chatterbox-tts_b200"))
sys.path.insert(0, "/root/reference")
from fake_backend import FakeModel, SCENARIOS, scenario_text  # noqa: E402
from cbx_b200.text_processing import SyntheticTokenizer, _sentences  # noqa: E402


def stub_modules():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("librosa")

    class Segmenter:
        def __init__(self, language="en", clean=False):
            pass

        def segment(self, text):
            return _sentences(text)

    mod("pysbd", Segmenter=Segmenter)

    def drop_invalid_tokens(x):
        """upstream chatterbox.models.s3tokenizer.drop_invalid_tokens"""
        assert len(x.shape) <= 2 and (x.dim() == 1 or x.shape[0] == 1)
        x = x.reshape(1, -1)
        SOS, EOS = 6561, 6562
        s = int((x == SOS).nonzero(as_tuple=True)[1][0]) + 1 if (x == SOS).any() else 0
        e = int((x == EOS).nonzero(as_tuple=True)[1][0]) if (x == EOS).any() else None
        return x[0, s:e]

    class T3Cond:
        def __init__(self, **kw):
            self.__dict__.update(kw)

        def to(self, device=None):
            return self

    mod("chatterbox")
    mod("chatterbox.models")
    mod("chatterbox.models.t3", T3=object)
    mod("chatterbox.models.t3.modules")
    mod("chatterbox.models.t3.modules.cond_enc", T3Cond=T3Cond)
    mod("chatterbox.models.s3tokenizer", S3_SR=16000, drop_invalid_tokens=drop_invalid_tokens)
    mod("chatterbox.models.s3gen", S3GEN_SR=24000, S3Gen=object)
    mod("chatterbox.models.tokenizers", EnTokenizer=object)
    mod("chatterbox.models.voice_encoder", VoiceEncoder=object)
    mod("chatterbox.tts", ChatterboxTTS=object)
    return T3Cond


class FakeTTS:
    sr = 24000

    def __init__(self, T3Cond):
        hp = types.SimpleNamespace(start_text_token=255, stop_text_token=0, speech_cond_prompt_len=150)
        tts = self

        class T3:
            def inference_stream(self, t3_cond, text_tokens, max_new_tokens, temperature=0.8, cfg_weight=0.5):
                ids = text_tokens[0, 1:-1].tolist()
                for t in FakeModel.tokens(ids, max_new_tokens):
                    yield torch.tensor([[t]])

        class S3:
            def inference(self, speech_tokens, ref_dict, cache_source):
                return FakeModel.s3gen(speech_tokens, cache_source if cache_source.shape[-1] else None)

        self.t3 = T3()
        self.t3.hp = hp
        self.s3gen = S3()
        self.tokenizer = SyntheticTokenizer()
        self.conds = types.SimpleNamespace(t3=T3Cond(), gen={})


async def run(engine, sc):
    out = b""
    tok = engine_mod.CancellationToken(asyncio.get_running_loop())
    async for chunk in engine.stream(text=scenario_text(sc["words"]), output_format="raw_pcm", voice_id=None, cfg_guidance_weight=0.5,
                                     synthesis_temperature=0.8, text_processing_chunk_size=sc["chunk"], audio_tokens_per_slice=sc["slice"],
                                     remove_trailing_milliseconds=sc["trail"], remove_leading_milliseconds=sc["lead"],
                                     chunk_overlap_strategy=sc["overlap"], crossfade_duration_milliseconds=sc["fade"],
                                     request_id=sc["name"], cancellation_token=tok):
        out += chunk
    return np.frombuffer(out, dtype=np.int16)


if __name__ == "__main__":
    os.environ.setdefault("API_KEY", "x")
    os.chdir(tempfile.mkdtemp())
    T3Cond = stub_modules()
    import src.tts_streaming as engine_mod  # the reference, unmodified
    eng = engine_mod.TextToSpeechEngine("cpu")
    eng.tts = FakeTTS(T3Cond)
    eng._initialization_state = engine_mod.InitializationState.READY
    res = {}
    for sc in SCENARIOS:
        res[sc["name"]] = asyncio.run(run(eng, sc))
        print(sc["name"], res[sc["name"]].shape, int(np.abs(res[sc["name"]]).max()))
    import zlib
    out = {}
    for k, a in res.items():   # compact fixture: length, CRC32 of all bytes, the first 4000 samples and every 53rd sample
        out[k + "_len"] = np.array([a.shape[0]])
        out[k + "_crc"] = np.array([zlib.crc32(a.tobytes())], dtype=np.uint32)
        out[k + "_head"] = a[:4000]
        out[k + "_stride"] = a[::53]
    np.savez_compressed(os.path.join(HERE, "engine_pcm.npz"), **out)
