"""chatterbox.tts.ChatterboxTTS (reference call sites src/tts_streaming.py:44, :252-259, :399-404)."""
import os
import types

from ._backend import Backend
from .models.s3gen import S3Gen, S3GEN_SR
from .models.t3 import T3
from .models.t3.modules.cond_enc import T3Cond
from .models.tokenizers import EnTokenizer
from .models.voice_encoder import VoiceEncoder


class ChatterboxTTS:
    """Attributes the reference engine reads: sr, t3, s3gen, ve, tokenizer, conds (truthy, with .t3 / .gen)."""

    def __init__(self, backend: Backend, ckpt_dir: str):
        from cbx_b200.weights import synthetic_conditionals
        self.backend = backend
        self.sr = S3GEN_SR
        self.device = backend.device
        self.t3 = T3(backend)
        self.s3gen = S3Gen(backend)
        self.ve = VoiceEncoder(backend)
        self.tokenizer = EnTokenizer(os.path.join(ckpt_dir or "", "tokenizer.json"), backend.cfg.t3.text_vocab)
        c = synthetic_conditionals(backend.cfg)          # stands in for conds.pt (built-in voice)
        self.conds = types.SimpleNamespace(t3=T3Cond(**c["t3"]), gen=dict(c["gen"]))

    @classmethod
    def from_local(cls, ckpt_dir, device) -> "ChatterboxTTS":
        return cls(Backend(str(ckpt_dir) if ckpt_dir is not None else "", str(device)), str(ckpt_dir) if ckpt_dir is not None else "")
