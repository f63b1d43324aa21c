"""chatterbox.models.voice_encoder.VoiceEncoder.embeds_from_wavs (reference call site src/tts_streaming.py:374)."""
import numpy as np
import torch

from ..._backend import seed_of


class VoiceEncoder(torch.nn.Module):
    """Speaker embedding (k, 256), L2-normalised.  Seeded from the clip until the encoder of SURVEY 8f.1 exists on the GPU."""

    def __init__(self, backend=None):
        super().__init__()
        self._b = backend

    def embeds_from_wavs(self, wavs, sample_rate=16000, as_spk=False, **kw):
        if self._b is not None:
            self._b.require_synthetic("ve.embeds_from_wavs")
        out = []
        for w in wavs:
            w = np.asarray(w, dtype=np.float32).reshape(-1)
            g = torch.Generator().manual_seed(seed_of(w[:4000], [len(w)]))
            e = torch.randn(256, generator=g)
            out.append((e / e.norm()).numpy())
        return np.stack(out).astype(np.float32)
