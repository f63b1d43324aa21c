"""chatterbox.models.voice_encoder.VoiceEncoder.embeds_from_wavs (reference call site src/tts_streaming.py:374)."""
import numpy as np
import torch


class VoiceEncoder(torch.nn.Module):
    """Speaker embedding (k, 256), L2-normalised: 40-mel partials through the 3 x LSTM-256 on the GPU (cbx_b200/conditioning.py)."""

    def __init__(self, backend=None):
        super().__init__()
        self._b = backend

    def embeds_from_wavs(self, wavs, sample_rate=16000, as_spk=False, **kw):
        if self._b is None:
            raise RuntimeError("VoiceEncoder needs the engine backend (ChatterboxTTS.from_local builds it)")
        enc = self._b.encoders("ve.embeds_from_wavs")
        out = []
        for w in wavs:
            w = np.asarray(w.detach().cpu().numpy() if torch.is_tensor(w) else w, dtype=np.float32).reshape(-1)
            x = enc.resample(enc._dev_wave(w), int(sample_rate), 16000)
            out.append(enc.voice_embed(x.contiguous()).cpu().numpy())
        return np.stack(out).astype(np.float32)
