"""chatterbox.models.s3tokenizer: constants and drop_invalid_tokens (reference call sites src/tts_streaming.py:39, :667)."""
import torch

S3_SR = 16_000
SPEECH_VOCAB_SIZE = 6561
SOS = SPEECH_VOCAB_SIZE
EOS = SPEECH_VOCAB_SIZE + 1


def drop_invalid_tokens(x):
    """Keeps what lies between the first SOS (exclusive) and the first EOS (exclusive)."""
    assert len(x.shape) <= 2 and (x.dim() == 1 or x.shape[0] == 1), "only batch size of one allowed for now"
    x = x.reshape(1, -1)
    s = int((x == SOS).nonzero(as_tuple=True)[1][0]) + 1 if (x == SOS).any() else 0
    e = int((x == EOS).nonzero(as_tuple=True)[1][0]) if (x == EOS).any() else None
    return x[0, s:e]
