"""Deterministic stand-in for the model, used ONLY to pin the engine's host logic (slice schedule, overlap
handling, trimming, crossfade, PCM conversion) against the reference's own src/tts_streaming.py.
`FakeModel` defines the token stream and the token->audio map; `FakeTTS` exposes it through the
`chatterbox` object surface for the reference engine, `FakeNative` through NativeEngine's interface
for ours.  Nothing here is a CPU fallback of the product: it produces sinusoids, not speech."""
import numpy as np
import torch

SPACE_ID = 685   # SyntheticTokenizer id of ' '


class FakeModel:
    @staticmethod
    def n_tokens(text_ids):
        ids = [int(t) for t in text_ids]
        return 10 * (sum(1 for t in ids if t == SPACE_ID) + 1)

    @staticmethod
    def tokens(text_ids, max_new):
        ids = [int(t) for t in text_ids]
        n = min(FakeModel.n_tokens(ids), max_new)
        s = sum(ids) * 31
        out = []
        for k in range(n):
            t = (s + 7 * k * k + 3 * k) % 6561
            if k % 23 == 11:
                t = 6600 + (k % 50)          # invalid speech ids: must be filtered (< 6561 rule)
            out.append(t)
        if s % 5 == 0 and n > 60:
            out[57] = 6562                   # an early EOS inside a stream
            out = out[:58]
        return out

    @staticmethod
    def s3gen(tokens, cache_source):
        tok = torch.as_tensor(tokens, dtype=torch.long).reshape(-1)
        n = tok.numel()
        j = torch.arange(960 * n, dtype=torch.float32)
        tj = tok.repeat_interleave(960).float()
        src = 0.3 * torch.sin(0.01 * j + 0.1 * tj)
        m = 0 if cache_source is None else cache_source.shape[-1]
        if m:
            src[:m] = cache_source.reshape(-1)[:m].float()
        wav = 0.8 * torch.sin(0.002 * j * (1 + tj % 5)) + 0.5 * src
        return wav.reshape(1, -1), src.reshape(1, 1, -1)


class FakeNative:
    """NativeEngine's interface over FakeModel (CPU tensors)."""
    is_fake = True
    device = 0

    def __init__(self):
        self.streams, self.next_slot = {}, 0

    def t3_open(self, voice, text_ids, cfg_weight, temperature, rep_penalty, min_p, top_p, seed, max_new):
        slot = self.next_slot
        self.next_slot += 1
        self.streams[slot] = {"all": FakeModel.tokens(text_ids[1:-1], max_new), "n": 0}   # engine framed the ids with SOT/EOT
        return slot

    def t3_step(self, slots, n_steps=1, noise=None):
        for s in slots:
            st = self.streams[s]
            st["n"] = min(len(st["all"]), st["n"] + n_steps)

    def t3_poll(self, slot):
        st = self.streams[slot]
        return st["n"], st["n"] >= len(st["all"])

    def t3_tokens(self, slot, start, count):
        return np.asarray(self.streams[slot]["all"][start:start + count], dtype=np.int32)

    def t3_close(self, slot):
        self.streams.pop(slot, None)

    def s3gen_infer(self, voice, tokens, cache_source=None, seed=0, **kw):
        return FakeModel.s3gen(tokens, cache_source)

    def crossfade_pcm(self, cur, n_out, prev_tail=None, fade_len=0, **kw):
        x = cur[:n_out].clone()
        if prev_tail is not None and fade_len > 0:
            t = torch.linspace(0, 1, fade_len)
            x[:fade_len] = prev_tail * torch.cos(t * 0.5 * torch.pi) + cur[:fade_len] * torch.sin(t * 0.5 * torch.pi)
        return (torch.clamp(x, -1.0, 1.0) * 32767).to(torch.int16)

    def voice_put(self, key, t3, gen):
        return 0

    def prepare_conditionals(self, wav, sr):
        """Deterministic stand-in conditioning of the right shapes (host-logic tests only: the real encoders are CUDA kernels,
        tests/test_gpu_conditioning.py)."""
        import numpy as np
        n = max(3, min(int(len(wav) / float(sr) * 25), 250))
        g = torch.Generator().manual_seed(int(np.abs(np.asarray(wav[:4000], dtype=np.float64)).sum() * 1e3) & 0x7FFFFFFF)
        spk = torch.randn(1, 256, generator=g)
        return {"t3": {"speaker_emb": spk / spk.norm(), "cond_prompt_speech_tokens": torch.randint(0, 6561, (1, min(n, 150)), generator=g),
                       "emotion_adv": 0.5 * torch.ones(1, 1, 1)},
                "gen": {"prompt_token": torch.randint(0, 6561, (1, n), generator=g), "prompt_token_len": torch.tensor([n]),
                        "prompt_feat": torch.randn(1, 2 * n, 80, generator=g), "prompt_feat_len": None, "embedding": torch.randn(1, 192, generator=g)}}

    def voice_drop(self, key):
        pass

    def close(self):
        pass


SCENARIOS = [
    dict(name="full_fade30", words=40, chunk=150, slice=35, overlap="full", fade=30, lead=0, trail=0),
    dict(name="zero_fade30", words=40, chunk=150, slice=35, overlap="zero", fade=30, lead=0, trail=0),
    dict(name="full_nofade", words=30, chunk=150, slice=35, overlap="full", fade=0, lead=0, trail=0),
    dict(name="trims_slice20", words=26, chunk=100, slice=20, overlap="full", fade=30, lead=50, trail=80),
    dict(name="short", words=3, chunk=150, slice=35, overlap="full", fade=30, lead=0, trail=0),
    dict(name="long_eos", words=70, chunk=150, slice=35, overlap="full", fade=30, lead=0, trail=0),
]


def scenario_text(words):
    import random
    r = random.Random(99)
    out = []
    for i in range(words):
        w = "".join(r.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(5))
        out.append(w + ("." if (i + 1) % 12 == 0 or i == words - 1 else ""))
    return " ".join(out)
