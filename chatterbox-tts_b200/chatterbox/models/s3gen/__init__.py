"""chatterbox.models.s3gen: S3Gen.inference / embed_ref / tokenizer (reference call sites src/tts_streaming.py:316-320,
:366, :370-372, :586-590)."""
import numpy as np
import torch

from ..._backend import new_key, seed_of

S3GEN_SR = 24_000


class _S3Tokenizer:
    """s3gen.tokenizer.forward([wav16k], max_len) -> (tokens, lens).  The conditioning encoders are the next scope row
    (SURVEY 8f.1): until they exist on the GPU the clip determines the SHAPES (25 tokens per second), contents are seeded."""

    def __init__(self, backend):
        self._b = backend

    def forward(self, wavs, max_len=None):
        self._b.require_synthetic("s3gen.tokenizer.forward")
        out, lens = [], []
        for w in wavs:
            w = np.asarray(w, dtype=np.float32).reshape(-1)
            n = max(1, int(len(w) / 16000.0 * 25))
            if max_len:
                n = min(n, int(max_len))
            g = torch.Generator().manual_seed(seed_of(w[:4000]))
            out.append(torch.randint(0, 6561, (n,), generator=g))
            lens.append(n)
        m = max(lens)
        toks = torch.stack([torch.nn.functional.pad(t, (0, m - len(t))) for t in out])
        return toks, torch.tensor(lens)


class S3Gen(torch.nn.Module):
    def __init__(self, backend):
        super().__init__()
        self._b = backend
        self.tokenizer = _S3Tokenizer(backend)

    def embed_ref(self, ref_wav, ref_sr, device="auto", ref_fade_out=True):
        """-> ref_dict (opaque to the engine apart from .to() on tensor values, :115-117).  Shapes follow the clip (25 prompt
        tokens / 50 mel frames per second, at most 10 s); contents are seeded until the encoders of SURVEY 8f.1 exist."""
        self._b.require_synthetic("s3gen.embed_ref")
        w = np.asarray(ref_wav.detach().cpu().numpy() if torch.is_tensor(ref_wav) else ref_wav, dtype=np.float32).reshape(-1)
        fc = self._b.cfg.flow
        n = max(3, min(int(len(w) / float(ref_sr) * 25), 250))
        g = torch.Generator().manual_seed(seed_of(w[:4000], [len(w)]))
        return {"prompt_token": torch.randint(0, fc.vocab, (1, n), generator=g), "prompt_token_len": torch.tensor([n]),
                "prompt_feat": torch.randn(1, 2 * n, fc.mel, generator=g) * 2.0 - 5.0, "prompt_feat_len": None,
                "embedding": torch.randn(1, fc.spk_dim, generator=g), "_cbx_key": new_key("gen")}

    def inference(self, speech_tokens, ref_dict, cache_source=None, finalize=True):
        """(wav (1, 960 n), source (1, 1, 960 n)) on the current CUDA stream (:583-590)."""
        b = self._b
        toks = torch.as_tensor(speech_tokens).reshape(-1).to("cpu").tolist()
        if "_cbx_key" not in ref_dict:
            ref_dict["_cbx_key"] = new_key("gen")
        voice = b.slot_for(ref_dict["_cbx_key"], gen=ref_dict)
        cs = None
        if cache_source is not None and torch.is_tensor(cache_source) and cache_source.shape[-1] > 0:
            cs = cache_source
        return b.native.s3gen_infer(voice, toks, cache_source=cs, seed=seed_of(toks))
