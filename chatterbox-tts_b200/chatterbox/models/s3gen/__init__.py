"""chatterbox.models.s3gen: S3Gen.inference / embed_ref / tokenizer (reference call sites src/tts_streaming.py:316-320,
:366, :370-372, :586-590)."""
import numpy as np
import torch

from ..._backend import new_key, seed_of

S3GEN_SR = 24_000


class _S3Tokenizer:
    """s3gen.tokenizer.forward([wav16k], max_len) -> (tokens (B, T), lens): S3Tokenizer-v2 on the GPU (cbx_b200/conditioning.py)."""

    def __init__(self, backend):
        self._b = backend

    def forward(self, wavs, max_len=None):
        enc = self._b.encoders("s3gen.tokenizer.forward")
        out = []
        for w in wavs:
            w = np.asarray(w.detach().cpu().numpy() if torch.is_tensor(w) else w, dtype=np.float32).reshape(-1)
            out.append(enc.s3_tokens_from_wav(enc._dev_wave(w), max_len=max_len).long().cpu())
        lens = [len(t) for t in out]
        m = max(lens)
        toks = torch.stack([torch.nn.functional.pad(t, (0, m - len(t))) for t in out])
        return toks, torch.tensor(lens)


class S3Gen(torch.nn.Module):
    def __init__(self, backend):
        super().__init__()
        self._b = backend
        self.tokenizer = _S3Tokenizer(backend)

    def embed_ref(self, ref_wav, ref_sr, device="auto", ref_fade_out=True):
        """-> ref_dict (opaque to the engine apart from .to() on tensor values, :115-117): 24 kHz prompt mel, S3Tokenizer prompt
        tokens and CAMPPlus x-vector of the clip, computed on the GPU."""
        enc = self._b.encoders("s3gen.embed_ref")
        w = np.asarray(ref_wav.detach().cpu().numpy() if torch.is_tensor(ref_wav) else ref_wav, dtype=np.float32).reshape(-1)
        x = enc.resample(enc._dev_wave(w), int(ref_sr), S3GEN_SR)
        g = enc.embed_ref(x.contiguous())
        n = g["prompt_token"].shape[0]
        return {"prompt_token": g["prompt_token"].long().cpu()[None], "prompt_token_len": torch.tensor([n]),
                "prompt_feat": g["prompt_feat"].cpu()[None], "prompt_feat_len": None,
                "embedding": g["embedding"].cpu()[None], "_cbx_key": new_key("gen")}

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
