"""The `chatterbox`-named shim on the real CUDA engine: T3.inference_stream and S3Gen.inference must return exactly what the
C-ABI calls underneath return (same tokens, same audio), in the shapes the reference engine consumes
(src/tts_streaming.py:420-435, :539, :583-590)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_shim_over_native_engine(tiny_cfg):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from conftest import bf16_round
    import chatterbox
    from chatterbox.tts import ChatterboxTTS
    from chatterbox._backend import seed_of
    from cbx_b200.native import NativeEngine
    from cbx_b200.weights import random_state_dict

    def factory(ckpt, device):
        eng = NativeEngine(tiny_cfg, max_streams=4, max_s3_tokens=200, n_lanes=1, n_voices=8)
        eng.load_state_dict(bf16_round(random_state_dict(tiny_cfg, 0)))
        eng.encoder_sd = random_state_dict(tiny_cfg, 0, parts=("cond",))
        return eng

    chatterbox.set_backend_factory(factory)
    try:
        tts = ChatterboxTTS.from_local("", "cuda:0")
        nat = tts.backend.native
        dev = torch.device("cuda", 0)
        text = tts.tokenizer.text_to_tokens("hello there").to(dev)
        text = torch.nn.functional.pad(torch.nn.functional.pad(text, (1, 0), value=tts.t3.hp.start_text_token), (0, 1), value=tts.t3.hp.stop_text_token)
        text2 = torch.cat([text, text], dim=0)                                   # the reference duplicates the row for CFG (:475-478)
        conds = tts.conds
        conds.t3 = conds.t3.to(device=dev)
        items = list(tts.t3.inference_stream(t3_cond=conds.t3, text_tokens=text2, max_new_tokens=20, temperature=0.8, cfg_weight=0.5))
        assert len(items) == 20 and all(t.shape == (1, 1) and t.is_cuda and t.dtype == torch.long for t in items)
        toks = torch.cat(items, dim=1)[0].tolist()
        # the same stream opened directly through the native interface
        ids = text[0].cpu().tolist()
        voice = tts.backend.slot_for(conds.t3._cbx_key)
        slot = nat.t3_open(voice, ids, 0.5, 0.8, 1.2, 0.05, 0.95, seed_of(ids, [20]), 20)
        nat.t3_step([slot], 20)
        direct = nat.t3_tokens(slot, 0, 20).tolist()
        nat.t3_close(slot)
        assert toks == direct
        # S3Gen through the shim == native call; cache_source threading as the reference does it (:694-699)
        sp = torch.tensor([t for t in toks if t < 6561][:9] or [1, 2, 3], device=dev)
        wav, src = tts.s3gen.inference(speech_tokens=sp, ref_dict=conds.gen, cache_source=torch.zeros(1, 1, 0, device=dev))
        torch.cuda.synchronize()
        n = sp.numel()
        assert wav.shape == (1, 960 * n) and src.shape == (1, 1, 960 * n) and wav.is_cuda and torch.isfinite(wav).all()
        wav2, src2 = tts.s3gen.inference(speech_tokens=sp, ref_dict=conds.gen, cache_source=src)
        torch.cuda.synchronize()
        assert torch.equal(src2, src) and torch.allclose(wav2, wav, atol=1e-6)
        # the three conditioning call sites of prepare_conditionals (:366, :370-372, :374) run the GPU encoders and agree with the oracle
        from oracle import cond as OC
        enc_sd = nat.encoder_sd
        tw = torch.arange(24000 * 3) / 24000.0
        w24 = (0.3 * torch.sin(2 * np.pi * 200 * tw) * torch.clamp(torch.sin(2 * np.pi * 2 * tw), min=0) + 0.02 * torch.randn(tw.shape[0], generator=torch.Generator().manual_seed(1))).float()
        w16 = OC.resample(w24, 24000, 16000)
        ref = tts.s3gen.embed_ref(w24.numpy(), 24000, device="cuda:0")
        oref = OC.embed_ref(enc_sd, w24)
        assert ref["prompt_feat"].shape == (1, 150, 80) and torch.equal(ref["prompt_token"][0], oref["prompt_token"])
        assert float((ref["embedding"][0] - oref["embedding"]).norm() / oref["embedding"].norm()) < 2e-4
        tk, ln = tts.s3gen.tokenizer.forward([w16.numpy()[: 6 * 16000]], max_len=150)
        assert int(ln[0]) == tk.shape[1] == 75 and torch.equal(tk[0], OC.s3_tokens_from_wav(enc_sd, w16[: 6 * 16000], max_len=150))
        ve = tts.ve.embeds_from_wavs([w16.numpy()], sample_rate=16000)
        assert ve.shape == (1, 256) and float(np.linalg.norm(ve[0] - OC.voice_embed(enc_sd, w16).numpy())) < 2e-4
        # an abandoned generator gives its KV pages back
        gen = tts.t3.inference_stream(t3_cond=conds.t3, text_tokens=text2, max_new_tokens=50)
        next(gen)
        gen.close()
        slots = [nat.t3_open(voice, ids, 0.5, 0.8, 1.2, 0.05, 0.95, 1, 50) for _ in range(4)]     # all 4 stream slots are free again
        for s in slots:
            nat.t3_close(s)
        nat.close()
    finally:
        chatterbox.set_backend_factory(None)
