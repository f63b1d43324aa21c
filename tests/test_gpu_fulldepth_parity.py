"""Full-depth parity: the CUDA path (through the C-ABI) against the fp32 oracle at the BENCHMARKED model size --
30 T3 layers, 6+4 conformer blocks, 14 x 4 CFM transformer blocks, 10 Euler steps, trump.wav-shaped conditioning
(194 prompt tokens / 388 prompt frames) -- on BASELINE.json's shapes (147 text tokens, 35- and 140-token slices,
8 concurrent streams).  The oracle runs on the same GPU in fp32 (TF32 off) so the whole file takes about a minute.

Stated tolerances (BASELINE.md "Parity tolerances"; north_star: logits <= 1e-2 relative, sampled ids bit-exact on
identical logits, mel / wav "within a stated tolerance"):
  T3 logits      rel-L2 per step and CFG row           <= 1e-2     (TOL_LOGITS)
  sampled ids    oracle sampler on the kernel's logits  ==         (bit-exact)
  mel            rel-L2 over the generated frames      <= 1e-2     (TOL_MEL; SNR >= 40 dB), max-abs reported
  waveform       rel-L2, oracle mel + source forced    <= 1e-2     (TOL_WAV; SNR >= 40 dB)
(measured in round 2: logits 7.6e-3 .. 8.3e-3, mel 2.1e-3 .. 2.8e-3 = 51 dB, wav 4.1e-3 = 47.7 dB)
Every measured figure is appended to gpurun_out/parity_fulldepth.json for DESIGN.md / BASELINE.md.
"""
import json
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL_LOGITS, TOL_MEL, TOL_WAV = 1e-2, 1e-2, 1e-2
V = 8194
_REPORT = {}


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _snr_db(a, b):
    return -20.0 * math.log10(max(_rel(a, b), 1e-30))


def _report(key, val):
    _REPORT[key] = val
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
        with open(os.path.join(root, "gpurun_out", "parity_fulldepth.json"), "w") as f:
            json.dump(_REPORT, f, indent=1, sort_keys=True)
    except OSError:
        pass


@pytest.fixture(scope="module")
def full():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from conftest import bf16_round
    from cbx_b200.config import ModelConfig
    from cbx_b200.native import NativeEngine
    from cbx_b200.weights import random_state_dict, synthetic_conditionals
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda", 0)
    cfg = ModelConfig()
    sd = bf16_round(random_state_dict(cfg, 0))
    eng = NativeEngine(cfg, max_streams=16, n_lanes=1, max_text=512)
    eng.load_state_dict(sd)
    conds = synthetic_conditionals(cfg)              # trump.wav shapes: 194 prompt tokens, 388 mel frames
    conds["gen"]["prompt_feat"] = conds["gen"]["prompt_feat"].to(torch.bfloat16).float()
    voice = eng.voice_put("default", conds["t3"], conds["gen"])
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    yield eng, sd_dev, cfg, conds, voice, dev
    eng.close()


def _text(n, k):
    return [255] + [(7 * i + 13 * k) % 700 + 1 for i in range(n)] + [0]


def _oracle_logits(sd_dev, cfg, cond_dev, text, toks, dev, cfg_w=0.5):
    """Logits of every decode step from ONE causal oracle forward, teacher-forced on `toks` (the ids the CUDA path sampled):
    step i sees [cond prefix | text | BOS BOS | emb(tok_0)+pos(1) ... emb(tok_{i-1})+pos(i)] (oracle/t3.py inference_stream)."""
    from oracle import t3 as O
    tt = torch.tensor([text, text] if cfg_w > 0 else [text], device=dev)
    with torch.no_grad():
        x = O.prepare_input_embeds(sd_dev, cfg.t3, cond_dev, tt, cfg_w)
        Lx = x.shape[1]
        n = len(toks)
        if n > 1:
            ids = torch.tensor(toks[:-1], device=dev)
            e = sd_dev["t3.speech_emb.weight"][ids] + sd_dev["t3.speech_pos_emb.emb.weight"][1:n]
            x = torch.cat([x, e[None].expand(x.shape[0], -1, -1)], dim=1)
        h, _ = O.llama_forward(sd_dev, cfg.t3, x)
        return O.speech_logits(sd_dev, h[:, Lx - 1: Lx - 1 + n]).transpose(0, 1).contiguous()   # (n, rows, V)


def _run_streams(eng, voice, texts, steps, dev, seed0=40, noise_seed=11):
    """Decodes `steps` tokens for all texts in ONE batched pass per step with explicit sampling noise; returns per stream the
    per-step logits (steps, 2, V), the sampled ids, and the noise used."""
    n = len(texts)
    g = torch.Generator().manual_seed(noise_seed)
    noise = torch.empty(steps, n, V).exponential_(generator=g).to(dev)
    slots = [eng.t3_open(voice, t, seed=seed0 + i, max_new=steps) for i, t in enumerate(texts)]
    logits = [[] for _ in range(n)]
    for i in range(steps):
        eng.t3_step(slots, 1, noise=noise[i].contiguous())
        for k, s in enumerate(slots):
            logits[k].append(torch.from_numpy(eng.t3_logits(s)))
    toks = [eng.t3_tokens(s, 0, steps).tolist() for s in slots]
    for s in slots:
        eng.t3_close(s)
    return [torch.stack(l) for l in logits], toks, noise


def _check_t3(full, texts, steps, label, persistent=False):
    from oracle import t3 as O
    eng, sd_dev, cfg, conds, voice, dev = full
    cond_dev = {k: v.to(dev) for k, v in conds["t3"].items()}
    eng.t3_set_persistent(persistent)
    try:
        logits, toks, noise = _run_streams(eng, voice, texts, steps, dev)
    finally:
        eng.t3_set_persistent(False)
    worst, worst_at = 0.0, None
    for k, text in enumerate(texts):
        ref = _oracle_logits(sd_dev, cfg, cond_dev, text, toks[k], dev)
        hist = [cfg.t3.start_speech_token]
        for i in range(steps):
            lg = logits[k][i].to(dev)
            for row in range(2):
                r = _rel(lg[row], ref[i, row])
                if r > worst:
                    worst, worst_at = r, (k, i, row)
            # sampler: the oracle's processors + argmax(p/q) on the KERNEL's logits and the same noise must give the same id
            fl = O.process_logits(lg, hist, 0.5, 0.8, 1.2, 0.05, 0.95)
            assert O.sample_from(fl, noise[i, k]) == toks[k][i], f"{label}: sampled id differs at stream {k} step {i}"
            hist.append(toks[k][i])
    _report(f"t3_logits_rel_{label}", {"worst": worst, "at_stream_step_row": worst_at, "steps": steps, "streams": len(texts)})
    assert worst < TOL_LOGITS, f"{label}: logits rel-L2 {worst:.3e} at {worst_at}"


def test_t3_logits_one_stream_64_steps(full):
    """configs[1] shape: 34 cond + 149 text ids (147 + SOT/EOT) + BOS, 64 decode steps, 2 rows."""
    _check_t3(full, [_text(147, 0)], 64, "gemv_1stream")


def test_t3_logits_eight_streams_64_steps(full):
    """configs[2] shape: 8 streams = 16 rows in one batched pass, ragged text lengths."""
    _check_t3(full, [_text(60 + 12 * i, i) for i in range(8)], 64, "gemv_8streams")


def test_t3_logits_sixteen_streams_one_pass(full):
    """16 streams = 32 rows decode in ONE pass over the weights (the 32-row GEMV instance; the down projection as two
    16-row passes over L2-resident weights), ragged text lengths."""
    from cbx_b200 import lib as L
    lib = L.load()
    before = lib.cbx_t3_tc_launches()
    _check_t3(full, [_text(40 + 9 * i, i + 5) for i in range(16)], 24, "gemv_16streams")
    # batched rows take the tcgen05 projections (swap-AB, t3_gemv_tc.cu): every layer but the first QKV, every step
    assert lib.cbx_t3_tc_launches() - before >= 24 * (4 * 30 - 1)


def test_t3_logits_persistent_kernel(full):
    _check_t3(full, [_text(147, 1)], 64, "persistent_1stream", persistent=True)
    _check_t3(full, [_text(50 + 20 * i, i + 3) for i in range(4)], 24, "persistent_4streams", persistent=True)


def test_t3_logits_position_beyond_1000(full):
    """Long context: 502 text ids -> prefill 537 positions, 480 decode steps -> KV positions up to 1017; the last 32 steps
    (positions >= 985) are compared, the sampler on all of them."""
    from oracle import t3 as O
    eng, sd_dev, cfg, conds, voice, dev = full
    cond_dev = {k: v.to(dev) for k, v in conds["t3"].items()}
    text = _text(500, 5)
    steps, tail = 480, 32
    g = torch.Generator().manual_seed(5)
    slot = eng.t3_open(voice, text, seed=77, max_new=steps)
    eng.t3_step([slot], steps - tail)                       # Philox noise for the bulk
    noise = torch.empty(tail, 1, V).exponential_(generator=g).to(dev)
    lgs = []
    for i in range(tail):
        eng.t3_step([slot], 1, noise=noise[i].contiguous())
        lgs.append(torch.from_numpy(eng.t3_logits(slot)))
    toks = eng.t3_tokens(slot, 0, steps).tolist()
    n, done = eng.t3_poll(slot)
    eng.t3_close(slot)
    assert n == steps and done
    ref = _oracle_logits(sd_dev, cfg, cond_dev, text, toks, dev)
    worst = 0.0
    for i in range(tail):
        step = steps - tail + i
        lg = lgs[i].to(dev)
        worst = max(worst, _rel(lg, ref[step]))
        fl = O.process_logits(lg, [cfg.t3.start_speech_token] + toks[:step], 0.5, 0.8, 1.2, 0.05, 0.95)
        assert O.sample_from(fl, noise[i, 0]) == toks[step]
    _report("t3_logits_rel_pos1000", {"worst": worst, "positions": [34 + 502 + 1 + steps - tail, 34 + 502 + steps]})
    assert worst < TOL_LOGITS, f"logits rel-L2 {worst:.3e} at KV positions ~1000"


# ------------------------------------------------------------------------------------------------ flow (K7-K9)
def _flow_oracle(sd_dev, cfg, conds, toks, dev):
    from oracle import flow as F
    ref = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in conds["gen"].items()}
    with torch.no_grad():
        return F.flow_inference(sd_dev, cfg.flow, torch.as_tensor(toks, device=dev), ref)[0].t().contiguous()   # (2n, 80)


@pytest.mark.parametrize("n", [35, 140])
def test_flow_mel_full_depth_single(full, n):
    """One S3Gen token->mel call on the first slice (35 tokens, T = 458 frames) and on four accumulated slices (140, T = 668)."""
    eng, sd_dev, cfg, conds, voice, dev = full
    g = torch.Generator().manual_seed(100 + n)
    toks = torch.randint(0, 6561, (n,), generator=g).numpy().astype(np.int32)
    mel_o = _flow_oracle(sd_dev, cfg, conds, toks, dev)
    mel = eng.flow_infer(voice, toks)
    torch.cuda.synchronize()
    r, mx = _rel(mel, mel_o), float((mel - mel_o).abs().max())
    _report(f"flow_mel_single_n{n}", {"rel_l2": r, "max_abs": mx, "snr_db": _snr_db(mel, mel_o), "ref_abs_max": float(mel_o.abs().max()), "T": 2 * (194 + n)})
    assert torch.isfinite(mel).all()
    assert r < TOL_MEL, f"mel rel-L2 {r:.3e} (max abs {mx:.3e})"


def test_flow_mel_full_depth_ragged_batch_of_8(full):
    """Eight calls of different lengths through ONE batched token->mel pass (cbx_s3gen_infer_batch), each against its own
    oracle run."""
    eng, sd_dev, cfg, conds, voice, dev = full
    g = torch.Generator().manual_seed(8)
    lens = [35, 70, 105, 140, 36, 12, 3, 77]
    toks = [torch.randint(0, 6561, (n,), generator=g).numpy().astype(np.int32) for n in lens]
    outs = eng.s3gen_infer_batch([(voice, t, None, 9) for t in toks], return_mel=True)
    torch.cuda.synchronize()
    worst, rows = 0.0, []
    for n, t, o in zip(lens, toks, outs):
        mel_o = _flow_oracle(sd_dev, cfg, conds, t, dev)
        r = _rel(o[2], mel_o)
        rows.append({"n": n, "rel_l2": r, "max_abs": float((o[2] - mel_o).abs().max())})
        worst = max(worst, r)
        assert o[0].shape == (1, 960 * n) and torch.isfinite(o[0]).all()
    _report("flow_mel_ragged_batch8", {"worst_rel_l2": worst, "calls": rows})
    assert worst < TOL_MEL, f"batched mel rel-L2 {worst:.3e}"


# ------------------------------------------------------------------------------------------------ HiFT (K10-K12)
@pytest.mark.parametrize("T", [70, 280])
def test_wav_teacher_forced_full_size(full, T):
    """Vocoder with the oracle's mel AND source teacher-forced (cache_source, the reference's own mechanism,
    src/tts_streaming.py:694-699) at the frame counts of one and four slices."""
    from oracle import hift as H
    eng, sd_dev, cfg, conds, voice, dev = full
    hc = cfg.hift
    g = torch.Generator().manual_seed(T)
    toks = torch.randint(0, 6561, (T // 2,), generator=g).numpy().astype(np.int32)
    mel = _flow_oracle(sd_dev, cfg, conds, toks, dev)          # a mel with the statistics the vocoder really sees
    phase = (torch.rand(9, generator=g) * 2 - 1) * math.pi
    noise = torch.randn(9, T * 480, generator=g).to(dev)
    with torch.no_grad():
        wav_o, s_o = H.hift_inference(sd_dev, hc, mel.t()[None].contiguous(), None, phase.to(dev), noise)
    wav, s2 = eng.hift_infer(mel, cache_source=s_o.contiguous())
    torch.cuda.synchronize()
    assert torch.equal(s2, s_o)
    r, mx = _rel(wav, wav_o), float((wav - wav_o).abs().max())
    _report(f"wav_teacher_forced_T{T}", {"rel_l2": r, "max_abs": mx, "snr_db": _snr_db(wav, wav_o), "ref_abs_max": float(wav_o.abs().max())})
    assert r < TOL_WAV, f"wav rel-L2 {r:.3e} (max abs {mx:.3e})"


# ------------------------------------------------------------------------------------------------ natural stop (K6)
def test_natural_eos_stops_the_stream_and_releases_pages(full):
    """With random-init weights the stop token (6562) is never sampled.  Here its speech-head row is CALIBRATED so that the
    stop logit is very low for steps 0..K-1 and very high at step K of one fixed trajectory (min-norm solution of those
    constraints on the oracle's hidden states, teacher-forced on the tokens the CUDA path sampled); the head is then
    re-uploaded and the same stream replayed.  It must sample 6562 at exactly step K, report `done` before max_new, stay
    stopped on further steps, agree with the oracle sampler at every step, and give its KV pages back on close."""
    from conftest import bf16_round
    from oracle import t3 as O
    from cbx_b200 import lib as L
    from cbx_b200.config import ModelConfig
    from cbx_b200.native import NativeEngine
    from cbx_b200.pack import frag_order
    from cbx_b200.weights import random_state_dict
    _, _, _, conds, _, dev = full
    cfg = ModelConfig.tiny()                      # the stop path does not depend on depth
    sd = bf16_round(random_state_dict(cfg, 3))
    sd["t3.speech_head.weight"][6562] = 0.0
    eng = NativeEngine(cfg, max_streams=4, max_s3_tokens=64, n_lanes=1)
    K, steps = 9, 24
    try:
        eng.load_state_dict(sd)
        voice = eng.voice_put("v", conds["t3"], conds["gen"])
        free0, open0 = eng.t3_stats()
        text = _text(20, 2)
        g = torch.Generator().manual_seed(1)
        noise = torch.empty(steps, 1, V).exponential_(generator=g).to(dev)

        def run(n_steps):
            slot = eng.t3_open(voice, text, seed=3, max_new=steps)
            out = []
            for i in range(n_steps):
                eng.t3_step([slot], 1, noise=noise[i].contiguous())
                n, done = eng.t3_poll(slot)
                out.append((n, done, torch.from_numpy(eng.t3_logits(slot)).to(dev), int(eng.t3_tokens(slot, min(i, n - 1), 1)[0])))
            return slot, out

        # pass 1: the trajectory without a usable stop row
        slot, out = run(K + 1)
        eng.t3_close(slot)
        toks = [o[3] for o in out]
        assert cfg.t3.stop_speech_token not in toks
        # calibrate the stop row on the oracle's final hidden states of steps 0..K (both CFG rows)
        sd_dev = {k: v.to(dev) for k, v in sd.items()}
        cond_dev = {k: v.to(dev) for k, v in conds["t3"].items()}
        with torch.no_grad():
            tt = torch.tensor([text, text], device=dev)
            x = O.prepare_input_embeds(sd_dev, cfg.t3, cond_dev, tt, 0.5)
            Lx = x.shape[1]
            ids = torch.tensor(toks[:K], device=dev)
            e = sd_dev["t3.speech_emb.weight"][ids] + sd_dev["t3.speech_pos_emb.emb.weight"][1:K + 1]
            h, _ = O.llama_forward(sd_dev, cfg.t3, torch.cat([x, e[None].expand(2, -1, -1)], dim=1))
            H = h[:, Lx - 1: Lx + K].reshape(-1, 1024).double()                     # rows: (cfg row, step)
            y = torch.full((2, K + 1), -12.0, dtype=torch.float64, device=dev)
            y[:, K] = 40.0
            w = (torch.linalg.pinv(H) @ y.reshape(-1)).float()
        sd["t3.speech_head.weight"][6562] = w.cpu().to(torch.bfloat16).float()
        head = sd["t3.speech_head.weight"]
        vpad = (head.shape[0] + 15) // 16 * 16
        hf = frag_order(torch.nn.functional.pad(head, (0, 0, 0, vpad - head.shape[0])).to(torch.bfloat16)).contiguous()
        L.check(eng.lib.cbx_tensor_upload(eng.h, b"t3.head_f", hf.data_ptr(), hf.numel() * 2))
        # pass 2: same stream, same noise -> same ids up to K-1, then the stop token
        slot, out = run(K + 4)
        free1, open1 = eng.t3_stats()
        assert free1 < free0 and open1 == open0 + 1
        hist = [cfg.t3.start_speech_token]
        for i in range(K + 1):
            n, done, lg, tok = out[i]
            fl = O.process_logits(lg, hist, 0.5, 0.8, 1.2, 0.05, 0.95)
            assert O.sample_from(fl, noise[i, 0]) == tok, f"step {i}"
            hist.append(tok)
            if i < K:
                assert tok == toks[i] and not done and n == i + 1
            else:
                assert tok == cfg.t3.stop_speech_token and done and n == K + 1, f"stop logit {float(lg[0, 6562]):.1f}, |w| {float(w.norm()):.1f}"
        for i in range(K + 1, K + 4):
            assert out[i][0] == K + 1 and out[i][1], "a stopped stream must not advance"
        assert eng.t3_tokens(slot, 0, K + 1).tolist() == toks[:K] + [cfg.t3.stop_speech_token]
        eng.t3_close(slot)
        assert eng.t3_stats() == (free0, open0), "closing the stream must return every KV page"
        _report("natural_eos", {"stopped_at_step": K, "stop_row_norm": float(w.norm())})
    finally:
        eng.close()
