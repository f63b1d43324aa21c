"""Alignment-based EOS control on the GPU (include/cbx_b200.h::cbx_t3_set_alignment_eos) against the oracle's
restatement of the analyzer (oracle/t3.py::AlignmentAnalyzer).
  * the analyzer kernel alone over given alignment matrices: every per-frame decision identical (integer work: exact);
  * inside the T3 decode step: the alignment rows (head-averaged attention over the text span at the probe layer) within
    2e-2 relative of the fp32 oracle's (bf16 KV cache), the analyzer state identical to the oracle analyzer fed the DEVICE rows,
    the sampled ids identical to the oracle sampler on the kernel's logits edited by that decision."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
V = 8194


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda", 0)


def _text(L, seed=3):
    g = torch.Generator().manual_seed(seed)
    t = torch.randint(1, 700, (1, L), generator=g)
    return torch.nn.functional.pad(torch.nn.functional.pad(t, (1, 0), value=255), (0, 1), value=0)


def _matrix(kind, S, seed):
    """Alignment matrices that drive the analyzer through its branches: a monotone walk over the text, then a long tail on the
    last token / a jump back to an early token / neither; `noisy` adds off-diagonal mass and discontinuities."""
    g = torch.Generator().manual_seed(seed)
    rows = []
    speed = 1 + int(torch.randint(1, 4, (1,), generator=g))
    frames = speed * S + int(torch.randint(0, 6, (1,), generator=g))
    for f in range(frames):
        r = torch.rand(S, generator=g) * (0.02 if kind != "noisy" else 0.12)
        col = min(f // speed, S - 1)
        if kind == "noisy" and float(torch.rand(1, generator=g)) < 0.15:
            col = int(torch.randint(0, S, (1,), generator=g))
        r[col] += 0.4 + 0.5 * float(torch.rand(1, generator=g))
        rows.append(r)
    extra = 30
    for f in range(extra):
        r = torch.rand(S, generator=g) * 0.02
        if kind == "tail":
            r[S - 1 - (f % 2)] += 0.8
        elif kind == "repeat":
            r[f % 2] += 0.7
        elif kind == "noisy":
            r[int(torch.randint(0, S, (1,), generator=g))] += 0.6
        else:
            r[S - 1] += 0.2
        rows.append(r)
    return torch.stack(rows)


@pytest.mark.parametrize("kind", ["tail", "repeat", "plain", "noisy"])
@pytest.mark.parametrize("S,n_pre", [(7, 1), (24, 1), (24, 0), (300, 1)])
def test_analyzer_kernel_matches_oracle(dev, kind, S, n_pre):
    from oracle.t3 import AlignmentAnalyzer
    from cbx_b200 import lib as L
    lib = L.load()
    A = _matrix(kind, S, seed=S * 7 + n_pre)
    an = AlignmentAnalyzer(S)
    want = []
    for f, r in enumerate(range(n_pre, A.shape[0])):
        ctl = an.step(A[:n_pre + 1] if f == 0 else A[r:r + 1])
        want.append([ctl, an.text_position, int(an.started), int(an.complete)])
    frames = A.shape[0] - n_pre
    got = np.zeros((frames, 4), np.int32)
    Ad = A.to(dev).contiguous()
    L.check(lib.cbx_op_alignment_run(Ad.data_ptr(), A.shape[0], S, n_pre, got.ctypes.data, None))
    want = np.asarray(want, np.int32)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, f"first differing frame {bad[0]}: kernel {got[bad[0]]} oracle {want[bad[0]]}"
    if kind in ("tail", "repeat"):
        assert (want[:, 0] & 2).any(), "the scenario must reach a forced EOS"
    assert (want[:, 0] & 1).any() and want[-1, 3] == 1


def _engine(tiny_cfg, sd):
    from cbx_b200.native import NativeEngine
    eng = NativeEngine(tiny_cfg, max_streams=4, max_s3_tokens=64, n_lanes=1)
    eng.load_state_dict(sd)
    return eng


def _run(eng, sd, cfg, conds, voice, dev, text, steps, cfg_w, check_rows):
    """Steps one stream with alignment control on, beside the oracle generator.  Returns (tokens, ctl per step, worst row error)."""
    from oracle import t3 as O
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    cond_dev = {k: v.to(dev) for k, v in conds["t3"].items()}
    g = torch.Generator().manual_seed(11)
    nd = torch.empty(steps, V).exponential_(generator=g).to(dev)
    tt = torch.cat([text, text]).to(dev) if cfg_w > 0 else text.to(dev)
    S = text.shape[1]
    trace = []
    gen = O.inference_stream(sd_dev, cfg.t3, cond_dev, tt, steps, temperature=0.8, cfg_weight=cfg_w, rep_penalty=1.2, min_p=0.05, top_p=0.95,
                             noise_fn=lambda i: nd[i], return_logits=True, alignment_layer=9, trace=trace)
    an = O.AlignmentAnalyzer(S)     # fed with the DEVICE rows: its state must equal the kernel's
    slot = eng.t3_open(voice, text[0].numpy(), seed=5, max_new=steps, rep_penalty=1.2, min_p=0.05, top_p=0.95, cfg_weight=cfg_w, temperature=0.8)
    st, _, pre = eng.t3_alignment_peek(slot)
    assert st["on"] == 1 and st["S"] == S and st["i0"] == 34 and st["has_pre"] == (1 if cfg_w > 0 else 0) and st["rows"] == 0
    hist, toks, ctls, worst, oracle_live = [cfg.t3.start_speech_token], [], [], 0.0, True
    with torch.no_grad():
        for i in range(steps):
            eng.t3_step([slot], 1, noise=nd[i:i + 1].contiguous())
            lg = torch.from_numpy(eng.t3_logits(slot)).to(dev)
            n, done = eng.t3_poll(slot)
            tok = int(eng.t3_tokens(slot, i, 1)[0])
            st, cur, _ = eng.t3_alignment_peek(slot)
            rows = torch.from_numpy(np.stack([pre, cur]) if (i == 0 and cfg_w > 0) else cur[None])
            ctl = an.step(rows)
            assert (st["ctl"], st["text_pos"], st["started"], st["complete"], st["frame_pos"], st["cur_posn"]) == \
                   (ctl, an.text_position, int(an.started), int(an.complete), an.curr_frame_pos, an.cur_text_posn), f"step {i}: {st}"
            assert st["rows"] == an.alignment.shape[0]
            used = O.AlignmentAnalyzer.apply(lg[:2 if cfg_w > 0 else 1], ctl, cfg.t3.stop_speech_token)
            fl = O.process_logits(used, hist, cfg_w, 0.8, 1.2, 0.05, 0.95)
            assert O.sample_from(fl, nd[i]) == tok, f"step {i}: sampled id differs from the oracle sampler on the edited logits"
            hist.append(tok); toks.append(tok); ctls.append(ctl)
            if oracle_live:
                otok, _ = next(gen)
                orows, _ = trace[-1]
                if check_rows:
                    err = float((rows.double() - orows.double()).norm() / orows.double().norm())
                    worst = max(worst, err)
                oracle_live = otok == tok and otok != cfg.t3.stop_speech_token
            if done:
                break
    eng.t3_close(slot)
    return toks, ctls, worst


@pytest.mark.parametrize("cfg_w,L", [(0.5, 21), (0.0, 9)])
def test_alignment_rows_and_decisions_in_the_decode_step(dev, tiny_cfg, cfg_w, L):
    from conftest import bf16_round
    from cbx_b200.weights import random_state_dict, synthetic_conditionals
    sd = bf16_round(random_state_dict(tiny_cfg, 0))
    conds = synthetic_conditionals(tiny_cfg, 1234, prompt_tokens=40)
    eng = _engine(tiny_cfg, sd)
    try:
        voice = eng.voice_put("v", conds["t3"], conds["gen"])
        eng.t3_set_alignment_eos(True, 9)     # the tiny trunk has two layers: the probe clamps to the last one, as the oracle does
        toks, ctls, worst = _run(eng, sd, tiny_cfg, conds, voice, dev, _text(L), 14, cfg_w, check_rows=True)
        assert worst < 2e-2, f"alignment rows: relative error {worst}"
        assert all(c & 1 for c in ctls[:4]), "EOS must be suppressed while the alignment is at the start of the text"
        # streams opened with the control off are not touched
        eng.t3_set_alignment_eos(False, 9)
        slot = eng.t3_open(voice, _text(L)[0].numpy(), seed=5, max_new=4)
        eng.t3_step([slot], 2)
        assert eng.t3_alignment_peek(slot, rows=False)[0]["on"] == 0
        eng.t3_close(slot)
    finally:
        eng.close()


def test_forced_eos_ends_the_stream(dev, tiny_cfg):
    """With random weights the attention over the text is too flat to ever cross the analyzer's thresholds, so the stream's
    analyzer is placed just short of a repetition verdict (test hook): the next frame must raise the force bit, and the first
    frame whose own argmax is past the last-three mark (no suppress bit) must sample 6562, end the stream and leave it stopped.
    While both bits are set every logit is -2^15 (ties: the sampled id is implementation-defined and not compared)."""
    from conftest import bf16_round
    from oracle import t3 as O
    from cbx_b200.weights import random_state_dict, synthetic_conditionals
    sd = bf16_round(random_state_dict(tiny_cfg, 0))
    conds = synthetic_conditionals(tiny_cfg, 1234, prompt_tokens=40)
    eng = _engine(tiny_cfg, sd)
    eos, steps = tiny_cfg.t3.stop_speech_token, 200
    try:
        voice = eng.voice_put("v", conds["t3"], conds["gen"])
        eng.t3_set_alignment_eos(True, 9)
        free0, open0 = eng.t3_stats()
        text = _text(6)
        g = torch.Generator().manual_seed(4)
        nd = torch.empty(steps, V).exponential_(generator=g).to(dev)
        slot = eng.t3_open(voice, text[0].numpy(), seed=5, max_new=steps)
        hist = [tiny_cfg.t3.start_speech_token]
        for i in range(3):
            eng.t3_step([slot], 1, noise=nd[i:i + 1].contiguous())
            hist.append(int(eng.t3_tokens(slot, i, 1)[0]))
        st, _, _ = eng.t3_alignment_peek(slot)
        assert not st["complete"] and not (st["ctl"] & 2)
        st.update(complete=1, completed_at=st["rows"], text_pos=st["S"] - 1, rep_sum=4.9999)
        eng.t3_alignment_poke(slot, st)
        ended = None
        for i in range(3, steps):
            eng.t3_step([slot], 1, noise=nd[i:i + 1].contiguous())
            st, _, _ = eng.t3_alignment_peek(slot)
            n, done = eng.t3_poll(slot)
            tok = int(eng.t3_tokens(slot, i, 1)[0])
            assert st["ctl"] & 2 and st["rep_sum"] > 5.0, f"step {i}: {st}"
            assert (st["ctl"] & 1) == (1 if st["cur_posn"] < st["S"] - 3 else 0)
            if st["ctl"] == 2:
                lg = torch.from_numpy(eng.t3_logits(slot)).to(dev)
                fl = O.process_logits(O.AlignmentAnalyzer.apply(lg, 2, eos), hist, 0.5, 0.8, 1.2, 0.05, 0.95)
                assert O.sample_from(fl, nd[i]) == tok == eos and done and n == i + 1
                ended = i
                break
            assert tok != eos and not done, "EOS is suppressed while the frame's argmax is short of the last three text tokens"
            hist.append(tok)
        assert ended is not None, "no frame without the suppress bit"
        eng.t3_step([slot], 2)
        assert eng.t3_poll(slot) == (ended + 1, True), "a stopped stream must not advance"
        eng.t3_close(slot)
        assert eng.t3_stats() == (free0, open0)
    finally:
        eng.close()
