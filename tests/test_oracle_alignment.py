"""The oracle's restatement of the alignment-based EOS analyzer (oracle/t3.py::AlignmentAnalyzer) on hand-built
alignment matrices: no upstream golden exists for it (parity unpinned), so these pin the behaviour the restatement documents."""
import torch

from oracle.t3 import AlignmentAnalyzer


def _row(S, col, mass=0.9):
    r = torch.full((S,), (1.0 - mass) / S)
    r[col] += mass
    return r


def _walk(S, frames, hold_last=0, jump_back=None):
    """Monotone alignment: frame f attends text column ~f/2, then stays on the last column (or jumps back)."""
    rows = []
    for f in range(frames):
        rows.append(_row(S, min(f // 2, S - 1)))
    for f in range(hold_last):
        rows.append(_row(S, S - 1 if jump_back is None else jump_back))
    return torch.stack(rows)


def test_eos_is_suppressed_until_the_last_three_text_tokens():
    S = 16
    A = _walk(S, 2 * S)
    an = AlignmentAnalyzer(S)
    ctl = [an.step(A[:2])] + [an.step(A[i:i + 1]) for i in range(2, A.shape[0])]
    first_free = next(i for i, c in enumerate(ctl) if not (c & 1))
    # frame i (i >= 1) is row i + 1, which attends column (i + 1) // 2: EOS is released when that reaches S - 3
    assert first_free == 2 * (S - 3) - 1
    assert all(c & 1 for c in ctl[:first_free]) and not any(c & 2 for c in ctl)
    assert an.started and an.complete and an.text_position == S - 1


def test_long_tail_forces_eos():
    S = 12
    A = _walk(S, 2 * S, hold_last=20)
    an = AlignmentAnalyzer(S)
    ctl = [an.step(A[:2])] + [an.step(A[i:i + 1]) for i in range(2, A.shape[0])]
    forced = [i for i, c in enumerate(ctl) if c & 2]
    assert forced, "a final column summing to >= 10 after completion must force EOS"
    # the column's mass per row is ~0.9 + 0.1/12, so twelve post-completion rows are needed; not before
    assert forced[0] >= an.completed_at + 10
    assert ctl[forced[0]] == 2


def test_repetition_forces_eos():
    S = 14
    A = _walk(S, 2 * S, hold_last=8, jump_back=2)
    an = AlignmentAnalyzer(S)
    ctl = [an.step(A[:2])] + [an.step(A[i:i + 1]) for i in range(2, A.shape[0])]
    forced = [i for i, c in enumerate(ctl) if c & 2]
    assert forced and forced[0] >= 2 * S - 1 + 5, "row maxima over the earlier columns must sum to > 5 first"
    # the jump back is a discontinuity (-4 < d < 7 fails): the text position stays at the end, but the frame's own argmax is
    # early again, so both bits are set -- upstream then leaves every logit at -2^15
    assert ctl[forced[0]] == 3 and an.text_position == S - 1


def test_false_start_and_masking():
    S = 10
    an = AlignmentAnalyzer(S)
    # attention on the last text tokens at the very start is masked away (columns above the frame counter are zeroed), the
    # first columns hold nothing: not started
    c = an.step(torch.stack([_row(S, S - 1), _row(S, S - 1)]))
    assert not an.started and c & 1 and an.text_position == 0
    assert float(an.alignment[:, 1:].abs().max()) == 0.0
    an.step(_row(S, 0)[None])
    assert an.started and an.started_at == 3


def test_apply_edits():
    lg = torch.randn(2, 50)
    out = AlignmentAnalyzer.apply(lg, 1, 7)
    assert float(out[0, 7]) == -2 ** 15 and torch.equal(out[:, :7], lg[:, :7])
    out = AlignmentAnalyzer.apply(lg, 2, 7)
    assert float(out[1, 7]) == 2 ** 15 and float(out[0, 3]) == -2 ** 15
    out = AlignmentAnalyzer.apply(lg, 3, 7)
    assert float(out.max()) == -2 ** 15
    assert torch.equal(AlignmentAnalyzer.apply(lg, 0, 7), lg)
