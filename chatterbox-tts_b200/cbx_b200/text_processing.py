"""Text front end of the engine: chunking and tokenisation (CPU string work).

Mirrors the behaviour of the reference's src/text_processing.py:114-196 (`split_text_into_chunks`:
normalise punctuation, segment into sentences, pack sentences up to `max_length` characters, split
oversized sentences by `;:` then `,` then words, merge chunks of fewer than two words) without its
`pysbd` dependency, which is not installed here: sentences are segmented on terminal punctuation
followed by whitespace.  This is the caller side of the hot path, kept small on purpose.
"""
import re
from typing import List

import torch

_SENT_END = re.compile(r"(?<=[.!?])[\"')\]]*\s+")


def _sentences(text: str) -> List[str]:
    return [s.strip() for s in _SENT_END.split(text) if s and s.strip()]


def _merge_small(chunks: List[str], min_words: int, max_len: int, slack: float = 0.10) -> List[str]:
    out, i = [], 0
    while i < len(chunks):
        cur = chunks[i]
        if len(cur.split()) < min_words:
            if out and len(out[-1]) + len(cur) + 1 <= max_len * (1 + slack):
                out[-1] = out[-1] + " " + cur
            elif i + 1 < len(chunks) and len(cur) + len(chunks[i + 1]) + 1 <= max_len * (1 + slack):
                out.append(cur + " " + chunks[i + 1])
                i += 1
            else:
                out.append(cur)
        else:
            out.append(cur)
        i += 1
    return out


def _split_keep(text: str, delims: str) -> List[str]:
    parts, cur = [], ""
    for ch in text:
        cur += ch
        if ch in delims:
            if cur.strip() == ch and parts:
                parts[-1] += ch
            else:
                parts.append(cur.strip())
            cur = ""
    if cur.strip():
        parts.append(cur.strip())
    return parts


def _split_oversized(text: str, max_len: int) -> List[str]:
    mid = []
    for seg in _split_keep(text, ";:"):
        if len(seg) <= max_len:
            mid.append(seg)
            continue
        for sub in _split_keep(seg, ","):
            if len(sub) <= max_len:
                mid.append(sub)
                continue
            words, cur, wc = sub.split(), "", []
            for w in words:
                if len(cur) + len(w) + 1 <= max_len:
                    cur += (" " if cur else "") + w
                else:
                    if cur:
                        wc.append(cur)
                    cur = w
            if cur:
                wc.append(cur)
            mid.extend(_merge_small(wc, 2, max_len))
    return [c.strip() for c in _merge_small(mid, 2, max_len) if c.strip()]


_NORMALISE = (("...", ". "), ("\u2026", ". "), (" - ", ", "), ("\u2014", "-"), ("\u2013", "-"), (" ,", ","),
              ("\u201c", '"'), ("\u201d", '"'), ("\u2018", "'"), ("\u2019", "'"))
_ENDERS = (".", "!", "?", "-")


def split_text_into_chunks(text: str, max_length: int = None) -> List[str]:
    """Same observable behaviour as the reference (src/text_processing.py:114-196): collapse whitespace, normalise
    punctuation, capitalise the first letter, segment into sentences, make every sentence end with punctuation, pack
    sentences while `len(chunk) + len(sentence) + 1 <= max_length`, split oversized sentences, merge one-word chunks."""
    if not text or not text.strip():
        return []
    text = " ".join(text.split())
    for old, new in _NORMALISE:
        text = text.replace(old, new)
    if text and text[0].islower():
        text = text[0].upper() + text[1:]
    sentences = []
    for s in _sentences(text):
        s = s.strip()
        if s:
            sentences.append(s if s.endswith(_ENDERS) else s + ".")
    if max_length is None:
        return sentences
    chunks, cur = [], ""
    for s in sentences:
        if len(s) > max_length:
            if cur:
                chunks.append(cur)
                cur = ""
            chunks.extend(_split_oversized(s, max_length))
        elif len(cur) + len(s) + 1 <= max_length:
            cur = (cur + " " + s) if cur else s
        else:
            if cur:
                chunks.append(cur)
            cur = s
    if cur:
        chunks.append(cur)
    return [c.strip() for c in _merge_small(chunks, 2, max_length) if c.strip()]


class SyntheticTokenizer:
    """Deterministic char -> [1, 703] map used when no tokenizer.json ships with the checkpoint
    (SURVEY 8d synthetic inputs).  Same call surface as the fork's EnTokenizer.text_to_tokens
    (reference call sites src/tts_streaming.py:282, :464): str -> IntTensor (1, L) on the CPU."""

    def __init__(self, vocab: int = 704):
        self.vocab = vocab

    def text_to_tokens(self, text: str) -> torch.Tensor:
        ids = [1 + (ord(ch) * 131 + 7) % (self.vocab - 1) for ch in text]
        ids = [i if i != 255 else 254 for i in ids]   # keep SOT (255) out of the body
        return torch.tensor(ids or [1], dtype=torch.int32).unsqueeze(0)


class JsonTokenizer:
    """tokenizer.json present in the checkpoint directory: upstream EnTokenizer behaviour ([SPACE] for ' ')."""

    def __init__(self, path: str):
        from tokenizers import Tokenizer
        self.tk = Tokenizer.from_file(path)

    def text_to_tokens(self, text: str) -> torch.Tensor:
        ids = self.tk.encode(text.replace(" ", "[SPACE]")).ids
        return torch.tensor(ids, dtype=torch.int32).unsqueeze(0)
