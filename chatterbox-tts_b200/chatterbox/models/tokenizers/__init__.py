"""chatterbox.models.tokenizers.EnTokenizer.text_to_tokens (reference call sites src/tts_streaming.py:282, :464)."""
import os


class EnTokenizer:
    def __init__(self, vocab_file_path=None, text_vocab=704):
        from cbx_b200.text_processing import JsonTokenizer, SyntheticTokenizer
        self._tk = JsonTokenizer(vocab_file_path) if vocab_file_path and os.path.exists(vocab_file_path) else SyntheticTokenizer(text_vocab)

    def text_to_tokens(self, text: str):
        return self._tk.text_to_tokens(text)
