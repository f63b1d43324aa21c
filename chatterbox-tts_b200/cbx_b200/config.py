"""Hyper-parameters of the Chatterbox hot path (T3 -> S3Gen flow -> HiFT).

The reference repo (akashdeep000/chatterbox-tts) holds no model code; these
dimensions restate the public `resemble-ai/chatterbox` package that the
reference imports at src/tts_streaming.py:38-44.
`tiny()` variants exist only so CPU tests and small parity cases finish in
seconds; the benchmark always uses the full configuration.
"""
from dataclasses import dataclass, field, asdict
from typing import List


@dataclass
class T3Config:
    # token ids (reference touches these at src/tts_streaming.py:283, :369, :477, :606)
    start_text_token: int = 255
    stop_text_token: int = 0
    text_vocab: int = 704
    max_text_tokens: int = 2048
    start_speech_token: int = 6561
    stop_speech_token: int = 6562
    speech_vocab: int = 8194
    max_speech_tokens: int = 4096
    speech_cond_prompt_len: int = 150
    speaker_embed_size: int = 256
    perceiver_queries: int = 32
    perceiver_heads: int = 4
    # Llama_520M trunk
    dim: int = 1024
    n_layers: int = 30
    n_heads: int = 16
    head_dim: int = 64
    ffn: int = 4096
    rms_eps: float = 1e-5
    rope_theta: float = 500000.0
    rope_factor: float = 8.0
    rope_low_freq_factor: float = 1.0
    rope_high_freq_factor: float = 4.0
    rope_orig_max_pos: int = 8192

    @property
    def cond_len(self) -> int:  # speaker + perceiver queries + emotion
        return 1 + self.perceiver_queries + 1

    @staticmethod
    def tiny() -> "T3Config":
        return T3Config(n_layers=2)


@dataclass
class FlowConfig:
    vocab: int = 6561
    enc_dim: int = 512
    enc_heads: int = 8
    enc_ffn: int = 2048
    enc_blocks: int = 6
    up_blocks: int = 4
    pre_lookahead: int = 3
    spk_dim: int = 192
    mel: int = 80
    token_mel_ratio: int = 2
    # CFM estimator (ConditionalDecoder, causal, channels=[256])
    ch: int = 256
    heads: int = 8
    head_dim: int = 64
    n_blocks: int = 4          # transformer blocks per resnet stage
    n_mid: int = 12
    n_timesteps: int = 10
    cfg_rate: float = 0.7
    noise_len: int = 15000

    @property
    def in_ch(self) -> int:
        return 4 * self.mel

    @property
    def time_dim(self) -> int:
        return 4 * self.ch

    @staticmethod
    def tiny() -> "FlowConfig":
        return FlowConfig(enc_blocks=1, up_blocks=1, n_blocks=1, n_mid=1, n_timesteps=2)


@dataclass
class HiFTConfig:
    sr: int = 24000
    mel: int = 80
    base_ch: int = 512
    nb_harmonics: int = 8
    nsf_alpha: float = 0.1
    nsf_sigma: float = 0.003
    voiced_threshold: float = 10.0
    upsample_rates: List[int] = field(default_factory=lambda: [8, 5, 3])
    upsample_kernels: List[int] = field(default_factory=lambda: [16, 11, 7])
    resblock_kernels: List[int] = field(default_factory=lambda: [3, 7, 11])
    resblock_dilations: List[int] = field(default_factory=lambda: [1, 3, 5])
    source_resblock_kernels: List[int] = field(default_factory=lambda: [7, 7, 11])
    n_fft: int = 16
    hop: int = 4
    lrelu_slope: float = 0.1
    audio_limit: float = 0.99
    f0_ch: int = 512
    f0_layers: int = 5

    @property
    def upsample_total(self) -> int:  # samples per mel frame
        p = self.hop
        for u in self.upsample_rates:
            p *= u
        return p


@dataclass
class CondConfig:
    """Voice-conditioning encoders (reference src/tts_streaming.py:357-384; upstream S3TokenizerV2, CAMPPlus, VoiceEncoder)."""
    tok_mels: int = 128
    tok_dim: int = 1280
    tok_heads: int = 20
    tok_layers: int = 6
    tok_fsmn_kernel: int = 31
    xv_feat: int = 80
    xv_dim: int = 192
    xv_growth: int = 32
    xv_init: int = 128
    xv_blocks: tuple = ((12, 3, 1), (24, 3, 2), (16, 3, 2))    # (layers, kernel, dilation) of the three dense TDNN blocks
    ve_mels: int = 40
    ve_hidden: int = 256
    ve_layers: int = 3
    ve_embed: int = 256

    @staticmethod
    def tiny() -> "CondConfig":
        return CondConfig(tok_layers=2, xv_blocks=((3, 3, 1), (4, 3, 2), (2, 3, 2)))


@dataclass
class ModelConfig:
    t3: T3Config = field(default_factory=T3Config)
    flow: FlowConfig = field(default_factory=FlowConfig)
    hift: HiFTConfig = field(default_factory=HiFTConfig)
    cond: CondConfig = field(default_factory=CondConfig)

    @staticmethod
    def tiny() -> "ModelConfig":
        return ModelConfig(t3=T3Config.tiny(), flow=FlowConfig.tiny(), hift=HiFTConfig(), cond=CondConfig.tiny())

    def to_dict(self):
        return asdict(self)


S3_SR = 16000
S3GEN_SR = 24000
SPEECH_VOCAB_SIZE = 6561
