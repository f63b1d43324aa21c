"""Host side of the B200-native Chatterbox hot path (ctypes over libcbx_b200.so)."""
from .config import ModelConfig, T3Config, FlowConfig, HiFTConfig, S3_SR, S3GEN_SR  # noqa: F401
