"""The C-ABI library loads, exports every symbol include/cbx_b200.h declares, its tensor registry matches the
host packer, and it refuses to run without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest
import torch

from cbx_b200 import lib as L
from cbx_b200.config import ModelConfig
from cbx_b200.pack import pack_state_dict
from cbx_b200.weights import random_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "cbx_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cbx_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = L.load()
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/cbx_b200.h but not exported"
        assert s in L.SIGNATURES, f"{s} has no ctypes signature"
    assert lib.cbx_abi_version() == 2


def _manifest(cfg):
    lib = L.load()
    f = cfg.flow
    cc = L.CbxConfig(cfg.t3.n_layers, f.enc_blocks, f.up_blocks, f.n_blocks, f.n_mid, f.n_timesteps, f.cfg_rate, 8, 1536, 512, 1056, 250, 8, 2)
    h = C.c_void_p()
    L.check(lib.cbx_manifest_create(C.byref(cc), C.byref(h)))
    name = C.create_string_buffer(256)
    numel, dt = C.c_int64(), C.c_int()
    out = []
    for i in range(lib.cbx_tensor_count(h)):
        L.check(lib.cbx_tensor_info(h, i, name, 256, C.byref(numel), C.byref(dt)))
        out.append((name.value.decode(), numel.value, dt.value))
    lib.cbx_engine_destroy(h)
    return out


def test_packer_covers_registry_tiny():
    cfg = ModelConfig.tiny()
    man = _manifest(cfg)
    packed = pack_state_dict(random_state_dict(cfg, 0), cfg)
    assert {m[0] for m in man} == set(packed)
    for name, numel, dt in man:
        t = packed[name]
        assert t.numel() == numel and (t.dtype == torch.float32) == (dt == 0), name


def test_registry_full_model_sizes():
    man = _manifest(ModelConfig())
    names = {m[0]: m for m in man}
    assert names["t3.l29.wqkv_f"][1] == 3 * 1024 * 1024 and names["t3.head_f"][1] == 8208 * 1024
    assert names["cfm.r13.c1.w"][1] == 256 * 3 * 512          # up-stage resnet consumes the 512-channel skip concat
    t3_bytes = sum(n * (4 if d == 0 else 2) for k, n, d in man if k.startswith("t3.l") and k.endswith("_f"))
    assert t3_bytes == 30 * 16777216 * 2                        # decode streams exactly the bf16 projection weights


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = L.load()
    cfg = ModelConfig.tiny()
    f = cfg.flow
    cc = L.CbxConfig(cfg.t3.n_layers, f.enc_blocks, f.up_blocks, f.n_blocks, f.n_mid, f.n_timesteps, f.cfg_rate, 8, 1536, 512, 1056, 250, 8, 2)
    h = C.c_void_p()
    assert lib.cbx_engine_create(C.byref(cc), 0, C.byref(h)) != 0
    assert b"no CUDA device" in lib.cbx_last_error()
    from cbx_b200.native import NativeEngine
    from cbx_b200.engine import TextToSpeechEngine
    with pytest.raises(RuntimeError):
        NativeEngine(cfg)
    with pytest.raises(RuntimeError):
        TextToSpeechEngine("cpu")
