"""`chatterbox`-named shim over the B200-native engine: the object surface the reference engine imports and calls
(src/tts_streaming.py:38-44, :254-257, :283-292, :316-320, :366-381, :420-435, :586-590, :667; SURVEY 8b level 2), so the
reference's tts_streaming.py / worker.py run UNMODIFIED on libcbx_b200.so.  Put `chatterbox-tts_b200/` on PYTHONPATH in
front of (or instead of) the upstream package.  See INTEGRATION.md section C.

This path keeps the reference's orchestration (one generator per request, one S3Gen call per slice) and therefore none of
the cross-request batching of cbx_b200.engine; it exists for drop-in compatibility, the engine class is the fast way."""
from ._backend import set_backend_factory  # noqa: F401
