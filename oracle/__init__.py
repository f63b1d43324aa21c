"""ORACLE package — test infrastructure only (see headers of t3.py / flow.py / hift.py).
PARITY UNPINNED: the reference's arithmetic is in an un-vendored, unpinned dependency."""
import os, sys
_pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "chatterbox-tts_b200")
if _pkg not in sys.path:
    sys.path.insert(0, _pkg)
