"""Opt-in `chromadb`-named shim: put `<repo>/shim` on PYTHONPATH and the reference
app (api/app.py:87-91), its scripts and its tests import this package in place
of chromadb==0.5.3 and run unmodified on the B200 engine."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)

from local_rag_system_b200 import (Client, Collection, EphemeralClient,  # noqa: E402,F401
                                   PersistentClient)
from . import utils  # noqa: E402,F401

__version__ = "0.5.3+b200"
