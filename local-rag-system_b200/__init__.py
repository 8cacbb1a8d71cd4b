"""local-rag-system_b200: B200-native exact dense retrieval behind the Chroma-style
collection API that akak0487521/Local-RAG-System calls.

The directory name carries a hyphen (repo convention); import it as
`local_rag_system_b200` -- the tiny loader module of that name at the repo
root maps one onto the other.
"""
from .collection import Client, Collection, EphemeralClient, PersistentClient  # noqa: F401
from .engine import DeviceStore, ShardedDeviceStore, merge_keys_device  # noqa: F401
from . import embedding_functions  # noqa: F401

__version__ = "0.2.0"
__all__ = ["Client", "Collection", "EphemeralClient", "PersistentClient", "DeviceStore", "ShardedDeviceStore",
           "merge_keys_device", "embedding_functions"]
