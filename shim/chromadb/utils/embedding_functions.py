"""chromadb.utils.embedding_functions -- re-exported so that
`from chromadb.utils.embedding_functions import SentenceTransformerEmbeddingFunction`
(api/app.py:88) resolves, and stays monkeypatch-able as a module attribute."""
from local_rag_system_b200.embedding_functions import (  # noqa: F401
    DefaultEmbeddingFunction, EmbeddingFunction, SentenceTransformerEmbeddingFunction)
