"""Embedding-function hooks (the step *before* the hot path).

The reference builds `SentenceTransformerEmbeddingFunction(model_name=...)`
(api/app.py:90, scripts/build_index.py:16) and hands it to
get_or_create_collection; Chroma calls it with a list of texts whenever
`documents=` / `query_texts=` arrive without embeddings.  Text embedding is out
of scope for this engine (it is a transformer forward pass, not retrieval), so
this class simply defers to the sentence-transformers package when it is
installed and otherwise fails with a clear message on first use.  It stays a
plain module attribute so tests can monkeypatch it (tests/test_kb_crud.py:62-66
in the reference).
"""
from __future__ import annotations

from typing import List, Sequence


class EmbeddingFunction:
    def __call__(self, input: Sequence[str]) -> List[List[float]]:  # noqa: A002
        raise NotImplementedError


class SentenceTransformerEmbeddingFunction(EmbeddingFunction):
    def __init__(self, model_name: str = "all-MiniLM-L6-v2", device: str = "cpu",
                 normalize_embeddings: bool = False, **kwargs):
        self.model_name, self.device = model_name, device
        self.normalize_embeddings = normalize_embeddings
        self._kwargs = kwargs
        self._model = None

    def _load(self):
        if self._model is None:
            try:
                from sentence_transformers import SentenceTransformer
            except ImportError as e:   # not installable in the offline build image
                raise ValueError(
                    "sentence-transformers is not installed: pass embeddings= / query_embeddings= "
                    "directly or supply your own embedding_function") from e
            self._model = SentenceTransformer(self.model_name, device=self.device, **self._kwargs)
        return self._model

    def __call__(self, input: Sequence[str]) -> List[List[float]]:  # noqa: A002
        model = self._load()
        return model.encode(list(input), convert_to_numpy=True,
                            normalize_embeddings=self.normalize_embeddings).tolist()


class DefaultEmbeddingFunction(SentenceTransformerEmbeddingFunction):
    pass
