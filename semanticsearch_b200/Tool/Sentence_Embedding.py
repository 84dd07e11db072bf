"""Boundary stub for the transformer encoder (reference: Tool/Sentence_Embedding.py:75-150).

The encoder forward pass is outside the hot path; this module only keeps
the call signature ``sentence_embedding(text_list, model_name, batch_size, device_preference)``
and lets the caller plug any encoder in.  The default backend loads ``sentence_transformers``
lazily when it is installed; otherwise the call fails loudly.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import numpy as np

EmbeddingBackend = Callable[..., np.ndarray]

_backend: Optional[EmbeddingBackend] = None
loaded_models: Dict[str, object] = {}


def set_embedding_backend(fn: Optional[EmbeddingBackend]) -> None:
    """Install ``fn(text_list, model_name, batch_size, device_preference) -> ndarray[N, d]``."""
    global _backend
    _backend = fn


def _default_backend(text_list, model_name, batch_size=32, device_preference=None):
    try:
        from sentence_transformers import SentenceTransformer  # type: ignore
    except Exception as exc:  # pragma: no cover - depends on the environment
        raise RuntimeError(
            "no embedding backend installed: call semanticsearch_b200.Tool.Sentence_Embedding."
            "set_embedding_backend(fn) or install sentence-transformers") from exc
    model = loaded_models.get(model_name)
    if model is None:
        model = SentenceTransformer(model_name, device="cuda")
        loaded_models[model_name] = model
    return model.encode(list(text_list), batch_size=batch_size, show_progress_bar=False)


def sentence_embedding(text_list: list, model_name: str, batch_size: int = 32, device_preference: Optional[str] = None):
    """Embed ``text_list``; same signature as the reference (Tool/Sentence_Embedding.py:75)."""
    fn = _backend or _default_backend
    return fn(text_list, model_name, batch_size=batch_size, device_preference=device_preference)
