"""TEST INFRASTRUCTURE — import shim for the *unmodified* reference under /root/reference.

Only usable inside the build container (``/root/reference`` does not exist on the GPU box).
It is used by ``oracle/gen_golden.py`` to generate the committed fixtures in ``tests/golden/``
and by the ``-m "not gpu"`` tests that pin ``oracle/*.py`` against the reference's own
functions when the reference tree is present.  Nothing in the product package imports it.

The reference's hot-path modules pull in leaf dependencies that are absent from this image
(``sentence_transformers``, ``spacy``, ``rank_bm25``, ``pyopenie``); SURVEY.md §8(c) describes
the four stubs.  The encoder is replaced by a lookup table (sentence string -> vector) so the
reference's real numpy / scikit-learn arithmetic runs on synthetic embeddings we control:

* ``Tool.Sentence_Embedding.sentence_embedding``      (Tool/Sentence_Embedding.py:75)
* ``Method.semantic_common.embed_text_list``           (Method/semantic_common.py:28)
* ``extract_sentences_spacy`` in the Grouping / Splitter namespaces
  (Method/Semantic_Grouping_Optimized.py:7, Method/Semantic_Splitter_Optimized.py import)
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types
from typing import Dict, List, Optional

import numpy as np

REFERENCE_ROOT = os.environ.get("SEMSEARCH_REFERENCE_ROOT", "/root/reference")
SENT_DELIM = " || "

_TABLE: Dict[str, np.ndarray] = {}
_loaded = None


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "Method"))


def _install_stubs() -> None:
    if "sentence_transformers" not in sys.modules:
        m = types.ModuleType("sentence_transformers")

        class SentenceTransformer:  # noqa: D401 - stub
            def __init__(self, *a, **k):
                self.device = "cpu"

            def encode(self, text_list, **_k):
                return fake_sentence_embedding(list(text_list), "stub")

        class CrossEncoder:  # dead code in the reference imports it lazily
            def __init__(self, *a, **k):
                raise RuntimeError("CrossEncoder stub")

        m.SentenceTransformer = SentenceTransformer
        m.CrossEncoder = CrossEncoder
        sys.modules["sentence_transformers"] = m
    if "pyopenie" not in sys.modules:
        m = types.ModuleType("pyopenie")

        class OpenIE5:
            def __init__(self, *a, **k):
                pass

            def extract(self, *_a, **_k):
                return []

        m.OpenIE5 = OpenIE5
        sys.modules["pyopenie"] = m
    if "rank_bm25" not in sys.modules:
        m = types.ModuleType("rank_bm25")

        class BM25Okapi:
            """Stand-in: BM25 is lexical and off the dense path."""

            def __init__(self, corpus, epsilon=0.25):
                self.n = len(corpus)

            def get_scores(self, _q):
                return np.zeros(self.n, dtype=float)

        m.BM25Okapi = BM25Okapi
        sys.modules["rank_bm25"] = m
    if "spacy" not in sys.modules:
        m = types.ModuleType("spacy")

        def load(_name, *a, **k):
            raise OSError("spacy stub: no model")

        m.load = load
        sys.modules["spacy"] = m


def fake_sentence_embedding(text_list, model_name=None, batch_size: int = 32, device_preference=None):
    """Replacement encoder: looks each string up in the injected table."""
    if not text_list:
        return np.zeros((0, 0), dtype=np.float32)
    return np.stack([_TABLE[t] for t in text_list]).astype(np.float32, copy=False)


def fake_extract_sentences(text: str, max_sent_length: int = 1000) -> List[str]:
    if not text:
        return []
    return [s.strip() for s in text.split(SENT_DELIM.strip()) if s.strip()]


def set_embeddings(sentences: List[str], vectors: np.ndarray) -> None:
    _TABLE.clear()
    for s, v in zip(sentences, vectors):
        _TABLE[s] = np.asarray(v, dtype=np.float32)


def make_doc(vectors: np.ndarray, tag: str = "d"):
    """Register one synthetic document; returns (passage_text, sentences)."""
    sents = [f"{tag}s{i:05d}x" for i in range(len(vectors))]
    for s, v in zip(sents, vectors):
        _TABLE[s] = np.asarray(v, dtype=np.float32)
    return SENT_DELIM.join(sents), sents


class _Ref(types.SimpleNamespace):
    pass


def load_reference() -> Optional[_Ref]:
    """Import the reference hot-path modules (cached).  Returns None if the tree is absent."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        return None
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):
        import Tool.Sentence_Embedding as ref_embed  # noqa: E402
        import Tool.Sentence_Segmenter as ref_seg  # noqa: E402
        import Method.semantic_common as ref_common  # noqa: E402
        import Method.Semantic_Grouping_Optimized as ref_group  # noqa: E402
        import Method.Semantic_Splitter_Optimized as ref_split  # noqa: E402
        import Tool.rank_chunks_optimized as ref_rank  # noqa: E402
    ref_embed.sentence_embedding = fake_sentence_embedding
    ref_common.embed_text_list = fake_sentence_embedding
    ref_group.extract_sentences_spacy = fake_extract_sentences
    ref_split.extract_sentences_spacy = fake_extract_sentences
    _loaded = _Ref(common=ref_common, group=ref_group, split=ref_split, rank=ref_rank,
                   embed=ref_embed, seg=ref_seg)
    return _loaded
