"""Sentence segmentation boundary (reference: Tool/Sentence_Segmenter.py:126-177, fallback :99-124).

Text segmentation is a CPU text stage outside the hot path; it only determines the ragged
sentence counts.  spaCy is used when it is installed with its English model, otherwise a
punctuation rule with the same contract as the reference's fallback: split after ``. ! ?`` when
followed by whitespace and an upper-case letter, drop fragments shorter than 10 characters, and
make every sentence end with punctuation.
"""
from __future__ import annotations

import re
from typing import Callable, List, Optional

_SPLIT_RE = re.compile(r"(?<=[.!?])\s+(?=[A-Z])")
_END_RE = re.compile(r"[.!?]$")
_nlp = None
_nlp_failed = False
_override: Optional[Callable[[str], List[str]]] = None


def set_sentence_splitter(fn: Optional[Callable[[str], List[str]]]) -> None:
    """Install a custom splitter (used by tests and by callers that already hold sentences)."""
    global _override
    _override = fn


def _rule_split(text: str) -> List[str]:
    if not text or not isinstance(text, str):
        return []
    flat = re.sub(r"\s+", " ", text.strip())
    out = []
    for piece in _SPLIT_RE.split(flat):
        piece = piece.strip()
        if len(piece) < 10:
            continue
        out.append(piece if _END_RE.search(piece) else piece + ".")
    return out


def extract_sentences_spacy(text: str, max_sent_length: int = 1000) -> List[str]:
    global _nlp, _nlp_failed
    if _override is not None:
        return _override(text)
    if not text or not isinstance(text, str):
        return []
    if _nlp is None and not _nlp_failed:
        try:
            import spacy  # type: ignore
            _nlp = spacy.load("en_core_web_sm")
        except Exception:
            _nlp_failed = True
    if _nlp is None:
        return _rule_split(text)
    sents = []
    for sent in _nlp(text).sents:
        s = sent.text.strip()
        if s and len(s) >= 10:
            sents.append(s[:max_sent_length])
    return sents


def count_tokens_spacy(text: str) -> int:
    return len(text.split()) if text else 0
