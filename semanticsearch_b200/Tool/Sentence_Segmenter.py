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


_SUBSPLIT_RE = re.compile(r"(?<=[.!?;])\s+")
_TOKEN_RE = re.compile(r"\b\w+\b|[^\w\s]")


def _post_process(sent_texts, max_sent_length: int) -> List[str]:
    """The reference's treatment of spaCy's sentences (:150-170): fragments under 10 characters are dropped, sentences
    longer than ``max_sent_length`` are re-split after ``. ! ? ;`` (pieces under 10 characters dropped), and every
    emitted sentence ends with ``. ! ?`` (a full stop is appended otherwise).  Nothing is truncated."""
    out: List[str] = []
    for raw in sent_texts:
        s = raw.strip()
        if len(s) < 10:
            continue
        if len(s) > max_sent_length:
            for piece in _SUBSPLIT_RE.split(s):
                piece = piece.strip()
                if len(piece) >= 10:
                    out.append(piece if _END_RE.search(piece) else piece + ".")
        else:
            out.append(s if _END_RE.search(s) else s + ".")
    return out


def extract_sentences_spacy(text: str, max_sent_length: int = 1000) -> List[str]:
    global _nlp, _nlp_failed
    if _override is not None:
        return _override(text)
    if not text or not isinstance(text, str) or not text.strip():
        return []
    if _nlp is None and not _nlp_failed:
        try:
            import spacy  # type: ignore
            _nlp = spacy.load("en_core_web_sm")
        except Exception:
            _nlp_failed = True
    if _nlp is None:
        return _rule_split(text)
    try:
        return _post_process((sent.text for sent in _nlp(text).sents), max_sent_length)
    except Exception:
        return _rule_split(text)   # reference :175-177


def count_tokens_spacy(text: str) -> int:
    """Reference :179-200: spaCy's non-space tokens, or the regex count (words and single punctuation marks)."""
    if not text or not isinstance(text, str):
        return 0
    if _nlp is not None:
        try:
            return len([t for t in _nlp(text) if not t.is_space])
        except Exception:
            pass
    return len(_TOKEN_RE.findall(text.strip()))
