"""Host tests of the sentence-segmentation boundary against the reference's Tool/Sentence_Segmenter.py: the rule-based
fallback, and the post-processing of spaCy's sentences (over-long sentences are re-split, never truncated; every
sentence ends with punctuation) driven through a fake spaCy pipeline on both sides."""
import types

import pytest

from semanticsearch_b200.Tool import Sentence_Segmenter as seg

TEXT = ("Alpha beta gamma delta. Short. The second real sentence has no end mark "
        "Another one follows here! Is this the fourth sentence? yes it is; and it goes on and on and on; until the very end "
        "of a long tail without punctuation")


class _FakeSent:
    def __init__(self, text):
        self.text = text


class _FakeNlp:
    """Splits on ' | ' — stands in for spaCy's sentence boundaries on both sides."""
    def __call__(self, text):
        return types.SimpleNamespace(sents=[_FakeSent(t) for t in text.split(" | ")])


def _reference_module():
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference tree not mounted")
    ref_shim.load_reference()
    import importlib
    return importlib.import_module("Tool.Sentence_Segmenter")


def test_rule_based_fallback_matches_reference():
    ref = _reference_module()
    for text in (TEXT, "", "   ", "One. Two. Three.", "No capital after the stop. then lower case. And Upper again here."):
        assert seg._rule_split(text) == ref._fallback_sentence_split(text)


def test_spacy_branch_post_processing_matches_reference(monkeypatch):
    ref = _reference_module()
    fake = _FakeNlp()
    monkeypatch.setattr(ref, "_ensure_spacy_model", lambda: True)
    monkeypatch.setattr(ref, "_nlp_model", fake, raising=False)
    monkeypatch.setattr(seg, "_nlp", fake)
    monkeypatch.setattr(seg, "_nlp_failed", False)
    docs = [
        "A first sentence without end | tiny | " + "word " * 60 + "end; " + "more words here " * 20 + "stop. Tail piece after the stop | Final one!",
        "x" * 50 + " | " + "; ".join("clause number %d of a very long sentence" % i for i in range(40)),
        "",
    ]
    for doc in docs:
        for limit in (1000, 120, 40):
            assert seg.extract_sentences_spacy(doc, limit) == ref.extract_sentences_spacy(doc, limit), (doc[:30], limit)
    # nothing is truncated: every character of a long sentence survives in some piece
    long_doc = "; ".join("clause number %d of a very long sentence" % i for i in range(40))
    pieces = seg.extract_sentences_spacy(long_doc, 100)
    assert sum(len(p) for p in pieces) >= len(long_doc) - 2 * len(pieces)


def test_token_count_fallback_matches_reference(monkeypatch):
    ref = _reference_module()
    monkeypatch.setattr(ref, "_ensure_spacy_model", lambda: False)
    monkeypatch.setattr(seg, "_nlp", None)
    for text in ("Hello, world! It's 3.5 degrees.", "", "a  b\tc"):
        assert seg.count_tokens_spacy(text) == ref.count_tokens_spacy(text)
