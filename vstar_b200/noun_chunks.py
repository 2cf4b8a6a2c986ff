"""Noun-chunk extraction for the context-cue branch of the guided search.

Mirror of /root/reference/visual_search.py:54-112 (`tranverse`, `get_noun_chunks`, `filter_chunk_list`,
`extract_noun_chunks`): the weak-cue branch asks the VSM where the target would be, and when the answer's location phrase
holds exactly ONE noun chunk that chunk becomes the segmentation prompt (visual_search.py:435-440).  The reference parses
with spaCy `en_core_web_sm` loaded at import time (visual_search.py:9-10); here the pipeline is loaded on first use.  If
spaCy (or the model) is not installed the call RAISES — silently answering "no chunks" would change the segmentation prompt,
hence the heat-map, hence the search trajectory.

Any callable `nlp(text) -> doc` with spaCy's Token surface (`.i`, `.pos_`, `.dep_`, `.children`, `doc[a:b].text`) can be
injected with `set_nlp` (tests use a deterministic rule parser).
"""
from __future__ import annotations

_nlp = None

LEFT_DEPS = ("amod", "compound", "poss")      # modifiers pulled into the chunk on the left of the head noun
RIGHT_DEPS = ("relcl", "prep")                # attachments pulled in on the right


class NounChunkerUnavailable(RuntimeError):
    pass


def set_nlp(nlp):
    """inject the parser (spaCy Language object or anything with the same call/Token surface); None = reload lazily"""
    global _nlp
    _nlp = nlp


def get_nlp():
    global _nlp
    if _nlp is None:
        try:
            import spacy
            _nlp = spacy.load("en_core_web_sm")
        except Exception as e:          # ImportError, or OSError when the model package is missing
            raise NounChunkerUnavailable(
                "the weak-cue branch of visual_search() needs spaCy + en_core_web_sm (requirements.txt:7,33 of the reference; "
                "visual_search.py:9-10) to build the context-cue phrase; install them or inject a parser with "
                f"vstar_b200.noun_chunks.set_nlp(...)  [{e!r}]") from e
    return _nlp


def subtree_span(token):
    """(leftmost, rightmost) token index of the dependency subtree under `token` (reference: `tranverse`)"""
    lo = hi = token.i
    stack = list(token.children)
    while stack:
        t = stack.pop()
        lo, hi = min(lo, t.i), max(hi, t.i)
        stack.extend(t.children)
    return lo, hi


def get_noun_chunks(token):
    """span of the chunk headed by `token`: contiguous run of amod/compound/poss children directly left of the head (nearest
    first, stop at the first other dependency), and of relcl/prep children on the right (visual_search.py:65-87)"""
    left = [c for c in token.children if c.i < token.i]
    right = [c for c in token.children if not (c.i < token.i)]
    start = token.i
    for child in reversed(left):
        if child.dep_ not in LEFT_DEPS:
            break
        start, _ = subtree_span(child)
    end = token.i
    for child in right:
        if child.dep_ not in RIGHT_DEPS:
            break
        _, end = subtree_span(child)
    return start, end


def filter_chunk_list(chunks):
    """longest chunks first; drop any chunk that touches/overlaps an already kept one (overlap length >= 0 counts, i.e. sharing
    one token); result ordered by start (visual_search.py:89-102)"""
    kept = []
    for c in sorted(chunks, key=lambda c: c[1] - c[0], reverse=True):      # stable, like the reference's sorted()
        if all(min(k[1], c[1]) - max(k[0], c[0]) < 0 for k in kept):
            kept.append(c)
    return sorted(kept, key=lambda c: c[0])


def extract_noun_chunks(expression, nlp=None):
    """visual_search.py:104-112: chunks headed by every NOUN / PRON token, filtered, as text"""
    doc = (nlp if nlp is not None else get_nlp())(expression)
    spans = [get_noun_chunks(t) for t in doc if t.pos_ in ("NOUN", "PRON")]
    return [doc[a:b + 1].text for a, b in filter_chunk_list(spans)]
