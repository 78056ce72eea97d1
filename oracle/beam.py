"""Oracle restatement of the host beam search (reference: src/beam.rs).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pure Python, generic over the token type and
the ``next`` / ``is_finished`` closures like the reference; tie-breaks are the reference's:

  * get_top_elements (beam.rs:81-110): ascending insertion list; a candidate equal to the
    current minimum of a full list is inserted at 0 and immediately evicted, so on exact
    ties the EARLIER element wins; for k = 1 this is first-index argmax.
  * beam_search (beam.rs:9-37): Rust ``Iterator::max_by`` keeps the LAST maximum.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, List, Sequence, Tuple


@dataclass
class BeamNode:
    """beam.rs:3-7."""
    seq: List[Any]
    log_prob: float


def _max_by_last(beams: Sequence[BeamNode]):
    """Iterator::max_by with partial_cmp: returns the last element among equal maxima."""
    best = None
    for b in beams:
        if best is None or not (b.log_prob < best.log_prob):
            best = b
    return best


def get_top_elements(elems: Sequence[Any], score: Callable[[Any], float], num: int) -> List[Any]:
    """beam.rs:81-110."""
    top_elems: List[Any] = []
    scores: List[float] = []
    for elem in elems:
        s = score(elem)
        if len(top_elems) == num:
            if s < scores[0]:
                continue
        idx = None
        for i, existing in enumerate(scores):
            if existing >= s:
                idx = i
                break
        if idx is not None:
            top_elems.insert(idx, elem)
            scores.insert(idx, s)
        else:
            top_elems.append(elem)
            scores.append(s)
        if len(top_elems) > num:
            top_elems.pop(0)
            scores.pop(0)
    return top_elems


def beam_search_step(beams: List[BeamNode], next_fn, is_finished, beam_size: int) -> List[BeamNode]:
    """beam.rs:39-79.  ``next_fn`` is evaluated for ALL beams incl. finished ones (:53)."""
    finished_beams: List[BeamNode] = []
    new_beams: List[BeamNode] = []
    continuations = next_fn(beams)
    for beam_node, conts in zip(beams, continuations):
        if is_finished(beam_node.seq):
            finished_beams.append(beam_node)
        else:
            for tok, log_prob in get_top_elements(conts, lambda c: c[1], beam_size):
                new_beams.append(BeamNode(seq=beam_node.seq + [tok], log_prob=log_prob))
    return get_top_elements(new_beams, lambda b: b.log_prob, beam_size) \
        + get_top_elements(finished_beams, lambda b: b.log_prob, beam_size)


def beam_search(initial_beams: List[BeamNode], next_fn, is_finished, beam_size: int,
                max_depth: int, trace: list | None = None) -> List[Any]:
    """beam.rs:9-37."""
    beams = initial_beams
    for _ in range(max_depth):
        best = _max_by_last(beams)
        if best is not None and is_finished(best.seq):
            break
        beams = beam_search_step(beams, next_fn, is_finished, beam_size)
        if trace is not None:
            trace.append([(list(b.seq), b.log_prob) for b in beams])
    best = _max_by_last(beams)
    return list(best.seq) if best is not None else []
