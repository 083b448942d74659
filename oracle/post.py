"""Post-inference oracle: activation -> top-k + threshold -> range/species mask -> threshold -> sort.

Test infrastructure only (see oracle/__init__.py).  Paths cited are relative to
/root/reference.  The order of operations is SURVEY.md §0 F6: the mask is applied to
the already-truncated top-5 list, then ``>= min_confidence`` is tested a second time.

activation/top-k live in the third-party crate birdnet-onnx 2.0.0-rc.16 (not vendored):
restated as "conf = act(score); keep conf >= min_conf; order by conf desc, ties by lower
class index; take top_k" — PARITY UNPINNED for that step.  The mask step is in-tree
(src/inference/geomodel_filter.rs:45-79) and pinned by its unit tests.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

f32 = np.float32

ACT_NONE, ACT_SIGMOID, ACT_SOFTMAX = 0, 1, 2     # mirrors bb_activation


@dataclass
class Prediction:
    """birdnet_onnx::Prediction as consumed at src/pipeline/processor.rs:372-383."""
    species: str
    confidence: np.float32
    index: int


@dataclass
class FilterSettings:
    """src/inference/geomodel_filter.rs:13-35."""
    threshold: float = 0.01
    keep_unmatched: bool = True     # UnmatchedPolicy::Keep
    rerank: bool = False

    def keeps_unmatched(self) -> bool:
        return self.keep_unmatched and not self.rerank


def activate(scores: np.ndarray, act: int) -> np.ndarray:
    """[B, C] f32 scores -> confidences (f32).  Sigmoid for BirdNET v2.4 logits, softmax
    for Perch v2 (manifests/Perch-v2-Models.models.json:15), identity when the graph
    already applies it (BirdNET v3.0)."""
    s = np.asarray(scores, dtype=np.float32)
    if act == ACT_NONE:
        return s
    if act == ACT_SIGMOID:
        with np.errstate(over="ignore"):
            return (f32(1.0) / (f32(1.0) + np.exp(-s.astype(np.float64)))).astype(np.float32)
    if act == ACT_SOFTMAX:
        z = s.astype(np.float64)
        z = np.exp(z - z.max(axis=-1, keepdims=True))
        return (z / z.sum(axis=-1, keepdims=True)).astype(np.float32)
    raise ValueError("unknown activation")


def top_k_threshold(conf_row: np.ndarray, top_k: int, min_conf: float) -> List[Tuple[int, np.float32]]:
    """At most ``top_k`` (index, confidence) with ``confidence >= min_conf``, descending,
    ties broken by the lower class index.  Configured at src/inference/classifier.rs:269-273."""
    c = np.asarray(conf_row, dtype=np.float32)
    cand = np.nonzero(c >= f32(min_conf))[0]
    if cand.size == 0:
        return []
    order = np.lexsort((cand, -c[cand].astype(np.float64)))   # primary: conf desc, secondary: index asc
    sel = cand[order][:top_k]
    return [(int(i), f32(c[i])) for i in sel]


def filter_row(preds: List[Tuple[int, np.float32]], mask: Optional[np.ndarray],
               settings: FilterSettings, species_keep: Optional[np.ndarray] = None
               ) -> List[Tuple[int, np.float32]]:
    """``apply_range_filter`` on one result.  src/inference/classifier.rs:587-645 ->
    src/inference/geomodel_filter.rs:45-79.

    ``mask``: dense [C] f32 projection of GeomodelScores; NaN = the label has no entry
    (``score_of`` -> None).  ``species_keep``: [C] bool for the static species list
    (mutually exclusive with the range filter, src/lib.rs:502-520)."""
    if mask is not None:
        out: List[Tuple[int, np.float32]] = []
        for idx, conf in preds:
            s = f32(mask[idx])
            if np.isnan(s):
                if settings.keeps_unmatched():
                    out.append((idx, conf))
            elif s >= f32(settings.threshold):
                out.append((idx, f32(conf * s) if settings.rerank else conf))
            # else: mapped but out of range -> drop
        if settings.rerank:
            # sort_unstable_by(total_cmp) descending; oracle breaks ties by original order
            out = [p for _, p in sorted(enumerate(out), key=lambda t: (-float(t[1][1]), t[0]))]
        return out
    if species_keep is not None:
        return [(i, c) for i, c in preds if bool(species_keep[i])]
    return list(preds)


def post_process(scores: np.ndarray, valid: int, act: int, min_conf: float, top_k: int = 5,
                 mask: Optional[np.ndarray] = None, settings: Optional[FilterSettings] = None,
                 species_keep: Optional[np.ndarray] = None
                 ) -> List[List[Tuple[int, np.float32]]]:
    """Rows [0, valid) of a padded batch -> per-segment detection candidates, i.e. what
    survives src/pipeline/processor.rs:268-277, :317 and the test at :374."""
    settings = settings or FilterSettings()
    conf = activate(scores, act)
    out = []
    for r in range(valid):
        p = top_k_threshold(conf[r], top_k, min_conf)
        p = filter_row(p, mask, settings, species_keep)
        p = [(i, c) for i, c in p if c >= f32(min_conf)]      # processor.rs:374
        out.append(p)
    return out


# --------------------------------------------------------------------------- A9: label-space projection
def scientific_name(label: str) -> str:
    """src/inference/geomodel.rs:28-33."""
    if "_" in label:
        prefix = label.split("_", 1)[0]
        if " " in prefix:
            return prefix
    return label


def species_key(label: str) -> str:
    """src/inference/geomodel.rs:36-38."""
    return scientific_name(label).lower()


class SpeciesMapping:
    """src/inference/geomodel.rs:41-127 — first classifier label wins on a key collision."""

    def __init__(self, geomodel_labels: Sequence[str], classifier_labels: Sequence[str]):
        by_key: Dict[str, str] = {}
        for lab in classifier_labels:
            k = species_key(lab)
            if k not in by_key:
                by_key[k] = lab
        self.by_species_key: Dict[str, str] = {}
        for g in geomodel_labels:
            k = species_key(g)
            if k in by_key:
                self.by_species_key[k] = by_key[k]
        self.total = len(classifier_labels)

    def classifier_label_for(self, geomodel_label: str) -> Optional[str]:
        return self.by_species_key.get(species_key(geomodel_label))

    def mapped_count(self) -> int:
        return len(self.by_species_key)

    def unmatched_count(self) -> int:
        return max(self.total - self.mapped_count(), 0)


class GeomodelScores:
    """src/inference/geomodel.rs:129-180 — label -> occurrence score, mapped-but-unreported = 0."""

    def __init__(self, scores: Iterable[Tuple[str, float]], mapping: SpeciesMapping):
        self.by_label: Dict[str, np.float32] = {lab: f32(0.0) for lab in mapping.by_species_key.values()}
        for species, score in scores:
            lab = mapping.classifier_label_for(species)
            if lab is not None:
                self.by_label[lab] = f32(score)

    def score_of(self, label: str) -> Optional[np.float32]:
        return self.by_label.get(label)

    def in_range_count(self, threshold: float) -> int:
        return sum(1 for v in self.by_label.values() if v >= f32(threshold))

    def is_empty(self) -> bool:
        return not self.by_label

    def dense_mask(self, classifier_labels: Sequence[str]) -> np.ndarray:
        """The GPU-side form: [C] f32, NaN where ``score_of(label)`` is None.  Two classifier
        rows with the SAME label string share one hash entry in the reference, so they share
        the score here too; a second label with the same scientific name but a different
        string has no entry (NaN)."""
        m = np.full(len(classifier_labels), np.nan, dtype=np.float32)
        for i, lab in enumerate(classifier_labels):
            v = self.by_label.get(lab)
            if v is not None:
                m[i] = v
        return m


def filter_predictions(preds: List[Prediction], scores: GeomodelScores,
                       settings: FilterSettings) -> List[Prediction]:
    """Label-keyed form of the filter, literal to src/inference/geomodel_filter.rs:45-79;
    used to check the dense-mask form against the reference's own unit tests."""
    out: List[Prediction] = []
    for p in preds:
        s = scores.score_of(p.species)
        if s is None:
            if settings.keeps_unmatched():
                out.append(Prediction(p.species, p.confidence, p.index))
        elif s >= f32(settings.threshold):
            c = f32(f32(p.confidence) * s) if settings.rerank else p.confidence
            out.append(Prediction(p.species, c, p.index))
    if settings.rerank:
        out = [p for _, p in sorted(enumerate(out), key=lambda t: (-float(t[1].confidence), t[0]))]
    return out


# --------------------------------------------------------------------------- A10
@dataclass
class Detection:
    """src/output/types.rs:8-23 (fields used by the writers)."""
    scientific_name: str
    common_name: str
    confidence: np.float32
    start_time: np.float32
    end_time: np.float32
    index: int = -1
    segment: int = -1


def split_label(label: str) -> Tuple[str, str]:
    """``Detection::from_label`` — split at the first '_'.  src/output/types.rs:58-79."""
    if "_" in label:
        a, b = label.split("_", 1)
        return a, b
    return label, label


def extract_detections(rows: List[List[Tuple[int, np.float32]]], start_time: np.ndarray,
                       end_time: np.ndarray, labels: Optional[Sequence[str]] = None,
                       min_conf: float = 0.0, first_segment: int = 0) -> List[Detection]:
    """src/pipeline/processor.rs:363-385."""
    dets: List[Detection] = []
    for r, preds in enumerate(rows):
        seg = first_segment + r
        for idx, conf in preds:
            if conf >= f32(min_conf):
                sci, com = split_label(labels[idx]) if labels is not None else (str(idx), str(idx))
                dets.append(Detection(sci, com, f32(conf), f32(start_time[seg]), f32(end_time[seg]), idx, seg))
    return dets


def sort_detections(dets: List[Detection]) -> List[Detection]:
    """src/pipeline/processor.rs:178-187 — (start_time asc, confidence desc).  The reference
    sort is unstable; the oracle fixes tie order (stable) so results are reproducible."""
    return sorted(dets, key=lambda d: (float(d.start_time), -float(d.confidence)))
