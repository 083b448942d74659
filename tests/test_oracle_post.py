"""Mask / mapping oracle against the reference's unit tests (CPU only).
Ports src/inference/geomodel_filter.rs:119-294 and src/inference/geomodel.rs:182-410."""
import numpy as np

from oracle import post
from oracle.post import FilterSettings, GeomodelScores, Prediction, SpeciesMapping, filter_predictions

f32 = np.float32


def scores_of(entries):
    labels = [s for s, _ in entries]
    mapping = SpeciesMapping(labels, labels)
    return GeomodelScores(entries, mapping)


def pred(species, conf, index=0):
    return Prediction(species, f32(conf), index)


def st(keep, rerank):
    return FilterSettings(threshold=0.01, keep_unmatched=keep, rerank=rerank)


def test_keeps_mapped_species_above_threshold():
    out = filter_predictions([pred("Parus major_x", 0.8)], scores_of([("Parus major_x", 0.5)]), st(True, False))
    assert len(out) == 1 and out[0].confidence == f32(0.8)


def test_drops_mapped_species_below_threshold():
    assert filter_predictions([pred("Parus major_x", 0.9)], scores_of([("Parus major_x", 0.005)]), st(True, False)) == []


def test_keeps_species_exactly_at_threshold():
    out = filter_predictions([pred("Parus major_x", 0.9)], scores_of([("Parus major_x", 0.01)]), st(True, False))
    assert len(out) == 1


def test_keep_policy_passes_unmatched_species_through():
    out = filter_predictions([pred("Dog_Dog", 0.7)], scores_of([("Parus major_x", 0.5)]), st(True, False))
    assert len(out) == 1 and out[0].confidence == f32(0.7)


def test_drop_policy_removes_unmatched_species():
    assert filter_predictions([pred("Dog_Dog", 0.7)], scores_of([("Parus major_x", 0.5)]), st(False, False)) == []


def test_rerank_scales_confidence_by_score():
    out = filter_predictions([pred("Parus major_x", 0.8)], scores_of([("Parus major_x", 0.5)]), st(True, True))
    assert abs(float(out[0].confidence) - 0.4) < 1e-6


def test_rerank_drops_unmatched_even_under_keep_policy():
    out = filter_predictions([pred("Parus major_x", 0.8), pred("Dog_Dog", 0.9)],
                             scores_of([("Parus major_x", 0.5)]), st(True, True))
    assert [p.species for p in out] == ["Parus major_x"]


def test_rerank_orders_plausible_species_above_implausible_one():
    out = filter_predictions([pred("Rara avis_y", 0.80), pred("Parus major_x", 0.70)],
                             scores_of([("Parus major_x", 0.9), ("Rara avis_y", 0.02)]), st(True, True))
    assert out[0].species == "Parus major_x" and len(out) == 2


def test_no_rerank_preserves_input_order():
    out = filter_predictions([pred("Aaa aaa_x", 0.5), pred("Bbb bbb_y", 0.4)],
                             scores_of([("Aaa aaa_x", 0.2), ("Bbb bbb_y", 0.9)]), st(True, False))
    assert [p.species for p in out] == ["Aaa aaa_x", "Bbb bbb_y"]


def test_prediction_index_survives_filtering():
    out = filter_predictions([pred("Parus major_x", 0.8, 42)], scores_of([("Parus major_x", 0.5)]), st(True, True))
    assert out[0].index == 42


def test_empty_and_all_unmatched():
    assert filter_predictions([], scores_of([("Aaa aaa_x", 0.9)]), st(True, True)) == []
    out = filter_predictions([pred("Dog_Dog", 0.7), pred("Siren_Siren", 0.6)], scores_of([]), st(True, False))
    assert len(out) == 2


# ---- geomodel.rs mapping tests ---------------------------------------------------------
def test_scientific_name():
    sn = post.scientific_name
    assert sn("Parus major_Great Tit") == "Parus major"
    assert sn("Parus major_Talitiainen") == "Parus major"
    assert sn("Parus major") == "Parus major"
    assert sn("Accelerating_and_revving_and_vroom") == "Accelerating_and_revving_and_vroom"
    assert sn("Accordion") == "Accordion" and sn("Dog_Dog") == "Dog_Dog"
    assert sn("Parus major_Great_Tit") == "Parus major"
    assert sn("") == ""


def test_mapping_tables():
    m = SpeciesMapping(["Parus major_Great Tit"], ["Parus major_Talitiainen"])
    assert (m.mapped_count(), m.unmatched_count()) == (1, 0)
    assert m.classifier_label_for("Parus major_Great Tit") == "Parus major_Talitiainen"
    assert SpeciesMapping(["Parus major_Great Tit"], ["Parus major"]).mapped_count() == 1
    assert SpeciesMapping(["parus major_Great Tit"], ["Parus Major_Talitiainen"]).mapped_count() == 1
    m = SpeciesMapping(["Parus major_Great Tit"],
                       ["Parus major_Great Tit", "Accipiter gentilis_Northern Goshawk", "Dog_Dog"])
    assert (m.mapped_count(), m.unmatched_count(), m.total) == (1, 2, 3)
    m = SpeciesMapping(["Parus major_Great Tit", "Petaurista albiventer_White-bellied Giant Flying Squirrel"],
                       ["Parus major_Great Tit"])
    assert (m.mapped_count(), m.unmatched_count()) == (1, 0)
    m = SpeciesMapping(["Parus major_Great Tit"], ["Parus major_First", "Parus major_Second"])
    assert m.mapped_count() == 1 and m.classifier_label_for("Parus major_Great Tit") == "Parus major_First"
    m = SpeciesMapping([], [])
    assert (m.mapped_count(), m.unmatched_count(), m.total) == (0, 0, 0)


def test_projection_tables():
    m = SpeciesMapping(["Parus major_Great Tit"], ["Parus major_Talitiainen"])
    p = GeomodelScores([("Parus major_Great Tit", 0.8)], m)
    assert p.score_of("Parus major_Talitiainen") == f32(0.8) and p.score_of("Parus major_Great Tit") is None
    m = SpeciesMapping(["Parus major_Great Tit"], ["Parus major_Great Tit"])
    assert GeomodelScores([], m).score_of("Parus major_Great Tit") == f32(0.0)
    m = SpeciesMapping(["Parus major_Great Tit"], ["Dog_Dog"])
    p = GeomodelScores([("Parus major_Great Tit", 0.8)], m)
    assert p.score_of("Dog_Dog") is None and p.is_empty()
    g = ["Aaa aaa_X", "Bbb bbb_Y", "Ccc ccc_Z"]
    p = GeomodelScores([("Aaa aaa_X", 0.9), ("Bbb bbb_Y", 0.005), ("Ccc ccc_Z", 0.02)], SpeciesMapping(g, g))
    assert (p.in_range_count(0.01), p.in_range_count(0.5), p.in_range_count(0.99)) == (2, 1, 0)


def test_dense_mask_equals_label_filter():
    """The dense [C] mask (GPU form) must give what the label-keyed reference form gives."""
    rng = np.random.default_rng(5)
    cls = [f"Gen{i} sp{i}_Common {i}" for i in range(40)] + ["Dog_Dog", "Siren_Siren", "Gen3 sp3_Second name"]
    geo = [f"Gen{i} sp{i}_English {i}" for i in range(0, 40, 2)] + ["Vulpes vulpes_Red Fox"]
    mapping = SpeciesMapping(geo, cls)
    scores = GeomodelScores([(g, float(rng.random()) ** 3) for g in geo[:-3]], mapping)
    mask = scores.dense_mask(cls)
    assert np.isnan(mask[40]) and np.isnan(mask[42]) and np.isnan(mask[1]) and mask[36] == 0.0
    for keep in (True, False):
        for rerank in (True, False):
            s = FilterSettings(0.01, keep, rerank)
            idx = rng.choice(len(cls), 5, replace=False)
            conf = np.sort(rng.random(5).astype(np.float32))[::-1]
            preds = [Prediction(cls[i], f32(c), int(i)) for i, c in zip(idx, conf)]
            want = [(p.index, p.confidence) for p in filter_predictions(preds, scores, s)]
            got = post.filter_row([(int(i), f32(c)) for i, c in zip(idx, conf)], mask, s)
            assert want == got


def test_top_k_threshold_and_second_threshold():
    c = np.array([0.05, 0.9, 0.9, 0.3, 0.1, 0.2, 0.8, 0.7, 0.6], np.float32)
    assert [i for i, _ in post.top_k_threshold(c, 5, 0.1)] == [1, 2, 6, 7, 8]
    assert post.top_k_threshold(c, 5, 0.95) == []
    assert [i for i, _ in post.top_k_threshold(c, 5, 0.75)] == [1, 2, 6]
    # rerank can push a survivor below min_conf: the second test (processor.rs:374) drops it
    logits = np.full((1, 6), -20.0, np.float32)
    logits[0, 2] = 0.0   # conf 0.5
    mask = np.full(6, 0.1, np.float32)
    rows = post.post_process(logits, 1, post.ACT_SIGMOID, 0.1, 5, mask, FilterSettings(0.01, True, True))
    assert rows == [[]]
    rows = post.post_process(logits, 1, post.ACT_SIGMOID, 0.1, 5, mask, FilterSettings(0.01, True, False))
    assert [i for i, _ in rows[0]] == [2]


def test_fixture_geomodel_kat():
    """tests/fixtures/make_fixture_geomodel.py:20-28 weights; tests/geomodel_range_filter.rs:218-254
    asserts species index 3 scores < 0.01 at Helsinki in June whatever week encoding is used."""
    W = np.array([[0.010, -0.020, 0.030, 0.001, 0.050], [0.005, 0.010, -0.015, 0.002, 0.020],
                  [0.100, 0.050, -0.200, 0.010, 0.150]], np.float32)
    B = np.array([0.5, -3.0, 0.2, -9.0, 1.0], np.float32)
    for week in (22.0, 23.0, 24.0):
        x = np.array([60.1699, 24.9384, week], np.float32)
        s = post.activate((x @ W + B)[None, :], post.ACT_SIGMOID)[0]
        assert s[3] < 0.01 and s[0] > 0.5


def test_sort_detections():
    d = post.extract_detections([[(1, f32(0.5)), (2, f32(0.9))], [(3, f32(0.2))]],
                                np.array([1.5, 0.0], np.float32), np.array([4.5, 3.0], np.float32),
                                ["a_b", "c_d", "e f_g h", "i"], 0.1)
    s = post.sort_detections(d)
    assert [(float(x.start_time), x.index) for x in s] == [(0.0, 3), (1.5, 2), (1.5, 1)]
    assert (s[1].scientific_name, s[1].common_name) == ("e f", "g h") and s[0].common_name == "i"
