"""Boundary pieces that need no GPU: label-space projection (bb_mask_build), the watchdog, the timeout rule and the
exception firewall.  The mask tests are the reference's own unit tests (src/inference/geomodel.rs:182-410) restated
against the C function — the product entry point, not the oracle — plus a cross-check with the oracle's projection."""
import math
import threading
import time

import numpy as np
import pytest

import birda_b200 as b
from birda_b200 import _lib
from birda_b200.api import mask_build, rules


def score_of(mask, labels, label):
    """GeomodelScores::score_of through the dense mask: None where the mask holds NaN or the label is not a row."""
    if label not in labels:
        return None
    v = float(mask[labels.index(label)])
    return None if math.isnan(v) else v


# ---- scientific_name (geomodel.rs:198-236) -------------------------------------------------------------------------
@pytest.mark.parametrize("label,want", [
    ("Parus major_Great Tit", "Parus major"),
    ("Parus major_Talitiainen", "Parus major"),
    ("Parus major", "Parus major"),
    ("Accelerating_and_revving_and_vroom", "Accelerating_and_revving_and_vroom"),
    ("Accordion", "Accordion"),
    ("Dog_Dog", "Dog_Dog"),
    ("Parus major_Great_Tit", "Parus major"),
    ("", ""),
])
def test_scientific_name(label, want):
    assert rules.scientific_name(label) == want


# ---- SpeciesMapping::build (geomodel.rs:238-330) -------------------------------------------------------------------
def test_mapping_matches_localized_classifier_labels():
    cl = ["Parus major_Talitiainen"]
    mask, mapped, unmatched = mask_build(cl, ["Parus major_Great Tit"], [("Parus major_Great Tit", 0.5)])
    assert (mapped, unmatched) == (1, 0)
    assert score_of(mask, cl, "Parus major_Talitiainen") == 0.5          # classifier_label_for -> the Finnish label


def test_mapping_matches_bare_binomial_perch_labels():
    _, mapped, _ = mask_build(["Parus major"], ["Parus major_Great Tit"], [])
    assert mapped == 1


def test_mapping_is_case_insensitive():
    _, mapped, _ = mask_build(["Parus Major_Talitiainen"], ["parus major_Great Tit"], [])
    assert mapped == 1


def test_mapping_counts_unmatched_classifier_species():
    cl = ["Parus major_Great Tit", "Accipiter gentilis_Northern Goshawk", "Dog_Dog"]
    mask, mapped, unmatched = mask_build(cl, ["Parus major_Great Tit"], [])
    assert (mapped, unmatched, len(mask)) == (1, 2, 3)
    assert mask[0] == 0.0 and np.isnan(mask[1]) and np.isnan(mask[2])


def test_mapping_ignores_geomodel_species_absent_from_the_classifier():
    _, mapped, unmatched = mask_build(["Parus major_Great Tit"],
                                      ["Parus major_Great Tit", "Petaurista albiventer_White-bellied Giant Flying Squirrel"], [])
    assert (mapped, unmatched) == (1, 0)


def test_mapping_keeps_the_first_of_two_colliding_classifier_labels():
    cl = ["Parus major_First", "Parus major_Second"]
    mask, mapped, _ = mask_build(cl, ["Parus major_Great Tit"], [("Parus major_Great Tit", 0.7)])
    assert mapped == 1
    assert score_of(mask, cl, "Parus major_First") == pytest.approx(0.7)
    assert score_of(mask, cl, "Parus major_Second") is None              # keyed by label STRING: the namesake has no entry


def test_mapping_of_empty_label_sets_is_empty():
    mask, mapped, unmatched = mask_build([], [], [])
    assert (len(mask), mapped, unmatched) == (0, 0, 0)


# ---- GeomodelScores::project (geomodel.rs:332-410) -----------------------------------------------------------------
def test_projection_keys_by_classifier_label():
    cl = ["Parus major_Talitiainen"]
    mask, _, _ = mask_build(cl, ["Parus major_Great Tit"], [("Parus major_Great Tit", 0.8)])
    assert score_of(mask, cl, "Parus major_Talitiainen") == np.float32(0.8)
    assert score_of(mask, cl, "Parus major_Great Tit") is None


def test_projection_includes_mapped_species_the_geomodel_omitted():
    cl = ["Parus major_Great Tit"]
    mask, _, _ = mask_build(cl, cl, [])
    assert score_of(mask, cl, cl[0]) == 0.0


def test_projection_omits_unmatched_species():
    mask, mapped, _ = mask_build(["Dog_Dog"], ["Parus major_Great Tit"], [])
    assert np.isnan(mask[0]) and mapped == 0                              # is_empty()


def test_projection_drops_geomodel_species_with_no_classifier_match():
    cl = ["Parus major_Great Tit"]
    mask, _, _ = mask_build(cl, ["Parus major_Great Tit", "Vulpes vulpes_Red Fox"],
                            [("Parus major_Great Tit", 0.8), ("Vulpes vulpes_Red Fox", 0.9)])
    assert score_of(mask, cl, "Parus major_Great Tit") == np.float32(0.8)
    assert score_of(mask, cl, "Vulpes vulpes_Red Fox") is None


def test_in_range_count_applies_the_threshold():
    g = ["Aaa aaa_X", "Bbb bbb_Y", "Ccc ccc_Z"]
    mask, _, _ = mask_build(g, g, [("Aaa aaa_X", 0.9), ("Bbb bbb_Y", 0.005), ("Ccc ccc_Z", 0.02)])
    count = lambda thr: int(np.sum(mask >= np.float32(thr)))
    assert (count(0.01), count(0.5), count(0.99)) == (2, 1, 0)


def test_later_scores_overwrite_and_identical_label_strings_share_the_entry():
    cl = ["Parus major_Great Tit", "Parus major_Great Tit", "Corvus corax"]
    mask, mapped, unmatched = mask_build(cl, ["Parus major_X", "Corvus corax_Raven"],
                                         [("Parus major_X", 0.1), ("PARUS MAJOR_Y", 0.6)])
    assert mask[0] == np.float32(0.6) and mask[1] == np.float32(0.6)     # same label string -> same map entry
    assert mask[2] == 0.0 and (mapped, unmatched) == (2, 1)


def test_non_ascii_case_folding():
    cl = ["Émberiza Ćitrinella_Keltasirkku", "ΑΒΓ δεζ_greek"]
    mask, mapped, _ = mask_build(cl, ["émberiza ćitrinella_Yellowhammer", "αβγ δεζ_g"], [("ÉMBERIZA ĆITRINELLA_x", 0.25)])
    assert mapped == 2 and mask[0] == 0.25 and mask[1] == 0.0


def test_mask_build_matches_the_oracle_projection_at_model_size():
    """6522 classifier labels (305 without a geomodel entry), 12 012 geomodel species: the C function and the oracle's
    label-keyed projection (oracle/post.py) agree on every row."""
    from oracle import post as opost
    rng = np.random.default_rng(5)
    geo = [f"Genus{i} species{i}_English {i}" for i in range(12_012)]
    pick = rng.permutation(12_012)[:6217]
    cl = [f"Genus{i} species{i}_Nimi {i}" for i in pick] + [f"Noise_{i}" for i in range(305)]
    order = rng.permutation(len(cl)); cl = [cl[i] for i in order]
    reported = rng.permutation(12_012)[:9000]
    scores = [(geo[i], float(np.float32(rng.random()))) for i in reported]
    mask, mapped, unmatched = mask_build(cl, geo, scores)
    assert (mapped, unmatched) == (6217, 305)
    ref = opost.GeomodelScores(scores, opost.SpeciesMapping(geo, cl)).dense_mask(cl)
    assert np.array_equal(np.isnan(mask), np.isnan(ref))
    assert np.array_equal(mask[~np.isnan(mask)], ref[~np.isnan(ref)])


def test_mask_build_rejects_null_labels():
    import ctypes as C
    arr = (C.c_char_p * 1)(None)
    out = np.zeros(1, np.float32)
    rc = _lib.lib.bb_mask_build(arr, 1, arr, 0, arr, out.ctypes.data_as(_lib.f32p), 0, out.ctypes.data_as(_lib.f32p), None, None)
    assert rc == -1


# ---- watchdog (src/gpu/watchdog.rs:22-66) and its timeout rule (processor.rs:194-211) ------------------------------
@pytest.mark.parametrize("value,want", [
    (None, 10), ("", 10), ("30", 30), ("1", 1), ("3600", 3600), ("0", 10), ("3601", 10), ("abc", 10), ("-5", 10),
    (" 30", 10), ("30 ", 10), ("+45", 45), ("1.5", 10), ("99999999999999999999999", 10), ("+", 10),
])
def test_inference_timeout_rule(value, want):
    assert rules.inference_timeout_secs(value) == want


def test_watchdog_cancelled_when_dropped():
    """watchdog.rs:73-84: cancel right away, sleep past the deadline, nothing fires."""
    fired = []
    w = b.Watchdog(150, 32, on_fire=lambda secs, batch: fired.append((secs, batch)))
    w.cancel()
    time.sleep(0.4)
    assert fired == []


def test_watchdog_fires_with_timeout_and_batch():
    ev = threading.Event(); got = []
    w = b.Watchdog(1100, 64, on_fire=lambda secs, batch: (got.append((secs, batch)), ev.set()))
    assert ev.wait(5.0)
    assert got == [(1, 64)]                                               # Duration::as_secs, batch size as given
    w.cancel()


def test_watchdog_default_action_terminates_the_process():
    """on_fire == NULL is the reference: its FATAL block on stderr, exit status 1 (watchdog.rs:31-49)."""
    import subprocess
    import sys
    code = ("import time, birda_b200 as b\n"
            "w = b.Watchdog(100, 32)\n"
            "time.sleep(5)\n"
            "print('survived')\n")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "survived" not in r.stdout
    assert "FATAL: Inference timeout after 0s (batch size: 32)" in r.stderr
    assert "birda -b 16 <input>" in r.stderr and "Terminating process to prevent system lockup." in r.stderr


# ---- exception firewall --------------------------------------------------------------------------------------------
def test_injected_allocation_failure_stops_at_the_boundary(tmp_path):
    """bb_debug_inject_alloc_failure(1): the next guarded entry point throws std::bad_alloc inside the library; the
    caller sees BB_ERR_OOM and a message, not an abort.  The call after it works again."""
    import ctypes as C
    lib = _lib.lib
    info = _lib.WavInfo() if hasattr(_lib, "WavInfo") else None
    path = str(tmp_path / "missing.wav").encode()
    lib.bb_debug_inject_alloc_failure(1)
    n = C.c_int32(-1)
    rc = lib.bb_device_count(C.byref(n))
    assert rc == -6, rc                                                   # BB_ERR_OOM
    assert b"out of host memory" in lib.bb_last_error(None)
    rc = lib.bb_device_count(C.byref(n))
    assert rc in (0, -8)
    # second entry from now
    lib.bb_debug_inject_alloc_failure(2)
    lib.bb_device_count(C.byref(n))
    out = np.zeros(1, np.float32)
    arr = (C.c_char_p * 1)(b"Parus major")
    rc = lib.bb_mask_build(arr, 1, arr, 1, arr, out.ctypes.data_as(_lib.f32p), 0, out.ctypes.data_as(_lib.f32p), None, None)
    assert rc == -6
    rc = lib.bb_mask_build(arr, 1, arr, 1, arr, out.ctypes.data_as(_lib.f32p), 0, out.ctypes.data_as(_lib.f32p), None, None)
    assert rc == 0 and out[0] == 0.0
    lib.bb_debug_inject_alloc_failure(0)
    del info, path


def test_every_extern_c_body_is_guarded():
    """Source check: each extern "C" function in the host-side files that returns a status opens with BB_TRY."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    unguarded = []
    # one-line delegations to a guarded implementation, and a getter that only copies three integers
    nothrow = {"bb_ctx_create", "bb_ctx_create_on_stream", "bb_melspec_info", "bb_wav_read", "bb_standin_classify", "bb_standin_weights"}
    for fn in ("capi.cu", "pipeline.cpp", "pool.cpp", "wav.cpp", "mask.cpp", "watchdog.cpp", "k5_melspec.cu", "k6_flac.cu", "standin.cu"):
        p = os.path.join(root, "birda_b200", "csrc", fn)
        if not os.path.exists(p):
            continue
        src = open(p).read()
        for m in re.finditer(r"^int32_t\s+(bb_[a-z0-9_]+)\s*\([^)]*\)\s*\{\s*\n?\s*(\S+)", src, flags=re.M):
            if m.group(2) != "BB_TRY" and m.group(1) not in nothrow:
                unguarded.append(f"{fn}:{m.group(1)}")
    assert not unguarded, unguarded
