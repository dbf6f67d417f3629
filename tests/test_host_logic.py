"""Host-side logic of the drop-in (no kernels): operators keep the reference's random stream,
infill selection, MOBO encoding, synthetic data.  CPU only."""
import random

import numpy as np
import pytest

from conftest import records
from cmoop_audio_processing_b200 import nsga, surrogate, synth
from oracle import extract as ex


def test_select_infill_points_matches_reference_golden(golden):
    for case in golden("infill")["cases"]:
        idx, _ = surrogate.select_infill_points(records(case["objs"], case["cv"]), case["k"])
        assert idx == case["indices"]


def test_mobo_encoding_matches_reference_golden(golden):
    g = golden("mobo")
    for item in g["encoding"]:
        np.testing.assert_array_equal(surrogate.hparams_to_vector(item["hp"]), item["vec"])
        assert surrogate.vector_to_hparams(np.array(item["vec"])) == item["hp"]
    for case in g["cases"]:
        best = np.array(case["candidates"])[case["argmax"]]
        assert surrogate.vector_to_hparams(best) == case["decoded"]


def test_lambda_schedule(golden):
    g = golden("nsga")["lambda_schedule"]
    assert [nsga.get_lambda(i, g["max_gen"]) for i in range(g["max_gen"])] == g["values"]


@pytest.mark.skipif(not ex.reference_available(), reason="/root/reference not present")
def test_operators_consume_random_stream_like_the_reference():
    ref = ex.extract("ablation_study/sa_nsga_local.py",
                     ["initialize_population", "dominates", "tournament_selection", "crossover", "mutate",
                      "perturb_hparams"])
    for seed in range(5):
        random.seed(seed)
        pop_r = ref["initialize_population"](9)
        a = ref["crossover"](pop_r[0], pop_r[1])
        b = [ref["mutate"](p) for p in pop_r]
        c = [ref["perturb_hparams"](p) for p in pop_r]
        recs = [{"hparams": p, "objs": [random.random() for _ in range(3)], "CV": random.choice([0.0, 0.3])} for p in pop_r]
        d = [ref["tournament_selection"](recs, 3.0) for _ in range(20)]
        tail_r = random.random()
        random.seed(seed)
        pop_o = nsga.initialize_population(9)
        a2 = nsga.crossover(pop_o[0], pop_o[1])
        b2 = [nsga.mutate(p) for p in pop_o]
        c2 = [surrogate.perturb_hparams(p) for p in pop_o]
        recs2 = [{"hparams": p, "objs": [random.random() for _ in range(3)], "CV": random.choice([0.0, 0.3])} for p in pop_o]
        d2 = [nsga.tournament_selection(recs2, 3.0) for _ in range(20)]
        tail_o = random.random()
        assert (pop_r, a, b, c, d, tail_r) == (pop_o, a2, b2, c2, d2, tail_o)


def test_synthetic_clips_are_deterministic_and_bounded():
    w1, l1 = synth.make_clips(24, 12, seed=1234)
    w2, l2 = synth.make_clips(24, 12, seed=1234)
    assert w1.shape == (24, 16000) and w1.dtype == np.float32
    np.testing.assert_array_equal(w1, w2)
    np.testing.assert_array_equal(l1, l2)
    assert np.abs(w1).max() <= 1.0 and sorted(set(l1.tolist())) == list(range(12))
    u = synth.uniform_clips(4)
    assert u.shape == (4, 16000) and u.min() >= -1 and u.max() <= 1
