"""BASELINE configs[0] through the drop-in on the GPU and through the CPU oracle: nsga_penalty.py constrained NSGA-II
(CNN variant A, thresholds 0.9 / 2.5 / 0.1, accuracy = history['val_accuracy'][-1], the argmax(y_val) quirk of :387),
population 8 x 2 generations, random.seed(0), synthetic GSC-shaped clips through the MFCC front-end.

The search loop runs ONCE on the GPU (drivers.nsga2 = nsga_penalty.py:610-776 with the CUDA evaluator and the CUDA
NDS / crowding); every true evaluation it made (genotype, harness seed) is then repeated by the torch-CPU oracle on the
same initial parameters, shuffles and dropout masks, and compared objective by objective; the populations' ranks are
compared bit-exactly with the oracle's NDS on the same records.  (Feeding two arithmetics into two separate searches would
compare chaotic trajectories, not implementations.)  Test size: 12 x 16 training / 12 x 8 validation clips, epoch cap 2 --
the CPU oracle needs about a minute for the 24 variant-A trainings; tools/run_config0.py runs the same comparison at
larger sizes (logs under profiles/).

Tolerance: a few dozen Adam steps into training the accuracy is in its take-off phase, where the fp32 CUDA path and the
fp64 oracle -- identical inputs, different summation orders -- can already sit several points apart for deep BN stacks
(the oracle's OWN fp32 run does the same against its fp64 run, tests/test_gpu_cnn.py), so the per-evaluation statement
is: exact size and epoch count, median |d acc| <= 0.02, >= 75 % of the evaluations within 0.05 accuracy, >= 90 % within
0.01 FPR (measured at 12 x 32 clips / cap 3, profiles/r02_config0_small.json: median 0.008, 79 % within 0.05, 100 % within
0.01 FPR, 12 of 24 evaluations identical to the last sample, ranks and crowding distances bit-exact).
"""
import os
import random
import sys

import numpy as np
import pytest

from oracle import cnn_ref, nsga_ref

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu


def test_config0_search_on_gpu_matches_the_oracle_evaluation_by_evaluation():
    import run_config0 as rc
    report = rc.run(per_class_train=16, per_class_val=8, epoch_cap=2, pop=8, gens=2, oracle_dtype="float64")
    assert report["evaluations"] == 8 + 2 * 8
    assert report["size_exact"] and report["epochs_equal"]
    assert report["median_abs_d_acc"] <= 0.02 and report["frac_d_acc_within_0.05"] >= 0.75
    assert report["frac_d_fpr_within_0.01"] >= 0.9
    assert report["ranks_bit_exact"] and report["crowding_bit_exact"]
    assert report["robust_rank_disagreements"] == 0
