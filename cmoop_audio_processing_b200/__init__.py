"""cmoop_audio_processing_b200 -- B200-native population-fitness hot path.

Drop-in replacements (same names, records and error behaviour) for the functions the
reference drivers nsga2() / run_mobo() call beneath their Python seam, backed by
hand-written sm_100a CUDA behind a C ABI (include/cmoop_b200.h):

    nsga       fast_non_dominated_sort, crowding_distance, environmental_selection (+ host operators)
    surrogate  SurrogateManager, select_infill_points, perform_local_search, train_gps, predict_gps, ...
    quality    hypervolume, generational_distance, inverted_gd, spread_metric, coverage_metric
    features   MfccFrontEnd (log-mel / MFCC)
    problem    evaluate_individual, compute_objectives_and_constraints (candidate-CNN train + score)

Importing the package never touches the GPU; the shared library is loaded on first use and
there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .nsga import (EPSILON, crowding_distance, dominates, environmental_selection,  # noqa: F401
                   fast_non_dominated_sort, get_lambda, tournament_selection, crossover, mutate,
                   initialize_population, HPARAM_SPACE)

__all__ = ["fast_non_dominated_sort", "crowding_distance", "environmental_selection", "dominates", "get_lambda",
           "tournament_selection", "crossover", "mutate", "initialize_population", "HPARAM_SPACE", "EPSILON"]
__version__ = "0.1.0"
