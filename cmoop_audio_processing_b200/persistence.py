"""Run logs in the on-disk formats of the reference, so that ``compare.ipynb``-style analysis
and the xlsx-seeded second stage of the ablation scripts work on the new outputs.

* per-generation table  nsga_penalty.py:700-722      columns Generation, Accuracy, Size_MB, FPR, CV + the six genes
* Pareto-set table      nsga_penalty.py:746-763,800-820   columns Accuracy, Size_MB, FPR + the six genes
* all generations       nsga_penalty.py:785-788      one sheet ``Gen_<i>`` per generation (xlsx); without openpyxl one
                                                      CSV with the same columns (the Generation column keeps the split)
* seeded initial population   ablation_study/psi_sa_nsga_local.py:255-269   rows of a Pareto table -> evaluated records,
                                                      CV recomputed from the thresholds

Host-side Python only (pandas); nothing here touches the GPU.
"""
from __future__ import annotations

import os

GENES = ["filters", "kernel_size", "use_bn", "residual_blocks", "fc_layers", "use_dropout"]


def generation_records(gen: int, pop_data):
    """Row dicts exactly as nsga_penalty.py:700-718 builds them."""
    return [{"Generation": gen, "Accuracy": -ind["objs"][0], "Size_MB": ind["objs"][1], "FPR": ind["objs"][2],
             "CV": ind["CV"], **ind["hparams"]} for ind in pop_data]


def generation_frame(gen: int, pop_data):
    import pandas as pd
    return pd.DataFrame(generation_records(gen, pop_data))


def pareto_records(pareto_set):
    """nsga_penalty.py:746-758 / 800-815."""
    return [{"Accuracy": -sol["objs"][0], "Size_MB": sol["objs"][1], "FPR": sol["objs"][2], **sol["hparams"]}
            for sol in pareto_set]


def save_pareto_csv(pareto_set, path: str) -> str:
    import pandas as pd
    pd.DataFrame(pareto_records(pareto_set)).to_csv(path, index=False)
    return path


def save_generations(gen_dfs, path: str) -> str:
    """``all_generations.xlsx`` with one sheet per generation when an Excel writer is installed, else a single CSV next to
    the requested path.  Returns the file written."""
    import pandas as pd
    try:
        import openpyxl  # noqa: F401
        with pd.ExcelWriter(path) as writer:
            for i, df in enumerate(gen_dfs):
                df.to_excel(writer, sheet_name=f"Gen_{i}", index=False)
        return path
    except ImportError:
        csv_path = os.path.splitext(path)[0] + ".csv"
        pd.concat(list(gen_dfs), ignore_index=True).to_csv(csv_path, index=False)
        return csv_path


def load_generations(path: str):
    """Inverse of save_generations: list of per-generation frames."""
    import pandas as pd
    if path.lower().endswith(".csv"):
        df = pd.read_csv(path)
        return [g.reset_index(drop=True) for _, g in df.groupby("Generation", sort=True)]
    sheets = pd.read_excel(path, sheet_name=None)
    return [sheets[k] for k in sorted(sheets, key=lambda s: int(s.split("_")[1]))]


def load_seed_population(path: str, min_accuracy: float, max_model_size: float, max_fpr: float):
    """Evaluated records from a Pareto table (csv or xlsx), as psi_sa_nsga_local.py:255-269 does for its second stage:
    genes cast to int / bool, objectives [-Accuracy, Size_MB, FPR], CV recomputed from the thresholds."""
    import pandas as pd
    df = pd.read_csv(path) if path.lower().endswith(".csv") else pd.read_excel(path)
    pop_data = []
    for _, r in df.iterrows():
        hp = {}
        for k in GENES:
            v = r[k]
            if k in ("use_bn", "use_dropout"):
                hp[k] = (v.strip().lower() == "true") if isinstance(v, str) else bool(v)
            else:
                hp[k] = int(v)
        objs = [-float(r["Accuracy"]), float(r["Size_MB"]), float(r["FPR"])]
        cv = max(0, min_accuracy - (-objs[0])) + max(0, objs[1] - max_model_size) + max(0, objs[2] - max_fpr)
        pop_data.append({"hparams": hp, "objs": objs, "CV": cv})
    return pop_data
