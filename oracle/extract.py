"""AST extraction of individual functions/classes/constants from the read-only
reference scripts (TEST INFRASTRUCTURE; only usable where /root/reference exists).

The reference scripts cannot be imported: they load data from hard-coded paths at
module import (nsga_penalty.py:157,167), three of them run the whole search at
import (nsga_penalty.py:783, mobo_penalty.py:492) and they import TensorFlow.
We therefore ``ast.parse`` the file, keep only the requested top-level
``FunctionDef`` / ``ClassDef`` / simple ``Assign`` nodes and ``exec`` them in a
namespace pre-loaded with numpy / pandas / sklearn.  No reference source text is
copied into this repository; the code is executed from where it lies.
"""
from __future__ import annotations

import ast
import json
import os
import random
import time
from copy import deepcopy

REFERENCE_ROOT = os.environ.get("CMOOP_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "nsga_penalty.py"))


def _base_namespace() -> dict:
    import numpy as np
    import pandas as pd
    from scipy.spatial.distance import cdist
    from sklearn.compose import ColumnTransformer
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel
    from sklearn.metrics import confusion_matrix
    from sklearn.preprocessing import OneHotEncoder, StandardScaler

    return dict(
        np=np, pd=pd, random=random, deepcopy=deepcopy, time=time, cdist=cdist,
        ColumnTransformer=ColumnTransformer, OneHotEncoder=OneHotEncoder,
        StandardScaler=StandardScaler, confusion_matrix=confusion_matrix,
        GaussianProcessRegressor=GaussianProcessRegressor, Matern=Matern,
        WhiteKernel=WhiteKernel, ConstantKernel=ConstantKernel, C=ConstantKernel,
    )


def _simple_value(node: ast.AST) -> bool:
    """True for literal-ish right-hand sides (numbers, lists, dicts of names/literals)."""
    for sub in ast.walk(node):
        if isinstance(sub, (ast.Call, ast.Attribute, ast.Subscript, ast.Lambda, ast.Await)):
            return False
    return True


def extract(rel_path: str, names: list[str] | None = None, *, constants: bool = True,
            extra_ns: dict | None = None, quiet: bool = True) -> dict:
    """Return a namespace holding the requested top-level defs of ``rel_path``.

    names=None keeps every FunctionDef/ClassDef.  Simple ALL-CAPS constant
    assignments (EPSILON, HPARAM_SPACE, thresholds ...) are kept when
    ``constants`` is true so the functions see the globals they read.
    """
    path = os.path.join(REFERENCE_ROOT, rel_path)
    with open(path, "r", encoding="utf-8") as fh:
        tree = ast.parse(fh.read(), filename=path)
    keep = []
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
            if names is None or node.name in names:
                keep.append(node)
        elif constants and isinstance(node, ast.Assign) and len(node.targets) == 1 \
                and isinstance(node.targets[0], ast.Name) and node.targets[0].id.isupper() \
                and _simple_value(node.value):
            keep.append(node)
    mod = ast.Module(body=keep, type_ignores=[])
    ns = _base_namespace()
    if quiet:
        ns["print"] = lambda *a, **k: None
    if extra_ns:
        ns.update(extra_ns)
    exec(compile(mod, path, "exec"), ns)
    return ns


def extract_notebook(rel_path: str, names: list[str]) -> dict:
    """Pull named function defs out of a notebook's code cells (compare.ipynb)."""
    path = os.path.join(REFERENCE_ROOT, rel_path)
    with open(path, "r", encoding="utf-8") as fh:
        nb = json.load(fh)
    keep = []
    for cell in nb["cells"]:
        if cell.get("cell_type") != "code":
            continue
        src = "".join(cell["source"])
        try:
            tree = ast.parse(src)
        except SyntaxError:
            continue
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and node.name in names:
                keep.append(node)
    ns = _base_namespace()
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns
