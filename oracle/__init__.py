"""CPU oracle for the population-fitness hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a *checker*: a plain NumPy / C / torch-CPU
restatement of what the reference (sumansamui/CMOOP_Audio_Processing) computes
on this path, each function citing the reference ``file:line`` it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import, link or execute anything from here.  The
product package (``cmoop_audio_processing_b200``) never does; it fails loudly
when its CUDA library is missing instead of falling back to this code.

Pin status (see DESIGN.md "Oracle"):

* NDS / crowding / dominance / infill / local search / FPR / GD / IGD / Spread /
  C-metric / MOBO helpers  -- PINNED: golden vectors in ``tests/golden`` were made
  by running the reference's own functions (AST-extracted from the read-only
  sources by ``oracle/extract.py`` + ``oracle/make_golden.py``).
* GP posterior                -- PINNED against installed scikit-learn (the
  reference's third-party dependency; pinned there at 1.5.2, here 1.9.0).
* model size                  -- PINNED by the closed-form Keras parameter count
  (hand-checked 19 674 params for variant A / (16,3,F,1,1) / 10 classes).
* MFCC / log-mel              -- PARITY UNPINNED: the reference has no feature code
  (features are loaded pre-computed, nsga_penalty.py:64-71).  ``mfcc_ref.py``
  *defines* the spec (fp64) and is cross-checked against torch/torchaudio on CPU.
* CNN train/score             -- PARITY UNPINNED: TensorFlow/Keras are not
  installable here; ``cnn_ref.py`` restates build_model/evaluate_individual in
  torch-CPU fp32 with Keras-default numerics.
* hypervolume                 -- PARITY UNPINNED (pygmo absent); exact 3-D HV
  restatement validated on hand-computable cases and Monte-Carlo.
"""
