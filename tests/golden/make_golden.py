"""Generates tests/golden/lda_golden.json from the CPU oracle (oracle/lda_oracle.c), which is itself pinned to the
reference's own tests (tests/test_oracle_golden.py).  The reference is Go and cannot run in this image, so these are
restatement-derived vectors: regression pins for the oracle (CPU suite) and fixed targets for the CUDA path (GPU suite).
    python tests/golden/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import corpus_synth  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.fixtures import FIXTURE_A, FIXTURE_B  # noqa: E402

out = {"about": "oracle-derived (float64 C restatement of lda.go); seed 0 unless stated", "cases": []}


def run(name, m, k, **f):
    o = O.LatentDirichletAllocation(k, seed=0)
    for key, v in f.items():
        setattr(o, key, v)
    theta = o.FitTransform(m)
    nphi, nz = o.get_state()
    out["cases"].append({
        "name": name, "K": k, "fields": f, "iterations_run": int(o.iterationsRun),
        "perplexity_trace": o.perplexity_trace.tolist(), "nz": nz.tolist(),
        "theta_first_docs": theta[:, :4].tolist(), "theta_checksum": float(np.abs(theta).sum()),
        "components_diag": np.diag(o.Components()[:, :k]).tolist(),
        "counters": [o.rhoPhiT, o.rhoThetaT, o.wordsInCorpus], "rnd_state": list(o.rnd_state),
    })


run("fixture_A_defaults", FIXTURE_A, 3)
run("fixture_B_defaults", FIXTURE_B, 3)
run("fixture_A_trace5", FIXTURE_A, 3, PerplexityEvaluationFrequency=1, PerplexityTolerance=0.0, Iterations=5)
indptr, ind, data = corpus_synth.generate(200, 600, block=50, mean_distinct=30, n_topics=16)
out["synthetic"] = {"D": 200, "W": 600, "block": 50, "mean_distinct": 30, "n_topics": 16, "nnz": int(indptr[-1]),
                    "ind_checksum": int(ind.sum()), "data_checksum": float(data.sum())}
run("synthetic_200x600_K12_b50", (600, 200, indptr, ind, data), 12, PerplexityEvaluationFrequency=1,
    PerplexityTolerance=0.0, Iterations=8, BatchSize=50)
run("synthetic_200x600_K12_sync3", (600, 200, indptr, ind, data), 12, PerplexityEvaluationFrequency=1,
    PerplexityTolerance=0.0, Iterations=4, BatchSize=50, SyncGroup=3)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lda_golden.json"), "w"), indent=1)
print("wrote", len(out["cases"]), "cases")
