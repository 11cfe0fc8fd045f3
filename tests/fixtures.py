"""Inline fixtures of the reference's own LDA tests (lda_test.go), shared by the oracle
golden tests (CPU) and the CUDA parity tests (GPU)."""
import numpy as np

# lda_test.go:26-36 / :101-111 / :188-198  (9x9 dense term x document, K=3)
FIXTURE_A = np.array([
    3, 3, 3, 0, 0, 0, 0, 0, 0,
    3, 3, 3, 0, 0, 0, 0, 0, 0,
    3, 3, 3, 0, 0, 0, 0, 0, 0,
    0, 0, 0, 3, 3, 3, 0, 0, 0,
    0, 0, 0, 3, 3, 3, 0, 0, 0,
    0, 0, 0, 3, 3, 3, 0, 0, 0,
    0, 0, 0, 0, 0, 0, 4, 4, 4,
    0, 0, 0, 0, 0, 0, 4, 4, 4,
    0, 0, 0, 0, 0, 0, 4, 4, 4,
], dtype=np.float64).reshape(9, 9)

# lda_test.go:46-56 / :127-137 / :203-213
FIXTURE_B = np.array([
    3, 3, 3, 0, 0, 0, 0, 0, 0,
    3, 3, 3, 0, 0, 0, 0, 0, 0,
    3, 3, 3, 0, 0, 0, 0, 0, 0,
    0, 0, 0, 3, 5, 1, 0, 0, 0,
    0, 0, 0, 3, 5, 0, 0, 0, 0,
    0, 0, 0, 3, 5, 0, 0, 0, 0,
    0, 0, 0, 0, 0, 0, 4, 4, 4,
    0, 0, 0, 0, 0, 0, 4, 4, 4,
    0, 0, 0, 0, 0, 0, 4, 4, 4,
], dtype=np.float64).reshape(9, 9)

# lda_test.go:37-41
EXPECTED_TOPICS_A = np.array([
    [0.33, 0.33, 0.33, 0, 0, 0, 0, 0, 0],
    [0, 0, 0, 0, 0, 0, 0.33, 0.33, 0.33],
    [0, 0, 0, 0.33, 0.33, 0.33, 0, 0, 0],
])
# lda_test.go:57-61
EXPECTED_TOPICS_B = np.array([
    [0.33, 0.33, 0.33, 0, 0, 0, 0, 0, 0],
    [0, 0, 0, 0, 0, 0, 0.33, 0.33, 0.33],
    [0, 0, 0, 0.428, 0.285, 0.285, 0, 0, 0],
])
# lda_test.go:112-122 and :138-148 (docs x topics)
EXPECTED_DOCS = np.array([
    [1, 0, 0], [1, 0, 0], [1, 0, 0],
    [0, 0, 1], [0, 0, 1], [0, 0, 1],
    [0, 1, 0], [0, 1, 0], [0, 1, 0],
], dtype=np.float64)

FIT_CASES = [(FIXTURE_A, EXPECTED_TOPICS_A), (FIXTURE_B, EXPECTED_TOPICS_B)]
