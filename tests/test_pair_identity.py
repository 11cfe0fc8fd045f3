"""The algebra behind the two-tokens-per-butterfly kernels (kernels.cuh, PAIR), checked on the CPU in float32.

Per token (lda.go:234-249 with c_k = (nPhi[w,k]+eta)/(nZ[k]+eta W), theta_k = nTheta[j,k]+alpha, q = 1-(1-rho)^v):
    S = sum_k c_k theta_k ;  theta'_k = theta_k (keep + mix c_k) + alpha q ,  keep = 1-q, mix = q wc / S.
For the next token (row c'), S' = sum_k c'_k theta'_k = keep Y + mix X + alpha q Z with Y = sum c' theta, X = sum c' c theta,
Z = sum c' -- none of which needs S -- so one reduction tree serves two tokens.  All terms are non-negative, so the regrouped
sum differs from the sequential one by fp32 rounding only."""
import numpy as np

f32 = np.float32


def test_paired_normaliser_equals_sequential_one():
    rng = np.random.default_rng(7)
    worst = 0.0
    for K in (100, 256, 1024):
        for _ in range(300):
            c = (rng.random(K) ** 4 * 1e-3).astype(f32)
            c2 = (rng.random(K) ** 4 * 1e-3).astype(f32)
            theta = (rng.random(K) * 50 + 0.1).astype(f32)
            q, wc, alpha = f32(rng.random() * 0.9), f32(rng.integers(1, 2000)), f32(0.1)
            S = f32((c * theta).sum(dtype=f32))
            keep, mix, aq = f32(1) - q, q * wc / S, alpha * q
            theta1 = theta * (mix * c + keep) + aq                     # sequential update (lda.go:244-249)
            u = c * theta
            Y, X, Z = f32((c2 * theta).sum(dtype=f32)), f32((c2 * u).sum(dtype=f32)), f32(c2.sum(dtype=f32))
            paired = mix * X + keep * Y + aq * Z
            exact = float((c2.astype(np.float64) * theta1.astype(np.float64)).sum())
            worst = max(worst, abs(float(paired) - exact) / exact)
    assert worst < 2e-6, worst


def test_paired_document_pass_matches_sequential_pass():
    """a whole burn-in pass over a 301-token document (odd: the last token goes alone), float32 both ways"""
    rng = np.random.default_rng(11)
    K, n = 256, 301
    rows = (rng.random((n, K)) ** 6 * 1e-2 + 1e-7).astype(f32)
    qs = (rng.random(n) * 0.2).astype(f32)
    wc, alpha = f32(n), f32(0.1)
    theta0 = (rng.random(K) * 3 + alpha).astype(f32)

    def step(theta, c, q):
        S = f32((c * theta).sum(dtype=f32))
        return theta * ((q * wc / S) * c + (f32(1) - q)) + alpha * q

    seq = theta0.copy()
    for t in range(n):
        seq = step(seq, rows[t], qs[t])
    par = theta0.copy()
    t = 0
    while t + 1 < n:
        c, c2, qa, qb = rows[t], rows[t + 1], qs[t], qs[t + 1]
        u = c * par
        Sa, Y, X, Z = (f32(x.sum(dtype=f32)) for x in (u, c2 * par, c2 * u, c2))
        keepa, mixa, aqa = f32(1) - qa, qa * wc / Sa, alpha * qa
        Sb = mixa * X + keepa * Y + aqa * Z
        par = par * (mixa * c + keepa) + aqa
        par = par * ((qb * wc / Sb) * c2 + (f32(1) - qb)) + alpha * qb
        t += 2
    if t < n:
        par = step(par, rows[t], qs[t])
    assert np.abs(par - seq).max() <= 1e-4 * np.abs(seq).max()
