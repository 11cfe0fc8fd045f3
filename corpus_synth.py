"""Seeded synthetic Zipfian corpora of the shapes BASELINE.json names (test / bench infrastructure).

Documents come from a planted LDA: a Zipf(s = 1.07) base distribution over a vocabulary of W words, tilted per
planted topic by log-normal noise (so topics differ but the marginal stays Zipfian), a few topics per document with
Dirichlet weights, log-normal document lengths.  Output is the CSC of the term x document matrix exactly as
`sparse.ToCSC()` would hand it to lda.go:464-466: per document the distinct word ids in ascending order with their
counts.  Generation is blockwise-deterministic: block m (documents [m*block, (m+1)*block)) depends only on
(seed, m), so a rank can generate just the minibatches it owns.  Runs on CPU (tests) or on a GPU (bench, via torch
-- plumbing only).
"""
import math

import numpy as np
import torch

SHAPES = {
    # name: (D, W, K, mean distinct words per document)
    "C1": (1_000, 5_000, 10, 100),
    "C2": (18_846, 50_000, 100, 133),
    "C3": (1_000_000, 100_000, 256, 200),
    "C4": (200_000, 100_000, 1024, 200),
}


class Planted:
    def __init__(self, W, seed=12345, n_topics=64, topics_per_doc=4, mean_distinct=200, device="cpu"):
        self.W, self.seed, self.T, self.tpd = W, seed, n_topics, topics_per_doc
        self.device = torch.device(device)
        g = torch.Generator(device="cpu").manual_seed(seed)
        base = 1.0 / (torch.arange(W, dtype=torch.float64) + 2.7) ** 1.07
        tilt = torch.exp(1.5 * torch.randn(n_topics, W, generator=g, dtype=torch.float64))
        phi = base[None, :] * tilt
        phi /= phi.sum(1, keepdim=True)
        cdf = torch.cumsum(phi, 1)
        cdf[:, -1] = 1.0
        # flattened per-topic CDFs: entry [t*W + w] = t + cdf_t(w), so one searchsorted serves all topics
        self.cdf_flat = (cdf + torch.arange(n_topics, dtype=torch.float64)[:, None]).reshape(-1).to(self.device)
        # tokens needed for ~mean_distinct distinct words: collisions on the Zipf head cost roughly a third
        self.mean_tokens = mean_distinct * 1.58
        self.sigma = 0.6

    def block(self, m, n_docs):
        """documents of block m -> (lengths int64[n_docs], ind int64[nnz], data float64[nnz]) as torch tensors"""
        dev = self.device
        g = torch.Generator(device=dev).manual_seed((self.seed * 1_000_003 + m) & 0x7FFFFFFFFFFFFFFF)
        mu = math.log(self.mean_tokens) - 0.5 * self.sigma ** 2
        n_tok = torch.exp(mu + self.sigma * torch.randn(n_docs, generator=g, device=dev)).clamp_(2, 20000).long()
        topics = torch.randint(0, self.T, (n_docs, self.tpd), generator=g, device=dev)
        wts = -torch.log(torch.rand(n_docs, self.tpd, generator=g, device=dev, dtype=torch.float64)) ** 2.0
        wcdf = torch.cumsum(wts / wts.sum(1, keepdim=True), 1)
        doc = torch.repeat_interleave(torch.arange(n_docs, device=dev), n_tok)
        total = doc.numel()
        u = torch.rand(total, generator=g, device=dev, dtype=torch.float64)
        slot = (u[:, None] > wcdf[doc]).sum(1).clamp_(max=self.tpd - 1)
        z = topics[doc, slot]
        u2 = torch.rand(total, generator=g, device=dev, dtype=torch.float64)
        word = torch.searchsorted(self.cdf_flat, z.double() + u2) - z * self.W
        word.clamp_(0, self.W - 1)
        key, cnt = torch.unique(doc * self.W + word, sorted=True, return_counts=True)
        kd = torch.div(key, self.W, rounding_mode="floor")
        lengths = torch.bincount(kd, minlength=n_docs)
        return lengths, key - kd * self.W, cnt.double()


def generate(D, W, seed=12345, block=8192, mean_distinct=200, device="cpu", owned=None, n_topics=64):
    """CSC (indptr, ind, data) numpy arrays of a D-document corpus.  owned(m) -> bool selects the blocks whose
    tokens are materialised (other documents get zero entries in this rank's view); None = all."""
    p = Planted(W, seed=seed, mean_distinct=mean_distinct, device=device, n_topics=n_topics)
    lens, inds, datas = [], [], []
    n_blocks = (D + block - 1) // block
    for m in range(n_blocks):
        n = min(block, D - m * block)
        if owned is not None and not owned(m):
            lens.append(np.zeros(n, dtype=np.int64))
            continue
        ln, ind, dat = p.block(m, n)
        lens.append(ln.cpu().numpy())
        inds.append(ind.cpu().numpy())
        datas.append(dat.cpu().numpy())
    indptr = np.concatenate(([0], np.cumsum(np.concatenate(lens)))).astype(np.int64)
    ind = np.concatenate(inds).astype(np.int64) if inds else np.zeros(0, dtype=np.int64)
    data = np.concatenate(datas).astype(np.float64) if datas else np.zeros(0, dtype=np.float64)
    return indptr, ind, data


def shape(name):
    return SHAPES[name]
