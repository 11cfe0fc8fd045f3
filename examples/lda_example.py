"""ExampleLatentDirichletAllocation (lda_test.go:237-284) on the B200 path, followed by the step after it: a nearest
neighbour search over the documents' topic vectors (index.go).  The vectoriser is out of scope here (it stays in Go);
a few lines of Python stand in for `nlp.NewCountVectoriser(stopWords...)` and produce the same term x document CSR the
Go pipeline would hand to LDA.

    python examples/lda_example.py          # needs a B200 (no CPU fallback)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import nlp_b200  # noqa: E402
from nlp_b200 import rand  # noqa: E402
from nlp_b200.lda import CSR  # noqa: E402
from nlp_b200.measures import pairwise  # noqa: E402

corpus = [
    "The quick brown fox jumped over the lazy dog",
    "The cow jumped over the moon",
    "The little dog laughed to see such fun",
]
stop_words = {"the", "over", "to", "such"}


def count_vectorise(docs):
    """term x document counts as CSR (rows = vocabulary in order of first appearance, like vectorisers.go:60-80)"""
    vocab, cells = {}, {}
    for j, doc in enumerate(docs):
        for word in doc.lower().split():
            if word in stop_words:
                continue
            i = vocab.setdefault(word, len(vocab))
            cells[(i, j)] = cells.get((i, j), 0) + 1
    indptr, ind, data = [0], [], []
    for i in range(len(vocab)):
        for (r, j), v in sorted(cells.items()):
            if r == i:
                ind.append(j)
                data.append(float(v))
        indptr.append(len(ind))
    m = CSR(len(vocab), len(docs), np.array(indptr, dtype=np.int64), np.array(ind, dtype=np.int64), np.array(data))
    return m, sorted(vocab, key=vocab.get)


m, vocab = count_vectorise(corpus)
lda = nlp_b200.NewLatentDirichletAllocation(2)
lda.Rnd = rand.New(rand.NewSource(0))
docs_over_topics = lda.FitTransform(m)                      # K x D, lda.go:541

for doc in range(docs_over_topics.shape[1]):
    print("Topic distribution for document: '%s' -" % corpus[doc],
          ", ".join("Topic #%d=%f" % (t, docs_over_topics[t, doc]) for t in range(docs_over_topics.shape[0])))
topics_over_words = lda.Components()                         # K x W, lda.go:415
for topic in range(topics_over_words.shape[0]):
    print("Word distribution for Topic #%d -" % topic,
          ", ".join("'%s'=%f" % (vocab[w], topics_over_words[topic, w]) for w in range(topics_over_words.shape[1])))

index = nlp_b200.NewLinearScanIndex(pairwise.CosineDistance)
index.IndexColumns(docs_over_topics, ids=corpus)             # ColDo(theta, func(j, v) { index.Index(v, corpus[j]) })
for match in index.Search(docs_over_topics[:, 0], 2):       # the documents nearest to the first one
    print("distance %.6f  %s" % (match.Distance, match.ID))
