"""`pairwise.Comparer` functions of the reference (measures/pairwise/comparisons.go) as they are used with the B200
index: each name is a token the index kernels are specialised on (nlp_b200.index.LinearScanIndex(pairwise.CosineDistance)),
and calling one on two vectors evaluates it on the GPU through a one-vector index -- there is no host arithmetic here."""
import numpy as np


class Comparer:
    def __init__(self, name, code, line):
        self.name, self.code, self.__doc__ = name, code, "measures/pairwise/comparisons.go:%s" % line

    def __call__(self, a, b):
        from ..index import LinearScanIndex
        a = np.asarray(a, dtype=np.float64).ravel()
        b = np.asarray(b, dtype=np.float64).ravel()
        if a.size != b.size:
            raise ValueError("nlp: vectors differ in length")   # the reference panics on a dimension mismatch
        ix = LinearScanIndex(self)
        try:
            ix.Index(b, 0)
            return ix.Search(a, 1)[0].Distance
        finally:
            ix.Close()

    def __repr__(self):
        return "pairwise." + self.name


CosineSimilarity = Comparer("CosineSimilarity", 0, "17-29")
CosineDistance = Comparer("CosineDistance", 1, "40-42")
AngularDistance = Comparer("AngularDistance", 2, "48-55")
AngularSimilarity = Comparer("AngularSimilarity", 3, "59-61")
HammingDistance = Comparer("HammingDistance", 4, "69-84")
HammingSimilarity = Comparer("HammingSimilarity", 5, "89-91")
EuclideanDistance = Comparer("EuclideanDistance", 6, "96-100")
ManhattenDistance = Comparer("ManhattenDistance", 7, "104-108")
VectorLenSimilarity = Comparer("VectorLenSimilarity", 8, "111-117")

ALL = [CosineSimilarity, CosineDistance, AngularDistance, AngularSimilarity, HammingDistance, HammingSimilarity,
       EuclideanDistance, ManhattenDistance, VectorLenSimilarity]
