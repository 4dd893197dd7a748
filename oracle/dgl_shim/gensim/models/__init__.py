import numpy as np


class KeyedVectors:
    """The two uses the reference makes of gensim (TEST INFRASTRUCTURE ONLY): `KeyedVectors.load_word2vec_format(path)` with
    `.vectors` / `kv[key]` (data_loader/dataset.py:127-128,154-155) and the `KeyedVectors(vector_size)` + `.add` of the masked
    dataset (:228-229).  Text word2vec format: a header line "<count> <dim>", then "<key> <v1> ... <vdim>" per line, float32."""

    def __init__(self, vector_size=0):
        self.vector_size = vector_size
        self.index = {}
        self.vectors = np.zeros((0, vector_size), dtype=np.float32)

    @classmethod
    def load_word2vec_format(cls, path):
        with open(path) as f:
            count, dim = (int(v) for v in f.readline().split())
            kv = cls(dim)
            kv.vectors = np.zeros((count, dim), dtype=np.float32)
            for i in range(count):
                parts = f.readline().rstrip().split(" ")
                kv.index[parts[0]] = i
                kv.vectors[i] = np.asarray(parts[1:1 + dim], dtype=np.float32)
        return kv

    def __getitem__(self, key):
        return self.vectors[self.index[key]]

    def add(self, keys, vectors):
        self.keys, self.added = list(keys), vectors

    def distances(self, *a, **k):
        raise NotImplementedError("gensim shim: test_topk needs real gensim")
