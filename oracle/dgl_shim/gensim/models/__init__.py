class KeyedVectors:
    def __init__(self, vector_size=0):
        self.vector_size = vector_size

    def add(self, keys, vectors):
        self.keys, self.vectors = list(keys), vectors

    def distances(self, *a, **k):
        raise NotImplementedError("gensim shim: test_topk needs real gensim")
