"""Word-embedding pair for the skip-gram demo (reference: extenncor/embed.py:1-59): a words -> vector dense layer and its
vector -> words counterpart over two caller-visible weight variables, plus vocabulary look-ups."""
import numpy as np

import tenncor_b200 as tc


class Embedding(object):
    def __init__(self, vweight, wweight, words):
        nwords, vecsize = tuple(vweight.shape())
        # both layers take the given variables as their kernels (the initialiser ignores the requested shape)
        self.embedding = tc.api.layer.dense([nwords], [vecsize], kernel_init=lambda shape, label: vweight, bias_init=None)
        self.exbedding = tc.api.layer.dense([vecsize], [nwords], kernel_init=lambda shape, label: wweight, bias_init=None)
        self.weight = vweight
        self.idx2word = words
        self.word2idx = {word: i for i, word in enumerate(words)}

    def __getitem__(self, idx):
        w = self.weight.get()
        return w[idx] if idx < len(w) else None

    def __len__(self):
        return self.weight.shape()[0]

    def get_vec(self, word):
        idx = self.word2idx.get(word)
        return None if idx is None else self[idx]

    def onehot(self, word):
        idx = self.word2idx.get(word)
        if idx is None:
            return None
        vec = [0] * len(self.idx2word)
        vec[idx] = 1
        return vec


def embedding_kernel_init(shape, label):
    return tc.variable(np.random.uniform(-1, 1, tuple(shape)), label)


def make_embedding(words, vecsize):
    """words: list of strings, the index is the label; vecsize: length of the mapped vector"""
    nwords = len(words)
    return Embedding(embedding_kernel_init([nwords, vecsize], "to_vec"), embedding_kernel_init([vecsize, nwords], "to_word"), words)


def vdistance(v1, v2):
    """cosine of the angle between two vectors"""
    return np.dot(v1, v2) / (np.linalg.norm(v1) * np.linalg.norm(v2))
