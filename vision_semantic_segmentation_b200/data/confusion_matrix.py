"""Confusion matrix of the segmentation network, as loaded by the mapper.

Mirrors ``src/data/confusion_matrix.py:6-63``: ``C[i, j]`` counts observations of true class i
predicted as j; ``get_submatrix`` selects the mapped classes, row-normalises to probabilities and
optionally takes the log.  Host-side, C x C float64.
"""
import numpy as np

__all__ = ["ConfusionMatrix"]


class ConfusionMatrix(object):
    def __init__(self, load_path):
        self._cfn_mtx = np.load(load_path)
        rows, cols = self._cfn_mtx.shape
        assert rows == cols
        self.num_class = rows

    def get_submatrix(self, indices, to_probability=False, use_log=False):
        if len(indices) == 0:
            return []
        if len(indices) > self.num_class:
            raise ValueError("The number of indices is greater than the number of classes in the confusion matrix!")
        for i in indices:
            if i < 0 or i >= self.num_class:
                raise ValueError("Invalid index!", i)
        sub = self._cfn_mtx[np.ix_(indices, indices)]
        if to_probability:
            sub = sub / np.sum(sub, axis=1)[:, np.newaxis]
            if use_log:
                sub = np.log(sub)
        return sub

    def __str__(self):
        return str(self._cfn_mtx)

    def __len__(self):
        return self.num_class

    def __getitem__(self, item):
        return self._cfn_mtx[item]
