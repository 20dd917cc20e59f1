"""``StatServer`` as the container the scoring functions take (sidekit/statserver.py:202-304) plus the six
helpers on the hot path (:656-684 align_*, :797-817 norm / rotate / center, :1357-1374 mean per model).
GMM / i-vector statistics, EM and HDF5 IO are out of scope.
"""
import copy

import numpy

STAT_TYPE = numpy.float64


def _first_index(have, wanted):
    """Index of the FIRST occurrence of every wanted id (``numpy.argwhere(have == v)[0][0]``), O(N) hash join."""
    pos = {}
    for i, v in enumerate(have.tolist()):
        if v not in pos:
            pos[v] = i
    try:
        return numpy.fromiter((pos[v] for v in wanted.tolist()), dtype=numpy.int64, count=wanted.shape[0])
    except KeyError as e:       # the reference raises IndexError from argwhere(...)[0]
        raise IndexError("id %s not found" % (e,))


class StatServer:
    def __init__(self, statserver_file_name=None):
        if statserver_file_name is not None and not isinstance(statserver_file_name, StatServer):
            raise NotImplementedError("StatServer file IO / IdMap construction is out of scope")
        self.modelset = numpy.empty(0, dtype="|O")
        self.segset = numpy.empty(0, dtype="|O")
        self.start = numpy.empty(0, dtype="|O")
        self.stop = numpy.empty(0, dtype="|O")
        self.stat0 = numpy.array([], dtype=STAT_TYPE)
        self.stat1 = numpy.array([], dtype=STAT_TYPE)

    @classmethod
    def from_embeddings(cls, ids, embeddings):
        """What ``extract_embeddings`` builds (xvector.py:1905-1914): stat1 = embeddings, stat0 = ones."""
        s = cls()
        s.modelset = numpy.asarray(ids)
        s.segset = numpy.asarray(ids)
        s.start = numpy.empty(len(ids), dtype="|O")
        s.stop = numpy.empty(len(ids), dtype="|O")
        s.stat0 = numpy.ones((len(ids), 1), dtype=STAT_TYPE)
        s.stat1 = numpy.asarray(embeddings, dtype=STAT_TYPE)
        return s

    def validate(self, warn=False):
        ok = self.modelset.ndim == 1 and self.modelset.shape == self.segset.shape == self.start.shape == self.stop.shape
        ok &= self.stat0.shape[0] == self.stat1.shape[0] == self.modelset.shape[0]
        return bool(ok)

    def _take(self, indx):
        self.segset = self.segset[indx]
        self.modelset = self.modelset[indx]
        self.start = self.start[indx]
        self.stop = self.stop[indx]
        self.stat0 = self.stat0[indx, :]
        self.stat1 = self.stat1[indx, :]

    def align_segments(self, segment_list):
        self._take(_first_index(self.segset, numpy.asarray(segment_list)))

    def align_models(self, model_list):
        self._take(_first_index(self.modelset, numpy.asarray(model_list)))

    def norm_stat1(self):
        vect_norm = numpy.clip(numpy.linalg.norm(self.stat1, axis=1), 1e-08, numpy.inf)
        self.stat1 = (self.stat1.transpose() / vect_norm).transpose()

    def rotate_stat1(self, R):
        self.stat1 = numpy.dot(self.stat1, R)

    def center_stat1(self, mu):
        dim = self.stat1.shape[1] / self.stat0.shape[1]
        index_map = numpy.repeat(numpy.arange(self.stat0.shape[1]), dim)
        self.stat1 = self.stat1 - (self.stat0[:, index_map] * mu.astype(STAT_TYPE))

    def mean_stat_per_model(self):
        out = StatServer()
        out.modelset, inv = numpy.unique(self.modelset, return_inverse=True)
        out.segset = copy.deepcopy(out.modelset)
        n = out.modelset.shape[0]
        counts = numpy.bincount(inv, minlength=n).astype(STAT_TYPE)[:, None]
        out.stat0 = numpy.zeros((n, self.stat0.shape[1]), dtype=STAT_TYPE)
        out.stat1 = numpy.zeros((n, self.stat1.shape[1]), dtype=STAT_TYPE)
        numpy.add.at(out.stat0, inv, self.stat0)
        numpy.add.at(out.stat1, inv, self.stat1)
        out.stat0 /= counts
        out.stat1 /= counts
        out.start = numpy.empty(out.segset.shape, "|O")
        out.stop = numpy.empty(out.segset.shape, "|O")
        return out
