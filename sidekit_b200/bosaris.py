"""Light BOSARIS containers used as argument / return types of the scoring functions
(sidekit/bosaris/ndx.py:48-182, sidekit/bosaris/scores.py:52-81, :304-313): same attribute names,
same ``validate()`` rules, no HDF5 / plotting (out of scope of the hot path).
"""
import numpy


class Ndx:
    """Trial index: ``modelset`` (M,), ``segset`` (S,), ``trialmask`` (M, S) bool."""

    def __init__(self, ndx_file_name="", models=numpy.array([]), testsegs=numpy.array([])):
        self.modelset = numpy.empty(0, dtype="|O")
        self.segset = numpy.empty(0, dtype="|O")
        self.trialmask = numpy.array([], dtype="bool")
        if ndx_file_name != "":
            raise NotImplementedError("Ndx file IO is out of scope; set modelset / segset / trialmask directly")
        if len(models):
            # every (model, segment) pair listed is a trial (ndx.py:73-79), vectorised
            modelset, mi = numpy.unique(models, return_inverse=True)
            segset, si = numpy.unique(testsegs, return_inverse=True)
            mask = numpy.zeros((modelset.shape[0], segset.shape[0]), dtype="bool")
            mask[mi, si] = True
            self.modelset, self.segset, self.trialmask = modelset, segset, mask

    def filter(self, modlist, seglist, keep):
        """Same semantics as ndx.py:128-165 (order of the Ndx kept, every occurrence kept), O(N) with hash sets."""
        mods, segs = set(numpy.asarray(modlist).tolist()), set(numpy.asarray(seglist).tolist())
        keepmod = numpy.fromiter(((m in mods) == bool(keep) for m in self.modelset.tolist()), dtype=bool,
                                 count=self.modelset.shape[0])
        keepseg = numpy.fromiter(((s in segs) == bool(keep) for s in self.segset.tolist()), dtype=bool,
                                 count=self.segset.shape[0])
        out = Ndx()
        out.modelset = self.modelset[keepmod]
        out.segset = self.segset[keepseg]
        out.trialmask = self.trialmask[keepmod, :][:, keepseg]
        assert out.validate(), "Wrong Ndx format"
        return out

    def validate(self):
        ok = isinstance(self.modelset, numpy.ndarray)
        ok &= isinstance(self.segset, numpy.ndarray)
        ok &= isinstance(self.trialmask, numpy.ndarray)
        ok &= self.modelset.ndim == 1
        ok &= self.segset.ndim == 1
        ok &= self.trialmask.ndim == 2
        ok &= self.trialmask.shape == (self.modelset.shape[0], self.segset.shape[0])
        return bool(ok)


class Scores:
    """Score matrix container (scores.py:52-81).  ``scoremat`` is materialised lazily: the CUDA result stays on
    the device (``scoremat_device``, a torch tensor) until the numpy array is first read."""

    def __init__(self, scores_file_name=""):
        if scores_file_name != "":
            raise NotImplementedError("Scores file IO is out of scope")
        self.modelset = numpy.empty(0, dtype="|O")
        self.segset = numpy.empty(0, dtype="|O")
        self.scoremask = numpy.array([], dtype="bool")
        self._scoremat = numpy.array([])
        self.scoremat_device = None

    @property
    def scoremat(self):
        if self._scoremat is None:
            self._scoremat = self.scoremat_device.cpu().numpy()
        return self._scoremat

    @scoremat.setter
    def scoremat(self, value):
        self._scoremat = value

    def validate(self):
        ok = self.scoremat.shape == self.scoremask.shape
        ok &= (self.scoremat.shape[0] == self.modelset.shape[0])
        ok &= (self.scoremat.shape[1] == self.segset.shape[0])
        return bool(ok)
