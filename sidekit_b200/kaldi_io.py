"""Kaldi ark / scp tables of float matrices: the on-disk format of ``sidekit/bin/extract_xvectors.py:115-146, :158-173``
(``kaldiio.WriteHelper('ark,scp:...')`` / ``ReadHelper('scp:...')``; ``kaldiio`` is not in this image).

Binary ark entry (Kaldi's ``Matrix<float>::Write(binary=true)``):

    <key> ' ' '\\0' 'B' 'F' 'M' ' ' '\\4' <int32 rows> '\\4' <int32 cols> <rows*cols float32, row-major, little endian>

(a vector is ``'F' 'V' ' ' '\\4' <int32 dim> <dim float32>``); the scp line of an entry is ``<key> <ark path>:<offset>``
with ``offset`` pointing at the ``'\\0'`` of the binary marker.  The reference writes each x-vector as the (1, E) matrix
``Xtractor.forward`` returns, so that is what :class:`ArkScpWriter` writes by default; readers accept both records.
PARITY UNPINNED against ``kaldiio`` itself (absent here): the layout above is Kaldi's published one and is round-tripped
in ``tests/test_feed_path.py``.
"""
import os
import struct

import numpy


class ArkScpWriter:
    """``with ArkScpWriter(ark_path, scp_path) as w: w(key, array)`` -- same call shape as ``kaldiio.WriteHelper``."""

    def __init__(self, ark_path, scp_path=None):
        self.ark_path = os.path.realpath(ark_path)
        self._ark = open(self.ark_path, "wb")
        self._scp = open(scp_path, "w") if scp_path else None

    def __call__(self, key, array):
        key = str(key)
        if not key or any(c.isspace() for c in key):
            raise ValueError("Kaldi keys must be non-empty and free of whitespace: %r" % (key,))
        a = numpy.ascontiguousarray(numpy.asarray(array), dtype="<f4")
        self._ark.write(key.encode() + b" ")
        offset = self._ark.tell()
        if a.ndim == 1:
            self._ark.write(b"\0BFV \4" + struct.pack("<i", a.shape[0]))
        elif a.ndim == 2:
            self._ark.write(b"\0BFM \4" + struct.pack("<i", a.shape[0]) + b"\4" + struct.pack("<i", a.shape[1]))
        else:
            raise ValueError("only vectors and matrices can be written")
        self._ark.write(a.tobytes())
        if self._scp:
            self._scp.write("%s %s:%d\n" % (key, self.ark_path, offset))

    def close(self):
        self._ark.close()
        if self._scp:
            self._scp.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def _read_record(f):
    if f.read(2) != b"\0B":
        raise ValueError("not a binary Kaldi record")
    tag = f.read(3)
    if tag not in (b"FM ", b"FV "):
        raise ValueError("unsupported Kaldi type %r (float matrices / vectors only)" % (tag,))

    def dim():
        if f.read(1) != b"\4":
            raise ValueError("corrupt Kaldi header")
        return struct.unpack("<i", f.read(4))[0]
    if tag == b"FV ":
        n = dim()
        return numpy.frombuffer(f.read(4 * n), dtype="<f4").copy()
    r, c = dim(), dim()
    return numpy.frombuffer(f.read(4 * r * c), dtype="<f4").reshape(r, c).copy()


def read_scp(scp_path):
    """Generator of ``(key, array)`` in file order (``kaldiio.ReadHelper('scp:...')``)."""
    with open(scp_path) as scp:
        for line in scp:
            line = line.strip()
            if not line:
                continue
            key, where = line.split(None, 1)
            path, _, offset = where.rpartition(":")
            with open(path, "rb") as f:
                f.seek(int(offset))
                yield key, _read_record(f)


def read_ark(ark_path):
    """Generator of ``(key, array)`` over a binary ark file."""
    with open(ark_path, "rb") as f:
        while True:
            key = b""
            while True:
                ch = f.read(1)
                if not ch:
                    if key:
                        raise ValueError("truncated ark file")
                    return
                if ch == b" ":
                    break
                key += ch
            yield key.decode(), _read_record(f)


def write_xvectors(statserver, scp_path, keys=None):
    """The table ``extract_xvectors.py`` leaves behind: ``<scp stem>.ark`` + the scp, one (1, E) row matrix per segment
    (keys default to ``segset``).  Returns the ark path."""
    ark = os.path.join(os.path.dirname(os.path.realpath(scp_path)), os.path.splitext(os.path.basename(scp_path))[0] + ".ark")
    keys = statserver.segset if keys is None else keys
    with ArkScpWriter(ark, scp_path) as w:
        for k, v in zip(keys, numpy.asarray(statserver.stat1)):
            w(k, v[None, :])
    return ark


def speaker_means(scp_path, spk2utt, out_scp_path):
    """``extract_xvectors.py:158-173``: per speaker the L2-normalised mean of its utterances' x-vectors, written as a
    second ark / scp pair (vectors, as the reference writes ``mean`` of shape (1, E) squeezed by numpy.mean(axis=0))."""
    table = dict(read_scp(scp_path))
    ark = os.path.join(os.path.dirname(os.path.realpath(out_scp_path)), os.path.splitext(os.path.basename(out_scp_path))[0] + ".ark")
    with ArkScpWriter(ark, out_scp_path) as w:
        for spk, utts in spk2utt.items():
            mean = numpy.mean([table[u] for u in utts], axis=0)
            mean = mean / numpy.linalg.norm(mean, ord=2)
            w(spk, mean)
    return ark
