"""Text formats either side of the path: ``IdMap`` / ``Ndx`` / ``Key`` / ``Scores`` ``read_txt`` / ``write_txt`` /
``save_txt`` (sidekit/bosaris/idmap.py:118-126, :312-340; ndx.py:115-126, :208-236; key.py:151-164, :262-304;
scores.py:118-131, :371-412).  Files written by the reference and the objects it reads back are committed under
tests/golden/text_io/ (oracle/make_golden.py text_io); with /root/reference present the same comparison runs live."""
import os

import numpy
import pytest

import sidekit_b200 as sk
from tests.helpers import GOLD

TXT = os.path.join(GOLD, "text_io")


def _objects():
    rng = numpy.random.default_rng(21)
    models = numpy.array(["m%02d" % i for i in rng.permutation(7)], dtype="|O")
    segs = numpy.array(["seg_%03d" % i for i in rng.permutation(11)], dtype="|O")
    key = sk.Key()
    key.modelset, key.segset = models, segs
    r = rng.random((7, 11))
    key.tar, key.non = r < 0.25, (r >= 0.25) & (r < 0.8)
    ndx = key.to_ndx()
    sc = sk.Scores()
    sc.modelset, sc.segset, sc.scoremask = models, segs, ndx.trialmask.copy()
    sc.scoremat = rng.standard_normal((7, 11)) * 3.0
    sc32 = sk.Scores()
    sc32.modelset, sc32.segset, sc32.scoremask = models, segs, ndx.trialmask.copy()
    sc32.scoremat = (rng.standard_normal((7, 11)) * 3.0).astype(numpy.float32)
    im = sk.IdMap()
    im.leftids = numpy.array(["spk%d" % (i % 3) for i in range(6)], dtype="|O")
    im.rightids = numpy.array(["dir/file%d" % i for i in range(6)], dtype="|O")
    im.start, im.stop = numpy.arange(6) * 10, numpy.arange(6) * 10 + 300
    return key, ndx, sc, sc32, im


def _read(name):
    with open(os.path.join(TXT, name)) as f:
        return f.read()


def test_writers_are_byte_identical_to_the_reference_files(tmp_path):
    key, ndx, sc, sc32, im = _objects()
    out = lambda n: os.path.join(tmp_path, "sub", n) if n.startswith("scores") else os.path.join(tmp_path, n)
    key.write_txt(out("key.txt"))
    ndx.save_txt(out("ndx.txt"))
    sc.write_txt(out("scores64.txt"))            # creates the missing directory, like the reference
    sc32.write_txt(out("scores32.txt"))
    im.write_txt(out("idmap4.txt"))
    im.start = numpy.array([None] * 6, dtype="|O")
    im.stop = numpy.array([None] * 6, dtype="|O")
    im.write_txt(out("idmap_none.txt"))
    for n in ("key.txt", "ndx.txt", "scores64.txt", "scores32.txt", "idmap4.txt", "idmap_none.txt"):
        with open(out(n)) as f:
            assert f.read() == _read(n), n


def test_readers_match_the_objects_the_reference_reads():
    g = numpy.load(os.path.join(TXT, "read_back.npz"), allow_pickle=False)
    key = sk.Key.read_txt(os.path.join(TXT, "key.txt"))
    assert numpy.array_equal(key.modelset, g["key_modelset"]) and numpy.array_equal(key.segset, g["key_segset"])
    assert numpy.array_equal(key.tar, g["key_tar"]) and numpy.array_equal(key.non, g["key_non"])
    ndx = sk.Ndx.read_txt(os.path.join(TXT, "ndx.txt"))
    assert numpy.array_equal(ndx.modelset, g["ndx_modelset"]) and numpy.array_equal(ndx.segset, g["ndx_segset"])
    assert numpy.array_equal(ndx.trialmask, g["ndx_trialmask"])
    for n in ("scores64", "scores32"):
        sc = sk.Scores.read_txt(os.path.join(TXT, n + ".txt"))
        assert numpy.array_equal(sc.modelset, g[n + "_modelset"]) and numpy.array_equal(sc.segset, g[n + "_segset"])
        assert numpy.array_equal(sc.scoremask, g[n + "_scoremask"])
        assert sc.scoremat.dtype == numpy.float64 and numpy.array_equal(sc.scoremat, g[n + "_scoremat"])
    im = sk.IdMap.read_txt(os.path.join(TXT, "idmap4.txt"))
    assert numpy.array_equal(im.leftids.astype("U"), g["idmap4_leftids"]) and numpy.array_equal(im.rightids.astype("U"), g["idmap4_rightids"])
    assert numpy.array_equal(im.start, g["idmap4_start"]) and numpy.array_equal(im.stop, g["idmap4_stop"])
    im2 = sk.IdMap.read_txt(os.path.join(TXT, "idmap2.txt"))
    assert numpy.array_equal(im2.leftids.astype("U"), g["idmap2_leftids"]) and all(v is None for v in im2.start) and all(v is None for v in im2.stop)
    with pytest.raises(ValueError):                 # 'None' boundaries written by write_txt do not parse back (reference quirk)
        sk.IdMap.read_txt(os.path.join(TXT, "idmap_none.txt"))


def test_round_trips_and_last_label_wins(tmp_path):
    key, ndx, sc, sc32, im = _objects()
    p = os.path.join(tmp_path, "k.txt")
    key.write_txt(p)
    back = sk.Key.read_txt(p)
    o_m, o_s = numpy.argsort(key.modelset), numpy.argsort(key.segset)
    keep_m = (key.tar | key.non).any(axis=1)[o_m]
    assert numpy.array_equal(back.modelset, key.modelset[o_m][keep_m])
    assert numpy.array_equal(back.tar, key.tar[o_m][:, o_s][keep_m][:, (key.tar | key.non).any(axis=0)[o_s]])
    with open(p, "w") as f:
        f.write("a s1 target\na s1 nontarget\nb s1 maybe\n")
    k2 = sk.Key.read_txt(p)
    assert k2.non[0, 0] and not k2.tar[0, 0] and not k2.tar[1, 0] and not k2.non[1, 0]
    p = os.path.join(tmp_path, "s.txt")
    sc.write_txt(p)
    back = sk.Scores.read_txt(p)
    assert numpy.array_equal(back.scoremat[back.scoremask], sc.scoremat[o_m][:, o_s][sc.scoremask[o_m][:, o_s]])


@pytest.mark.reference
def test_live_reference_agrees(tmp_path):
    from oracle.make_golden import text_io_reference
    ref_dir = os.path.join(tmp_path, "ref")
    os.makedirs(ref_dir)
    g = text_io_reference(ref_dir, _objects())
    for n in ("key.txt", "ndx.txt", "scores64.txt", "scores32.txt", "idmap4.txt", "idmap_none.txt"):
        with open(os.path.join(ref_dir, n)) as f:
            assert f.read() == _read(n), n
    committed = numpy.load(os.path.join(TXT, "read_back.npz"), allow_pickle=False)
    assert sorted(g.keys()) == sorted(committed.files)
    for k in g:
        assert numpy.array_equal(numpy.asarray(g[k]), committed[k]), k


@pytest.mark.reference
def test_container_utilities_match_the_reference():
    """IdMap.set / map_* / filter_on_* / merge / split, Key.filter, Scores.filter / get_score against the reference's own
    methods (bosaris/idmap.py:128-392, key.py:166-205, scores.py:262-302, :480-495) on the same objects."""
    from oracle import ref_import
    ref_import.import_reference()
    from sidekit.bosaris import IdMap as RIdMap, Key as RKey, Scores as RScores
    key, ndx, sc, sc32, im = _objects()

    def pair(cls_ref, mine, names):
        r = cls_ref()
        for n in names:
            setattr(r, n, getattr(mine, n).copy())
        return r

    f = ("leftids", "rightids", "start", "stop")
    rim = pair(RIdMap, im, f)
    same = lambda a, b: all(numpy.array_equal(numpy.asarray(getattr(a, n)), numpy.asarray(getattr(b, n))) for n in f)
    q = numpy.array(["spk2", "nobody", "spk0"], dtype="|O")
    assert numpy.array_equal(im.map_left_to_right(q), rim.map_left_to_right(q))
    q = numpy.array(["dir/file4", "dir/file1", "zzz"], dtype="|O")
    assert numpy.array_equal(im.map_right_to_left(q), rim.map_right_to_left(q))
    for keep in (True, False):
        assert same(im.filter_on_left(["spk1", "spk9"], keep), rim.filter_on_left(["spk1", "spk9"], keep))
        assert same(im.filter_on_right(["dir/file0", "dir/file5"], keep), rim.filter_on_right(["dir/file0", "dir/file5"], keep))
    other = sk.IdMap()
    other.set(numpy.array(["spk0", "new"], dtype="|O"), numpy.array(["dir/file0", "dir/file9"], dtype="|O"),
              numpy.array([0, 5]), numpy.array([300, 90]))
    rother = pair(RIdMap, other, f)
    assert same(sk.IdMap.merge(im, other), RIdMap.merge(rim, rother))
    for a, b in zip(im.split(4), rim.split(4)):
        assert same(a, b)
    rkey = pair(RKey, key, ("modelset", "segset", "tar", "non"))
    rsc = pair(RScores, sc, ("modelset", "segset", "scoremask", "scoremat"))
    mods, segs = ["m03", "m05", "zz"], ["seg_001", "seg_007", "seg_010", "qq"]
    for keep in (True, False):
        a, b = key.filter(mods, segs, keep), rkey.filter(mods, segs, keep)
        assert all(numpy.array_equal(getattr(a, n), getattr(b, n)) for n in ("modelset", "segset", "tar", "non"))
        a, b = sc.filter(mods, segs, keep), rsc.filter(mods, segs, keep)
        assert all(numpy.array_equal(getattr(a, n), getattr(b, n)) for n in ("modelset", "segset", "scoremask", "scoremat"))
    assert numpy.array_equal(sc.get_score("m03", "seg_007"), rsc.get_score("m03", "seg_007"))
    with pytest.raises(Exception):
        sc.get_score("nope", "seg_007")


def test_container_utilities_basic():
    key, ndx, sc, sc32, im = _objects()
    assert list(im.filter_on_left(["spk1"], True).rightids) == ["dir/file1", "dir/file4"]
    assert list(im.map_left_to_right(numpy.array(["spk2"], dtype="|O"))) == ["dir/file5"]          # the LAST pair of a left id
    assert [m.leftids.shape[0] for m in im.split(4)] == [2, 2, 1, 1]
    k = key.filter(key.modelset[:2], key.segset[:3], True)
    assert k.tar.shape == (2, 3) and numpy.array_equal(k.tar, key.tar[:2, :3])
    s = sc.filter(sc.modelset[:2], sc.segset[:3], False)
    assert s.scoremat.shape == (5, 8) and numpy.array_equal(s.scoremat, sc.scoremat[2:, 3:])
    assert sc.get_score(sc.modelset[1], sc.segset[2]).shape == (1, 1)


@pytest.mark.reference
def test_merges_match_the_reference():
    """Ndx.merge / Key.merge / Scores.merge (ndx.py:239-283, key.py:305-367, scores.py:414-467), sorted and unsorted sets."""
    import copy
    from oracle import ref_import
    ref_import.import_reference()
    from sidekit.bosaris import Key as RKey, Ndx as RNdx, Scores as RScores
    rng = numpy.random.default_rng(77)

    def fill(cls, src, names):
        o = cls()
        for n in names:
            setattr(o, n, copy.deepcopy(getattr(src, n)))
        return o

    def objects(models, segs, seed):
        r = numpy.random.default_rng(seed).random((len(models), len(segs)))
        k = sk.Key()
        k.modelset, k.segset = numpy.array(models, dtype="|O"), numpy.array(segs, dtype="|O")
        k.tar, k.non = r < 0.3, (r >= 0.3) & (r < 0.7)
        s = sk.Scores()
        s.modelset, s.segset, s.scoremask = k.modelset.copy(), k.segset.copy(), r < 0.7
        s.scoremat = numpy.random.default_rng(seed + 1).standard_normal(r.shape)
        return k, k.to_ndx(), s

    for shuffle in (False, True):
        m1, s1 = ["m%d" % i for i in range(5)], ["s%d" % i for i in range(6)]
        m2, s2 = ["m%d" % i for i in range(3, 9)], ["s%d" % i for i in range(6, 11)]          # disjoint segments: no clashes
        if shuffle:
            m1, s1 = list(rng.permutation(m1)), list(rng.permutation(s1))
        (k1, n1, c1), (k2, n2, c2) = objects(m1, s1, 5), objects(m2, s2, 9)
        rn1, rn2 = fill(RNdx, n1, ("modelset", "segset", "trialmask")), fill(RNdx, n2, ("modelset", "segset", "trialmask"))
        n1.merge([n2]); rn1.merge([rn2])
        assert all(numpy.array_equal(getattr(n1, a), getattr(rn1, a)) for a in ("modelset", "segset", "trialmask"))
        rk1, rk2 = fill(RKey, k1, ("modelset", "segset", "tar", "non")), fill(RKey, k2, ("modelset", "segset", "tar", "non"))
        k1.merge([k2]); rk1.merge([rk2])
        assert all(numpy.array_equal(getattr(k1, a), getattr(rk1, a)) for a in ("modelset", "segset", "tar", "non"))
        f = ("modelset", "segset", "scoremask", "scoremat")
        rc1, rc2 = fill(RScores, c1, f), fill(RScores, c2, f)
        c1.merge([c2]); rc1.merge([rc2])
        assert all(numpy.array_equal(getattr(c1, a), getattr(rc1, a)) for a in f)
    ka, _, _ = objects(["a"], ["x"], 1)
    kb, _, _ = objects(["a"], ["x"], 1)
    ka.tar[:], ka.non[:], kb.tar[:], kb.non[:] = True, False, False, True
    with pytest.raises(AssertionError):
        ka.merge([kb])


def test_ndx_filter_fast_paths_keep_the_reference_semantics():
    """Ndx.filter (ndx.py:128-165): identity / rows-only / columns-only / general selections give what two-step boolean
    indexing gives; the identity case shares the mask instead of copying it twice."""
    import sidekit_b200 as sk
    rng = numpy.random.default_rng(0)
    n = sk.Ndx()
    n.modelset = numpy.array(["m%d" % i for i in range(7)])
    n.segset = numpy.array(["s%d" % i for i in range(5)])
    n.trialmask = rng.random((7, 5)) < 0.5
    for mods, segs in ((n.modelset, n.segset), (n.modelset[[4, 1, 2]], n.segset), (n.modelset, n.segset[[3, 0]]),
                       (n.modelset[[6, 0]], n.segset[[1, 4, 2]]), (numpy.array(["zz"]), n.segset)):
        for keep in (True, False):
            out = n.filter(mods, segs, keep)
            km = numpy.isin(n.modelset, mods) == keep
            ks = numpy.isin(n.segset, segs) == keep
            assert numpy.array_equal(out.modelset, n.modelset[km]) and numpy.array_equal(out.segset, n.segset[ks])
            assert numpy.array_equal(out.trialmask, n.trialmask[km, :][:, ks]) and out.validate()
    same = n.filter(n.modelset, n.segset, True)
    assert numpy.shares_memory(same.trialmask, n.trialmask) and not numpy.shares_memory(same.modelset, n.modelset)


def test_get_tar_non_with_a_key_over_other_sets_equals_the_aligned_matrix():
    """scores.py:157-179: ``get_tar_non`` with a key whose model / segment sets differ from the scores' (other order, ids
    missing on both sides) returns what indexing the ``align_with_ndx`` copy returns -- gathered trial by trial here."""
    import sidekit_b200 as sk
    rng = numpy.random.default_rng(1)
    ids = numpy.array(["u%03d" % i for i in range(50)])
    S = sk.Scores()
    S.modelset, S.segset = ids[rng.permutation(50)][:40], ids[rng.permutation(50)][:45]
    S.scoremat, S.scoremask = rng.standard_normal((40, 45)), rng.random((40, 45)) < 0.8
    key = sk.Key(models=ids[rng.integers(0, 50, 300)], testsegs=ids[rng.integers(0, 50, 300)],
                 trials=numpy.where(rng.random(300) < 0.5, "target", "nontarget"))
    tar, non = S.get_tar_non(key)
    al = S.align_with_ndx(key)
    assert numpy.array_equal(tar, al.scoremat[key.tar & al.scoremask]) and numpy.array_equal(non, al.scoremat[key.non & al.scoremask])
    same = sk.Scores()
    same.modelset, same.segset = key.modelset, key.segset
    same.scoremat, same.scoremask = rng.standard_normal(key.tar.shape), rng.random(key.tar.shape) < 0.7
    tar, non = same.get_tar_non(key)
    assert numpy.array_equal(tar, same.scoremat[key.tar & same.scoremask]) and numpy.array_equal(non, same.scoremat[key.non & same.scoremask])
