"""oracle/_ref/libref_dbow2.so — the reference's OWN bag-of-words code (Thirdparty/DBoW2: TemplatedVocabulary.h, FORB.cpp, BowVector.cpp,
FeatureVector.cpp, ScoringObject.cpp, compiled unmodified and instantiated as include/ORBVocabulary.h does) against the restated oracle
(orc_vocab_transform / orc_bow_from_words), which is what the device path orbx_vocabulary_transform is tested against.
The vocabulary is a synthetic tree written in the ORBvoc.txt text format and read back by the reference's own loadFromTextFile
(ORBvoc.txt itself is absent from the checkout, D8).  CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
from openmavis_b200.synth import synth_descriptors

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def RV():
    global _lib
    if _lib is None:
        so = os.path.join(ROOT, 'oracle', '_ref', 'libref_dbow2.so')
        if not os.path.exists(so):
            O.ref_lib()                                      # runs make when /root/reference is mounted
        if not os.path.exists(so):
            raise RuntimeError('%s is missing and /root/reference is not mounted to build it' % so)
        _lib = C.CDLL(so)
        _lib.refv_load.restype = C.c_void_p
        _lib.refv_load.argtypes = [C.c_char_p]
        _lib.refv_destroy.argtypes = [C.c_void_p]
        _lib.refv_size.argtypes = [C.c_void_p]
        _lib.refv_word.argtypes = [C.c_void_p, C.c_void_p]
        _lib.refv_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
    return _lib


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def synthetic_tree(k, L, seed, irregular=False):
    """Random DBoW2-shaped tree in breadth-first order (the order of ORBvoc.txt): node 0 = root; leaves get word ids in node order."""
    rng = np.random.default_rng(seed)
    desc = [np.zeros(32, np.uint8)]; children = [[]]; parent = [0]
    frontier = [0]
    for lv in range(1, L + 1):
        nxt = []
        for q in frontier:
            nk = int(rng.integers(1, k + 1)) if irregular else k
            if irregular and lv > 1 and rng.random() < 0.15:
                continue                                           # early leaf
            for _ in range(nk):
                base = desc[q] if q else rng.integers(0, 256, 32, dtype=np.uint8)
                flips = np.packbits(rng.random(256) < (0.25 / lv), bitorder='little')
                desc.append(base ^ flips); children.append([]); parent.append(q)
                children[q].append(len(desc) - 1); nxt.append(len(desc) - 1)
        frontier = nxt
    n = len(desc)
    child_ptr = np.concatenate([[0], np.cumsum([len(c) for c in children])]).astype(np.int32)
    child = np.array([c for cs in children for c in cs], np.int32)
    word = np.full(n, -1, np.int32); weight = np.zeros(n, np.float64)
    leaves = [i for i in range(n) if not children[i]]
    word[leaves] = np.arange(len(leaves), dtype=np.int32)
    weight[leaves] = np.where(rng.random(len(leaves)) < 0.03, 0.0, rng.uniform(0.5, 9.0, len(leaves)))
    return child_ptr, child, np.stack(desc), word, weight, leaves, parent, children


def write_orbvoc(path, k, L, nd, weight, parent, children):
    """'k L scoring weighting' (L1_NORM = 0, TF_IDF = 0: the ORBvoc.txt settings), then one line per node 1..n-1:
    'parent isLeaf d0 .. d31 weight'.  No newline after the last line: loadFromTextFile loops on !eof() and would otherwise append an
    empty child to the root (TemplatedVocabulary.h:1377-1414)."""
    lines = ['%d %d 0 0' % (k, L)]
    for i in range(1, len(nd)):
        lines.append('%d %d %s %s' % (parent[i], 0 if children[i] else 1, ' '.join(str(int(x)) for x in nd[i]), repr(float(weight[i]))))
    with open(path, 'w') as f:
        f.write('\n'.join(lines))


@pytest.mark.parametrize('k,L,irregular,levelsup', [(10, 4, False, 4), (10, 4, False, 2), (10, 3, True, 2), (20, 2, False, 1), (3, 6, True, 5), (4, 6, False, 4),
                                                    (10, 5, True, 4)])
def test_transform_reference_equals_oracle(tmp_path, k, L, irregular, levelsup):
    """(Irregular trees only with L - levelsup <= 1: a leaf above the level whose node id transform() reports leaves the reference's
    `NodeId nid` uninitialised — TemplatedVocabulary.h:1147-1155 declares it inside the loop and :1245-1246 assigns it only when that level
    is reached — so the reference's FeatureVector is undefined there; the oracle and the device path report node 0.  ORBvoc.txt is a
    complete 10^6 tree, where this cannot happen.)"""
    Lr = RV(); Lo = O.lib()
    child_ptr, child, nd, word, weight, leaves, parent, children = synthetic_tree(k, L, 500 + k + L, irregular)
    path = str(tmp_path / 'voc.txt')
    write_orbvoc(path, k, L, nd, weight, parent, children)
    v = Lr.refv_load(path.encode())
    assert v, 'loadFromTextFile failed'
    try:
        assert Lr.refv_size(v) == len(leaves)
        rng = np.random.default_rng(5)
        n = 1200
        feat = nd[rng.choice(leaves, n)] ^ np.packbits(rng.random((n, 256)) < 0.08, axis=1, bitorder='little')
        feat[::7] = synth_descriptors((n + 6) // 7, 17)                # unrelated features
        feat[1] = feat[0]                                              # duplicates accumulate in the same word
        feat = np.ascontiguousarray(feat)
        # reference
        bi = np.zeros(n, np.int32); bv = np.zeros(n, np.float64); fn = np.zeros(n, np.int32); fp = np.zeros(n + 1, np.int32); fi = np.zeros(n, np.int32)
        nfn = C.c_int(0)
        nb = Lr.refv_transform(v, p(feat), n, levelsup, p(bi), p(bv), p(fn), p(fp), p(fi), C.byref(nfn))
        # oracle: words / nodes, then the two maps
        ow = np.zeros(n, np.int32); on = np.zeros(n, np.int32)
        Lo.orc_vocab_transform(p(child_ptr), p(child), p(nd), p(word), L, p(feat), n, levelsup, p(ow), p(on))
        leaf_of_word = np.zeros(len(leaves), np.int64); leaf_of_word[word[leaves]] = leaves
        w = np.ascontiguousarray(weight[leaf_of_word[ow]])
        ci = np.zeros(n, np.int32); cv = np.zeros(n, np.float64); cn = np.zeros(n, np.int32); cp = np.zeros(n + 1, np.int32); cx = np.zeros(n, np.int32)
        cfn = C.c_int(0)
        cb = Lo.orc_bow_from_words(p(ow), p(on), p(w), n, p(ci), p(cv), p(cn), p(cp), p(cx), C.byref(cfn))
        assert nb == cb and nb >= 3 and nfn.value == cfn.value
        assert np.array_equal(bi[:nb], ci[:cb])
        assert np.array_equal(bv[:nb].view(np.uint64), cv[:cb].view(np.uint64))          # the doubles to the last bit
        assert np.array_equal(fn[:nfn.value], cn[:cfn.value]) and np.array_equal(fp[:nfn.value + 1], cp[:cfn.value + 1])
        assert np.array_equal(fi[:fp[nfn.value]], cx[:cp[cfn.value]])
        # single-feature overload: the word ids themselves
        for i in range(0, n, 37):
            assert Lr.refv_word(v, p(feat[i])) == ow[i]
    finally:
        Lr.refv_destroy(v)


def test_forb_distance_is_the_popcount():
    Lr = RV()
    a = synth_descriptors(500, 3); b = synth_descriptors(500, 4)
    b[:5] = a[:5]; b[5:10] = ~a[5:10]
    for i in range(500):
        assert Lr.refv_distance(p(a[i]), p(b[i])) == O.hamming(a[i], b[i])
