"""GPU parity of the DBoW2 transform against the oracle: words, FeatureVector nodes and the BowVector doubles, bit-exact."""
import numpy as np
import pytest

from cubemapslam_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,L,levelsup", [(10, 4, 2), (6, 5, 4)])
def test_bow_transform(oracle, k, L, levelsup):
    from cubemapslam_b200.vocabulary import Vocabulary
    voc = synth.vocabulary(k=k, L=L, seed=k + L)
    gv = Vocabulary(k, L, voc["parent"], voc["is_leaf"], voc["desc"], voc["weight"], max_frames=2, max_features=2048)
    ov = oracle.Vocabulary(k, L, voc["parent"], voc["is_leaf"], voc["desc"], voc["weight"])
    rng = np.random.default_rng(1)
    leaves = np.nonzero(voc["is_leaf"])[0]
    frames = []
    for nfeat in (2000, 700):
        f1 = voc["desc"][rng.choice(leaves, nfeat - 100)] ^ np.packbits(rng.random((nfeat - 100, 256)) < 0.05, axis=1, bitorder="little")
        frames.append(np.concatenate([f1, rng.integers(0, 256, (100, 32), dtype=np.uint8)]))
    stride = 2048
    d = np.zeros((2, stride, 32), np.uint8); n = np.array([len(f) for f in frames], np.int32)
    for i, f in enumerate(frames):
        d[i, :len(f)] = f
    res = gv.transform(d, n=n, levelsup=levelsup)
    for i, f in enumerate(frames):
        bw, bv, node, word = ov.transform(f, levelsup)
        gw, gn, gbw, gbv = res[i]
        assert np.array_equal(gw, word) and np.array_equal(gn, node)
        assert np.array_equal(gbw, bw) and np.array_equal(gbv, bv)
    gv.close()
