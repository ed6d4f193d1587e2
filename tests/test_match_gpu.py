"""GPU parity: K10 through the C-ABI vs the oracle (and live cv2) -- bit-exact indices and distances."""
import numpy as np
import pytest

from oracle import vo_restate as V

pytestmark = pytest.mark.gpu


def _check(ctx, q, t, gate=None):
    qi, ti, d = V.bf_match_crosscheck(q, t)
    if gate is not None:
        qi, ti, d = V.match_gate(qi, ti, d, gate)
        m = ctx.match_hamming(q, t, True, 2.0, 30.0 * gate)
    else:
        m = ctx.match_hamming(q, t, True)
    assert np.array_equal(m["queryIdx"], qi)
    assert np.array_equal(m["trainIdx"], ti)
    assert np.array_equal(m["distance"], d.astype(np.float32))
    assert (m["imgIdx"] == 0).all()
    return m


@pytest.mark.parametrize("nq,nt,seed", [(300, 500, 0), (500, 300, 1), (2000, 2000, 2), (1, 1, 3), (17, 1000, 4),
                                        (257, 129, 5), (4000, 4000, 6), (8192, 8000, 7)])
def test_match_random(pkg, gpu_ctx, nq, nt, seed):
    q, t = pkg.synth.synth_descriptor_pair(seed, nq, nt)
    _check(gpu_ctx, q, t)
    _check(gpu_ctx, q, t, gate=1.0)


def test_match_ties(pkg, gpu_ctx):
    q = pkg.synth.synth_descriptors(10, 700, dup_frac=0.3)
    t = np.concatenate([q[::2], pkg.synth.synth_descriptors(11, 300, dup_frac=0.3)])
    _check(gpu_ctx, q, t)
    _check(gpu_ctx, q, t, gate=2.0)


def test_match_vs_cv2(pkg, gpu_ctx):
    cv2 = pytest.importorskip("cv2")
    q, t = pkg.synth.synth_descriptor_pair(21, 1500, 1800)
    m = gpu_ctx.match_hamming(q, t, True)
    ref = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(q, t)
    assert [x.queryIdx for x in ref] == m["queryIdx"].tolist()
    assert [x.trainIdx for x in ref] == m["trainIdx"].tolist()
    assert [x.distance for x in ref] == m["distance"].tolist()


def test_match_empty_and_errors(pkg, gpu_ctx):
    q = pkg.synth.synth_descriptors(1, 10)
    assert len(gpu_ctx.match_hamming(q[:0], q, True)) == 0      # reference: UB on empty (vo.cpp:229-242)
    assert len(gpu_ctx.match_hamming(q, q[:0], True)) == 0
    big = pkg.synth.synth_descriptors(2, 9000)
    with pytest.raises(pkg.VslamError) as e:
        gpu_ctx.match_hamming(big, q, True)
    assert e.value.status == -2


def test_match_no_crosscheck(pkg, gpu_ctx):
    q, t = pkg.synth.synth_descriptor_pair(30, 600, 900)
    D = V.hamming_matrix(q, t)
    m = gpu_ctx.match_hamming(q, t, False)
    assert np.array_equal(m["queryIdx"], np.arange(600))
    assert np.array_equal(m["trainIdx"], D.argmin(1))


def test_match_properties_full_size(pkg, gpu_ctx):
    """size-independent properties at the BASELINE size: symmetry of mutual matching, distance recomputation."""
    q, t = pkg.synth.synth_descriptor_pair(40, 4000, 4000)
    a = gpu_ctx.match_hamming(q, t, True)
    b = gpu_ctx.match_hamming(t, q, True)
    # mutual NN is symmetric up to tie-breaking; with random 256-bit descriptors the matched pair set is identical
    sa = set(zip(a["queryIdx"].tolist(), a["trainIdx"].tolist()))
    sb = set(zip(b["trainIdx"].tolist(), b["queryIdx"].tolist()))
    assert len(sa ^ sb) <= 0.002 * len(sa)
    x = np.bitwise_xor(q[a["queryIdx"]], t[a["trainIdx"]])
    d = np.unpackbits(x, axis=1).sum(1)
    assert np.array_equal(d.astype(np.float32), a["distance"])
    assert (np.diff(a["queryIdx"]) > 0).all()
