"""The D x D host algebra of the PLDA-family scorers (sidekit_b200/iv_scoring.py) against (a) the DEFINITION of the
log-likelihood ratio under the two-covariance model and (b) the reference's own formulas (sidekit/iv_scoring.py:193-197,
:296-313, :429-448), restated here.  CPU only."""
import numpy
import scipy.linalg

from sidekit_b200 import iv_scoring as I
from sidekit_b200 import synth


def _logpdf0(x, cov):
    """log N(x; 0, cov)."""
    sign, logdet = numpy.linalg.slogdet(cov)
    return -0.5 * (x.shape[0] * numpy.log(2 * numpy.pi) + logdet + x @ numpy.linalg.solve(cov, x))


def test_simplified_plda_terms_reproduce_the_likelihood_ratio_by_definition():
    """llr(e, t) = log N([e; t]; 0, [[T, B], [B, T]]) - log N(e; 0, T) - log N(t; 0, T) with B = F F', T = B + Sigma."""
    rng = numpy.random.default_rng(3)
    _, F, Sigma = synth.synth_plda(24, 10, seed=5)
    Phi, Psi, cst = I._simplified_plda_terms(F, Sigma, 1.0)
    B = F @ F.T
    T = B + Sigma
    joint = numpy.block([[T, B], [B, T]])
    for _ in range(5):
        e, t = rng.standard_normal(24), rng.standard_normal(24)
        llr = _logpdf0(numpy.concatenate([e, t]), joint) - _logpdf0(e, T) - _logpdf0(t, T)
        got = e @ Psi @ t + 0.5 * (e @ Phi @ e + t @ Phi @ t) + cst
        assert abs(got - llr) < 1e-9 * max(1.0, abs(llr))
    # no speaker subspace: the two hypotheses coincide
    Phi0, Psi0, cst0 = I._simplified_plda_terms(F[:, :0], Sigma, 1.0)
    assert cst0 == 0.0 and numpy.abs(Phi0).max() < 1e-12 and numpy.abs(Psi0).max() < 1e-12


def test_plda_terms_match_the_reference_formulas():
    _, F, Sigma = synth.synth_plda(32, 12, seed=7)
    G = 0.2 * numpy.random.default_rng(1).standard_normal((32, 6))
    for sf in (1.0, 0.6):
        # sidekit/iv_scoring.py:429-448
        invSigma = scipy.linalg.inv(Sigma)
        I_spk = numpy.eye(F.shape[1])
        K = F.T.dot(invSigma * sf).dot(F)
        K1, K2 = scipy.linalg.inv(K + I_spk), scipy.linalg.inv(2 * K + I_spk)
        cst = numpy.linalg.slogdet(K2)[1] / 2.0 - numpy.linalg.slogdet(K1)[1]
        Sac = F.dot(F.T)
        St = Sac + Sigma
        Sti = scipy.linalg.inv(St)
        Tmp = numpy.linalg.inv(St - Sac.dot(Sti).dot(Sac))
        Phi, Psi, c = I._simplified_plda_terms(F, Sigma, sf)
        assert numpy.abs(Phi - (Sti - Tmp)).max() < 1e-12 and numpy.abs(Psi - Sti.dot(Sac).dot(Tmp)).max() < 1e-12 and abs(c - cst) < 1e-11
        # sidekit/iv_scoring.py:296-313 (channel subspace)
        A = numpy.linalg.inv(G.T.dot(invSigma * sf).dot(G) + numpy.eye(G.shape[1]))
        Bm = F.T.dot(invSigma * sf).dot(numpy.eye(32) - G.dot(A).dot(G.T).dot(invSigma * sf))
        K = Bm.dot(F)
        K1, K2 = scipy.linalg.inv(K + I_spk), scipy.linalg.inv(2 * K + I_spk)
        Psi_f, Phi_f, c_f = I._full_plda_terms(F, G, Sigma, sf)
        assert numpy.abs(Psi_f - Bm.T.dot(K2).dot(Bm)).max() < 1e-12 and numpy.abs(Phi_f - Bm.T.dot(K2 - K1).dot(Bm)).max() < 1e-12
        assert abs(c_f - (numpy.linalg.slogdet(K2)[1] / 2.0 - numpy.linalg.slogdet(K1)[1])) < 1e-11
