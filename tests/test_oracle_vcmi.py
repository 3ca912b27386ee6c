"""Pin the VCMI oracle to the reference (Model.py:150-225)."""
import numpy as np
import pytest

from conftest import cfg_of, load_golden, rel_err
from oracle import params as P
from oracle import vcmi_oracle as V

VCMI = load_golden("vcmi")


def inputs(rec):
    c = cfg_of(rec)
    seed = int(rec["seed"])
    stack = P.vcmi_params(seed, c["embed"], c["hidden"])
    if c["act"] == "hardtanh":
        stack[-1] = (stack[-1][0] * 0.5, stack[-1][1] + 0.5)
    fx = P.features(seed + 1, c["bs"], c["embed"], c["scale"])
    fy = P.features(seed + 2, c["bs"], c["wy"], c["scale"])
    fz = P.features(seed + 3, c["bs"], c["embed"], c["scale"])
    kx = P.features(seed + 4, c["nprod"], c["embed"], c["scale"])
    ky = P.features(seed + 5, c["nprod"], c["embed"], c["scale"])
    kz = P.features(seed + 6, c["nprod"], c["embed"], c["scale"])
    return c, stack, (fx, fy, fz, kx, ky, kz)


@pytest.mark.parametrize("case", sorted(VCMI))
def test_vcmi_matches_reference(case):
    rec = VCMI[case]
    c, stack, ins = inputs(rec)
    for tag, wc, wl in (("gl", 0.0, 1.0), ("gc", 1.0, 0.0)):
        r = V.vcmi_estimator(stack, c["act"], c["embed"], *ins, w_cmi=wc, w_loss=wl)
        assert abs(r["cmi"] - rec["cmi"]) <= 5e-5 * max(1.0, abs(rec["cmi"]))
        assert abs(r["loss"] - rec["loss"]) <= 5e-5 * max(1.0, abs(rec["loss"]))
        for nm in ("fx", "fy", "fz", "kx", "ky", "kz"):
            ref = rec[f"{tag}_{nm}"]
            got = r["grads"][nm]
            assert np.abs(got - ref).max() <= 5e-5 * np.abs(ref).max() + 1e-8, (tag, nm)
        for k, v in rec.items():
            if k.startswith(tag + "p__"):
                assert np.abs(r["pg"][k[5:]] - v).max() <= 5e-5 * np.abs(v).max() + 1e-7, k
            elif k.startswith(tag + "s__"):
                g = r["pg"][k[5:]].ravel()
                got = np.array([np.abs(g).sum(), np.sqrt((g ** 2).sum())])
                assert np.allclose(got, v[1:], rtol=1e-4, atol=2e-6), k
