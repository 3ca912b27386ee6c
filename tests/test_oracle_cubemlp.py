"""Pin the CubeMLP oracle to the reference (MLPProcess.py:126-137)."""
import ast

import numpy as np
import pytest

from conftest import load_golden, rel_err
from oracle import cubemlp_oracle as C
from oracle import params as P

CUBE = load_golden("cubemlp")


def case_inputs(rec):
    c = ast.literal_eval(str(rec["cfg"]))
    seed = int(rec["seed"])
    blocks = P.cubemlp_params(seed, c["d_in"], c["d_hiddens"], c["d_outs"], c["bias"], c["ln_first"], c["res"])
    x = P.features(seed + 1, c["bs"] * c["d_in"][0] * c["d_in"][1], c["d_in"][2]).reshape(
        c["bs"], c["d_in"][0], c["d_in"][1], c["d_in"][2])
    oshape = (c["bs"], *c["d_outs"][-1])
    w = P.features(seed + 2, int(np.prod(oshape[:-1])), oshape[-1]).reshape(oshape)
    return c, blocks, x, w


@pytest.mark.parametrize("case", sorted(CUBE))
def test_cubemlp_matches_reference(case):
    rec = CUBE[case]
    c, blocks, x, w = case_inputs(rec)
    y, caches = C.encoder_forward(blocks, x, c["act"], c["ln_first"], c["res"])
    gx, pg = C.encoder_backward(caches, w.astype(np.float64), c["ln_first"], c["res"])
    big = c["d_in"][2] >= 128
    ys, gxs = (y[:, :, :, ::8], gx[:, ::5, :, ::8]) if big else (y, gx)
    assert rel_err(ys, rec["y"]) < 5e-5
    assert rel_err(gxs, rec["gx"]) < 1e-4
    for k, v in rec.items():
        if k.startswith("pg__"):
            assert np.abs(pg[k[4:]] - v).max() <= 1e-4 * np.abs(v).max() + 1e-6, k
        elif k.startswith("pgs__"):
            g = pg[k[5:]].ravel()
            got = np.array([np.abs(g).sum(), np.sqrt((g ** 2).sum())])
            assert np.allclose(got, v[1:], rtol=2e-4, atol=1e-5), k
