"""pytest configuration: the ``gpu`` marker and golden-fixture helpers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(group):
    """tests/golden/<group>.npz -> {case: {key: ndarray}} (written by oracle/gen_golden.py)."""
    z = np.load(os.path.join(GOLDEN, group + ".npz"), allow_pickle=False)
    out = {}
    for k in z.files:
        case, key = k.split("::", 1)
        out.setdefault(case, {})[key] = z[k]
    return out


def cfg_of(rec):
    return {k[4:]: (v.item() if v.shape == () else v) for k, v in rec.items() if k.startswith("cfg_")}


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(np.abs(b).max() if b.size else 0.0, 1e-30)
    return (np.abs(a - b).max() if b.size else 0.0) / denom


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(group):
        if group not in cache:
            cache[group] = load_golden(group)
        return cache[group]
    return get


TINY_BERT = dict(vocab_size=120, hidden_size=32, num_hidden_layers=1, num_attention_heads=2, intermediate_size=64,
                 max_position_embeddings=64, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)


@pytest.fixture
def tiny_bert():
    """``from_pretrained`` -> a tiny random-init BERT (the one oracle/gen_golden.py built the model goldens with); the
    BERT encoder is out of scope and there are no cached weights offline."""
    import transformers
    classes = (transformers.BertConfig, transformers.BertModel)
    saved = [cls.__dict__.get("from_pretrained") for cls in classes]
    transformers.BertConfig.from_pretrained = staticmethod(
        lambda *a, **k: transformers.BertConfig(output_hidden_states=True, **TINY_BERT))
    transformers.BertModel.from_pretrained = staticmethod(lambda *a, config=None, **k: transformers.BertModel(config))
    yield TINY_BERT
    for cls, old in zip(classes, saved):                      # inherited classmethod: remove the override again
        if old is None:
            delattr(cls, "from_pretrained")
        else:
            setattr(cls, "from_pretrained", old)
