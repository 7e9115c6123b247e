"""Golden fixture of the conv front-end restatement (oracle/frontend_oracle.py, numpy fp64): `python -m
tests.golden.make_golden_frontend` rewrites tests/golden/frontend_17x13.json.  Like the other fixtures it pins the ORACLE
against regressions, not TensorFlow (no TF in this image).  Variables are regenerated from the seed by
tests.test_frontend._randomised (he_normal kernels, perturbed biases / gammas / betas); the image from a torch CPU generator."""
from __future__ import annotations

import json
import os

import torch

from tests.golden.make_golden import _rec

HERE = os.path.dirname(os.path.abspath(__file__))
CONFIG = {"scope": "Generator/Generator", "seed": 31, "image_seed": 5, "shape": [2, 17, 13, 3]}


def compute():
    from oracle import frontend_oracle as FO
    from tests.test_frontend import _randomised
    net = _randomised(CONFIG["scope"], CONFIG["seed"])
    images = torch.randn(*CONFIG["shape"], generator=torch.Generator().manual_seed(CONFIG["image_seed"]), dtype=torch.float64)
    out = FO.front_end({k: v.numpy() for k, v in net.tf_variables().items()}, CONFIG["scope"], images.numpy())
    return {"config": CONFIG, "output_shape": list(out.shape), "downsampled": _rec("downsampled", torch.from_numpy(out))}


if __name__ == "__main__":
    with open(os.path.join(HERE, "frontend_17x13.json"), "w") as f:
        json.dump(compute(), f, indent=1)
    print("wrote frontend_17x13.json")
