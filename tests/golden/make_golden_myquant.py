"""Generates tests/golden/myquant.npz from the reference's own NumPy restatement of the 80-candidate MSE clip search
(/root/reference/myQuant.py:6-45, hard-wired to 4 bits) — an independent cross-check of K2a / oracle.mse_search that does
not go through torch. Run in the build container only:  python tests/golden/make_golden_myquant.py
NumPy >= 2 keeps `float32_scalar * python_float` in float32 (NEP 50), so these values follow the same fp32 discipline as
the torch path except for the score reduction (np.mean over fp32 = pairwise fp32 sum)."""
import contextlib
import io
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import REF, save  # noqa: E402


def main():
    sys.path.insert(0, REF)
    import myQuant
    rng = np.random.default_rng(424)
    w = (rng.standard_normal((12, 8, 3, 3)) * 0.05).astype(np.float32)
    w[3] *= 4.0; w[7] += 0.02                     # a wide channel and an offset one
    with contextlib.redirect_stdout(io.StringIO()):
        delta, zp = myQuant.init_delta(w)
    save("myquant", w=w, delta=np.asarray(delta, dtype=np.float64).reshape(-1), zero_point=np.asarray(zp, dtype=np.float64).reshape(-1),
         numpy_version=np.array(np.__version__))


if __name__ == "__main__":
    main()
