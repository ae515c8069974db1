import glob
import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

warnings.filterwarnings("ignore", message=".*Sparse CSR tensor support is in beta.*")
warnings.filterwarnings("ignore", message=".*Sparse invariant checks.*")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


@pytest.fixture(params=GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def golden(request):
    z = np.load(request.param)
    return {k: z[k] for k in z.files}
