"""A reduced, fixed-seed slice of every job of tests/fuzz_gpu.py under ``pytest -m gpu`` (the full sweep --
``python tests/fuzz_gpu.py 300`` -- stays a manual run): randomised target assignment (ragged batches, duplicate / touching /
tiny / out-of-image GT, label and encode modes, culled and dense), top-k and NMS with tied scores and duplicate boxes under the
three criteria, tie blocks around the selection cut on every cluster width, the fused detect pipeline, the MultiBox loss
selection and the WIDER AP counters -- each at the bar of the parity tests (indices, labels, keep lists and counters bit-exact;
coordinates rtol 1e-5 / atol 1e-6)."""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu

SLICE = {"assign": 12, "nms/topk": 12, "tieblock": 2, "detect": 12, "loss": 8, "eval": 8}


@pytest.fixture(scope="module")
def fuzz():
    spec = importlib.util.spec_from_file_location("fuzz_gpu", os.path.join(os.path.dirname(os.path.abspath(__file__)), "fuzz_gpu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("job", sorted(SLICE))
def test_fuzz_slice(fuzz, job):
    fuzz.run_job(job, SLICE[job])
