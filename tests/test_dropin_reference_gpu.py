"""GPU: the real drop-in.  The reference's OWN ``model/DrugLAMP*.py`` classes (unmodified; imported
from /root/reference in the build container, from the byte-compiled ``oracle/_ref`` on the GPU box)
are constructed after ``druglamp_b200.patch_reference()``, so ``DrugLAMPBase.__init__``
(model/basic_model.py:75-121) builds the sm_100a modules, and their ``forward`` (model/DrugLAMP.py:8-79,
model/DrugLAMP2C2P.py:8-89) plus the calls ``trainer.py:196-213`` makes run on cuda:0.  Results are
held against the fixtures recorded from the UNPATCHED reference on the CPU (make_golden.py).

Each case runs in its own process: patching rebinds names inside the reference's modules."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.util import load_golden

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(kind, fixture, mode):
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("neither /root/reference nor oracle/_ref (python oracle/build_ref.py) is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_runner.py"), kind, fixture, mode],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


@pytest.mark.parametrize("kind,fixture", [("DrugLAMP2C2P", "druglamp2c2p_train_b16.npz"),
                                          ("DrugLAMPwoLLM", "druglampwollm_train_b12.npz"),
                                          ("DrugLAMP", "druglamp_eval_b2.npz")])
def test_reference_forward_runs_on_patched_modules_fp32(kind, fixture):
    fx = load_golden(fixture)
    o = _run(kind, fixture, "f32")
    assert o["launches"] > 100, "the sm_100a kernels did not run"
    assert o["mask_equal"] and o["finite_grads"]
    sc = np.asarray(o["score"])
    ref = fx["score"].flatten()
    assert np.abs(sc - ref).max() <= 1e-3 * np.abs(ref).max(), (sc[:4], ref[:4])
    assert abs(o["loss"] - float(fx["loss"])) <= 1e-3 * abs(float(fx["loss"]))
    assert o["worst_grad"][1] <= 1e-2, o["worst_grad"]
    assert np.isfinite(o["ssl"]).all()
    if kind == "DrugLAMP2C2P":
        assert abs(o["cm_loss"] - float(fx["cm_losses"][0])) <= 2e-3 * abs(float(fx["cm_losses"][0]))
        assert abs(o["cm_margin_after_step"] - float(fx["cm_margins"][1])) < 1e-12


def test_reference_forward_runs_on_patched_modules_bf16():
    """patch_reference() + set_compute_dtype(bf16).  The replaced modules
    compute in bf16 and hand fp32 back to the reference's own fp32 layers between them
    (modules.RETURN_CALLER_DTYPE); eval-mode logits hold the 2e-2 bar."""
    fixture = "druglamp_eval_b2.npz"
    fx = load_golden(fixture)
    o = _run("DrugLAMP", fixture, "bf16")
    assert o["score_dtype"] == "torch.float32" and o["vd_dtype"] == "torch.float32"
    assert o["mask_equal"] and o["finite_grads"]
    sc = np.asarray(o["score"])
    ref = fx["score"].flatten()
    assert np.abs(sc - ref).max() <= 2e-2 * np.abs(ref).max(), (sc, ref)
    assert abs(o["loss"] - float(fx["loss"])) <= 2e-2 * abs(float(fx["loss"]))
